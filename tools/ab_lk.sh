timeout 900 python -m pytest tests/test_gpu_vision.py tests/test_gpu_fullsize.py tests/test_gpu_tracker.py tests/test_gpu_random.py tests/test_gpu_pairs_mc.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -3
timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); l=d['lifecycle']; print('c5', d['value'], 'lifecycle', l['value'], 'bgr', l['bgr_frames']['value'])"
for wl in c1 c4; do timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('$wl', d['value'], d['stage_ms_serial'], d['lifecycle_step']['resident_ms_per_frame'])"; done
