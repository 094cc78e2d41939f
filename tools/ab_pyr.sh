timeout 600 python -m pytest tests/test_gpu_vision.py tests/test_gpu_round2.py tests/test_gpu_tracker.py tests/test_gpu_fullsize.py tests/test_gpu_random.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
for i in 1 2; do
timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-mc --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('c2', round(d['value']), d['roofline']['stage_ms'], round(d['track_solve']['value']), round(d['lifecycle']['ms_per_frame'],5))"
done
timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); l=d['lifecycle']; print('c5', round(d['value']), 'lifecycle', round(l['value']), 'bgr', round(l['bgr_frames']['value']))"
