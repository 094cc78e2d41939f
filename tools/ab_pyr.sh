for st in 512 256; do
OFB_SELECT_THREADS=$st timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('c5 select_threads $st', round(d['value']), round(d['lifecycle']['value']))"
done
timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-mc --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('c2 default', round(d['value']), d['roofline']['stage_ms'], round(d['independent_pairs']['value']), round(d['e2e']['value']))"
timeout 600 python -m pytest tests/test_gpu_vision.py tests/test_gpu_fullsize.py tests/test_gpu_pairs_mc.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
