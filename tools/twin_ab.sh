mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vision.py tests/test_gpu_fullsize.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-mc 2>gpurun_out/bench.err | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['stage_ms'])
"
