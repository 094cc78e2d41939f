for i in 1 2; do
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); m=d['mc']; print('mc', m['value'], m.get('value_fp64'), m.get('value_without_R'), 'c2', d['value'])"
done
timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); l=d['lifecycle']; print('c5 lifecycle', l['value'], 'bgr', l['bgr_frames']['value'])"
timeout 300 python -m pytest tests/test_gpu_pairs_mc.py tests/test_gpu_round2.py tests/test_gpu_tracker.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
