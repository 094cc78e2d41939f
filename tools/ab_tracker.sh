mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tracker.py tests/test_gpu_round2.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -8
for sp in 0 1 0 1; do
export OFB_TRACKER_SPLIT_SOLVE=$sp
timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu --no-mc 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('split $sp c2 lifecycle', d['lifecycle']['ms_per_frame'], d['lifecycle']['min_tracked'], 'value', round(d['value']))"
timeout 300 python bench.py --workload c1 --steps 50 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); l=d['lifecycle_step']; print('split $sp c1 lifecycle', l['resident_ms_per_frame'], l['host_call_ms_p50'])"
done
for sp in 0 1; do
OFB_TRACKER_SPLIT_SOLVE=$sp timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); l=d['lifecycle']; print('split $sp c5 lifecycle', round(l['value']), l['ms_per_step'], 'bgr', round(l['bgr_frames']['value']), 'e2e', round(l['e2e']['value']))"
done
