mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tracker.py tests/test_gpu_round2.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -8
for mode in "0 0" "1 0" "1 1"; do
set -- $mode
export OFB_TRACKER_EARLY_PYR=$1 OFB_TRACKER_DEFER_TOPUP=$2
timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu --no-mc 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('early $1 defer $2 c2 lifecycle', d['lifecycle']['ms_per_frame'], d['lifecycle']['min_tracked'], 'value', d['value'])"
timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); l=d['lifecycle']; print('early $1 defer $2 c5 lifecycle', l['value'], l['ms_per_step'], 'bgr', l['bgr_frames']['value'], 'e2e', l['e2e']['value'])"
timeout 300 python bench.py --workload c1 --steps 50 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); l=d['lifecycle_step']; print('early $1 defer $2 c1 lifecycle', l['resident_ms_per_frame'], l['host_call_ms_p50'])"
done
