for sp in 0 1 0 1; do
OFB_TRACKER_SPLIT_SOLVE=$sp timeout 300 python bench.py --workload c1 --steps 400 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); l=d['lifecycle_step']; print('split $sp c1 lifecycle (400 steps)', l['resident_ms_per_frame'], l['host_call_ms_p50'])"
done
