for wl in c1 c4; do
OFB_SELECT_TRACE=1 timeout 100 python bench.py --workload $wl --steps 3 --warmup 1 --no-cpu 2>&1 | grep "select trace" | grep -v "ncand 0 " | tail -2
done
for cs in 2 4; do
OFB_SELECT_CLUSTER=$cs timeout 100 python bench.py --workload c1 --steps 50 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('c1 cluster $cs p50', d['value'], d['stage_ms_serial'])"
OFB_SELECT_TRACE=1 OFB_SELECT_CLUSTER=$cs timeout 100 python bench.py --workload c1 --steps 3 --warmup 1 --no-cpu 2>&1 | grep "select trace" | grep -v "ncand 0 " | tail -1
done
