for wv in 4 1 4 1; do
if [ $wv == 4 ]; then unset OFB200_LIB; else export OFB200_LIB=$PWD/drone-stabilisation-using-optical-flow-gps-and-inertial-sensors_b200/libofb200_w$wv.so; fi
timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-mc --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('mk_warps $wv', round(d['value']), d['roofline']['stage_ms'], round(d['independent_pairs']['value']))"
done
