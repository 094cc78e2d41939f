for v in base s512 base s512; do
if [ $v == base ]; then unset OFB200_LIB; else export OFB200_LIB=$PWD/drone-stabilisation-using-optical-flow-gps-and-inertial-sensors_b200/libofb200_$v.so; fi
for wl in c1 c4; do timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('$v $wl', round(d['value'],4), d['stage_ms_serial']['solve'], round(d['lifecycle_step']['resident_ms_per_frame'],4))"; done
timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-mc --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('$v c2', round(d['value']), d['roofline']['stage_ms']['solve'], round(d['track_solve']['value']), round(d['lifecycle']['ms_per_frame'],5))"
done
