timeout 600 python -m pytest tests/test_gpu_vision.py tests/test_gpu_round2.py tests/test_gpu_tracker.py tests/test_gpu_fullsize.py tests/test_gpu_random.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
for v in t128 t256 t128 t256; do
if [ $v == t128 ]; then unset OFB200_LIB; else export OFB200_LIB=$PWD/drone-stabilisation-using-optical-flow-gps-and-inertial-sensors_b200/libofb200_$v.so; fi
timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-mc --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('$v c2', round(d['value']), d['roofline']['stage_ms'], round(d['track_solve']['value']), round(d['lifecycle']['ms_per_frame'],5))"
done
