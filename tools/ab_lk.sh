# A/B of build-time variants (-DOFB_LKF_WARPS / -DOFB_MK_WARPS / -DOFB_SOLVE_THREADS / -DOFB_PYR_THREADS ...): build the
# variant object(s) by hand, link them into <package>/libofb200_<tag>.so and list the tags here; "base" is the in-tree library.
#   usage (on the GPU box): bash tools/ab_lk.sh base <tag> base <tag>
PKG=$PWD/drone-stabilisation-using-optical-flow-gps-and-inertial-sensors_b200
for v in "$@"; do
if [ $v == base ]; then unset OFB200_LIB; else export OFB200_LIB=$PKG/libofb200_$v.so; fi
timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-mc --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('$v c2', round(d['value']), d['roofline']['stage_ms'], round(d['track_solve']['value']), round(d['lifecycle']['ms_per_frame'],5))"
done
