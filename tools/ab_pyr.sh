for sp in 0 40 30 60 70; do
OFB_TWIN_SPLIT=$sp timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-mc --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('split $sp c2', round(d['value']), round(d['independent_pairs']['value']), round(d['track_solve']['value']), d['check']['max_abs_v_error_vs_truth'])"
done
