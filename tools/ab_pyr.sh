for wv in 3 2 3 2; do
OFB_EIG_WAVES=$wv timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-mc --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('waves $wv', round(d['value']), d['roofline']['stage_ms'], round(d['independent_pairs']['value']), round(d['e2e']['value']))"
done
