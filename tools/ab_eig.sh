mkdir -p gpurun_out
for wv in 2 3 4 6; do
OFB_EIG_WAVES=$wv timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-mc --no-cpu > gpurun_out/bench_w$wv.log 2>/dev/null; echo "exit $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_w$wv.log").read().strip().split("\n")[-1])
print("waves $wv", round(d["value"]), "pairs/s", d["roofline"]["stage_ms"], "e2e", round(d["e2e"]["value"]))
PY
done
