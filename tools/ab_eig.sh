mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vision.py tests/test_gpu_fullsize.py tests/test_gpu_tracker.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu_r2b.log; tail -5 gpurun_out/pytest_gpu_r2b.log
for cfg in "0 3" "1 3" "1 4" "1 5"; do
set -- $cfg
OFB_EIG_ORDER=$1 OFB_EIG_WAVES=$2 timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-mc --no-cpu > gpurun_out/bench_o$1w$2.log 2>/dev/null; echo "exit $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_o$1w$2.log").read().strip().split("\n")[-1])
print("order $1 waves $2", round(d["value"]), "pairs/s", d["roofline"]["stage_ms"], "e2e", round(d["e2e"]["value"]))
PY
done
