mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vision.py tests/test_gpu_fullsize.py tests/test_gpu_tracker.py tests/test_gpu_random.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu_r2b.log; tail -5 gpurun_out/pytest_gpu_r2b.log
for e in 0 1 0 1; do
OFB_EIG_EDGE_FAST=$e timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-mc --no-cpu > gpurun_out/bench_e$e.log 2>/dev/null; echo "exit $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_e$e.log").read().strip().split("\n")[-1])
print("edge_fast $e", round(d["value"]), "pairs/s", d["roofline"]["stage_ms"], "e2e", round(d["e2e"]["value"]))
PY
done
