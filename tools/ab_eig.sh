mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vision.py tests/test_gpu_fullsize.py tests/test_gpu_tracker.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu_r2b.log; tail -5 gpurun_out/pytest_gpu_r2b.log
for wv in 3 4; do
OFB_EIG_WAVES=$wv timeout 300 python bench.py --steps 10 --warmup 3 --no-mc --no-cpu > gpurun_out/bench_w$wv.log 2>&1; echo "exit $?"
done
python - <<'PY'
import json
for f in ("bench_w2","bench_w3","bench_w4"):
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().split("\n")[-1])
        print(f, round(d["value"]), "pairs/s", d["roofline"]["stage_ms"], "e2e", round(d["e2e"]["value"]))
    except Exception as e: print(f, "ERR", e)
PY
