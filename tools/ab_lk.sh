mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vision.py tests/test_gpu_fullsize.py tests/test_gpu_tracker.py tests/test_gpu_pairs_mc.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu_r2c.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu_r2c.log; tail -12 gpurun_out/pytest_gpu_r2c.log
OFB_LK_V1=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-mc --no-cpu > gpurun_out/bench_lkv1.log 2>&1; echo "exit $?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-mc --no-cpu > gpurun_out/bench_lkv2.log 2>&1; echo "exit $?"
python - <<'PY'
import json
for f in ("bench_lkv1","bench_lkv2"):
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().split("\n")[-1])
        print(f, round(d["value"]), "pairs/s", d["roofline"]["stage_ms"], "e2e", round(d["e2e"]["value"]), "trk", round(d["track_solve"]["value"]), d["check"])
    except Exception as e: print(f, "ERR", e)
PY
bash tools/prof_eig.sh "lk_track" 1 prof_r2_lk
