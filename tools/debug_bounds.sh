# The whole GPU suite against the debug build (device-side index assertions, OFB_DEV_ASSERT in csrc/common.cuh)
mkdir -p gpurun_out
export OFB200_LIB=$PWD/drone-stabilisation-using-optical-flow-gps-and-inertial-sensors_b200/libofb200_dbg.so
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu_dbg.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu_dbg.log
grep -c OFB_DEV_ASSERT gpurun_out/pytest_gpu_dbg.log; tail -6 gpurun_out/pytest_gpu_dbg.log
unset OFB200_LIB
timeout 600 python -m pytest tests/test_gpu_pairs_mc.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); m=d['mc']; print('mc', m['value'], m.get('value_fp64'), m.get('value_without_R'), 'c2', d['value'])"
