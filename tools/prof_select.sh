# ncu --set full capture of the corner-selection kernel in the C4 (1080p, 5000 corners) configuration
mkdir -p gpurun_out
PROF="python bench.py --workload c4 --steps 3 --warmup 1 --no-cpu"
timeout 300 $PROF > gpurun_out/prof_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:select_kernel -s 2 -c 1 -f -o gpurun_out/${1:-prof_select_r2} $PROF > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
