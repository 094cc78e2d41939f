for wl in c1 c4; do
OFB_SELECT_TRACE=1 timeout 100 python bench.py --workload $wl --steps 3 --warmup 1 --no-cpu 2>&1 | grep "select trace" | grep -v "ncand 0 " | tail -2
done
