OFB_SELECT_TRACE=1 timeout 100 python bench.py --workload c4 --steps 3 --warmup 1 --no-cpu 2>&1 | grep "select trace" | grep -v "ncand 0 " | tail -3
