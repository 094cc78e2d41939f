mkdir -p gpurun_out
for wv in 2 3 4 6; do
echo "== waves $wv"
OFB_TWIN_CHUNKS=0 OFB_EIG_WAVES=$wv timeout 300 ncu --metrics gpu__time_duration.sum,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct,launch__grid_size -k regex:eig_march --clock-control none -c 4 --csv python tools/profile_pairs.py --batch 32 --steps 1 --mc-trials 1000 2>/dev/null | grep -v "^==" | python -c "
import sys,csv
rows=list(csv.reader(sys.stdin))
hdr=None
for r in rows:
    if 'Metric Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); print(d['ID'], d['Metric Name'], d['Metric Value'])
"
done
