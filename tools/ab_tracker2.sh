for mode in "0 0" "1 0" "1 1"; do
set -- $mode
export OFB_TRACKER_EARLY_PYR=$1 OFB_TRACKER_DEFER_TOPUP=$2
echo "== early $1 defer $2"
timeout 200 python tools/tracker_latency.py 1920 1080 1000 2>&1 | head -4
done
