for ts in 8 16 24 40 64; do
NMB_ITEM_TS=$ts NMB_STAGE_TIMING=1 timeout 120 python bench.py --steps 200 --warmup 5 --no-cpu 2>&1 | grep -E "stage timing" | sed -n "2p;4p" | sed "s/^/[item_ts=$ts] /"
done
