for mb in 3 4 5 6; do
NMB_EVAL_MINB=$mb NMB_STAGE_TIMING=1 timeout 120 python bench.py --steps 200 --warmup 5 --no-cpu 2>&1 | grep -E "stage timing" | sed -n "2p;4p" | sed "s/^/[minb=$mb] /"
done
