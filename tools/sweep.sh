run() { echo "=== $*"; env "$@" DGE_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | grep -E "sc\(n=3|ms_per_step" | tail -8 | grep -E "${PAT:-l2 |ms_per_step}" | sed -E 's/\[dge\]   sc\(n=[0-9]+\) //; s/.*"ms_per_step": ([0-9.]+).*/ms_per_step \1/' | tr '\n' ' '; echo; }
run DGE_L2_BUCKET=2
run DGE_L2_BUCKET=3
run DGE_L2_BUCKET=4
run DGE_L2_BUCKET=1
