python -m pytest tests -m gpu -x -q 2>&1 | tail -1
run() { echo "=== $*"; env "$@" DGE_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | grep -E "sc\(n=3" | tail -7 | grep -E "${PAT:-l1 hist|plan|l2 hist|l2 scan|dedup}" ; }
run DGE_TILE=2
run DGE_TILE=1
run DGE_TILE=0
run DGE_TILE=4
run DGE_TILE=3
