python -m pytest tests -m gpu -x -q 2>&1 | tail -1
run() { echo "=== $*"; env "$@" DGE_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | grep -E "sc\(n=3" | tail -7 | grep -E "${PAT:-l1 hist|l2 hist|l2 scan|dedup}" | tr '\n' ' '; echo; }
run DGE_STAGED=0
run DGE_STAGED=3
run DGE_STAGED=3 DGE_STILE=1
run DGE_STAGED=3 DGE_STILE=2
run DGE_STAGED=3 DGE_S2TILE=1
run DGE_STAGED=3 DGE_S2TILE=2
run DGE_STAGED=3 DGE_L1_TARGET=375000
run DGE_STAGED=3 DGE_L1_TARGET=375000 DGE_S2TILE=2
run DGE_STAGED=3 DGE_L1_TARGET=750000
run DGE_STAGED=3 DGE_L1_TARGET=750000 DGE_S2TILE=2
