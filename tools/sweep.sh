run() { echo "=== $*"; env "$@" DGE_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | grep -E "sc\(n=3" | tail -7 | grep -E "dedup" | sed -E 's/\[dge\]   sc\(n=[0-9]+\) //'; }
run DGE_MS_VARIANT=1
run DGE_MS_VARIANT=2
DGE_MS_VARIANT=2 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "medium or golden or invariants_20m" 2>&1 | tail -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -1
