python -m pytest tests -m gpu -x -q 2>&1 | tail -1
run() { echo "=== $*"; env "$@" DGE_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | grep -E "sc\(n=3" | tail -7 | grep -E "${PAT:-plan|l2 hist|l2 scan|dedup|compact}" | sed -E 's/\[dge\]   sc\(n=[0-9]+\) //' | tr '\n' ' '; echo; }
run DGE_MS_NO_WARP=1
run DGE_SC_TARGET=832
run DGE_SC_TARGET=448
run DGE_SC_TARGET=416
run DGE_SC_TARGET=384
run DGE_SC_TARGET=352
run DGE_SC_TARGET=416 DGE_MS_BPS=256
