run() { echo "=== $*"; env "$@" DGE_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | grep -E "sc\(n=3|ms_per_step" | tail -8 | grep -E "l2 hist|l2 scan|dedup|compact|ms_per_step" | sed -E "s/.*\"ms_per_step\": ([0-9.]+).*/ms_per_step \1/"; }
run DGE_SC_TARGET=1024 DGE_SC_HT=2048
run DGE_SC_TARGET=768 DGE_SC_HT=2048
run DGE_SC_TARGET=640 DGE_SC_HT=1024
run DGE_SC_TARGET=512 DGE_SC_HT=1024
run DGE_SC_TARGET=384 DGE_SC_HT=1024
run DGE_SC_TARGET=320 DGE_SC_HT=512
