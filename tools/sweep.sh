run() { echo "=== $*"; env "$@" DGE_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | grep -E "sc\(n=3" | tail -7 | grep -E "l2 hist|l2 scan|dedup|compact"; }
run DGE_DEDUP_THREADS=256
run DGE_DEDUP_THREADS=512
run DGE_DEDUP_THREADS=1024
run DGE_DEDUP_THREADS=512 DGE_SC_TARGET=1536 DGE_SC_HT=4096
run DGE_DEDUP_THREADS=1024 DGE_SC_TARGET=1536 DGE_SC_HT=4096
run DGE_DEDUP_THREADS=512 DGE_SC_TARGET=512 DGE_SC_HT=1024
