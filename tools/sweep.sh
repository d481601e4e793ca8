timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
run() { echo "=== $*"; env "$@" DGE_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | grep -E "\[dge\]|ms_per_step" | tail -22 | sed -E 's/.*"ms_per_step": ([0-9.]+).*"stage_ms_per_step": (\{[^}]*\}).*/ms_per_step \1 \2/' ; }
run A=1
run DGE_SC_TARGET=1024 DGE_SC_HT=2048
run DGE_SC_TARGET=768 DGE_SC_HT=2048
run DGE_SC_TARGET=512 DGE_SC_HT=1024
