python -m pytest tests -m gpu -x -q 2>&1 | tail -1
run2() { echo "=== $*"; env "$@" timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel'], round(d['roofline']['ms_per_launch'],3), round(d['roofline']['frac'],3), d['roofline_second']['kernel'], round(d['roofline_second']['ms_per_launch'],3))"; }
run2 DGE_FILL_VARIANT=0
run2 DGE_FILL_VARIANT=1
run2 DGE_FILL_VARIANT=2
