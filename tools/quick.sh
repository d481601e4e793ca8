# quick GPU check: parity tests, then a traced bench (stage marks on stderr) -- run through gpurun
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
DGE_TRACE=1 python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/q_bench.log 2> gpurun_out/q_bench.err
grep -E "^\[dge\]" gpurun_out/q_bench.err | tail -${TAILN:-38}
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','stage_ms_per_step')}); print(d['roofline'])"
