# ncu launch list (per-launch gpu__time_duration, cold-cache + serialised: shares only) of one full-size step
O=${1:-launches}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/$O.csv \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/$O.log 2>&1
python tools/launch_summary.py gpurun_out/$O.csv k_synth | head -${2:-45}
