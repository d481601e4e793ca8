set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dedup_sort -s 0 -c 1 -f -o gpurun_out/prof_dedup_r1b \
   python bench.py --reads 100000000 --cells 2500 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_dedup_r1b.log 2>&1
tail -2 gpurun_out/prof_dedup_r1b.log
