# Full-size (BASELINE configs[1]) ncu evidence for the round: launch list + --set full captures of the two heaviest kernels.
# Numbers printed by runs under ncu are never bench values.
set -x
bash tools/launches.sh r1_launches_c2 60 > gpurun_out/r1_launches_c2.summary.txt 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_sort_dedup|k_dedup_sort<false|k_fill_compact" -c 10 -f -o gpurun_out/r1_full_c2 \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r1_full_c2.log 2>&1
ncu -i gpurun_out/r1_full_c2.ncu-rep --page raw --csv > gpurun_out/r1_full_c2.csv 2>/dev/null
ls -la gpurun_out/r1_full_c2.*
rm -f gpurun_out/r1_full_c2.ncu-rep
