# ncu --set full on the heaviest kernels of one 100M-read step (1 GPU; never a timing run)
set -x
for k in k_dedup_sort k_l1_scatter k_fill_compact k_l2_pass; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 2 -f -o gpurun_out/prof_$k \
     python bench.py --reads 100000000 --cells 2500 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_$k.log 2>&1
  tail -2 gpurun_out/prof_$k.log
done
ls -la gpurun_out
