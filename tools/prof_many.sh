# usage: prof_many.sh name:kernel-regex[:skip] ...   (ncu --set full, one launch each, 100M-read replica)
for spec in "$@"; do
  O=${spec%%:*}; rest=${spec#*:}; K=${rest%%:*}; S=${rest#*:}; [ "$S" = "$rest" ] && S=0
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o gpurun_out/$O \
     python bench.py --reads 100000000 --cells 2500 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/$O.log 2>&1
  ncu -i gpurun_out/$O.ncu-rep --page raw --csv > gpurun_out/$O.csv 2>/dev/null
  rm -f gpurun_out/$O.ncu-rep
  ls -la gpurun_out/$O.csv
done
