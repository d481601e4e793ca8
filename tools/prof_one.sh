# usage: prof_one.sh <kernel-regex> <out-name> [env...]   -- one ncu --set full capture (never a timing run)
K=$1; O=$2; shift 2
env "$@" timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/$O \
   python bench.py --reads 100000000 --cells 2500 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/$O.log 2>&1
tail -2 gpurun_out/$O.log | cut -c1-300
ncu -i gpurun_out/$O.ncu-rep --page raw --csv > gpurun_out/$O.csv 2>/dev/null
ls -la gpurun_out/$O.*
