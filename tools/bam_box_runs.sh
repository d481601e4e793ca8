#!/bin/bash
# The GPU-box commands behind profiles/r2_bam_ingest_final*.log and r2_bam_ingest_name_mode.log (gpurun -- bash tools/bam_box_runs.sh <1|2|3>).
# gpurun_in/bam_rate.bam = tools/bam_rate.py 2000000 (its /tmp/bam_rate.bam); gpurun_in/name_mode.bam = the generator in the header of run 3.
case "$1" in
1)
  # BAM -> count matrix with the pipelined reader (10 M reads = the 2 M-read file five times), bulk vs one-read path on one file, then the GPU tests
  cd /root/repo
  EXE=dropest_b200/lib/test_bam_pipeline
  B=gpurun_in/bam_rate.bam
  OUT=gpurun_out/r2_bam_final.log
  {
  nproc
  for rep in 1 2; do
    $EXE - 5 5 "" "" "" $B $B $B $B $B | grep -E "timing|stats|error" | sed 's/^/bulk x5      /'
  done
  DGE_BAM_ZLIB_INFLATE=1 $EXE - 5 5 "" "" "" $B $B $B $B $B | grep -E "timing|stats|error" | sed 's/^/bulk x5 zlib /'
  $EXE - 5 5 "" "" "" $B | grep -E "timing|stats|error" | sed 's/^/bulk x1      /'
  DGE_BAM_ONE_BY_ONE=1 $EXE - 5 5 "" "" "" $B | grep -E "timing|stats|error" | sed 's/^/one-read x1  /'
  $EXE - 5 5 "" "" "" $B | grep -v timing | md5sum
  DGE_BAM_ONE_BY_ONE=1 $EXE - 5 5 "" "" "" $B | grep -v timing | md5sum
  } > $OUT 2>&1
  timeout 100 python -m pytest tests/test_bam_ingest.py tests/test_bam_output.py tests/test_facade.py tests/test_n_reads.py tests/test_rpupc.py -q -m gpu -x 2>&1 | tail -5 >> $OUT
  timeout 250 python -m pytest tests/test_gpu_parity.py tests/test_bench_shape.py tests/test_collisions.py tests/test_dist_sim.py -q -m gpu -x 2>&1 | tail -5 >> $OUT
  tail -30 $OUT
  ;;
2)
  cd /root/repo
  EXE=dropest_b200/lib/test_bam_pipeline
  B=gpurun_in/bam_rate.bam
  F=""; for i in $(seq 20); do F="$F $B"; done
  {
  nproc; lscpu | grep "Model name"
  for rep in 1 2 3; do $EXE - 5 5 "" "" "" $F | grep -E "timing|stats|error" | sed 's/^/bulk x20      /'; done
  DGE_BAM_ZLIB_INFLATE=1 $EXE - 5 5 "" "" "" $F | grep -E "timing|stats|error" | sed 's/^/bulk x20 zlib /'
  } > gpurun_out/r2_bam_final_x20.log 2>&1
  cat gpurun_out/r2_bam_final_x20.log
  ;;
3)
  cd /root/repo
  EXE=dropest_b200/lib/test_bam_pipeline
  B=gpurun_in/name_mode.bam
  G=tests/golden/ref_gtf/gtf_test.gtf.gz
  {
  for mode in "tag" "gtf"; do
    if [ $mode = gtf ]; then export DGE_BAM_GENES=$G; fi
    DGE_BAM_NAME_MODE=1 $EXE - 2 2 "" "" "" $B > /tmp/bulk.txt
    DGE_BAM_NAME_MODE=1 DGE_BAM_ONE_BY_ONE=1 $EXE - 2 2 "" "" "" $B > /tmp/one.txt
    echo "name mode + $mode: bulk $(grep -v timing /tmp/bulk.txt | md5sum | cut -c1-12) one-read $(grep -v timing /tmp/one.txt | md5sum | cut -c1-12) lines $(wc -l < /tmp/bulk.txt)"
    grep -E "timing|stats|error" /tmp/bulk.txt | sed 's/^/  bulk     /'
    grep -E "timing|stats|error" /tmp/one.txt | sed 's/^/  one-read /'
  done
  } > gpurun_out/r2_bam_name_mode.log 2>&1
  cat gpurun_out/r2_bam_name_mode.log
  ;;
esac
