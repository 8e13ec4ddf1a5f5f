#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for spec in "bunny4k 1 0" "bunny4k 2 0" "bunny4k 4 0" "bunny4k 8 0" "synthetic10m 1 0" "synthetic10m 8 0" "bunny4k 1 32" "synthetic10m 1 32"; do
  set -- $spec
  python tools/one_frame.py $1 2 $2 $3 > /dev/null 2>&1 || { echo "one_frame $spec failed"; continue; }
  timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02b_ncu_$1_w$2_f$3.csv python tools/one_frame.py $1 2 $2 $3 > gpurun_out/r02b_ncu_$1_w$2_f$3.log 2>&1
  echo "ncu $spec exit $?"; tail -1 gpurun_out/r02b_ncu_$1_w$2_f$3.log
done
python tools/one_frame.py bunny4k 2 1 0 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:pixel_kernel -s 1 -c 1 -o gpurun_out/r02b_pixel_bunny4k python tools/one_frame.py bunny4k 2 1 0 > gpurun_out/r02b_ncu_full.log 2>&1; echo "ncu full exit $?"
python tools/one_frame.py synthetic10m 2 1 0 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:pixel_kernel -s 1 -c 1 -o gpurun_out/r02b_pixel_synthetic10m python tools/one_frame.py synthetic10m 2 1 0 > gpurun_out/r02b_ncu_full2.log 2>&1; echo "ncu full2 exit $?"
