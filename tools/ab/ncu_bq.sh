#!/bin/bash
# ncu --set full of the bank-sorted fused pass (one launch per mesh size); raw CSV + digest into gpurun_out/
O=gpurun_out
for nh in 256 1024; do
  name=r02_bq_nh$nh
  ncu --set full --clock-control none --import-source on -k regex:k_vp_pass_bq -s 4 -c 1 -f -o $O/$name python tools/ab/nh_run.py 100000000 $nh 0 0 > /dev/null 2>&1
  ncu -i $O/$name.ncu-rep --page raw --csv > $O/${name}_raw.csv 2>/dev/null
  python tools/ncu_summary.py $O/${name}_raw.csv
done
