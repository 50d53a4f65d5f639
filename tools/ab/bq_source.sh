#!/bin/bash
# per-instruction counts of the bank-sorted pass (source page of one ncu capture) for the mesh sizes given
O=gpurun_out
for nh in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:k_vp_pass_bq -s 4 -c 1 -f -o /tmp/bqsrc_$nh python tools/ab/nh_run.py 100000000 $nh 0 0 > /dev/null 2>&1
  ncu -i /tmp/bqsrc_$nh.ncu-rep --page source --csv 2>/dev/null | python tools/ncu_source_digest.py > $O/r02_bq_source_nh$nh.csv
  ncu -i /tmp/bqsrc_$nh.ncu-rep --page raw --csv 2>/dev/null > $O/r02_bq_raw_nh$nh.csv
  python tools/ncu_summary.py $O/r02_bq_raw_nh$nh.csv
done
