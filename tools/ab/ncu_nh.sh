#!/bin/bash
# ncu --set full of the fused pass at mid-size meshes (one launch each); raw CSV + digest into gpurun_out/
O=gpurun_out
for cfg in "32 1 0" "64 1 0" "64 4 0" "128 8 6"; do
  set -- $cfg
  name=nh$1_p$2
  ncu --set full --clock-control none -k regex:k_vp_pass -s 4 -c 1 -f -o /tmp/prof_$name python tools/ab/nh_run.py 100000000 $1 $2 $3 > /dev/null 2>&1
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > $O/ncu_r01d_${name}_raw.csv 2>/dev/null
  python tools/ncu_summary.py $O/ncu_r01d_${name}_raw.csv
done
