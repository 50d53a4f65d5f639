#!/bin/bash
# ncu --set full (+ source counters) of the limb-atomic fused pass, one launch per mesh size; raw CSV, source CSV and digest into gpurun_out/
O=gpurun_out
for cfg in "256 af_ctas=1" "1024 af_ctas=1" "16 af_ctas=1"; do
  set -- $cfg
  name=af_nh$1
  ncu --set full --import-source on --clock-control none -k regex:k_vp_pass -s 4 -c 1 -f -o /tmp/prof_$name python tools/ab/tune_run.py 100000000 $1 af=1 $2 > /dev/null 2>&1
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > $O/r02c_ncu_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$name.ncu-rep --page source --csv 2>/dev/null | python tools/ncu_source_digest.py > $O/r02c_ncu_${name}_source.csv
  python tools/ncu_summary.py $O/r02c_ncu_${name}_raw.csv
done
