#!/bin/bash
# ncu evidence for profiles/ (run under gpurun on ONE GPU; numbers printed under ncu are not bench values).
# Only CSV exports (and the main kernel's report) are kept: gpurun brings back at most 64 MiB.
R=${1:-r01b}
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_$R.csv python bench.py --steps 5 --warmup 3 --no-cpu > $O/bench_under_ncu_$R.log 2>&1
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o /tmp/prof_$name "$@" > /dev/null 2>&1
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > $O/ncu_${R}_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$name.ncu-rep --page details --csv > $O/ncu_${R}_${name}_details.csv 2>/dev/null
}
cap vp_pass k_vp_pass 4 2 python bench.py --steps 5 --warmup 3 --no-cpu
cp /tmp/prof_vp_pass.ncu-rep $O/prof_${R}_vp_pass.ncu-rep
ncu -i /tmp/prof_vp_pass.ncu-rep --page source --csv > $O/ncu_${R}_vp_pass_source.csv 2>/dev/null
cap deposit k_vp_pass 0 2 python tools/sweep.py --what deposit --nh 16 --orders 4
cap lb_stage k_lb_stage 1 4 python tools/sweep.py --what lb --nknots 41
cap lb_rhs "k_v_rhs|k_v_moments|k_v_deposit" 0 3 python tools/sweep.py --what lb --nknots 41
ls -la $O
