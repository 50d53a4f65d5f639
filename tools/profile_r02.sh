#!/bin/bash
# Round-2 evidence run on ONE GPU: ncu launch list of the bench command, full captures (with source) of the fused pass
# at n_h = 16 (lane-private), 256 and 1024 (bank-sorted queues), the fixed-point variant, and the v-space kernels.
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_bench_1gpu.csv python bench.py --steps 5 --warmup 3 --repeats 1 --no-cpu --no-parity --no-ceilings > $O/bench_under_ncu_r02.log 2>&1
for nh in 16 256 1024; do
  ncu --set full --clock-control none --import-source on -k regex:"k_vp_pass" -s 4 -c 1 -f -o $O/r02_ncu_pass_nh$nh python tools/ab/nh_run.py 100000000 $nh 0 0 > /dev/null 2>&1
  ncu -i $O/r02_ncu_pass_nh$nh.ncu-rep --page raw --csv > $O/r02_ncu_pass_nh${nh}_raw.csv 2>/dev/null
  # per-instruction executed counts / stall samples of the kernel (source page), then drop the 30 MB report
  ncu -i $O/r02_ncu_pass_nh$nh.ncu-rep --page source --csv 2>/dev/null | cut -d, -f1-9 > $O/r02_ncu_pass_nh${nh}_source.csv
  rm -f $O/r02_ncu_pass_nh$nh.ncu-rep
done
ncu --set full --clock-control none -k regex:"k_lb_stage|k_v_rhs|k_v_moments|k_v_deposit" -s 2 -c 7 -f -o /tmp/prof_lb python tools/sweep.py --what lb --nknots 41 > /dev/null 2>&1
ncu -i /tmp/prof_lb.ncu-rep --page raw --csv > $O/r02_ncu_lb_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/r02_ncu_pass_nh16_raw.csv $O/r02_ncu_pass_nh256_raw.csv $O/r02_ncu_pass_nh1024_raw.csv $O/r02_ncu_lb_raw.csv > $O/r02_ncu_digest.txt
cat $O/r02_ncu_digest.txt
