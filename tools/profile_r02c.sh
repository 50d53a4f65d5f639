#!/bin/bash
# Round-2 (limb-atomic layout) evidence run on ONE GPU: tests, bench line, ncu launch list of the bench command, full
# captures (with per-instruction source counters) of the fused pass at 16 cells (lane-private) and 64 / 256 / 1024 cells
# (limb atomics with bank-steered replicas).
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > $O/bench_r02c_1gpu.log 2>&1; grep '^{' $O/bench_r02c_1gpu.log > $O/bench_r02c_1gpu.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02c_launches_bench_1gpu.csv python bench.py --steps 5 --warmup 3 --repeats 1 --no-cpu --no-parity --no-ceilings > $O/bench_under_ncu_r02c.log 2>&1
for nh in 16 64 256 1024; do
  ncu --set full --clock-control none --import-source on -k regex:"k_vp_pass" -s 4 -c 1 -f -o /tmp/r02c_nh$nh python tools/ab/tune_run.py 100000000 $nh > /dev/null 2>&1
  ncu -i /tmp/r02c_nh$nh.ncu-rep --page raw --csv > $O/r02c_ncu_pass_nh${nh}_raw.csv 2>/dev/null
  ncu -i /tmp/r02c_nh$nh.ncu-rep --page source --csv 2>/dev/null | python tools/ncu_source_digest.py > $O/r02c_ncu_pass_nh${nh}_source.csv
done
python tools/ncu_summary.py $O/r02c_ncu_pass_nh16_raw.csv $O/r02c_ncu_pass_nh64_raw.csv $O/r02c_ncu_pass_nh256_raw.csv $O/r02c_ncu_pass_nh1024_raw.csv > $O/r02c_ncu_digest.txt
cat $O/r02c_ncu_digest.txt
