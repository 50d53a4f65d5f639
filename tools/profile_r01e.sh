#!/bin/bash
# Round-1 (session e) evidence run on ONE GPU: tests, smoke, bench line, ncu launch list of the bench command, and
# full captures of the fused pass at n_h = 16 (headline) and n_h = 64 / 128 (deep lane-private tiers).
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > $O/bench_r01e.json 2> $O/bench_r01e.err
tail -c 600 $O/bench_r01e.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01e_launches_bench_1gpu.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-secondary > $O/bench_under_ncu_r01e.log 2>&1
for cfg in "16 0 0" "64 0 0" "128 0 0"; do
  set -- $cfg
  ncu --set full --clock-control none -k regex:k_vp_pass -s 4 -c 1 -f -o /tmp/prof_nh$1 python tools/ab/nh_run.py 100000000 $1 $2 $3 > /dev/null 2>&1
  ncu -i /tmp/prof_nh$1.ncu-rep --page raw --csv > $O/r01e_ncu_vp_pass_nh$1_raw.csv 2>/dev/null
done
python tools/ncu_summary.py $O/r01e_ncu_vp_pass_nh16_raw.csv $O/r01e_ncu_vp_pass_nh64_raw.csv $O/r01e_ncu_vp_pass_nh128_raw.csv > $O/r01e_ncu_digest.txt
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_r01e.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "roofline", d["roofline"]["frac"], "step", d["step_hbm_frac"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"], "launches", d["gpu_launches"], d["clocks"])
print(d["deposit"]); print(d["secondary"])
P
