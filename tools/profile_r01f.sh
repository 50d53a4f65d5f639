#!/bin/bash
# Round-1 final evidence run on ONE GPU: tests, smoke, bench line, ncu launch list of the bench command and full
# captures of the Lenard-Bernstein kernels (the x-space captures r01e_* were taken with the same pass kernels).
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > $O/bench_r01f.json 2> $O/bench_r01f.err
tail -c 300 $O/bench_r01f.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01f_launches_bench_1gpu.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-secondary > $O/bench_under_ncu_r01f.log 2>&1
ncu --set full --clock-control none -k regex:"k_lb_stage|k_v_rhs|k_v_moments|k_v_deposit" -s 2 -c 7 -f -o /tmp/prof_lb python tools/sweep.py --what lb --nknots 41 > /dev/null 2>&1
ncu -i /tmp/prof_lb.ncu-rep --page raw --csv > $O/r01f_ncu_lb_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/r01f_ncu_lb_raw.csv > $O/r01f_ncu_lb_digest.txt
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_r01f.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "roofline", d["roofline"]["frac"], "step", d["step_hbm_frac"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"], d["clocks"])
print(d["deposit"]["ms"], d["deposit"]["per_particle_weights"]); print(d["secondary"])
P
