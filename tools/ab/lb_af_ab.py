"""Lenard-Bernstein kernels under tuning variants, timed round-robin in one process:
    python tools/ab/lb_af_ab.py [N] name:key=value,... ...
Used once for an experimental port of the limb-atomic layout to the v-space deposit and the RK438 stage pass (tuning key
v_af, not kept): profiles/r02c_vspace_limb_atomic_ab.jsonl -- LB RHS 0.612 vs 0.610 ms, CLB RK438 step 3.59 vs 3.51 ms per
1e8 particles: at 8 B/particle the layout trades the lane-private pass's shared-memory wavefronts for instructions and
gains nothing, so the v-space passes keep the lane-private replicas."""
import json, sys
sys.path.insert(0, '.')
import numpy as np
from __graft_entry__ import load_package
vm = load_package()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
specs = sys.argv[2:] or ["default:", "pairs1:pairs=1"]
PEAK = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
state = []
for spec in specs:
    name, _, kvs = spec.partition(":")
    ctx = vm.Context(0)
    for kv in kvs.split(","):
        if kv:
            ctx.set_tuning(kv.split("=")[0], int(kv.split("=")[1]))
    p = vm.DeviceParticles(ctx, N)
    p.fill(vm._lib.VM_FILL_DOUBLE_MAXWELLIAN, [-10.0, 10.0, 2.0], 1)
    vs = vm.DeviceVSpline(ctx, -10.0, 10.0, 41, 4, 1)
    vs.lb_rhs(p, 1.0, True, to_host=False); vs.rk438_run(p, 1e-3, 1, 1.0, True, 0)
    state.append((name, ctx, p, vs))
def timed(ctx, fn, reps):
    ctx.sync(); ctx.event_record(0)
    for _ in range(reps):
        fn()
    ctx.event_record(1)
    return ctx.event_elapsed_ms(0, 1) / reps
res = {name: {"lb": [], "clb": [], "rk_lb": [], "rk_clb": []} for name, *_ in state}
for rnd in range(4):
    for name, ctx, p, vs in state:
        res[name]["lb"].append(timed(ctx, lambda: vs.lb_rhs(p, 1.0, False, to_host=False), 5))
        res[name]["clb"].append(timed(ctx, lambda: vs.lb_rhs(p, 1.0, True, to_host=False), 5))
        res[name]["rk_lb"].append(timed(ctx, lambda: vs.rk438_run(p, 1e-3, 3, 1.0, False, 0), 1) / 3)
        res[name]["rk_clb"].append(timed(ctx, lambda: vs.rk438_run(p, 1e-3, 3, 1.0, True, 0), 1) / 3)
for name, *_ in state:
    r = {k: float(np.median(v)) for k, v in res[name].items()}
    print(json.dumps({"variant": name, "lb_rhs_ms": r["lb"], "clb_rhs_ms": r["clb"], "rk438_lb_step_ms": r["rk_lb"], "rk438_clb_step_ms": r["rk_clb"],
                      "lb_rhs_frac": 24 * N / r["lb"] / 1e6 / PEAK, "clb_rhs_frac": 32 * N / r["clb"] / 1e6 / PEAK,
                      "rk438_clb_frac": 200 * N / r["rk_clb"] / 1e6 / PEAK}), flush=True)
