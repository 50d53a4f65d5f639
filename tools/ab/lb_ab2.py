"""Lenard-Bernstein kernels: v-space deposit with the deep tier (default) vs the round-1 loop (tuning pairs = 1)."""
import json, sys
sys.path.insert(0, '.')
import numpy as np
from __graft_entry__ import load_package
vm = load_package()
N = 100_000_000
for name, tune in (("deep", {}), ("round1", {"pairs": 1})):
    ctx = vm.Context(0)
    for k, v in tune.items():
        ctx.set_tuning(k, v)
    p = vm.DeviceParticles(ctx, N)
    p.fill(vm._lib.VM_FILL_DOUBLE_MAXWELLIAN, [-10.0, 10.0, 2.0], 1)
    for nknots in (41, 129):
        vs = vm.DeviceVSpline(ctx, -10.0, 10.0, nknots, 4, 1)
        def timed(fn, reps):
            fn(); ctx.sync(); ctx.event_record(0)
            for _ in range(reps):
                fn()
            ctx.event_record(1)
            return ctx.event_elapsed_ms(0, 1) / reps
        row = {"case": name, "nknots": nknots, "project_ms": timed(lambda: vs.project(p), 10),
               "lb_rhs_ms": timed(lambda: vs.lb_rhs(p, 1.0, False, to_host=False), 5),
               "clb_rhs_ms": timed(lambda: vs.lb_rhs(p, 1.0, True, to_host=False), 5),
               "clb_rk438_step_ms": timed(lambda: vs.rk438_run(p, 1e-3, 5, 1.0, True, 0), 2) / 5}
        row["lb_rhs_hbm_frac"] = 24 * N / row["lb_rhs_ms"] / 1e6 / 6463.3
        row["clb_rhs_hbm_frac"] = 32 * N / row["clb_rhs_ms"] / 1e6 / 6463.3
        print(json.dumps(row), flush=True)
        vs.close()
    ctx.close()
