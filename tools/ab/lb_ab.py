#!/usr/bin/env python
"""A/B of library builds (VLASOV_B200_LIB) on ONE box: x-space deposit and fused step with uniform and with
per-particle weights at n_h = 16, and the Lenard-Bernstein kernels.  One JSON object per line."""
import json
import math
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else ""
vm = load_package()
N = 100_000_000
L = 2 * math.pi / 0.3
ctx = vm.Context(0)
p = vm.DeviceParticles(ctx, N)


def timed(fn, reps):
    fn(); ctx.sync(); ctx.event_record(4)
    for _ in range(reps):
        fn()
    ctx.event_record(5)
    return round(ctx.event_elapsed_ms(4, 5) / reps, 4)


p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [0.03, 0.3, 0.1, 0.5, 4.5], 1)
fld = vm.DeviceField(ctx, 0.0, L, 4, 16, 0)
for gw in (0, 1):
    ctx.set_tuning("no_uniform_w", gw)
    print(json.dumps({"tag": tag, "n_h": 16, "general_weights": gw, "deposit_ms": timed(lambda: fld.deposit(p, 0), 10),
                      "step_ms": round(timed(lambda: fld.run(p, 0.1, 20, 0, 0, 1.0), 2) / 20, 4)}), flush=True)
ctx.set_tuning("no_uniform_w", 0)
fld.close()
p.fill(vm._lib.VM_FILL_DOUBLE_MAXWELLIAN, [-10.0, 10.0, 2.0], 2)
for nknots in (41, 129):
    vs = vm.DeviceVSpline(ctx, -10.0, 10.0, nknots, 4, 1)
    for rep in range(2 if nknots == 41 else 1):
        out = {"tag": tag, "lb_knots": nknots}
        for cons in (False, True):
            out["clb_rhs_ms" if cons else "lb_rhs_ms"] = timed(lambda: vs.lb_rhs(p, 1.0, cons, to_host=False), 5)
            out["clb_step_ms" if cons else "lb_step_ms"] = round(timed(lambda: vs.rk438_run(p, 1e-3, 5, 1.0, cons), 2) / 5, 4)
        print(json.dumps(out), flush=True)
    vs.close()
p.close(); ctx.close()
