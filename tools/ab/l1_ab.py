#!/usr/bin/env python
"""Does the shared-memory carve-out starve L1 of lines for in-flight streaming loads?  Same kernels, (a) built with
ld.global.L1::no_allocate instead of ld.global.cs (VLASOV_B200_LIB selects the build), (b) with fewer warps, i.e.
less shared memory and a larger L1.  One process, one box; one JSON object per line."""
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


def vp_ms(ctx, fld, p, steps=20):
    fld.run(p, 0.1, 3, 0, 0, 1.0)
    ctx.sync(); ctx.event_record(4)
    fld.run(p, 0.1, steps, 0, 0, 1.0)
    ctx.event_record(5)
    return ctx.event_elapsed_ms(4, 5) / steps


def dep_ms(ctx, fld, p, reps=5):
    fld.deposit(p, 0); ctx.sync(); ctx.event_record(4)
    for _ in range(reps):
        fld.deposit(p, 0)
    ctx.event_record(5)
    return ctx.event_elapsed_ms(4, 5) / reps


ctx = vm.Context(0)
p = vm.DeviceParticles(ctx, N)
p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [0.03, 0.3, 0.1, 0.5, 4.5], 1)
for nh, tunes in ((16, [{}]), (32, [{}, {"ctas_per_sm": 1, "threads_per_cta": 576}, {"ctas_per_sm": 1, "threads_per_cta": 448}]),
                  (64, [{}, {"ctas_per_sm": 1, "threads_per_cta": 288}]), (128, [{}]), (256, [{}])):
    for tune in tunes:
        for k in ("ctas_per_sm", "threads_per_cta"):
            ctx.set_tuning(k, tune.get(k, 0))
        fld = vm.DeviceField(ctx, 0.0, L, 4, nh, 0)
        try:
            print(json.dumps({"tag": tag, "n_h": nh, "tune": tune, "step_ms": round(vp_ms(ctx, fld, p), 4),
                              "deposit_ms": round(dep_ms(ctx, fld, p), 4)}), flush=True)
        except Exception as e:   # noqa: BLE001
            print(json.dumps({"tag": tag, "n_h": nh, "tune": tune, "error": str(e)}), flush=True)
        fld.close()
for k in ("ctas_per_sm", "threads_per_cta"):
    ctx.set_tuning(k, 0)
p.fill(vm._lib.VM_FILL_DOUBLE_MAXWELLIAN, [-10.0, 10.0, 2.0], 2)
vs = vm.DeviceVSpline(ctx, -10.0, 10.0, 41, 4, 1)
for rep in range(2):
    out = {"tag": tag, "lb": "41 knots"}
    for cons in (False, True):
        vs.lb_rhs(p, 1.0, cons, to_host=False); ctx.sync(); ctx.event_record(4)
        for _ in range(5):
            vs.lb_rhs(p, 1.0, cons, to_host=False)
        ctx.event_record(5)
        out["clb_rhs_ms" if cons else "lb_rhs_ms"] = round(ctx.event_elapsed_ms(4, 5) / 5, 4)
        vs.rk438_run(p, 1e-3, 2, 1.0, cons); ctx.sync(); ctx.event_record(4)
        vs.rk438_run(p, 1e-3, 10, 1.0, cons)
        ctx.event_record(5)
        out["clb_step_ms" if cons else "lb_step_ms"] = round(ctx.event_elapsed_ms(4, 5) / 10, 4)
    print(json.dumps(out), flush=True)
