#!/usr/bin/env python
"""Secondary measurements on one B200 (not the bench contract): per-kernel GB/s for the deposit-only
pass, the LB / CLB right-hand sides and RK438 steps, and an n_h / order / N sweep of the fused step.
Prints one JSON object per line; results are summarised under profiles/."""
from __future__ import annotations

import argparse
import json
import math
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402

PEAK = 6547.5
try:
    PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass


def timed(ctx, fn, reps):
    fn()
    ctx.sync()
    ctx.event_record(4)
    for _ in range(reps):
        fn()
    ctx.event_record(5)
    return ctx.event_elapsed_ms(4, 5) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000_000)
    ap.add_argument("--what", default="deposit,lb,vp")
    ap.add_argument("--nh", default="16,64,128,256,512,1024")
    ap.add_argument("--orders", default="3,4,5")
    ap.add_argument("--nknots", default="41,129,513")
    ap.add_argument("--tune", default="", help="comma list of key=value tuning knobs (ctas_per_sm, threads_per_cta, replicas)")
    args = ap.parse_args()
    vm = load_package()
    ctx = vm.Context(0)
    for kv in filter(None, args.tune.split(",")):
        k, v = kv.split("=")
        ctx.set_tuning(k, int(v))
    N = args.n
    L = 2 * math.pi / 0.3
    p = vm.DeviceParticles(ctx, N)
    what = args.what.split(",")

    if "deposit" in what or "vp" in what:
        p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [0.03, 0.3, 0.1, 0.5, 4.5], 1)
        for order in [int(o) for o in args.orders.split(",")]:
            for nh in [int(x) for x in args.nh.split(",")]:
                fld = vm.DeviceField(ctx, 0.0, L, order, nh, 0)
                if "deposit" in what:
                    for mode, name in ((0, "deterministic"),) + ((() if args.tune else ((1, "atomic"),))):
                        ms = timed(ctx, lambda: fld.deposit(p, mode), 5)
                        print(json.dumps({"kernel": "deposit", "tune": args.tune, "mode": name, "order": order, "n_h": nh, "N": N, "ms": ms,
                                          "GBps": 16 * N / ms / 1e6, "frac_of_measured_peak": 16 * N / ms / 1e6 / PEAK}), flush=True)
                if "vp" in what:
                    fld.run(p, 0.1, 3, 0, 0, 1.0)
                    ctx.sync(); ctx.event_record(4)
                    fld.run(p, 0.1, 20, 0, 0, 1.0)
                    ctx.event_record(5)
                    ms = ctx.event_elapsed_ms(4, 5) / 20
                    print(json.dumps({"kernel": "vp_step", "order": order, "n_h": nh, "N": N, "ms_per_step": ms,
                                      "particle_steps_per_s": N / ms * 1e3, "step_frac_of_measured_peak": 40 * N / ms / 1e6 / PEAK}), flush=True)
                fld.close()

    if "lb" in what:
        p.fill(vm._lib.VM_FILL_DOUBLE_MAXWELLIAN, [-10.0, 10.0, 2.0], 2)
        for nknots in [int(k) for k in args.nknots.split(",")]:
            vs = vm.DeviceVSpline(ctx, -10.0, 10.0, nknots, 4, 1)
            for cons in (False, True):
                ms = timed(ctx, lambda: vs.lb_rhs(p, 1.0, cons, to_host=False), 5)
                alg = 40 if cons else 32
                print(json.dumps({"kernel": "clb_rhs" if cons else "lb_rhs", "nknots": nknots, "N": N, "ms": ms,
                                  "rhs_evals_per_s": N / ms * 1e3, "GBps": alg * N / ms / 1e6,
                                  "frac_of_measured_peak": alg * N / ms / 1e6 / PEAK}), flush=True)
            for cons in (False, True):
                ms = timed(ctx, lambda: vs.rk438_run(p, 1e-3, 5, 1.0, cons, 0), 2) / 5
                alg = 200 if cons else 144
                print(json.dumps({"kernel": ("clb" if cons else "lb") + "_rk438_step", "nknots": nknots, "N": N, "ms_per_step": ms,
                                  "particle_steps_per_s": N / ms * 1e3, "GBps": alg * N / ms / 1e6,
                                  "frac_of_measured_peak": alg * N / ms / 1e6 / PEAK}), flush=True)
            vs.close()


if __name__ == "__main__":
    main()
