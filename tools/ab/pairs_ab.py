#!/usr/bin/env python
"""A/B of the software-pipeline depth (`pairs` tuning key) and of the lane-private warp floor
(`priv_min_warps`) for mid-size meshes, and of the pipelined RK438 stage pass -- all in ONE process on
ONE box.  Prints one JSON object per line."""
from __future__ import annotations

import argparse
import json
import math
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000_000)
    ap.add_argument("--nh", default="32,64,128")
    ap.add_argument("--what", default="vp,lb")
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    vm = load_package()
    ctx = vm.Context(0)
    N = args.n
    L = 2 * math.pi / 0.3
    p = vm.DeviceParticles(ctx, N)

    def vp_ms(fld, steps=20):
        fld.run(p, 0.1, 3, 0, 0, 1.0)
        ctx.sync(); ctx.event_record(4)
        fld.run(p, 0.1, steps, 0, 0, 1.0)
        ctx.event_record(5)
        return ctx.event_elapsed_ms(4, 5) / steps

    def dep_ms(fld, reps=5):
        fld.deposit(p, 0); ctx.sync(); ctx.event_record(4)
        for _ in range(reps):
            fld.deposit(p, 0)
        ctx.event_record(5)
        return ctx.event_elapsed_ms(4, 5) / reps

    if "vp" in args.what:
        p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [0.03, 0.3, 0.1, 0.5, 4.5], 1)
        for nh in [int(x) for x in args.nh.split(",")]:
            fld = vm.DeviceField(ctx, 0.0, L, 4, nh, 0)
            # (pairs, priv_min_warps, no_repg): the round-1 configuration, then the automatic choice without and
            # with the 16-fold gather table
            for pairs, pmw, no_repg in ((1, 12, 1), (0, 0, 1), (0, 0, 0)):
                ctx.set_tuning("pairs", pairs)
                ctx.set_tuning("priv_min_warps", pmw)
                ctx.set_tuning("no_repg", no_repg)
                try:
                    out = {"tag": args.tag, "n_h": nh, "pairs": pairs, "priv_min_warps": pmw, "no_repg": no_repg,
                           "step_ms": round(vp_ms(fld), 4), "deposit_ms": round(dep_ms(fld), 4)}
                except Exception as e:      # noqa: BLE001
                    out = {"tag": args.tag, "n_h": nh, "pairs": pairs, "priv_min_warps": pmw, "error": str(e)}
                print(json.dumps(out), flush=True)
            fld.close()
        ctx.set_tuning("pairs", 0)
        ctx.set_tuning("priv_min_warps", 0)
        ctx.set_tuning("no_repg", 0)

    if "lb" in args.what:
        p.fill(vm._lib.VM_FILL_DOUBLE_MAXWELLIAN, [-10.0, 10.0, 2.0], 2)
        vs = vm.DeviceVSpline(ctx, -10.0, 10.0, 41, 4, 1)
        for pairs in (1, 0, 0):
            try:
                ctx.set_tuning("pairs", pairs)
            except Exception:       # noqa: BLE001  (a baseline library without the key)
                pass
            out = {"tag": args.tag, "lb": "41 knots", "pairs": pairs}
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


if __name__ == "__main__":
    main()
