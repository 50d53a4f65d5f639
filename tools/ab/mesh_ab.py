"""Mesh-size sweep of the fused step and the deposit-only pass, bank-sorted pass (tuning bankq = 1) against the
round-1 variants (bankq = -1), one process, one box.  Prints one JSON line per (n_h, variant):
    python tools/ab/mesh_ab.py [N] [n_h ...]"""
import json, math, sys
sys.path.insert(0, '.')
import numpy as np
from __graft_entry__ import load_package
vm = load_package()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
meshes = [int(a) for a in sys.argv[2:]] or [32, 64, 128, 200, 256, 512, 1024]
L = 2 * math.pi / 0.3
PEAK = 6463.3
ref = {}
for bankq in (-1, 1):
    ctx = vm.Context(0)
    ctx.set_tuning("bankq", bankq)
    p = vm.DeviceParticles(ctx, N)
    for nh in meshes:
        p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [0.03, 0.3, 0.1, 0.5, 4.5], 1)
        fld = vm.DeviceField(ctx, 0.0, L, 4, nh, 0)
        plan = vm._lib.pass_plan(nh, 4, 1) if bankq >= 0 else None
        fld.deposit(p, 0)
        rhs = fld.rhs
        ts = []
        for rep in range(3):
            ctx.sync(); ctx.event_record(0)
            for _ in range(5):
                fld.deposit(p, 0)
            ctx.event_record(1)
            ts.append(ctx.event_elapsed_ms(0, 1) / 5)
        dep = float(np.median(ts))
        fld.run(p, 0.1, 3, 0, 0, 1.0)
        ts = []
        for rep in range(3):
            ctx.sync(); ctx.event_record(0)
            fld.run(p, 0.1, 10, 0, 0, 1.0)
            ctx.event_record(1)
            ts.append(ctx.event_elapsed_ms(0, 1) / 10)
        step = float(np.median(ts))
        d = fld.run(p, 0.1, 2, 2, 0, 1.0)
        key = nh
        err = None
        if bankq == -1:
            ref[key] = rhs
        else:
            err = float(np.max(np.abs(rhs - ref[key])) / np.max(np.abs(ref[key])))
        print(json.dumps({"n_h": nh, "bankq": bankq, "deposit_ms": dep, "step_ms": step, "step_hbm_frac": 32 * N / step / 1e6 / PEAK,
                          "rhs_rel_vs_round1_variant": err, "energy": float(d[-1, 0] + d[-1, 1])}), flush=True)
        fld.close()
    p.close(); ctx.close()
