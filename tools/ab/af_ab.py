"""Mesh-size sweep of the fused step and the deposit-only pass: the limb-atomic fixed-point pass (tuning af = 1, with
its CTA-shape / gather-table variants) against the layouts it replaces (af = -1: lane-private replicas up to 80 cells,
bank-sorted queues above), one process, one box.  Prints one JSON line per (n_h, variant):
    python tools/ab/af_ab.py [N] [n_h ...]"""
import json, math, sys
sys.path.insert(0, '.')
import numpy as np
from __graft_entry__ import load_package
vm = load_package()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
meshes = [int(a) for a in sys.argv[2:]] or [16, 32, 64, 128, 256, 512, 1024]
L = 2 * math.pi / 0.3
PEAK = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
variants = [("base", {"af": -1}), ("af", {"af": 1}), ("af_r8", {"af": 1, "af_replicas": 8}), ("af_r1", {"af": 1, "af_replicas": 1}), ("af_2cta", {"af": 1, "af_ctas": 2})]
ref = {}
for name, tun in variants:
    ctx = vm.Context(0)
    for k, val in tun.items():
        ctx.set_tuning(k, val)
    p = vm.DeviceParticles(ctx, N)
    for nh in meshes:
        p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [0.03, 0.3, 0.1, 0.5, 4.5], 1)
        fld = vm.DeviceField(ctx, 0.0, L, 4, nh, 0)
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
        err = None
        if name == "base":
            ref[nh] = rhs
        else:
            err = float(np.max(np.abs(rhs - ref[nh])) / np.max(np.abs(ref[nh])))
        print(json.dumps({"n_h": nh, "variant": name, "deposit_ms": dep, "step_ms": step, "step_hbm_frac": 32 * N / step / 1e6 / PEAK,
                          "rhs_rel_vs_base": err, "energy": float(d[-1, 0] + d[-1, 1])}), flush=True)
        fld.close()
    p.close(); ctx.close()
