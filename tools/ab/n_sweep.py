"""BASELINE configs[4] on one GPU: fused step over particle counts 1e6 .. 1e9 and 64 .. 1024 spline modes (uniform weights,
32 B per particle-step).  One JSON line per point:  python tools/ab/n_sweep.py"""
import json, math, sys
sys.path.insert(0, '.')
import numpy as np
from __graft_entry__ import load_package
vm = load_package()
L = 2 * math.pi / 0.3
PEAK = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
ctx = vm.Context(0)
for N in (1_000_000, 10_000_000, 100_000_000, 1_000_000_000):
    p = vm.DeviceParticles(ctx, N)
    p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [0.03, 0.3, 0.1, 0.5, 4.5], 1)
    for nh in (16, 64, 256, 1024):
        fld = vm.DeviceField(ctx, 0.0, L, 4, nh, 0)
        steps = 200 if N <= 10_000_000 else (20 if N <= 100_000_000 else 5)
        fld.run(p, 0.1, 3, 0, 0, 1.0)
        ts = []
        for rep in range(3):
            ctx.sync(); ctx.event_record(0)
            fld.run(p, 0.1, steps, 0, 0, 1.0)
            ctx.event_record(1)
            ts.append(ctx.event_elapsed_ms(0, 1) / steps)
        ms = float(np.median(ts))
        d = fld.run(p, 0.1, 2, 2, 0, 1.0)
        print(json.dumps({"N": N, "n_h": nh, "ms_per_step": ms, "particle_steps_per_s": N / ms * 1e3,
                          "step_hbm_frac": 32 * N / ms / 1e6 / PEAK, "energy": float(d[-1, 0] + d[-1, 1])}), flush=True)
        fld.close()
    p.close()
