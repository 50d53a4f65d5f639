"""n_h = 128 (and 96): lane-private 6 warps x 8 pairs (default) vs xor-shuffle 16 replicas x 13 warps vs bank-sorted."""
import json, math, sys
sys.path.insert(0, '.')
import numpy as np
from __graft_entry__ import load_package
vm = load_package()
N = 100_000_000
L = 2 * math.pi / 0.3
cases = [("default", {}), ("xor16_13w", {"replicas": 16, "threads_per_cta": 416, "ctas_per_sm": 1}),
         ("xor16_12w", {"replicas": 16, "threads_per_cta": 384, "ctas_per_sm": 1}), ("bankq", {"bankq": 1})]
for nh in (96, 128):
    for name, tune in cases:
        ctx = vm.Context(0)
        for k, v in tune.items():
            ctx.set_tuning(k, v)
        p = vm.DeviceParticles(ctx, N)
        p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [0.03, 0.3, 0.1, 0.5, 4.5], 1)
        fld = vm.DeviceField(ctx, 0.0, L, 4, nh, 0)
        try:
            fld.run(p, 0.1, 3, 0, 0, 1.0)
            ts = []
            for rep in range(3):
                ctx.sync(); ctx.event_record(0)
                fld.run(p, 0.1, 10, 0, 0, 1.0)
                ctx.event_record(1)
                ts.append(ctx.event_elapsed_ms(0, 1) / 10)
            d = fld.run(p, 0.1, 2, 2, 0, 1.0)
            print(json.dumps({"n_h": nh, "case": name, "step_ms": float(np.median(ts)), "frac": 32 * N / float(np.median(ts)) / 1e6 / 6463.3,
                              "energy": float(d[-1, 0] + d[-1, 1])}), flush=True)
        except Exception as e:
            print(json.dumps({"n_h": nh, "case": name, "error": str(e)[:200]}), flush=True)
        ctx.close()
