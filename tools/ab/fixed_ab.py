"""Cost of the order-independent fixed-point deposit (VM_RUN_FIXED_DEPOSIT) beside the fp64-accumulating default."""
import json, math, sys
sys.path.insert(0, '.')
import numpy as np
from __graft_entry__ import load_package
vm = load_package()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
L = 2 * math.pi / 0.3
ctx = vm.Context(0)
p = vm.DeviceParticles(ctx, N)
for nh in (16, 64, 256, 1024):
    fld = vm.DeviceField(ctx, 0.0, L, 4, nh, 0)
    row = {"n_h": nh}
    for name, flags, dmode in (("fp64", 0, 0), ("fixed", vm._lib.VM_RUN_FIXED_DEPOSIT, vm._lib.VM_DEPOSIT_FIXED)):
        p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [0.03, 0.3, 0.1, 0.5, 4.5], 1)
        fld.run(p, 0.1, 3, 0, flags, 1.0)
        ts = []
        for rep in range(3):
            ctx.sync(); ctx.event_record(0)
            fld.run(p, 0.1, 10, 0, flags, 1.0)
            ctx.event_record(1)
            ts.append(ctx.event_elapsed_ms(0, 1) / 10)
        row[name + "_step_ms"] = float(np.median(ts))
        ts = []
        for rep in range(3):
            ctx.sync(); ctx.event_record(0)
            for _ in range(5):
                fld.deposit(p, dmode)
            ctx.event_record(1)
            ts.append(ctx.event_elapsed_ms(0, 1) / 5)
        row[name + "_deposit_ms"] = float(np.median(ts))
    row["step_cost_ratio"] = row["fixed_step_ms"] / row["fp64_step_ms"]
    print(json.dumps(row), flush=True)
    fld.close()
