"""Burst vs sustained time of the fused step at one mesh size: the first 10 timed steps of a fresh process (cool GPU,
full clocks), then the median of 20 more batches of 10 steps (power-capped clocks):  python tools/ab/burst.py N n_h [key=value ...]"""
import sys, math, json
sys.path.insert(0, '.')
import numpy as np
from __graft_entry__ import load_package
vm = load_package()
L = 2 * math.pi / 0.3
ctx = vm.Context(0)
N = int(sys.argv[1]); nh = int(sys.argv[2])
for kv in sys.argv[3:]:
    k, val = kv.split("=")
    ctx.set_tuning(k, int(val))
fld = vm.DeviceField(ctx, 0.0, L, 4, nh, 0)
p = vm.DeviceParticles(ctx, N)
p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [0.03, 0.3, 0.1, 0.5, 4.5], 1)
fld.run(p, 0.1, 3, 0, 0, 1.0)
ts = []
for rep in range(21):
    ctx.sync(); ctx.event_record(0)
    fld.run(p, 0.1, 10, 0, 0, 1.0)
    ctx.event_record(1)
    ts.append(ctx.event_elapsed_ms(0, 1) / 10)
PEAK = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
print(json.dumps({"n_h": nh, "tuning": sys.argv[3:], "burst_ms": ts[0], "sustained_ms": float(np.median(ts[5:])),
                  "burst_frac": 32 * N / ts[0] / 1e6 / PEAK, "sustained_frac": 32 * N / float(np.median(ts[5:])) / 1e6 / PEAK}))
