"""Full-size check of the Sobol bump-on-tail loads (fill kinds 7 / 8) at 1e8 and 1e9 particles: fill time, total charge and
density modulation of the deposit.  Measured: 4 / 3 ms at 1e8, 34 / 29 ms at 1e9 particles; sum rhs / L = 1 to 12 digits."""
import sys, math, time
sys.path.insert(0, '.')
import numpy as np
from __graft_entry__ import load_package
vm = load_package()
ctx = vm.Context(0)
for N in (100_000_000, 1_000_000_000):
    p = vm.DeviceParticles(ctx, N)
    for kind in (7, 8):
        ctx.sync(); t0 = time.perf_counter()
        p.fill(kind, [0.03, 0.3, 0.1, 0.5, 4.5, -1.0], 11)
        ctx.sync(); dt = time.perf_counter() - t0
        f = vm.DeviceField(ctx, 0.0, 2 * math.pi / 0.3, 4, 16, 0)
        f.deposit(p, 0)
        r = f.rhs
        print(N, kind, "fill s %.3f" % dt, "sum rhs / L %.12f" % (r.sum() / (2 * math.pi / 0.3)), "rhs modulation %.5f" % ((r.max() - r.min()) / r.mean()))
        f.close()
    p.close()
