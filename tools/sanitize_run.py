"""Small run of every hot kernel family for compute-sanitizer (memcheck / racecheck / synccheck): deposits on meshes of
16 / 64 / 256 / 1024 cells (lane-private, limb-atomic, bank-sorted, fixed-point), a fused Strang run with diagnostics,
the LB / CLB right-hand sides and an RK438 step.  Sizes are tiny: the tools slow kernels down 10-100x.
    compute-sanitizer --tool racecheck python tools/sanitize_run.py"""
import math, sys
sys.path.insert(0, '.')
import numpy as np
from __graft_entry__ import load_package
vm = load_package()
rng = np.random.default_rng(3)
N = 6001
L = 2 * math.pi / 0.3
x = rng.uniform(0, L, N); v = rng.standard_normal(N); w = np.full(N, L / N)
for tun in ({}, {"af": -1}, {"bankq": 1}):
    ctx = vm.Context(0)
    for k, val in tun.items():
        ctx.set_tuning(k, val)
    p = vm.DeviceParticles(ctx, N)
    for nh in (16, 64, 256, 1024):
        p.upload(x, v, w)
        f = vm.DeviceField(ctx, 0.0, L, 4, nh, 0)
        for mode in (0, 2):
            f.deposit(p, mode)
        r = f.rhs
        assert abs(r.sum() - L) < 1e-9 * L, (tun, nh, r.sum())
        f.run(p, 0.1, 3, 1, 0, 1.0)
        f.close()
    p.upload(v=v, w=np.full(N, 1.0 / N))
    vs = vm.DeviceVSpline(ctx, -10.0, 10.0, 41, 4, 1)
    vs.lb_rhs(p, 1.0, True)
    vs.rk438_run(p, 1e-3, 1, 1.0, True, 0)
    vs.close(); p.close(); ctx.close()
print("sanitize_run: ok")
