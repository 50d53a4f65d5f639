import sys, math, json
sys.path.insert(0, '.')
from __graft_entry__ import load_package
vm = load_package()
ctx = vm.Context(0)
N = 100_000_000
L = 2 * math.pi / 0.3
fld = vm.DeviceField(ctx, 0.0, L, 4, 16, 0)
p = vm.DeviceParticles(ctx, N)
p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [0.03, 0.3, 0.1, 0.5, 4.5], 20240601)
fld.run(p, 0.1, 3, 0, 0, 1.0)
for rep in range(8):
    prof = rep in (3, 4)
    ctx.set_tuning("profile", 1 if prof else 0)
    ctx.profile_read()
    ctx.sync(); ctx.event_record(0)
    fld.run(p, 0.1, 100, 0, 0, 1.0)
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms(0, 1)
    kn, kms = ctx.profile_read()
    x, v, _ = (None, None, None)
    print(f"rep {rep} steps {3+100*rep}-{3+100*(rep+1)} prof={prof} ms/step {ms/100:.4f} bracketed_kernel {kms/max(kn,1):.4f}", flush=True)
xs, vs, _ = p.download(w=False)
import numpy as np
print("x range", xs.min(), xs.max(), "v range", vs.min(), vs.max(), "nan", np.isnan(xs).sum())
p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [0.03, 0.3, 0.1, 0.5, 4.5], 20240601)
ctx.sync(); ctx.event_record(0); fld.run(p, 0.1, 100, 0, 0, 1.0); ctx.event_record(1)
print("after refill ms/step", ctx.event_elapsed_ms(0, 1) / 100)
