"""One fused-loop run at a given mesh size with arbitrary tuning keys (target of ncu captures):
    python tools/ab/tune_run.py N n_h [key=value ...]"""
import sys, math
sys.path.insert(0, '.')
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
fld.run(p, 0.1, 8, 0, 0, 1.0)
ctx.sync()
