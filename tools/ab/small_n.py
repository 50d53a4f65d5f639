"""Fixed cost of a fused step: ms/step versus N (one GPU)."""
import sys, math
sys.path.insert(0, '.')
from __graft_entry__ import load_package
vm = load_package()
L = 2 * math.pi / 0.3
ctx = vm.Context(0)
for N in (1000, 10_000, 100_000, 1_000_000, 4_000_000, 12_500_000):
    fld = vm.DeviceField(ctx, 0.0, L, 4, 16, 0)
    p = vm.DeviceParticles(ctx, N)
    p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [0.03, 0.3, 0.1, 0.5, 4.5], 1)
    fld.run(p, 0.1, 20, 0, 0, 1.0)
    best = 1e9
    for rep in range(3):
        ctx.sync(); ctx.event_record(0)
        fld.run(p, 0.1, 500, 0, 0, 1.0)
        ctx.event_record(1)
        best = min(best, ctx.event_elapsed_ms(0, 1) / 500)
    print(f"N={N:>9d} us/step {best*1e3:8.2f}  particle-steps/s {N/best*1e3:.3e}", flush=True)
    p.close(); fld.close()
