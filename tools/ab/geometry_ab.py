"""A/B of CTA geometry for the fused step at two problem sizes (one box)."""
import sys, math
sys.path.insert(0, '.')
from __graft_entry__ import load_package
vm = load_package()
L = 2 * math.pi / 0.3
for N in (100_000_000, 12_500_000):
    for tune in ({}, {"ctas_per_sm": 1, "threads_per_cta": 1024}, {"ctas_per_sm": 2, "threads_per_cta": 512, "replicas": 32},
                 {"ctas_per_sm": 1, "threads_per_cta": 768}, {"ctas_per_sm": 3, "threads_per_cta": 320}):
        ctx = vm.Context(0)
        for k, v in tune.items():
            ctx.set_tuning(k, v)
        fld = vm.DeviceField(ctx, 0.0, L, 4, 16, 0)
        p = vm.DeviceParticles(ctx, N)
        p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [0.03, 0.3, 0.1, 0.5, 4.5], 1)
        try:
            fld.run(p, 0.1, 5, 0, 0, 1.0)
            best = 1e9
            for rep in range(3):
                ctx.sync(); ctx.event_record(0)
                fld.run(p, 0.1, 200 if N < 5e7 else 60, 0, 0, 1.0)
                ctx.event_record(1)
                best = min(best, ctx.event_elapsed_ms(0, 1) / (200 if N < 5e7 else 60))
            print(f"N={N:.3g} tune={tune} ms/step {best:.4f}", flush=True)
        except Exception as e:
            print(f"N={N:.3g} tune={tune} failed: {e}")
        p.close(); fld.close(); ctx.close()
