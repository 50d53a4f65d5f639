"""Fixed cost of a fused step for several meshes / tunings: us/step versus N (one GPU):
    python tools/ab/small_n2.py "n_h ..." name:key=value,... ..."""
import sys, math, json
sys.path.insert(0, '.')
from __graft_entry__ import load_package
vm = load_package()
L = 2 * math.pi / 0.3
meshes = [int(a) for a in sys.argv[1].split()]
for spec in sys.argv[2:]:
    name, _, kvs = spec.partition(":")
    ctx = vm.Context(0)
    for kv in kvs.split(","):
        if kv:
            ctx.set_tuning(kv.split("=")[0], int(kv.split("=")[1]))
    for nh in meshes:
        row = {"variant": name, "n_h": nh}
        for N in (1000, 1_000_000, 12_500_000):
            fld = vm.DeviceField(ctx, 0.0, L, 4, nh, 0)
            p = vm.DeviceParticles(ctx, N)
            p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [0.03, 0.3, 0.1, 0.5, 4.5], 1)
            fld.run(p, 0.1, 20, 0, 0, 1.0)
            best = 1e9
            for rep in range(3):
                ctx.sync(); ctx.event_record(0)
                fld.run(p, 0.1, 300, 0, 0, 1.0)
                ctx.event_record(1)
                best = min(best, ctx.event_elapsed_ms(0, 1) / 300)
            row[f"us_per_step_N{N}"] = round(best * 1e3, 2)
            p.close(); fld.close()
        print(json.dumps(row), flush=True)
    ctx.close()
