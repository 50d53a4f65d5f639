"""Interleaved A/B of the fused step over mesh sizes: every variant keeps its own context + particles, and the variants
are timed round-robin (5 rounds of 10 steps each) so that clock / power drift hits all of them alike.
    python tools/ab/af_ab2.py N "n_h n_h ..." name:key=value,key=value ..."""
import json, math, sys
sys.path.insert(0, '.')
import numpy as np
from __graft_entry__ import load_package
vm = load_package()
N = int(sys.argv[1])
meshes = [int(a) for a in sys.argv[2].split()]
variants = []
for spec in sys.argv[3:]:
    name, _, kvs = spec.partition(":")
    variants.append((name, dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in kvs.split(",") if kv)))
L = 2 * math.pi / 0.3
PEAK = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
state = []
for name, tun in variants:
    ctx = vm.Context(0)
    for k, val in tun.items():
        ctx.set_tuning(k, val)
    state.append((name, ctx, vm.DeviceParticles(ctx, N)))
for nh in meshes:
    flds = []
    for name, ctx, p in state:
        p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [0.03, 0.3, 0.1, 0.5, 4.5], 1)
        fld = vm.DeviceField(ctx, 0.0, L, 4, nh, 0)
        fld.run(p, 0.1, 3, 0, 0, 1.0)
        flds.append(fld)
    ts = {name: [] for name, _, _ in state}
    for rnd in range(5):
        for (name, ctx, p), fld in zip(state, flds):
            ctx.sync(); ctx.event_record(0)
            fld.run(p, 0.1, 10, 0, 0, 1.0)
            ctx.event_record(1)
            ts[name].append(ctx.event_elapsed_ms(0, 1) / 10)
    for name, _, _ in state:
        med = float(np.median(ts[name]))
        print(json.dumps({"n_h": nh, "variant": name, "step_ms_median": med, "step_ms_min": float(min(ts[name])),
                          "step_hbm_frac": 32 * N / med / 1e6 / PEAK}), flush=True)
    for fld in flds:
        fld.close()
