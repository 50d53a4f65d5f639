import sys, math, json, threading, time
sys.path.insert(0, '.')
import numpy as np
from __graft_entry__ import load_package
import pynvml
pynvml.nvmlInit(); H = pynvml.nvmlDeviceGetHandleByIndex(0)
vm = load_package()
ctx = vm.Context(0)
N = 100_000_000
L = 2 * math.pi / 0.3
fld = vm.DeviceField(ctx, 0.0, L, 4, 16, 0)
p = vm.DeviceParticles(ctx, N)
p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [0.03, 0.3, 0.1, 0.5, 4.5], 20240601)

def timed(label, steps=100):
    samples = []
    stop = threading.Event()
    def samp():
        while not stop.is_set():
            samples.append((pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(H) / 1000.0,
                            pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_MEM)))
            time.sleep(0.002)
    th = threading.Thread(target=samp); th.start()
    ctx.sync(); ctx.event_record(0)
    fld.run(p, 0.1, steps, 0, 0, 1.0)
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms(0, 1)
    stop.set(); th.join()
    a = np.array(samples[len(samples)//3:]) if len(samples) > 3 else np.array(samples)
    print(f"{label}: ms/step {ms/steps:.4f}  sm_mhz {np.median(a[:,0]):.0f} power_w {np.median(a[:,1]):.0f} mem_mhz {np.median(a[:,2]):.0f}", flush=True)

fld.run(p, 0.1, 3, 0, 0, 1.0)
timed("fresh 0-100")
timed("100-200")
timed("200-300")
fld.run(p, 0.1, 500, 0, 0, 1.0)
timed("800-900")
x, v, w = p.download()
# (a) wrap positions into the domain, keep v
p.upload(x=np.mod(x, L))
timed("after wrapping x into [0,L)")
# (b) original magnitude x but shuffled v? restore x, permute nothing: put back unwrapped x
p.upload(x=x, v=v)
timed("restored unwrapped state")
# (c) keep evolved x (unwrapped), fresh Maxwellian v
rng = np.random.default_rng(0)
p.upload(x=x, v=rng.standard_normal(N))
timed("evolved x, fresh v")
# (d) sorted-by-cell particles (coherent cells within a warp)
xs = np.mod(x, L); order = np.argsort(xs, kind="stable")
p.upload(x=xs[order], v=v[order], w=w[order])
timed("sorted by x")
