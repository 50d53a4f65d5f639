import sys, math, ctypes as C
sys.path.insert(0, '.')
from __graft_entry__ import load_package
vm = load_package()
lib = vm.lib()
L = 2 * math.pi / 0.3
ctx = vm.Context(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
fld = vm.DeviceField(ctx, 0.0, L, 4, 16, 0)
p = vm.DeviceParticles(ctx, N)
p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [0.03, 0.3, 0.1, 0.5, 4.5], 1)
fld.run(p, 0.1, 30, 0, 0, 1.0)
ctx.sync()
buf = (C.c_longlong * 32)()
lib.vm_debug_read.argtypes = [C.POINTER(C.c_longlong)]
lib.vm_debug_read(buf)
t = list(buf)
ghz = 1.965
names = {0: "start", 1: "zero-filled", 12: "after griddep wait", 2: "main loop done", 3: "flushed", 4: "finish enter", 5: "after fence", 6: "after ticket"}
print(f"N={N}: CTA 0 (since kernel start):")
for k in (1, 12, 2, 3, 4, 5, 6):
    print(f"  {names[k]:20s} {(t[k]-t[0])/ghz/1e3:8.2f} us")
print("  last CTA: fence %.2f us, reduce %.2f us, solve %.2f us" % ((t[9]-t[8])/ghz/1e3, (t[10]-t[9])/ghz/1e3, (t[11]-t[10])/ghz/1e3))
