#!/bin/bash
# ncu --set full of the fused limb-atomic pass at 512 cells and of the deposit-only limb-atomic pass at 128 cells (final code)
O=gpurun_out
ncu --set full --clock-control none -k regex:k_vp_pass -s 4 -c 1 -f -o /tmp/r02d_nh512 python tools/ab/tune_run.py 100000000 512 > /dev/null 2>&1
ncu -i /tmp/r02d_nh512.ncu-rep --page raw --csv > $O/r02d_ncu_pass_nh512_raw.csv 2>/dev/null
cat > /tmp/dep_run.py <<'PY'
import sys, math
sys.path.insert(0, '.')
from __graft_entry__ import load_package
vm = load_package()
ctx = vm.Context(0)
fld = vm.DeviceField(ctx, 0.0, 2 * math.pi / 0.3, 4, 128, 0)
p = vm.DeviceParticles(ctx, 100_000_000)
p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [0.03, 0.3, 0.1, 0.5, 4.5], 1)
for _ in range(6):
    fld.deposit(p, 0)
ctx.sync()
PY
ncu --set full --clock-control none -k regex:k_vp_pass -s 3 -c 1 -f -o /tmp/r02d_dep128 python /tmp/dep_run.py > /dev/null 2>&1
ncu -i /tmp/r02d_dep128.ncu-rep --page raw --csv > $O/r02d_ncu_deposit_only_nh128_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/r02d_ncu_pass_nh512_raw.csv $O/r02d_ncu_deposit_only_nh128_raw.csv > $O/r02d_ncu_digest.txt
cat $O/r02d_ncu_digest.txt
