#!/bin/bash
# A/B of kernel revisions on ONE box: same bench, different builds of the same ABI.
for L in "$@"; do
  export VLASOV_B200_LIB=$PWD/tools/ab/lib_$L.so
  python bench.py --steps 100 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$L', 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'])
"
  python tools/sweep.py --what deposit --nh 16 --orders 4 2>/dev/null | head -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('   deposit-only ms', round(d['ms'],4), 'frac', round(d['frac_of_measured_peak'],3))"
done
