#!/bin/bash
for rep in 1 2; do
for flag in "" "--no-pdl"; do
python bench.py --steps 100 --warmup 3 --no-cpu $flag 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('1gpu N=1e8 [$flag]', 'ms/step', round(d['ms_per_step'],4), 'bracketed', round(d['roofline']['ms_per_step_with_brackets'],4), 'kernel_ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3))
"
python bench.py --steps 200 --warmup 3 --no-cpu --particles 12500000 $flag 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('1gpu N=1.25e7 [$flag]', 'ms/step', round(d['ms_per_step'],4), 'bracketed', round(d['roofline']['ms_per_step_with_brackets'],4), 'kernel_ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3))
"
done; done
