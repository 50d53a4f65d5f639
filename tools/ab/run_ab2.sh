#!/bin/bash
for rep in 1 2; do
for L in prev head; do
  export VLASOV_B200_LIB=$PWD/tools/ab/lib_$L.so
  echo "== $L"
  python tools/sweep.py --what lb --nknots 41 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ', d['kernel'], round(d.get('ms', d.get('ms_per_step')),4))
"
  for fl in "" "--l2-prefetch"; do
  [ "$L" = "prev" ] && [ -n "$fl" ] && continue
  python bench.py --steps 100 --warmup 3 --no-cpu --no-secondary $fl 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('   bench [$fl] ms/step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['avg_launch_ms'],4))
"
  done
done; done
