#!/bin/bash
# Strong-scaling run of bench.py at the GPU counts given as arguments (uses torchrun for N > 1).
# usage: tools/scale.sh "1 2 4 8" [extra bench flags]
for n in $1; do
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps 100 --warmup 3 --no-cpu ${@:2} 2>/dev/null
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 100 --warmup 3 ${@:2} 2>/dev/null
  fi | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'n_gpus': d['n_gpus'], 'value': d['value'], 'ms_per_step': d['ms_per_step'], 'kernel_ms': d['roofline']['avg_launch_ms'], 'kernel_frac': d['roofline']['frac'], 'step_hbm_frac': d['step_hbm_frac'], 'collective': d['config'].get('collective'), 'launches': d['gpu_launches'], 'e2e': d['e2e']['value']}))
"
done
