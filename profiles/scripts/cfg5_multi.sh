#!/bin/bash
# config 5 (Swin-B shape) at N GPUs for the given degrees: bash profiles/scripts/cfg5_multi.sh N "1 1" "3 3"
N=$1; shift
for d in "$@"; do
  t=${d// /_}
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config 5 --degree $d --steps 20 --warmup 3 --no-extras 2>gpurun_out/multi_err.log | tail -1 > gpurun_out/r02_cfg5_deg${t}_n$N.json
  python -c "
import json; d=json.loads(open('gpurun_out/r02_cfg5_deg${t}_n$N.json').read()); print('cfg5 deg', d['config']['gpf_degree'], 'N=$N', round(d['value']), 'e2e', round(d['e2e']['value']), round(d['ms_per_step'],3))"
done
