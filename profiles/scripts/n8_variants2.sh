#!/bin/bash
N=${1:-8}
run() {
  tag="$1"; shift
  envs="$1"; shift
  env $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --no-extras "$@" 2>gpurun_out/n8_err.log | tail -1 > gpurun_out/n${N}_$tag.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/n${N}_$tag.json").read())
pr=d["per_rank"]
print("$tag", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],3), "median", round(d["ms_per_step_median"],3),
      "step_ms_mean min/max", round(min(pr["step_ms_mean"]),3), round(max(pr["step_ms_mean"]),3),
      "chain min/max", round(min(pr["ns_chain_ms"]),3), round(max(pr["ns_chain_ms"]),3), "clk", d["clocks"]["sm_mhz"])
PY
}
run noallreduce_b "" --no-allreduce
run chunk32 "" --bucket-mb 32
run chunk512 "" --bucket-mb 512
run chunk128 "" --bucket-mb 128
run chunk512_ctas8 "NCCL_MAX_CTAS=8" --bucket-mb 512
run chunk512_ctas16 "NCCL_MAX_CTAS=16" --bucket-mb 512
run chunk512_normal "" --bucket-mb 512 --nccl-normal-priority
