#!/bin/bash
# evidence table for the 8-GPU data-parallel step: one box, all variants back to back
N=${1:-8}
run() {
  tag="$1"; shift
  envs="$1"; shift
  echo "== $tag"
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
run default ""
run noallreduce "" --no-allreduce
run nooverlap "" --no-overlap
run static_sched "EGM_SCHED=0"
run normal_prio "" --nccl-normal-priority
run maxctas8 "NCCL_MAX_CTAS=8"
run default2 ""
echo "== N=1 on the same box"
python bench.py --steps 30 --warmup 5 --no-extras 2>/dev/null | tail -1 > gpurun_out/n${N}_single.json
python -c "
import json; d=json.loads(open('gpurun_out/n${N}_single.json').read()); print('single', round(d['value']), round(d['ms_per_step'],3), d['per_rank']['ns_chain_ms'], d['clocks']['sm_mhz'])"
