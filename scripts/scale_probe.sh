#!/bin/bash
# What does an 8-GPU box lose when all GPUs are busy, with and without the gradient exchange?  (a) 8 independent single-GPU
# benches at once (no NCCL at all), (b) the data-parallel bench, (c) the same with every bucket exchanged after backward
# (VITRS_DP_DEFER), (d) with NCCL capped at 2 CTAs.
mkdir -p gpurun_out
for i in 0 1 2 3 4 5 6 7; do
  CUDA_VISIBLE_DEVICES=$i python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/indep_$i.json 2>/dev/null &
done
wait
python - <<PY
import json
v = [json.loads(open(f"gpurun_out/indep_{i}.json").read().strip().splitlines()[-1]) for i in range(8)]
print("independent x8:", [round(d["value"]) for d in v], "sum", round(sum(d["value"] for d in v)), "clocks", [d["clocks"]["sm_mhz"] for d in v])
PY
runN() { tag=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-strong "$@" 2>/dev/null | tail -1 > gpurun_out/probe_$tag.json; python -c "
import json; d=json.load(open('gpurun_out/probe_$tag.json')); print('$tag', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['clocks']['sm_mhz'])"; }
runN default
VITRS_DP_DEFER=1 runN defer
runN ctas2 --nccl-max-ctas 2
runN zero1 --zero1
