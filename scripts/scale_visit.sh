#!/bin/bash
# One multi-GPU visit (N = ${N:-8} GPUs of one box): N = 1 reference on the same box, then the data-parallel bench of the headline
# config (with the strong-scaling leg), ViT-B/8 and ViT-S/16.  One bench line per run into gpurun_out/scale_<cfg>_n<N>.json.
N=${N:-8}
mkdir -p gpurun_out
run1() { python bench.py --config $1 --steps ${STEPS:-8} --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/scale_$1_n1.json 2>> gpurun_out/scale.err; }
runN() { cfg=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $cfg --steps ${STEPS:-8} --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/scale_${cfg}_n$N.json 2>> gpurun_out/scale.err; echo "$cfg N=$N exit=$?"; }
: > gpurun_out/scale.err
run1 b16
runN b16
for c in ${CONFIGS:-b8 s16}; do runN $c --no-strong; done
python - <<PY
import json, glob
for p in sorted(glob.glob("gpurun_out/scale_*_n*.json")):
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        print(p, d["value"], d["ms_per_step"], d.get("strong_scaling"), d["roofline"]["achieved"] if d.get("roofline") else None)
    except Exception as e:
        print(p, "unreadable", e)
PY
grep -v "^\*\|OMP_NUM\|warn" gpurun_out/scale.err | tail -5
