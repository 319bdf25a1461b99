#!/bin/bash
# One 8-GPU visit: N=1 reference, the data-parallel variants (VARIANTS) and the CUPTI timeline of rank 0.
N=${N:-8}
OUT=gpurun_out/scale_ab_$N.jsonl
: > $OUT; : > gpurun_out/scale_ab.err
python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline 2>>gpurun_out/scale_ab.err | grep '^{' | sed 's/^/{"variant": "n1", "line": /; s/$/}/' >> $OUT
run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>>gpurun_out/scale_ab.err | grep '^{' | sed "s/^/{\"variant\": \"$tag\", \"line\": /; s/\$/}/" >> $OUT; }
for v in ${VARIANTS:-default static ctas0 zero1}; do
  case $v in
    default) run default ;;
    static) VITRS_GEMM_STATIC=1 run static --no-strong ;;
    f32wire) run f32wire --comm f32 --no-strong ;;
    ctas0) run ctas0 --nccl-max-ctas 0 --no-strong ;;
    ctas4) run ctas4 --nccl-max-ctas 4 --no-strong ;;
    ctas16) run ctas16 --nccl-max-ctas 16 --no-strong ;;
    zero1) run zero1 --zero1 --no-strong ;;
  esac
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/timeline_dp.py > gpurun_out/timeline_dp$N.txt 2>> gpurun_out/scale_ab.err
grep -v "^\*\|OMP_NUM\|warn" gpurun_out/scale_ab.err | tail -5
