#!/bin/bash
# A/B of the gradient-exchange options at N GPUs of one box (N=1 reference first): one bench line per variant into gpurun_out/scale_ab_N.jsonl
N=${N:-2}
OUT=gpurun_out/scale_ab_$N.jsonl
: > $OUT
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/scale_ab.err | sed 's/^/{"variant": "n1", "line": /; s/$/}/' >> $OUT
run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>>gpurun_out/scale_ab.err | sed "s/^/{\"variant\": \"$tag\", \"line\": /; s/\$/}/" >> $OUT; }
for v in ${VARIANTS:-default f32wire ctas0 zero1}; do
  case $v in
    default) run default ;;
    f32wire) run f32wire --comm f32 --no-strong ;;
    ctas0) run ctas0 --nccl-max-ctas 0 --no-strong ;;
    ctas4) run ctas4 --nccl-max-ctas 4 --no-strong ;;
    ctas16) run ctas16 --nccl-max-ctas 16 --no-strong ;;
    zero1) run zero1 --zero1 --no-strong ;;
  esac
done
tail -3 gpurun_out/scale_ab.err
