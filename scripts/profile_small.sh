#!/bin/bash
# One gpurun call: ncu --set full of the kernels the ViT-B/16 layer profile does not contain — the streaming attention kernels
# (ViT-B/8, T = 785), im2col, patch_bwd_reduce, head_loss, the SIMT class-head GEMM, colsum, cls gather / scatter — taken from one
# training step of ViT-B/8 at batch 64, reduced to CSV on the box.  Follows a plain run of the same command that exited 0.
set -u
mkdir -p gpurun_out
CMD="python bench.py --config b8 --batch 64 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_small.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_small.log; exit 1; }
K='regex:attn_fwd_stream|attn_bwd_dkv|attn_bwd_dq|im2col|patch_bwd_reduce|head_loss|gemm_simt|colsum|cls_gather|cls_scatter'
# matches per step: 36 attention + im2col + patch_bwd_reduce + head_loss + 3 gemm_simt + 1 colsum + 2 cls = 45; skip the warm-up steps
ncu --set full --clock-control none -k "$K" -s ${SKIP:-$((3*45 + 10))} -c ${COUNT:-45} -f -o /tmp/prof_small $CMD > gpurun_out/ncu_small.log 2>&1
echo "small=$?"
ncu -i /tmp/prof_small.ncu-rep --page raw --csv > gpurun_out/prof_small.raw.csv 2>/dev/null
ls -la gpurun_out/prof_small.raw.csv
