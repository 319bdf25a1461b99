#!/bin/bash
# ncu --set full of the 4 forward + 8 backward tcgen05 GEMM launches of the first / last ViT-B/16 block of step 4, reduced to CSV on the box.
mkdir -p gpurun_out
CMD="python bench.py --config ${CFG:-b16} --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 439 -c 4 -o /tmp/g_fwd $CMD > gpurun_out/ncu_g1.log 2>&1; echo "fwd=$?"
ncu -i /tmp/g_fwd.ncu-rep --page raw --csv > gpurun_out/gemm_fwd.raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 487 -c 8 -o /tmp/g_bwd $CMD > gpurun_out/ncu_g2.log 2>&1; echo "bwd=$?"
ncu -i /tmp/g_bwd.ncu-rep --page raw --csv > gpurun_out/gemm_bwd.raw.csv 2>/dev/null
