#!/bin/bash
# Source-level ncu capture (warp-stall samples per SASS line) of one tcgen05 GEMM launch of step 4: IDX = index of the launch
# among the step's 146 GEMMs (0 patch embed; layer 0 forward: 1 qkv, 2 proj, 3 fc, 4 fcproj).
mkdir -p gpurun_out
CMD="python bench.py --config ${CFG:-b16} --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel -s $((438 + ${IDX:-3})) -c 1 -o /tmp/g_src $CMD > gpurun_out/ncu_src.log 2>&1; echo "ncu=$?"
ncu -i /tmp/g_src.ncu-rep --page source --csv > gpurun_out/gemm_${NAME:-fc}_source.csv 2>/dev/null
ncu -i /tmp/g_src.ncu-rep --page raw --csv > gpurun_out/gemm_${NAME:-fc}.raw.csv 2>/dev/null
ls -la gpurun_out | head
