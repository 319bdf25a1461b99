#!/bin/bash
# Source-level ncu capture (warp-stall samples per SASS line) of one launch of kernel K (regex) in step 4 of the bench.
# SKIP = launches of K to skip (3 warm-up steps x launches per step + index inside the step).
mkdir -p gpurun_out
CMD="python bench.py --config ${CFG:-b16} --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
ncu --set full --import-source on --clock-control none -k regex:$K -s ${SKIP:-36} -c 1 -o /tmp/k_src $CMD > gpurun_out/ncu_src.log 2>&1; echo "ncu=$?"
ncu -i /tmp/k_src.ncu-rep --page source --csv > gpurun_out/${NAME:-kernel}_source.csv 2>/dev/null
ncu -i /tmp/k_src.ncu-rep --page raw --csv > gpurun_out/${NAME:-kernel}.raw.csv 2>/dev/null
ls -la gpurun_out | head
