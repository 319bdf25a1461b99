#!/bin/bash
# ncu --set full (+ source-level stall samples) of one launch of an attention kernel inside scripts/bench_attn.py.
# K = kernel regex, NAME = output stem, SKIP = launches of K to skip (warm-up)
mkdir -p gpurun_out
CMD="python scripts/bench_attn.py"
$CMD > gpurun_out/plain_attn.log 2>&1 || { tail -5 gpurun_out/plain_attn.log; exit 1; }
ncu --set full --import-source on --clock-control none -k regex:$K -s ${SKIP:-4} -c 1 -f -o /tmp/k_attn $CMD > gpurun_out/ncu_attn.log 2>&1; echo "ncu=$?"
ncu -i /tmp/k_attn.ncu-rep --page source --csv > gpurun_out/${NAME:-attn}_source.csv 2>/dev/null
ncu -i /tmp/k_attn.ncu-rep --page raw --csv > gpurun_out/${NAME:-attn}.raw.csv 2>/dev/null
ncu -i /tmp/k_attn.ncu-rep --page details > gpurun_out/${NAME:-attn}_details.txt 2>/dev/null
cat gpurun_out/plain_attn.log
