#!/bin/bash
# One gpurun call: ncu launch list of one training step + full metric sets of one forward and one backward
# block, reduced to CSV on the box (the .ncu-rep files are too large to bring back).  Each ncu run follows a
# plain run of the same command that exited 0.
set -u
mkdir -p gpurun_out
CMD="python bench.py --config ${CFG:-b16} --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-822} -c ${COUNT:-269} --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list=$?"
K='regex:gemm_tc_kernel|attn_fwd_tc|attn_bwd_tc|ln_fwd|ln_bwd|colsum|adamw'
if [ -n "${FULL:-1}" ]; then
  ncu --set full --clock-control none -k "$K" -s ${FSKIP1:-687} -c ${FCOUNT1:-8} -o /tmp/prof_fwd $CMD > gpurun_out/ncu_full1.log 2>&1
  echo "full1=$?"
  ncu -i /tmp/prof_fwd.ncu-rep --page raw --csv > gpurun_out/prof_fwd_layer.raw.csv 2>/dev/null
  ncu --set full --clock-control none -k "$K" -s ${FSKIP2:-774} -c ${FCOUNT2:-16} -o /tmp/prof_bwd $CMD > gpurun_out/ncu_full2.log 2>&1
  echo "full2=$?"
  ncu -i /tmp/prof_bwd.ncu-rep --page raw --csv > gpurun_out/prof_bwd_layer.raw.csv 2>/dev/null
fi
ls -la gpurun_out | head -20
