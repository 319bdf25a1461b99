#!/bin/bash
# One gpurun call: ncu launch list of one training step + full metric sets of the first block's forward kernels and the
# last block's backward kernels, reduced to CSV on the box (the .ncu-rep files are too large to bring back).  Each ncu run
# follows a plain run of the same command that exited 0.
set -u
mkdir -p gpurun_out
CMD="python bench.py --config ${CFG:-b16} --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
# launches per step (ViT-B/16): 230; 3 warm-up steps + ~20 set-up kernels precede the timed step
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-710} -c ${COUNT:-232} --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list=$?"
if [ "${FULL:-1}" = "1" ]; then
  K='regex:gemm_tc_kernel|attn_fwd|attn_bwd|ln_fwd|ln_bwd|adamw'
  # matches per step: 146 GEMM + 12 + 12 + 25 + 25 + 1 = 221; forward = 1 + 12*7 + 1 = 86
  ncu --set full --clock-control none -k "$K" -s $((3*221+1)) -c 7 -o /tmp/prof_fwd $CMD > gpurun_out/ncu_full1.log 2>&1
  echo "full1=$?"
  ncu -i /tmp/prof_fwd.ncu-rep --page raw --csv > gpurun_out/prof_fwd_layer.raw.csv 2>/dev/null
  ncu --set full --clock-control none -k "$K" -s $((3*221+86)) -c 13 -o /tmp/prof_bwd $CMD > gpurun_out/ncu_full2.log 2>&1
  echo "full2=$?"
  ncu -i /tmp/prof_bwd.ncu-rep --page raw --csv > gpurun_out/prof_bwd_layer.raw.csv 2>/dev/null
  ncu --set full --clock-control none -k regex:adamw -s 3 -c 1 -o /tmp/prof_adamw $CMD > gpurun_out/ncu_full3.log 2>&1
  ncu -i /tmp/prof_adamw.ncu-rep --page raw --csv > gpurun_out/prof_adamw.raw.csv 2>/dev/null
fi
ls -la gpurun_out | head -20
