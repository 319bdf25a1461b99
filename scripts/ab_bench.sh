#!/bin/bash
# Same-box A/B of two builds of the library (boxes differ by +-3 % under the power cap): alternates
# build_ab/libvitrs_base.so (A, a copy of an earlier build) and the in-tree library (B) over REPS rounds of the training-step bench.
mkdir -p gpurun_out
for r in $(seq 1 ${REPS:-2}); do
  for v in A B; do
    if [ $v = A ]; then export VITRS_LIB=$PWD/build_ab/libvitrs_base.so; else unset VITRS_LIB; fi
    timeout 400 python bench.py --config ${CFG:-b16} --steps ${STEPS:-8} --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null > gpurun_out/ab_$v$r.json
    python - <<PY
import json
d=json.load(open("gpurun_out/ab_$v$r.json"))
print("$v$r", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["roofline"]["achieved"], d["roofline"]["gemm_ms_per_step"])
PY
  done
done
