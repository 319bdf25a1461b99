#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a short bench, and (only if the bench exited 0) the ncu launch list.
# Everything is logged under gpurun_out/ because only the tail of stdout comes back.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
for f in test_gpu_gemm test_gpu_ops test_gpu_model; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --timeout 300 --timeout-method thread -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit=$?" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/$f.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit=$?" | tee -a gpurun_out/summary.txt
tail -n 4 gpurun_out/smoke.log
for cfg in ${BENCH_CONFIGS:-ti16 b16}; do
  timeout 900 python bench.py --config $cfg --steps ${BENCH_STEPS:-5} --warmup 3 > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
  echo "bench $cfg exit=$?" | tee -a gpurun_out/summary.txt
  cat gpurun_out/bench_$cfg.json; tail -n 3 gpurun_out/bench_$cfg.err
done
if [ -n "$NCU_LIST" ]; then
  timeout 600 python bench.py --config ti16 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_ti16.csv \
    python bench.py --config ti16 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_run.log 2>&1
  echo "ncu exit=$?" | tee -a gpurun_out/summary.txt
fi
