#!/bin/bash
# One gpurun call: the whole -m gpu suite, then the bench line of every single-GPU BASELINE config (tracked copy: profiles/r2_configs.json
# is assembled from gpurun_out/cfg_*.json by scripts/collect_configs.py).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/gputests.log 2>&1; echo "tests exit=$?"; tail -n 6 gpurun_out/gputests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -n 3 gpurun_out/smoke.log
for cfg in b16 ti16 s16 b8; do
  timeout 400 python bench.py --config $cfg --steps ${STEPS:-10} --warmup 3 > gpurun_out/cfg_$cfg.json 2> gpurun_out/cfg_$cfg.err; echo "bench $cfg exit=$?"
  cut -c1-230 gpurun_out/cfg_$cfg.json
done
timeout 200 python bench.py --config tiny --mode f32 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/cfg_tiny_f32.json 2> gpurun_out/cfg_tiny_f32.err; echo "bench tiny exit=$?"; cut -c1-200 gpurun_out/cfg_tiny_f32.json
