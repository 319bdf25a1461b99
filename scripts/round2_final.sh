#!/bin/bash
# One gpurun call with the round's final evidence: the whole -m gpu suite, smoke, the bench line of every single-GPU BASELINE config,
# the ncu launch list of one ViT-B/16 step and of one ViT-Ti/16 step, the full metric set of one forward layer (ViT-B/16) and of the
# streaming attention kernels (ViT-B/8, batch 64).  Every ncu run follows a plain run of the same command that exited 0.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/gputests.log 2>&1; echo "tests exit=$?"; tail -n 2 gpurun_out/gputests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -n 2 gpurun_out/smoke.log
timeout 400 python bench.py --config b16 --steps 10 --warmup 3 > gpurun_out/cfg_b16.json 2> gpurun_out/cfg_b16.err; echo "bench b16 exit=$?"; cut -c1-200 gpurun_out/cfg_b16.json
for cfg in ti16 s16 b8; do
  timeout 300 python bench.py --config $cfg --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/cfg_$cfg.json 2> gpurun_out/cfg_$cfg.err; echo "bench $cfg exit=$?"
  cut -c1-200 gpurun_out/cfg_$cfg.json
done
timeout 200 python bench.py --config tiny --mode f32 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/cfg_tiny_f32.json 2> gpurun_out/cfg_tiny_f32.err; echo "bench tiny exit=$?"
# ---- ncu: launch list + one forward layer (ViT-B/16) ----
FULL=0 VITRS_NO_STEP_GRAPH=1 timeout 500 bash scripts/profile_round.sh > gpurun_out/profile_round.log 2>&1; tail -n 3 gpurun_out/profile_round.log | head -2
CMD="python bench.py --config b16 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
K='regex:gemm_tc_kernel|attn_fwd|attn_bwd|ln_fwd|ln_bwd|adamw'
VITRS_NO_STEP_GRAPH=1 timeout 300 ncu --set full --clock-control none -k "$K" -s $((3*221+1)) -c 7 -f -o /tmp/prof_fwd $CMD > gpurun_out/ncu_full1.log 2>&1; echo "full1=$?"
ncu -i /tmp/prof_fwd.ncu-rep --page raw --csv > gpurun_out/prof_fwd_layer.raw.csv 2>/dev/null
# ---- ncu: streaming attention and the small kernels (ViT-B/8) ----
VITRS_NO_STEP_GRAPH=1 timeout 400 bash scripts/profile_small.sh > gpurun_out/profile_small.log 2>&1; tail -n 2 gpurun_out/profile_small.log
# ---- ncu: launch list of a ViT-Ti/16 step ----
CMD="python bench.py --config ti16 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
VITRS_NO_STEP_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 710 -c 232 --csv --log-file gpurun_out/launches_ti16.csv $CMD > gpurun_out/ncu_list_ti16.log 2>&1; echo "list ti16=$?"
