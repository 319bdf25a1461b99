mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 300 -x -p no:cacheprovider > gpurun_out/gputests.log 2>&1; rc=$?; echo "tests exit=$rc"; tail -n 2 gpurun_out/gputests.log
[ $rc -ne 0 ] && { grep -n "Error\|assert\|FAILED" gpurun_out/gputests.log | head -20; exit 1; }
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -n 2 gpurun_out/smoke.log
timeout 400 python bench.py --config b16 --steps 10 --warmup 3 > gpurun_out/cfg_b16.json 2> gpurun_out/cfg_b16.err; echo "bench b16 exit=$?"; cut -c1-200 gpurun_out/cfg_b16.json
timeout 300 python bench.py --config ti16 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/cfg_ti16.json 2> gpurun_out/cfg_ti16.err; echo "bench ti16 exit=$?"; cut -c1-200 gpurun_out/cfg_ti16.json
FULL=0 VITRS_NO_STEP_GRAPH=1 timeout 500 bash scripts/profile_round.sh > gpurun_out/profile_round.log 2>&1; tail -n 3 gpurun_out/profile_round.log | head -2
CMD="python bench.py --config ti16 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
VITRS_NO_STEP_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 710 -c 232 --csv --log-file gpurun_out/launches_ti16.csv $CMD > gpurun_out/ncu_list_ti16.log 2>&1; echo "list ti16=$?"
