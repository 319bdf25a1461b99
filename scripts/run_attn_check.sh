mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "attention" --timeout 60 -x -p no:cacheprovider > gpurun_out/t_attn.log 2>&1; echo "attn tests exit=$?"; tail -n 4 gpurun_out/t_attn.log
VITRS_ATTN_BWD_OVERWRITE=1 timeout 120 python scripts/bench_attn.py
VITRS_ATTN_BWD_OVERWRITE=1 B=256 T=785 timeout 120 python scripts/bench_attn.py
VITRS_LIB=$PWD/build_ab/libvitrs_base.so VITRS_ATTN_BWD_OVERWRITE=1 B=256 T=785 TAG=base timeout 120 python scripts/bench_attn.py
KERNEL=bwd B=64 FROM=40000 TO=100000 timeout 120 python scripts/attn_trace.py > gpurun_out/trace_bwd.txt 2>&1; head -2 gpurun_out/trace_bwd.txt
KERNEL=fwd B=64 FROM=15000 TO=45000 timeout 120 python scripts/attn_trace.py > gpurun_out/trace_fwd.txt 2>&1; head -2 gpurun_out/trace_fwd.txt
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_b16_elect.json 2> gpurun_out/bench_b16_elect.err; echo "bench exit=$?"; cut -c1-260 gpurun_out/bench_b16_elect.json
