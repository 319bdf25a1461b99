mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "attention" --timeout 60 -x -p no:cacheprovider > gpurun_out/t_attn.log 2>&1; echo "attn tests exit=$?"; tail -n 4 gpurun_out/t_attn.log
VITRS_ATTN_BWD_OVERWRITE=1 timeout 120 python scripts/bench_attn.py
VITRS_ATTN_FWD_NOSPLIT=1 VITRS_ATTN_BWD_OVERWRITE=1 TAG=nosplit timeout 120 python scripts/bench_attn.py
for t in 144 160 224 256; do VITRS_ATTN_BWD_OVERWRITE=1 T=$t B=256 timeout 120 python scripts/bench_attn.py; done
KERNEL=fwd B=64 FROM=15000 TO=40000 WARPS=1,2,6,10,14 timeout 120 python scripts/attn_trace.py > gpurun_out/trace_fwd_split.txt 2>&1; head -2 gpurun_out/trace_fwd_split.txt
