mkdir -p gpurun_out
# 1. the diagnostic build first: its mbarrier waits are bounded, so a protocol bug traps instead of hanging the box
VITRS_LIB=$PWD/vit.rs_b200/libvitrs_trace.so timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "attention or contexts" --timeout 60 -x -p no:cacheprovider > gpurun_out/t_attn_trace.log 2>&1; rc=$?; echo "attn tests (trace build) exit=$rc"; tail -n 4 gpurun_out/t_attn_trace.log
[ $rc -ne 0 ] && exit 1
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "attention or contexts" --timeout 60 -x -p no:cacheprovider > gpurun_out/t_attn.log 2>&1; rc=$?; echo "attn tests exit=$rc"; tail -n 4 gpurun_out/t_attn.log
[ $rc -ne 0 ] && exit 1
VITRS_ATTN_BWD_OVERWRITE=1 timeout 120 python scripts/bench_attn.py
VITRS_ATTN_BWD_OVERWRITE=1 T=${T2:-785} B=${B2:-64} timeout 120 python scripts/bench_attn.py
