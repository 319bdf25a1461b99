// attention_tc.cu — tensor-core (tcgen05 / TMEM) fused attention for head size 64.
#include "tc_ptx.cuh"

int op_attention_forward_tc(vitrs_ctx* ctx, bf16* out, float* lse, const bf16* qkv, int b, int t, int c, int nh, int causal) {
    (void)ctx; (void)out; (void)lse; (void)qkv; (void)b; (void)t; (void)c; (void)nh; (void)causal;
    return VITRS_ERR_UNSUPPORTED;
}
int op_attention_backward_tc(vitrs_ctx* ctx, bf16* dqkv, const bf16* dout, const bf16* out, const bf16* qkv, const float* lse,
                             int b, int t, int c, int nh, int causal) {
    (void)ctx; (void)dqkv; (void)dout; (void)out; (void)qkv; (void)lse; (void)b; (void)t; (void)c; (void)nh; (void)causal;
    return VITRS_ERR_UNSUPPORTED;
}
