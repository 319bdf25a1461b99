// model.cu — the L2 model object: storage, forward, backward, update, the training step, data
// parallel gradient exchange and the llm.c-style checkpoint.  Mirrors `struct ViT` and
// `impl ViT { build_from_checkpoint, forward, backward }` of rusty_vit.rs:63-450 (and the
// #[repr(C)] flavour, train_vit.rs:65-374) with the ViT-specific pieces of DEVIATIONS.md D7/D8.
//
// Two modes share the storage conventions (flat fp32 params / grads / m / v in the reference's
// tensor-major order, rusty_vit.rs:105-122; activations stacked [L, ...], :150-174):
//   VITRS_MODE_F32   verify: the reference's op sequence, one launch per op, every activation
//                    and activation gradient materialised (incl. preatt / att), fp32 throughout.
//   VITRS_MODE_BF16  production: bf16 activations and weight shadows, fused GEMM epilogues
//                    (bias, bias+GELU, bias+residual, GELU-backward), flash-style attention that
//                    keeps only lse, LayerNorm-backward fused with the residual accumulation and
//                    the next bias gradient; one layer's worth of gradient scratch.
#include <stdlib.h>

#include "model.cuh"

int vitrs_nccl_allreduce_group(vitrs_ctx* ctx, float* const* bufs, const size_t* counts, int count);


namespace {


void param_sizes_of(const vitrs_config& cfg, size_t* s) {
    const size_t c = cfg.channels, l = cfg.num_layers, t = cfg.max_seq_len;
    const size_t kdim = 3u * cfg.patch_size * cfg.patch_size;
    s[P_PATCHW] = c * kdim; s[P_PATCHB] = c; s[P_CLS] = c; s[P_WPE] = t * c;
    s[P_LN1W] = l * c; s[P_LN1B] = l * c; s[P_QKVW] = l * 3 * c * c; s[P_QKVB] = l * 3 * c;
    s[P_ATTPROJW] = l * c * c; s[P_ATTPROJB] = l * c; s[P_LN2W] = l * c; s[P_LN2B] = l * c;
    s[P_FCW] = l * 4 * c * c; s[P_FCB] = l * 4 * c; s[P_FCPROJW] = l * c * 4 * c; s[P_FCPROJB] = l * c;
    s[P_LNFW] = c; s[P_LNFB] = c; s[P_HEADW] = (size_t)cfg.num_classes * c; s[P_HEADB] = cfg.num_classes;
}

// elements per image of each activation (rusty_vit.rs:150-174 with the batch factor restored)
void act_sizes_of(const vitrs_config& cfg, size_t* s) {
    const size_t T = cfg.max_seq_len, C = cfg.channels, L = cfg.num_layers, NH = cfg.num_heads, V = cfg.num_classes;
    s[A_ENCODED] = T * C; s[A_LN1] = L * T * C; s[A_LN1_MEAN] = L * T; s[A_LN1_RSTD] = L * T;
    s[A_QKV] = L * T * 3 * C; s[A_ATTY] = L * T * C; s[A_PREATT] = L * NH * T * T; s[A_ATT] = L * NH * T * T;
    s[A_ATTPROJ] = L * T * C; s[A_RESIDUAL2] = L * T * C; s[A_LN2] = L * T * C; s[A_LN2_MEAN] = L * T;
    s[A_LN2_RSTD] = L * T; s[A_FCH] = L * T * 4 * C; s[A_FCH_GELU] = L * T * 4 * C; s[A_FCPROJ] = L * T * C;
    s[A_RESIDUAL3] = L * T * C; s[A_LNF] = C; s[A_LNF_MEAN] = 1; s[A_LNF_RSTD] = 1;
    s[A_LOGITS] = V; s[A_PROBS] = V; s[A_LOSSES] = 1;
}

bool act_is_f32_always(int a) {
    return a == A_LN1_MEAN || a == A_LN1_RSTD || a == A_LN2_MEAN || a == A_LN2_RSTD || a == A_PREATT || a == A_ATT || a >= A_LNF;
}

// element size and byte offset of every view of an activation arena (host arithmetic; also vitrs_model_footprint); returns its size
size_t arena_layout(const vitrs_config& cfg, int max_batch, int mode, bool is_grad, size_t* per_image, int* elems, size_t* offs) {
    act_sizes_of(cfg, per_image);
    size_t off = 0;
    for (int a = 0; a < A_COUNT; ++a) {
        int elem = (mode == VITRS_MODE_F32 || act_is_f32_always(a)) ? 4 : 2;
        if (mode == VITRS_MODE_BF16) {
            // the fused path never materialises these; its gradient arena keeps only the head tensors
            if (a == A_PREATT || a == A_ATT || a == A_ATTPROJ || a == A_FCPROJ) elem = 0;
            if (is_grad && a < A_LNF) elem = 0;
        }
        elems[a] = elem;
        offs[a] = off;
        off += (per_image[a] * max_batch * elem + 255) / 256 * 256;
    }
    return off;
}

int arena_create(vitrs_ctx* ctx, Arena* ar, const vitrs_config& cfg, int max_batch, int mode, bool is_grad) {
    size_t offs[A_COUNT];
    const size_t off = arena_layout(cfg, max_batch, mode, is_grad, ar->per_image, ar->elem, offs);
    ar->bytes = off;
    ar->base = nullptr;
    if (off) VITRS_CUDA(ctx, cudaMalloc(&ar->base, off));
    for (int a = 0; a < A_COUNT; ++a) ar->view[a] = ar->elem[a] ? ar->base + offs[a] : nullptr;
    return VITRS_OK;
}

// ---- GEMM call shapes ----------------------------------------------------------------------
template <typename T>
int gemm_dx(vitrs_ctx* ctx, T* dinp, const T* dout, const T* w, long rows, int c, int oc, int kind, const T* aux, int accumulate) {
    GemmDesc g = {};
    g.A = dout; g.a_rs = oc; g.a_ks = 1;
    g.B = w; g.b_rs = 1; g.b_ks = c;
    g.M = (int)rows; g.N = c; g.K = oc;
    g.epi.kind = kind; g.epi.aux = aux; g.epi.accumulate = accumulate; g.epi.out = dinp; g.epi.ldo = c;
    return gemm_dispatch<T>(ctx, g);
}
template <typename T>
int gemm_dw(vitrs_ctx* ctx, float* dw, const T* dout, const T* inp, long rows, int c, int oc, float* dbias = nullptr) {
    GemmDesc g = {};
    g.a_colsum = dbias;  // bias gradient = column sums of dout, fused into the tensor-core weight-gradient GEMM
    g.A = dout; g.a_rs = 1; g.a_ks = oc;
    g.B = inp; g.b_rs = 1; g.b_ks = c;
    g.M = oc; g.N = c; g.K = (int)rows;
    g.epi.kind = EPI_ACCUM_F32; g.epi.out = dw; g.epi.ldo = c;
    return gemm_dispatch<T>(ctx, g);
}

struct Dims {
    int B, T, C, L, NH, V, img, patch, kdim;
    long btc;
};
Dims dims_of(const vitrs_model* m) {
    Dims d;
    d.B = m->batch; d.T = m->cfg.max_seq_len; d.C = m->cfg.channels; d.L = m->cfg.num_layers; d.NH = m->cfg.num_heads;
    d.V = m->cfg.num_classes; d.img = m->cfg.image_size; d.patch = m->cfg.patch_size; d.kdim = 3 * d.patch * d.patch;
    d.btc = (long)d.B * d.T * d.C;
    return d;
}

template <typename T> T* act(const vitrs_model* m, int a) { return reinterpret_cast<T*>(m->acts.view[a]); }
template <typename T> T* gact(const vitrs_model* m, int a) { return reinterpret_cast<T*>(m->gacts.view[a]); }

// ---- head: CLS rows -> final LayerNorm -> logits -> softmax / loss (rusty_vit.rs:335-347) --------
template <typename T>
int head_forward(vitrs_model* m, const T* last_residual, bool fused_loss) {
    vitrs_ctx* ctx = m->ctx;
    const Dims d = dims_of(m);
    VITRS_TRY(op_cls_gather<T>(ctx, m->cls_rows, last_residual, d.B, d.T, d.C));
    VITRS_TRY(op_layernorm_forward<float>(ctx, act<float>(m, A_LNF), act<float>(m, A_LNF_MEAN), act<float>(m, A_LNF_RSTD),
                                          m->cls_rows, P(m, P_LNFW), P(m, P_LNFB), d.B, d.C));
    VITRS_TRY((gemm_fwd<float>(ctx, act<float>(m, A_LOGITS), act<float>(m, A_LNF), P(m, P_HEADW), P(m, P_HEADB), d.B, d.C, d.V,
                               EPI_BIAS, nullptr, nullptr)));
    VITRS_CUDA(ctx, cudaMemsetAsync(m->d_mean_loss, 0, sizeof(float), ctx->stream));
    const float scale = m->dloss_scale != 0.f ? m->dloss_scale : 1.0f / (float)d.B;
    if (fused_loss) {
        // probs, losses, mean loss and dlogits in one pass (production)
        if (m->has_targets) VITRS_CUDA(ctx, cudaMemsetAsync(m->dlogits, 0, sizeof(float) * (size_t)d.B * d.V, ctx->stream));
        VITRS_TRY(op_head_loss(ctx, act<float>(m, A_PROBS), act<float>(m, A_LOSSES), m->d_mean_loss,
                               m->has_targets ? m->dlogits : nullptr, act<float>(m, A_LOGITS), m->has_targets ? m->labels : nullptr,
                               d.B, d.V, scale));
    } else {
        VITRS_TRY(op_softmax_forward(ctx, act<float>(m, A_PROBS), act<float>(m, A_LOGITS), d.B, d.V));
        if (m->has_targets) {
            VITRS_TRY(op_crossentropy_forward(ctx, act<float>(m, A_LOSSES), act<float>(m, A_PROBS), m->labels, d.B, d.V));
            VITRS_TRY(op_scaled_sum(ctx, m->d_mean_loss, act<float>(m, A_LOSSES), d.B, scale));
        }
    }
    return VITRS_OK;
}

// ---- verify mode: the reference's op sequence (rusty_vit.rs:282-347) -------------------------------
int forward_f32(vitrs_model* m) {
    vitrs_ctx* ctx = m->ctx;
    const Dims d = dims_of(m);
    const long btc = d.btc;
    float* patches = reinterpret_cast<float*>(m->patches);
    if (m->images_u8) VITRS_TRY(op_im2col_u8<float>(ctx, patches, m->images_u8, m->u8_layout, m->norm_mean, m->norm_std, d.B, d.img, d.patch));
    else VITRS_TRY(op_im2col<float>(ctx, patches, m->images, d.B, d.img, d.patch));
    {
        GemmDesc g = {};
        g.A = patches; g.a_rs = d.kdim; g.a_ks = 1;
        g.B = P(m, P_PATCHW); g.b_rs = d.kdim; g.b_ks = 1;
        g.M = d.B * d.T; g.N = d.C; g.K = d.kdim;
        g.epi.kind = EPI_PATCH; g.epi.bias = P(m, P_PATCHB); g.epi.cls = P(m, P_CLS); g.epi.pos = P(m, P_WPE); g.epi.np = d.T;
        g.epi.out = act<float>(m, A_ENCODED); g.epi.ldo = d.C;
        VITRS_TRY(gemm_simt_f32(ctx, g));
    }
    const int C = d.C, L = d.L;
    for (int l = 0; l < L; ++l) {
        const float* residual = l == 0 ? act<float>(m, A_ENCODED) : act<float>(m, A_RESIDUAL3) + (l - 1) * btc;
        const long lbt = (long)l * d.B * d.T, latt = (long)l * d.B * d.NH * d.T * d.T;
        float* ln1 = act<float>(m, A_LN1) + l * btc;
        float* qkv = act<float>(m, A_QKV) + l * btc * 3;
        float* atty = act<float>(m, A_ATTY) + l * btc;
        float* attproj = act<float>(m, A_ATTPROJ) + l * btc;
        float* residual2 = act<float>(m, A_RESIDUAL2) + l * btc;
        float* ln2 = act<float>(m, A_LN2) + l * btc;
        float* fch = act<float>(m, A_FCH) + l * btc * 4;
        float* fch_gelu = act<float>(m, A_FCH_GELU) + l * btc * 4;
        float* fcproj = act<float>(m, A_FCPROJ) + l * btc;
        float* residual3 = act<float>(m, A_RESIDUAL3) + l * btc;
        const long rows = (long)d.B * d.T;
        VITRS_TRY(op_layernorm_forward<float>(ctx, ln1, act<float>(m, A_LN1_MEAN) + lbt, act<float>(m, A_LN1_RSTD) + lbt, residual,
                                              P(m, P_LN1W) + l * C, P(m, P_LN1B) + l * C, rows, C));
        VITRS_TRY((gemm_fwd<float>(ctx, qkv, ln1, P(m, P_QKVW) + (long)l * 3 * C * C, P(m, P_QKVB) + l * 3 * C, rows, C, 3 * C,
                                   EPI_BIAS, nullptr, nullptr)));
        VITRS_TRY(op_attention_forward<float>(ctx, atty, act<float>(m, A_PREATT) + latt, act<float>(m, A_ATT) + latt,
                                              m->lse + (long)l * d.B * d.NH * d.T, qkv, d.B, d.T, C, d.NH, m->cfg.causal));
        VITRS_TRY((gemm_fwd<float>(ctx, attproj, atty, P(m, P_ATTPROJW) + (long)l * C * C, P(m, P_ATTPROJB) + l * C, rows, C, C,
                                   EPI_BIAS, nullptr, nullptr)));
        VITRS_TRY(op_residual_forward<float>(ctx, residual2, residual, attproj, btc));
        VITRS_TRY(op_layernorm_forward<float>(ctx, ln2, act<float>(m, A_LN2_MEAN) + lbt, act<float>(m, A_LN2_RSTD) + lbt, residual2,
                                              P(m, P_LN2W) + l * C, P(m, P_LN2B) + l * C, rows, C));
        VITRS_TRY((gemm_fwd<float>(ctx, fch, ln2, P(m, P_FCW) + (long)l * 4 * C * C, P(m, P_FCB) + l * 4 * C, rows, C, 4 * C,
                                   EPI_BIAS, nullptr, nullptr)));
        VITRS_TRY(op_gelu_forward<float>(ctx, fch_gelu, fch, btc * 4));
        VITRS_TRY((gemm_fwd<float>(ctx, fcproj, fch_gelu, P(m, P_FCPROJW) + (long)l * C * 4 * C, P(m, P_FCPROJB) + l * C, rows,
                                   4 * C, C, EPI_BIAS, nullptr, nullptr)));
        VITRS_TRY(op_residual_forward<float>(ctx, residual3, residual2, fcproj, btc));
    }
    return head_forward<float>(m, act<float>(m, A_RESIDUAL3) + (L - 1) * btc, false);
}

// rusty_vit.rs:354-449
int backward_f32(vitrs_model* m) {
    vitrs_ctx* ctx = m->ctx;
    const Dims d = dims_of(m);
    const long btc = d.btc;
    const int C = d.C, L = d.L, V = d.V;
    const long rows = (long)d.B * d.T;
    const float dloss = m->dloss_scale != 0.f ? m->dloss_scale : 1.0f / (float)d.B;
    VITRS_TRY(op_fill_const(ctx, gact<float>(m, A_LOSSES), d.B, dloss));  // rusty_vit.rs:366-369
    VITRS_TRY(op_crossentropy_softmax_backward(ctx, gact<float>(m, A_LOGITS), gact<float>(m, A_LOSSES), act<float>(m, A_PROBS),
                                               m->labels, d.B, V));
    VITRS_TRY((gemm_dx<float>(ctx, gact<float>(m, A_LNF), gact<float>(m, A_LOGITS), P(m, P_HEADW), d.B, C, V, EPI_NONE, nullptr, 1)));
    VITRS_TRY((gemm_dw<float>(ctx, G(m, P_HEADW), gact<float>(m, A_LOGITS), act<float>(m, A_LNF), d.B, C, V)));
    VITRS_TRY(op_colsum<float>(ctx, G(m, P_HEADB), gact<float>(m, A_LOGITS), d.B, V, V));
    VITRS_TRY(op_layernorm_backward<float>(ctx, m->dcls_rows, G(m, P_LNFW), G(m, P_LNFB), gact<float>(m, A_LNF), m->cls_rows,
                                           P(m, P_LNFW), act<float>(m, A_LNF_MEAN), act<float>(m, A_LNF_RSTD), d.B, C, nullptr));
    VITRS_TRY(op_cls_scatter_add<float>(ctx, gact<float>(m, A_RESIDUAL3) + (L - 1) * btc, m->dcls_rows, d.B, d.T, C));

    for (int l = L - 1; l >= 0; --l) {
        const float* residual = l == 0 ? act<float>(m, A_ENCODED) : act<float>(m, A_RESIDUAL3) + (l - 1) * btc;
        float* dresidual = l == 0 ? gact<float>(m, A_ENCODED) : gact<float>(m, A_RESIDUAL3) + (l - 1) * btc;
        const long lbt = (long)l * d.B * d.T, latt = (long)l * d.B * d.NH * d.T * d.T;
        // op order: rusty_vit.rs:436-445
        VITRS_TRY(op_residual_backward<float>(ctx, gact<float>(m, A_RESIDUAL2) + l * btc, gact<float>(m, A_FCPROJ) + l * btc,
                                              gact<float>(m, A_RESIDUAL3) + l * btc, btc));
        {
            const float* dout = gact<float>(m, A_FCPROJ) + l * btc;
            VITRS_TRY((gemm_dx<float>(ctx, gact<float>(m, A_FCH_GELU) + l * btc * 4, dout, P(m, P_FCPROJW) + (long)l * C * 4 * C, rows,
                                      4 * C, C, EPI_NONE, nullptr, 1)));
            VITRS_TRY((gemm_dw<float>(ctx, G(m, P_FCPROJW) + (long)l * C * 4 * C, dout, act<float>(m, A_FCH_GELU) + l * btc * 4, rows,
                                      4 * C, C)));
            VITRS_TRY(op_colsum<float>(ctx, G(m, P_FCPROJB) + l * C, dout, rows, C, C));
        }
        VITRS_TRY(op_gelu_backward<float>(ctx, gact<float>(m, A_FCH) + l * btc * 4, act<float>(m, A_FCH) + l * btc * 4,
                                          gact<float>(m, A_FCH_GELU) + l * btc * 4, btc * 4));
        {
            const float* dout = gact<float>(m, A_FCH) + l * btc * 4;
            VITRS_TRY((gemm_dx<float>(ctx, gact<float>(m, A_LN2) + l * btc, dout, P(m, P_FCW) + (long)l * 4 * C * C, rows, C, 4 * C,
                                      EPI_NONE, nullptr, 1)));
            VITRS_TRY((gemm_dw<float>(ctx, G(m, P_FCW) + (long)l * 4 * C * C, dout, act<float>(m, A_LN2) + l * btc, rows, C, 4 * C)));
            VITRS_TRY(op_colsum<float>(ctx, G(m, P_FCB) + l * 4 * C, dout, rows, 4 * C, 4 * C));
        }
        VITRS_TRY(op_layernorm_backward<float>(ctx, gact<float>(m, A_RESIDUAL2) + l * btc, G(m, P_LN2W) + l * C, G(m, P_LN2B) + l * C,
                                               gact<float>(m, A_LN2) + l * btc, act<float>(m, A_RESIDUAL2) + l * btc,
                                               P(m, P_LN2W) + l * C, act<float>(m, A_LN2_MEAN) + lbt, act<float>(m, A_LN2_RSTD) + lbt,
                                               rows, C, nullptr));
        VITRS_TRY(op_residual_backward<float>(ctx, dresidual, gact<float>(m, A_ATTPROJ) + l * btc, gact<float>(m, A_RESIDUAL2) + l * btc,
                                              btc));
        {
            const float* dout = gact<float>(m, A_ATTPROJ) + l * btc;
            VITRS_TRY((gemm_dx<float>(ctx, gact<float>(m, A_ATTY) + l * btc, dout, P(m, P_ATTPROJW) + (long)l * C * C, rows, C, C,
                                      EPI_NONE, nullptr, 1)));
            VITRS_TRY((gemm_dw<float>(ctx, G(m, P_ATTPROJW) + (long)l * C * C, dout, act<float>(m, A_ATTY) + l * btc, rows, C, C)));
            VITRS_TRY(op_colsum<float>(ctx, G(m, P_ATTPROJB) + l * C, dout, rows, C, C));
        }
        VITRS_TRY(op_attention_backward<float>(ctx, gact<float>(m, A_QKV) + l * btc * 3, gact<float>(m, A_PREATT) + latt,
                                               gact<float>(m, A_ATT) + latt, gact<float>(m, A_ATTY) + l * btc,
                                               act<float>(m, A_QKV) + l * btc * 3, act<float>(m, A_ATT) + latt, nullptr, d.B, d.T, C,
                                               d.NH, m->cfg.causal));
        {
            const float* dout = gact<float>(m, A_QKV) + l * btc * 3;
            VITRS_TRY((gemm_dx<float>(ctx, gact<float>(m, A_LN1) + l * btc, dout, P(m, P_QKVW) + (long)l * 3 * C * C, rows, C, 3 * C,
                                      EPI_NONE, nullptr, 1)));
            VITRS_TRY((gemm_dw<float>(ctx, G(m, P_QKVW) + (long)l * 3 * C * C, dout, act<float>(m, A_LN1) + l * btc, rows, C, 3 * C)));
            VITRS_TRY(op_colsum<float>(ctx, G(m, P_QKVB) + l * 3 * C, dout, rows, 3 * C, 3 * C));
        }
        VITRS_TRY(op_layernorm_backward<float>(ctx, dresidual, G(m, P_LN1W) + l * C, G(m, P_LN1B) + l * C,
                                               gact<float>(m, A_LN1) + l * btc, residual, P(m, P_LN1W) + l * C,
                                               act<float>(m, A_LN1_MEAN) + lbt, act<float>(m, A_LN1_RSTD) + lbt, rows, C, nullptr));
    }
    // patch embedding backward (D7)
    const float* denc = gact<float>(m, A_ENCODED);
    VITRS_TRY(op_patch_backward_reduce<float>(ctx, G(m, P_WPE), G(m, P_CLS), G(m, P_PATCHB), denc, d.B, d.T, C));
    VITRS_TRY((gemm_dw<float>(ctx, G(m, P_PATCHW), denc, reinterpret_cast<const float*>(m->patches), rows, d.kdim, C)));
    return VITRS_OK;
}

// ---- production mode -----------------------------------------------------------------------------
int forward_bf16(vitrs_model* m) {
    vitrs_ctx* ctx = m->ctx;
    const Dims d = dims_of(m);
    const long btc = d.btc;
    const int C = d.C, L = d.L;
    const long rows = (long)d.B * d.T;
    bf16* patches = reinterpret_cast<bf16*>(m->patches);
    if (m->images_u8) VITRS_TRY(op_im2col_u8<bf16>(ctx, patches, m->images_u8, m->u8_layout, m->norm_mean, m->norm_std, d.B, d.img, d.patch));
    else VITRS_TRY(op_im2col<bf16>(ctx, patches, m->images, d.B, d.img, d.patch));
    {
        GemmDesc g = {};
        g.A = patches; g.a_rs = d.kdim; g.a_ks = 1;
        g.B = S(m, P_PATCHW); g.b_rs = d.kdim; g.b_ks = 1;
        g.M = (int)rows; g.N = C; g.K = d.kdim;
        g.epi.kind = EPI_PATCH; g.epi.bias = P(m, P_PATCHB); g.epi.cls = P(m, P_CLS); g.epi.pos = P(m, P_WPE); g.epi.np = d.T;
        g.epi.out = act<bf16>(m, A_ENCODED); g.epi.ldo = C;
        VITRS_TRY(gemm_tc_bf16(ctx, g));
    }
    for (int l = 0; l < L; ++l) {
        const bf16* residual = l == 0 ? act<bf16>(m, A_ENCODED) : act<bf16>(m, A_RESIDUAL3) + (l - 1) * btc;
        const long lbt = (long)l * rows;
        bf16* ln1 = act<bf16>(m, A_LN1) + l * btc;
        bf16* qkv = act<bf16>(m, A_QKV) + l * btc * 3;
        bf16* atty = act<bf16>(m, A_ATTY) + l * btc;
        bf16* residual2 = act<bf16>(m, A_RESIDUAL2) + l * btc;
        bf16* ln2 = act<bf16>(m, A_LN2) + l * btc;
        bf16* fch = act<bf16>(m, A_FCH) + l * btc * 4;
        bf16* fch_gelu = act<bf16>(m, A_FCH_GELU) + l * btc * 4;
        bf16* residual3 = act<bf16>(m, A_RESIDUAL3) + l * btc;
        float* lse = m->lse + (long)l * d.B * d.NH * d.T;
        VITRS_TRY(op_layernorm_forward<bf16>(ctx, ln1, act<float>(m, A_LN1_MEAN) + lbt, act<float>(m, A_LN1_RSTD) + lbt, residual,
                                             P(m, P_LN1W) + l * C, P(m, P_LN1B) + l * C, rows, C));
        VITRS_TRY((gemm_fwd<bf16>(ctx, qkv, ln1, S(m, P_QKVW) + (long)l * 3 * C * C, P(m, P_QKVB) + l * 3 * C, rows, C, 3 * C,
                                  EPI_BIAS, nullptr, nullptr)));
        int r = op_attention_forward_tc(ctx, atty, lse, qkv, d.B, d.T, C, d.NH, m->cfg.causal);
        if (r == VITRS_ERR_UNSUPPORTED)
            r = op_attention_forward<bf16>(ctx, atty, nullptr, nullptr, lse, qkv, d.B, d.T, C, d.NH, m->cfg.causal);
        VITRS_TRY(r);
        // out-projection + residual_forward in one epilogue (rusty_vit.rs:325-326)
        VITRS_TRY((gemm_fwd<bf16>(ctx, residual2, atty, S(m, P_ATTPROJW) + (long)l * C * C, P(m, P_ATTPROJB) + l * C, rows, C, C,
                                  EPI_BIAS_RESIDUAL, residual, nullptr)));
        VITRS_TRY(op_layernorm_forward<bf16>(ctx, ln2, act<float>(m, A_LN2_MEAN) + lbt, act<float>(m, A_LN2_RSTD) + lbt, residual2,
                                             P(m, P_LN2W) + l * C, P(m, P_LN2B) + l * C, rows, C));
        // fc + gelu_forward (rusty_vit.rs:328-329): both fch and fch_gelu are kept for backward
        // (inference — no targets, rusty_vit.rs:339-350 — keeps only the activated output: no backward will read fch)
        if (m->has_targets)
            VITRS_TRY((gemm_fwd<bf16>(ctx, fch, ln2, S(m, P_FCW) + (long)l * 4 * C * C, P(m, P_FCB) + l * 4 * C, rows, C, 4 * C,
                                      EPI_BIAS_GELU, nullptr, fch_gelu)));
        else
            VITRS_TRY((gemm_fwd<bf16>(ctx, fch_gelu, ln2, S(m, P_FCW) + (long)l * 4 * C * C, P(m, P_FCB) + l * 4 * C, rows, C, 4 * C,
                                      EPI_BIAS_GELU_ONLY, nullptr, nullptr)));
        VITRS_TRY((gemm_fwd<bf16>(ctx, residual3, fch_gelu, S(m, P_FCPROJW) + (long)l * C * 4 * C, P(m, P_FCPROJB) + l * C, rows,
                                  4 * C, C, EPI_BIAS_RESIDUAL, residual2, nullptr)));
    }
    return head_forward<bf16>(m, act<bf16>(m, A_RESIDUAL3) + (L - 1) * btc, true);
}

int allreduce_layer(vitrs_model* m, int l);
int allreduce_tail(vitrs_model* m, bool head);

int backward_bf16(vitrs_model* m) {
    vitrs_ctx* ctx = m->ctx;
    const Dims d = dims_of(m);
    const long btc = d.btc;
    const int C = d.C, L = d.L, V = d.V;
    const long rows = (long)d.B * d.T;
    // head (fp32): dlogits was produced with the loss
    VITRS_TRY((gemm_dx<float>(ctx, m->dlnf, m->dlogits, P(m, P_HEADW), d.B, C, V, EPI_NONE, nullptr, 0)));
    VITRS_TRY((gemm_dw<float>(ctx, G(m, P_HEADW), m->dlogits, act<float>(m, A_LNF), d.B, C, V)));
    VITRS_TRY(op_colsum<float>(ctx, G(m, P_HEADB), m->dlogits, d.B, V, V));
    VITRS_CUDA(ctx, cudaMemsetAsync(m->dcls_rows, 0, sizeof(float) * (size_t)d.B * C, ctx->stream));
    VITRS_TRY(op_layernorm_backward<float>(ctx, m->dcls_rows, G(m, P_LNFW), G(m, P_LNFB), m->dlnf, m->cls_rows, P(m, P_LNFW),
                                           act<float>(m, A_LNF_MEAN), act<float>(m, A_LNF_RSTD), d.B, C, nullptr));
    bf16* dres = m->dres;
    VITRS_CUDA(ctx, cudaMemsetAsync(dres, 0, sizeof(bf16) * (size_t)btc, ctx->stream));
    VITRS_TRY(op_cls_scatter_add<bf16>(ctx, dres, m->dcls_rows, d.B, d.T, C));
    VITRS_TRY(allreduce_tail(m, true));

    for (int l = L - 1; l >= 0; --l) {
        const bf16* residual = l == 0 ? act<bf16>(m, A_ENCODED) : act<bf16>(m, A_RESIDUAL3) + (l - 1) * btc;
        const long lbt = (long)l * rows;
        const bf16* fch = act<bf16>(m, A_FCH) + l * btc * 4;
        const bf16* fch_gelu = act<bf16>(m, A_FCH_GELU) + l * btc * 4;
        const bf16* ln2 = act<bf16>(m, A_LN2) + l * btc;
        const bf16* residual2 = act<bf16>(m, A_RESIDUAL2) + l * btc;
        const bf16* atty = act<bf16>(m, A_ATTY) + l * btc;
        const bf16* qkv = act<bf16>(m, A_QKV) + l * btc * 3;
        const bf16* ln1 = act<bf16>(m, A_LN1) + l * btc;
        const float* lse = m->lse + (long)l * d.B * d.NH * d.T;
        bf16* dfch = m->dbig;
        bf16* dqkv = m->dbig;
        bf16* dln = m->dln;
        // dres == dresidual3[l] == dfcproj[l] (residual_backward, rusty_vit.rs:436)
        // fcproj matmul_backward + gelu_backward in one epilogue: dfch = (dres . Wfcproj) * gelu'(fch)
        VITRS_TRY((gemm_dx<bf16>(ctx, dfch, dres, S(m, P_FCPROJW) + (long)l * C * 4 * C, rows, 4 * C, C, EPI_GELU_BWD, fch, 0)));
        VITRS_TRY((gemm_dw<bf16>(ctx, G(m, P_FCPROJW) + (long)l * C * 4 * C, dres, fch_gelu, rows, 4 * C, C, G(m, P_FCPROJB) + l * C)));
        VITRS_TRY((gemm_dx<bf16>(ctx, dln, dfch, S(m, P_FCW) + (long)l * 4 * C * C, rows, C, 4 * C, EPI_NONE, nullptr, 0)));
        VITRS_TRY((gemm_dw<bf16>(ctx, G(m, P_FCW) + (long)l * 4 * C * C, dfch, ln2, rows, C, 4 * C, G(m, P_FCB) + l * 4 * C)));
        // dresidual2 = dres + LN2-backward(dln2)
        VITRS_TRY(op_layernorm_backward<bf16>(ctx, dres, G(m, P_LN2W) + l * C, G(m, P_LN2B) + l * C, dln, residual2,
                                              P(m, P_LN2W) + l * C, act<float>(m, A_LN2_MEAN) + lbt, act<float>(m, A_LN2_RSTD) + lbt,
                                              rows, C, nullptr));
        // datty = dres . Wattproj; with 64-wide heads the same epilogue also takes D = rowsum(datty * atty) per (image, head,
        // query), the row term of the softmax backward (it replaces a pass over datty and atty)
        const bool fuse_d = C == d.NH * 64;
        if (fuse_d) {
            VITRS_TRY(vitrs_ensure_scratch(ctx, (size_t)d.B * d.NH * d.T));
            GemmDesc g = {};
            g.A = dres; g.a_rs = C; g.a_ks = 1;
            g.B = S(m, P_ATTPROJW) + (long)l * C * C; g.b_rs = 1; g.b_ks = C;
            g.M = (int)rows; g.N = C; g.K = C;
            g.epi.kind = EPI_ROWDOT; g.epi.aux = atty; g.epi.out = dln; g.epi.out2 = ctx->scratch; g.epi.ldo = C; g.epi.np = d.T;
            VITRS_TRY(gemm_dispatch<bf16>(ctx, g));
        } else {
            VITRS_TRY((gemm_dx<bf16>(ctx, dln, dres, S(m, P_ATTPROJW) + (long)l * C * C, rows, C, C, EPI_NONE, nullptr, 0)));
        }
        VITRS_TRY((gemm_dw<bf16>(ctx, G(m, P_ATTPROJW) + (long)l * C * C, dres, atty, rows, C, C, G(m, P_ATTPROJB) + l * C)));
        int r = op_attention_backward_tc(ctx, dqkv, dln, atty, qkv, lse, d.B, d.T, C, d.NH, m->cfg.causal, 0, fuse_d ? ctx->scratch : nullptr);
        if (r == VITRS_ERR_UNSUPPORTED) {
            VITRS_CUDA(ctx, cudaMemsetAsync(dqkv, 0, sizeof(bf16) * (size_t)btc * 3, ctx->stream));
            r = op_attention_backward<bf16>(ctx, dqkv, nullptr, nullptr, dln, qkv, nullptr, lse, d.B, d.T, C, d.NH, m->cfg.causal);
        }
        VITRS_TRY(r);
        VITRS_TRY((gemm_dx<bf16>(ctx, dln, dqkv, S(m, P_QKVW) + (long)l * 3 * C * C, rows, C, 3 * C, EPI_NONE, nullptr, 0)));  // dln1
        VITRS_TRY((gemm_dw<bf16>(ctx, G(m, P_QKVW) + (long)l * 3 * C * C, dqkv, ln1, rows, C, 3 * C, G(m, P_QKVB) + l * 3 * C)));
        VITRS_TRY(op_layernorm_backward<bf16>(ctx, dres, G(m, P_LN1W) + l * C, G(m, P_LN1B) + l * C, dln, residual,
                                              P(m, P_LN1W) + l * C, act<float>(m, A_LN1_MEAN) + lbt, act<float>(m, A_LN1_RSTD) + lbt,
                                              rows, C, nullptr));
        // every gradient of block l is final
        VITRS_TRY(allreduce_layer(m, l));
    }
    VITRS_TRY(op_patch_backward_reduce<bf16>(ctx, G(m, P_WPE), G(m, P_CLS), G(m, P_PATCHB), dres, d.B, d.T, C));
    VITRS_TRY((gemm_dw<bf16>(ctx, G(m, P_PATCHW), dres, reinterpret_cast<const bf16*>(m->patches), rows, d.kdim, C)));
    VITRS_TRY(allreduce_tail(m, false));
    return VITRS_OK;
}

// ---- data parallel: bucketed gradient exchange on the comm stream, overlapped with backward ----------
int bucket_begin(vitrs_model* m) {
    vitrs_ctx* ctx = m->ctx;
    VITRS_CUDA(ctx, cudaEventRecord(m->ev_bucket, ctx->stream));
    VITRS_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, m->ev_bucket, 0));
    return VITRS_OK;
}

// Bucket b of the gradient exchange, in the order backward completes them: 0 = final LayerNorm + head,
// 1..L = blocks L-1..0 (the block's 12 slices of the tensor-major buffer), L+1 = patch / cls / position
// embeddings.  Returns the slice count; offsets / counts are in elements of the flat gradient buffer.
// big[i] = 1 for the GEMM weight matrices (patchw, qkvw, attprojw, fcw, fcprojw: 98.7 % of ViT-B/16), which the kernels
// read through the bf16 shadow and ZeRO-1 shards; every other tensor (gains, biases, embeddings, the class head) is read in
// fp32 by the kernels and stays replicated.
int bucket_slices(const vitrs_config& cfg, const size_t* sizes, const size_t* offs, int bucket, size_t* out_off, size_t* out_cnt,
                  int* big = nullptr) {
    const int L = cfg.num_layers;
    int dummy[12];
    if (!big) big = dummy;
    if (bucket == 0) {  // lnfw, lnfb, headw, headb are contiguous
        out_off[0] = offs[P_LNFW];
        out_cnt[0] = sizes[P_LNFW] + sizes[P_LNFB] + sizes[P_HEADW] + sizes[P_HEADB];
        big[0] = 0;
        return 1;
    }
    if (bucket == L + 1) {  // patchw | patchb, cls, wpe (contiguous)
        out_off[0] = offs[P_PATCHW]; out_cnt[0] = sizes[P_PATCHW]; big[0] = 1;
        out_off[1] = offs[P_PATCHB]; out_cnt[1] = offs[P_LN1W] - offs[P_PATCHB]; big[1] = 0;
        return 2;
    }
    const int l = L - bucket;
    int n = 0;
    for (int i = P_LN1W; i <= P_FCPROJB; ++i) {
        const size_t per = sizes[i] / L;
        out_off[n] = offs[i] + (size_t)l * per;
        out_cnt[n] = per;
        big[n] = i == P_QKVW || i == P_ATTPROJW || i == P_FCW || i == P_FCPROJW;
        ++n;
    }
    return n;
}

// The exchange buffer holds the buckets back to back ("Z order": the layer-major layout SURVEY section 7(v) asks for, as a view
// for the wire while the named parameter views stay tensor-major).  Inside a bucket's region the big slices come first, padded
// to a multiple of 8 * world elements so that the world shards are equal and 16-byte aligned (ZeRO-1 reduce-scatter), then
// the small slices, padded to 8.
void zplan_sizes(const vitrs_config& cfg, const size_t* sizes, const size_t* offs, int world, int bucket, size_t* big_len, size_t* small_len) {
    size_t off[12], cnt[12], tb = 0, ts = 0;
    int big[12];
    const int n = bucket_slices(cfg, sizes, offs, bucket, off, cnt, big);
    for (int i = 0; i < n; ++i) (big[i] ? tb : ts) += cnt[i];
    const size_t q = (size_t)8 * world;
    *big_len = (tb + q - 1) / q * q;
    *small_len = (ts + 7) / 8 * 8;
}

int ensure_zplan(vitrs_model* m) {
    vitrs_ctx* ctx = m->ctx;
    if (m->z_world == ctx->world && m->comm_buf) return VITRS_OK;
    VITRS_ARG(ctx, !m->zero1);  // the shards were cut for another world size
    VITRS_CUDA(ctx, cudaStreamSynchronize(ctx->comm_stream));
    if (m->comm_buf) VITRS_CUDA(ctx, cudaFree(m->comm_buf));
    m->comm_buf = nullptr;
    const int nb = m->cfg.num_layers + 2;
    free(m->z_off);
    m->z_off = (size_t*)calloc(5 * (size_t)nb, sizeof(size_t));
    m->z_len = m->z_off + nb; m->z_big = m->z_len + nb; m->z_shard = m->z_big + nb; m->s_off = m->z_shard + nb;
    size_t z = 0, sh = 0;
    for (int b = 0; b < nb; ++b) {
        size_t small = 0;
        zplan_sizes(m->cfg, m->param_sizes, m->param_off, ctx->world, b, &m->z_big[b], &small);
        m->z_len[b] = m->z_big[b] + small;
        m->z_shard[b] = m->z_big[b] / ctx->world;
        m->z_off[b] = z; m->s_off[b] = sh;
        z += m->z_len[b]; sh += m->z_shard[b];
    }
    m->z_size = z; m->s_total = sh; m->z_buckets = nb;
    VITRS_CUDA(ctx, cudaMalloc(&m->comm_buf, z * sizeof(bf16)));
    VITRS_CUDA(ctx, cudaMemsetAsync(m->comm_buf, 0, z * sizeof(bf16), ctx->comm_stream));  // the padding stays zero for ever
    m->z_world = ctx->world;
    return VITRS_OK;
}

// which: 0 = every slice, 1 = the big slices only, 2 = the small slices only (offsets in the exchange buffer are the same)
SliceTable table_of(const vitrs_model* m, int bucket, int which = 0) {
    SliceTable t;
    memset(&t, 0, sizeof(t));
    size_t off[12], cnt[12];
    int big[12];
    const int n = bucket_slices(m->cfg, m->param_sizes, m->param_off, bucket, off, cnt, big);
    size_t zb = m->z_off[bucket], zs = m->z_off[bucket] + m->z_big[bucket];
    for (int i = 0; i < n; ++i) {
        size_t& z = big[i] ? zb : zs;
        if (which == 0 || (which == 1) == (big[i] != 0)) {
            t.src_off[t.n] = off[i]; t.cnt[t.n] = cnt[i]; t.z_off[t.n] = z;
            t.n++;
        }
        z += cnt[i];
    }
    return t;
}

// Gradient exchange of one bucket.  The pack runs on the COMPUTE stream, right behind the kernels that produced the bucket's
// gradients, where it has the whole GPU (28 MB in, 14 MB out: ~10 us); only the collective itself goes to the comm stream, and
// the unpack of every bucket waits for comm_join.  (Packing and unpacking on the comm stream cost 45-95 us per bucket EACH:
// those kernels can only run on SMs the persistent GEMMs are not holding, and their 1184 blocks delayed the next GEMM's
// clusters by about as long as the collective itself — measured, profiles/r2_timeline_dp2_*.txt.)
int exchange_bucket(vitrs_model* m, int bk) {
    vitrs_ctx* ctx = m->ctx;
    if (m->comm_dtype == 0 && !m->zero1) {  // exact fp32 sums, slice by slice in place
        size_t off[12], cnt[12];
        float* bufs[12];
        const int n = bucket_slices(m->cfg, m->param_sizes, m->param_off, bk, off, cnt);
        for (int i = 0; i < n; ++i) bufs[i] = m->grads + off[i];
        VITRS_TRY(bucket_begin(m));
        return vitrs_nccl_allreduce_group(ctx, bufs, cnt, n);
    }
    // one contiguous bf16 message per bucket (bf16 on the wire in production)
    bf16* region = m->comm_buf + m->z_off[bk];
    VITRS_TRY(op_pack_f32_to_bf16(ctx, m->comm_buf, m->grads, table_of(m, bk), ctx->stream));
    VITRS_TRY(bucket_begin(m));
    m->unpack_pending |= 1ull << bk;
    if (m->zero1) {
        // big slices: each rank receives the sum of its 1/world shard (AdamW runs on the shard in update()); small slices
        // stay replicated: summed everywhere and unpacked into the fp32 gradient views
        const size_t small = m->z_len[bk] - m->z_big[bk];
        VITRS_TRY(vitrs_nccl_group(ctx, 1));
        if (m->z_shard[bk]) VITRS_TRY(vitrs_nccl_reduce_scatter(ctx, region, region + (size_t)ctx->rank * m->z_shard[bk], m->z_shard[bk], 1));
        if (small) VITRS_TRY(vitrs_nccl_allreduce(ctx, region + m->z_big[bk], region + m->z_big[bk], small, 1));
        return vitrs_nccl_group(ctx, 0);
    }
    return vitrs_nccl_allreduce(ctx, region, region, m->z_len[bk], 1);
}

int allreduce_bucket(vitrs_model* m, int bucket) {
    if (m->mode != VITRS_MODE_BF16 || (!m->ctx->nccl_comm && !m->zero1)) return VITRS_OK;
    // VITRS_DP_DEFER (tuning aid): every bucket is exchanged after the last gradient kernel instead of behind its block
    const bool defer = m->ctx->env_dp_defer != 0;
    const int last = m->cfg.num_layers + 1;
    if (defer && bucket != last) return VITRS_OK;
    if (m->comm_dtype != 0 || m->zero1) VITRS_TRY(ensure_zplan(m));
    for (int bk = defer ? 0 : bucket; bk <= bucket; ++bk) VITRS_TRY(exchange_bucket(m, bk));
    return VITRS_OK;
}

// block l: every slice of its bucket is final when the block's backward has been issued (the bias gradients come out of
// the weight-gradient GEMMs of the same block)
int allreduce_layer(vitrs_model* m, int l) { return allreduce_bucket(m, m->cfg.num_layers - l); }
int allreduce_tail(vitrs_model* m, bool head) { return allreduce_bucket(m, head ? 0 : m->cfg.num_layers + 1); }

int comm_join(vitrs_model* m) {
    vitrs_ctx* ctx = m->ctx;
    if (!ctx->nccl_comm && !m->zero1) return VITRS_OK;
    VITRS_CUDA(ctx, cudaEventRecord(m->ev_comm_done, ctx->comm_stream));
    VITRS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, m->ev_comm_done, 0));
    // the summed bf16 gradients go back into the fp32 views (ZeRO-1: of the small, replicated slices only)
    for (int bk = 0; bk < m->z_buckets && m->unpack_pending; ++bk) {
        if (!((m->unpack_pending >> bk) & 1ull)) continue;
        VITRS_TRY(op_unpack_bf16_to_f32(ctx, m->grads, m->comm_buf, table_of(m, bk, m->zero1 ? 2 : 0), ctx->stream));
    }
    m->unpack_pending = 0;
    return VITRS_OK;
}

// The local value is sum(losses) * dloss_scale; under data parallel the global mean is the sum over ranks.  It is reduced ONCE
// per forward, out of place, on the comm stream, so vitrs_model_mean_loss stays a pure local read (no collective hides in a
// getter: reading twice, or on one rank only, is safe).
int reduce_loss(vitrs_model* m) {
    vitrs_ctx* ctx = m->ctx;
    m->loss_reduced = 0;
    if (!ctx->nccl_comm || !m->has_targets) return VITRS_OK;
    VITRS_TRY(bucket_begin(m));
    VITRS_TRY(vitrs_nccl_allreduce(ctx, m->d_mean_loss, m->d_mean_loss + 1, 1, 0));
    m->loss_reduced = 1;
    return VITRS_OK;
}

// ---- ZeRO-1: fp32 master weights and AdamW moments of the GEMM weight matrices sharded 1/world per bucket ------
// The small tensors (everything the kernels read in fp32) keep replicated fp32 weights in the parameter view and replicated
// moments in two compact arrays; they are five contiguous runs of the tensor-major buffer.
struct SmallRun { size_t off, cnt, soff; };
int small_runs_of(const size_t* param_off, const size_t* param_sizes, SmallRun* r) {
    const int first[5] = {P_PATCHB, P_QKVB, P_ATTPROJB, P_FCB, P_FCPROJB}, last[5] = {P_LN1B, P_QKVB, P_LN2B, P_FCB, P_HEADB};
    size_t so = 0;
    for (int i = 0; i < 5; ++i) {
        r[i].off = param_off[first[i]];
        r[i].cnt = param_off[last[i]] + param_sizes[last[i]] - r[i].off;
        r[i].soff = so;
        so += (r[i].cnt + 3) / 4 * 4;  // 16-byte aligned starts
    }
    return 5;
}
int small_runs(const vitrs_model* m, SmallRun* r) { return small_runs_of(m->param_off, m->param_sizes, r); }

// (re)cut the shards from the full tensor-major buffers params / m / v, then drop the full moment buffers
int zero_shard_from_full(vitrs_model* m) {
    vitrs_ctx* ctx = m->ctx;
    m->zero1 = 0;
    VITRS_TRY(ensure_zplan(m));
    VITRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SmallRun runs[5];
    small_runs(m, runs);
    const size_t small_total = runs[4].soff + (runs[4].cnt + 3) / 4 * 4;
    if (!m->zp) {
        VITRS_CUDA(ctx, cudaMalloc(&m->zp, (m->s_total + 4) * sizeof(float)));
        VITRS_CUDA(ctx, cudaMalloc(&m->zm, (m->s_total + 4) * sizeof(float)));
        VITRS_CUDA(ctx, cudaMalloc(&m->zv, (m->s_total + 4) * sizeof(float)));
        VITRS_CUDA(ctx, cudaMalloc(&m->m_small, small_total * sizeof(float)));
        VITRS_CUDA(ctx, cudaMalloc(&m->v_small, small_total * sizeof(float)));
        VITRS_CUDA(ctx, cudaMemsetAsync(m->m_small, 0, small_total * sizeof(float), ctx->comm_stream));
        VITRS_CUDA(ctx, cudaMemsetAsync(m->v_small, 0, small_total * sizeof(float), ctx->comm_stream));
        m->small_total = small_total;
    }
    float* tmp = nullptr;
    VITRS_CUDA(ctx, cudaMalloc(&tmp, m->z_size * sizeof(float)));
    float* full[3] = {m->params, m->m, m->v};
    float* shard[3] = {m->zp, m->zm, m->zv};
    float* small[3] = {nullptr, m->m_small, m->v_small};
    for (int k = 0; k < 3; ++k) {
        if (!full[k]) continue;  // moments already dropped: they stay as they are
        VITRS_CUDA(ctx, cudaMemsetAsync(tmp, 0, m->z_size * sizeof(float), ctx->comm_stream));
        for (int b = 0; b < m->z_buckets; ++b) {
            if (!m->z_shard[b]) continue;
            VITRS_TRY(op_pack_f32_to_f32(ctx, tmp, full[k], table_of(m, b, 1), ctx->comm_stream));
            VITRS_CUDA(ctx, cudaMemcpyAsync(shard[k] + m->s_off[b], tmp + m->z_off[b] + (size_t)ctx->rank * m->z_shard[b],
                                            m->z_shard[b] * sizeof(float), cudaMemcpyDeviceToDevice, ctx->comm_stream));
        }
        if (small[k])
            for (int i = 0; i < 5; ++i)
                VITRS_CUDA(ctx, cudaMemcpyAsync(small[k] + runs[i].soff, full[k] + runs[i].off, runs[i].cnt * sizeof(float),
                                                cudaMemcpyDeviceToDevice, ctx->comm_stream));
    }
    VITRS_CUDA(ctx, cudaStreamSynchronize(ctx->comm_stream));
    VITRS_CUDA(ctx, cudaFree(tmp));
    if (m->m) { VITRS_CUDA(ctx, cudaFree(m->m)); m->m = nullptr; }
    if (m->v) { VITRS_CUDA(ctx, cudaFree(m->v)); m->v = nullptr; }
    m->zero1 = 1;
    return VITRS_OK;
}

// all-gather one sharded fp32 array (zp / zm / zv) into the big slices of a full tensor-major buffer; `small` (nullable): the
// compact replicated array that fills the small runs
int zero_gather_full(vitrs_model* m, const float* shards, const float* small, float* full) {
    vitrs_ctx* ctx = m->ctx;
    float* tmp = nullptr;
    VITRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    VITRS_CUDA(ctx, cudaMalloc(&tmp, m->z_size * sizeof(float)));
    for (int b = 0; b < m->z_buckets; ++b) {
        if (!m->z_shard[b]) continue;
        float* region = tmp + m->z_off[b];
        float* mine = region + (size_t)ctx->rank * m->z_shard[b];
        VITRS_CUDA(ctx, cudaMemcpyAsync(mine, shards + m->s_off[b], m->z_shard[b] * sizeof(float), cudaMemcpyDeviceToDevice, ctx->comm_stream));
        VITRS_TRY(vitrs_nccl_all_gather(ctx, mine, region, m->z_shard[b], 0));
        VITRS_TRY(op_unpack_f32_to_f32(ctx, full, tmp, table_of(m, b, 1), ctx->comm_stream));
    }
    if (small) {
        SmallRun runs[5];
        small_runs(m, runs);
        for (int i = 0; i < 5; ++i)
            VITRS_CUDA(ctx, cudaMemcpyAsync(full + runs[i].off, small + runs[i].soff, runs[i].cnt * sizeof(float), cudaMemcpyDeviceToDevice,
                                            ctx->comm_stream));
    }
    VITRS_CUDA(ctx, cudaStreamSynchronize(ctx->comm_stream));
    VITRS_CUDA(ctx, cudaFree(tmp));
    return VITRS_OK;
}

// AdamW on this rank's shard of every bucket (gradients: the reduce-scattered bf16 sums sitting in the exchange buffer), the
// updated bf16 weights all-gathered in place and scattered into the tensor-major shadow the GEMMs read; the small tensors are
// updated in place, replicated, from the summed fp32 gradients
int zero_update(vitrs_model* m) {
    vitrs_ctx* ctx = m->ctx;
    VITRS_TRY(bucket_begin(m));  // the comm stream continues behind whatever the compute stream has issued (hyper-parameters)
    for (int b = 0; b < m->z_buckets; ++b) {
        if (!m->z_shard[b]) continue;
        bf16* mine = m->comm_buf + m->z_off[b] + (size_t)ctx->rank * m->z_shard[b];
        VITRS_TRY(op_adamw_apply_shard(ctx, m->zp + m->s_off[b], mine, m->zm + m->s_off[b], m->zv + m->s_off[b], m->z_shard[b], ctx->comm_stream));
    }
    VITRS_TRY(vitrs_nccl_group(ctx, 1));
    for (int b = 0; b < m->z_buckets; ++b) {
        bf16* region = m->comm_buf + m->z_off[b];
        if (m->z_shard[b]) VITRS_TRY(vitrs_nccl_all_gather(ctx, region + (size_t)ctx->rank * m->z_shard[b], region, m->z_shard[b], 1));
    }
    VITRS_TRY(vitrs_nccl_group(ctx, 0));
    SmallRun runs[5];
    small_runs(m, runs);
    for (int i = 0; i < 5; ++i)
        VITRS_TRY(op_adamw_apply(ctx, m->params + runs[i].off, m->grads + runs[i].off, m->m_small + runs[i].soff, m->v_small + runs[i].soff,
                                 runs[i].cnt, m->shadow + runs[i].off, ctx->comm_stream));
    for (int b = 0; b < m->z_buckets; ++b)
        if (m->z_shard[b]) VITRS_TRY(op_unpack_bf16_to_bf16(ctx, m->shadow, m->comm_buf, table_of(m, b, 1), ctx->comm_stream));
    return comm_join(m);
}

int set_batch(vitrs_model* m, const float* images, const int* labels, int b) {
    vitrs_ctx* ctx = m->ctx;
    VITRS_ARG(ctx, images != nullptr && b >= 1 && b <= m->max_batch);
    m->batch = b;
    m->images = images;
    m->images_u8 = nullptr;
    m->labels = labels;
    m->has_targets = labels != nullptr;
    return VITRS_OK;
}

}  // namespace

extern "C" {

int vitrs_model_create(vitrs_ctx* ctx, const vitrs_config* cfg_in, int max_batch, int mode, vitrs_model** out) {
    if (!ctx) return VITRS_ERR_ARG;
    VITRS_ARG(ctx, cfg_in && out && max_batch >= 1 && (mode == VITRS_MODE_F32 || mode == VITRS_MODE_BF16));
    vitrs_config cfg = *cfg_in;
    VITRS_ARG(ctx, cfg.patch_size > 0 && cfg.image_size % cfg.patch_size == 0 && cfg.patch_size % 4 == 0);
    VITRS_ARG(ctx, cfg.channels > 0 && cfg.num_heads > 0 && cfg.channels % cfg.num_heads == 0 && cfg.channels % 8 == 0);
    VITRS_ARG(ctx, cfg.num_layers >= 1 && cfg.num_layers <= 62 && cfg.num_classes >= 1);  // (one bit per gradient bucket)
    VITRS_ARG(ctx, cfg.max_seq_len == 0 || cfg.max_seq_len == tokens(cfg));
    cfg.max_seq_len = tokens(cfg);
    cfg.vocab_size = cfg.num_classes;
    *out = nullptr;
    VITRS_CUDA(ctx, cudaSetDevice(ctx->device));
    vitrs_model* m = (vitrs_model*)calloc(1, sizeof(vitrs_model));
    m->ctx = ctx; m->cfg = cfg; m->mode = mode; m->max_batch = max_batch;
    m->comm_dtype = mode == VITRS_MODE_BF16 ? 1 : 0;
    for (int i = 0; i < 3; ++i) { m->norm_mean[i] = 0.5f; m->norm_std[i] = 0.5f; }  // uint8 -> [-1, 1]
    param_sizes_of(cfg, m->param_sizes);
    size_t off = 0;
    for (int i = 0; i < P_COUNT; ++i) { m->param_off[i] = off; off += m->param_sizes[i]; }
    m->num_params = off;
    const size_t pbytes = off * sizeof(float);
    const size_t B = max_batch, T = cfg.max_seq_len, C = cfg.channels, L = cfg.num_layers, NH = cfg.num_heads, V = cfg.num_classes;
    const size_t kdim = 3u * cfg.patch_size * cfg.patch_size;
    const size_t esz = mode == VITRS_MODE_F32 ? 4 : 2;
#define MODEL_CUDA(expr)                                                                                           \
    do {                                                                                                           \
        cudaError_t e__ = (expr);                                                                                  \
        if (e__ != cudaSuccess) {                                                                                  \
            int rc__ = vitrs_set_error(ctx, VITRS_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,           \
                                       cudaGetErrorString(e__));                                                   \
            vitrs_model_destroy(m);                                                                                \
            return rc__;                                                                                           \
        }                                                                                                          \
    } while (0)
    MODEL_CUDA(cudaMalloc(&m->params, pbytes));
    MODEL_CUDA(cudaMalloc(&m->grads, pbytes));
    MODEL_CUDA(cudaMalloc(&m->m, pbytes));
    MODEL_CUDA(cudaMalloc(&m->v, pbytes));
    MODEL_CUDA(cudaMemsetAsync(m->params, 0, pbytes, ctx->stream));
    MODEL_CUDA(cudaMemsetAsync(m->grads, 0, pbytes, ctx->stream));
    MODEL_CUDA(cudaMemsetAsync(m->m, 0, pbytes, ctx->stream));
    MODEL_CUDA(cudaMemsetAsync(m->v, 0, pbytes, ctx->stream));
    if (mode == VITRS_MODE_BF16) {
        MODEL_CUDA(cudaMalloc(&m->shadow, off * sizeof(bf16)));
        MODEL_CUDA(cudaMemsetAsync(m->shadow, 0, off * sizeof(bf16), ctx->stream));
    }
    if (arena_create(ctx, &m->acts, cfg, max_batch, mode, false) != VITRS_OK ||
        arena_create(ctx, &m->gacts, cfg, max_batch, mode, true) != VITRS_OK) {
        vitrs_model_destroy(m);
        return VITRS_ERR_CUDA;
    }
    MODEL_CUDA(cudaMalloc(&m->lse, sizeof(float) * L * B * NH * T));
    MODEL_CUDA(cudaMalloc(&m->cls_rows, sizeof(float) * B * C));
    MODEL_CUDA(cudaMalloc(&m->dcls_rows, sizeof(float) * B * C));
    MODEL_CUDA(cudaMalloc(&m->patches, esz * B * T * kdim));
    if (mode == VITRS_MODE_BF16) {
        MODEL_CUDA(cudaMalloc(&m->dres, sizeof(bf16) * B * T * C));
        MODEL_CUDA(cudaMalloc(&m->dln, sizeof(bf16) * B * T * C));
        MODEL_CUDA(cudaMalloc(&m->dbig, sizeof(bf16) * B * T * 4 * C));
        MODEL_CUDA(cudaMalloc(&m->dlogits, sizeof(float) * B * V));
        MODEL_CUDA(cudaMalloc(&m->dlnf, sizeof(float) * B * C));
    }
    MODEL_CUDA(cudaMalloc(&m->d_mean_loss, 2 * sizeof(float)));  // [0] local, [1] summed over ranks
    MODEL_CUDA(cudaMallocHost(&m->h_mean_loss, 2 * sizeof(float)));  // [0] loss, [1] device error flags
    *m->h_mean_loss = -1.0f;
    for (int i = 0; i < 2; ++i) {
        MODEL_CUDA(cudaEventCreateWithFlags(&m->stage_ready[i], cudaEventDisableTiming));
        MODEL_CUDA(cudaEventCreateWithFlags(&m->stage_free[i], cudaEventDisableTiming));
    }
    MODEL_CUDA(cudaEventCreateWithFlags(&m->ev_bucket, cudaEventDisableTiming));
    MODEL_CUDA(cudaEventCreateWithFlags(&m->ev_comm_done, cudaEventDisableTiming));
    MODEL_CUDA(cudaStreamSynchronize(ctx->stream));
#undef MODEL_CUDA
    *out = m;
    return VITRS_OK;
}

int vitrs_model_destroy(vitrs_model* m) {
    if (!m) return VITRS_OK;
    cudaSetDevice(m->ctx->device);
    cudaDeviceSynchronize();
    cudaFree(m->params); cudaFree(m->grads); cudaFree(m->m); cudaFree(m->v); cudaFree(m->shadow);
    cudaFree(m->acts.base); cudaFree(m->gacts.base);
    cudaFree(m->lse); cudaFree(m->cls_rows); cudaFree(m->dcls_rows); cudaFree(m->patches);
    cudaFree(m->dres); cudaFree(m->dln); cudaFree(m->dbig); cudaFree(m->dlogits); cudaFree(m->dlnf);
    cudaFree(m->d_mean_loss);
    cudaFree(m->comm_buf); cudaFree(m->zp); cudaFree(m->zm); cudaFree(m->zv); cudaFree(m->m_small); cudaFree(m->v_small);
    free(m->z_off);
    if (m->h_mean_loss) cudaFreeHost(m->h_mean_loss);
    for (int i = 0; i < 2; ++i) {
        cudaFree(m->stage_images[i]); cudaFree(m->stage_labels[i]);
        if (m->stage_ready[i]) cudaEventDestroy(m->stage_ready[i]);
        if (m->stage_free[i]) cudaEventDestroy(m->stage_free[i]);
    }
    if (m->ev_bucket) cudaEventDestroy(m->ev_bucket);
    if (m->ev_comm_done) cudaEventDestroy(m->ev_comm_done);
    for (StepGraph& g : m->step_graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    if (m->cap_stream) cudaStreamDestroy(m->cap_stream);
    free(m);
    return VITRS_OK;
}

static int refresh_shadow(vitrs_model* m) {
    if (m->mode != VITRS_MODE_BF16) return VITRS_OK;
    return op_cast_f32_bf16(m->ctx, m->shadow, m->params, m->num_params);
}

// init_parameters (rusty_vit.rs:864-903, DEVIATIONS D9): weights U[lo,hi) from the counter
// generator (stream = tensor index), LayerNorm gains 1, every bias 0
int vitrs_model_init_parameters(vitrs_model* m, uint64_t seed, int init_mode) {
    if (!m) return VITRS_ERR_ARG;
    vitrs_ctx* ctx = m->ctx;
    VITRS_CUDA(ctx, cudaMemsetAsync(m->params, 0, m->num_params * sizeof(float), ctx->stream));
    if (m->m) VITRS_CUDA(ctx, cudaMemsetAsync(m->m, 0, m->num_params * sizeof(float), ctx->stream));
    if (m->v) VITRS_CUDA(ctx, cudaMemsetAsync(m->v, 0, m->num_params * sizeof(float), ctx->stream));
    m->adam_step = 0;
    const float lo = init_mode == 1 ? -0.02f : 0.0f, hi = 0.02f;
    const int weight_ids[] = {P_PATCHW, P_CLS, P_WPE, P_QKVW, P_ATTPROJW, P_FCW, P_FCPROJW, P_HEADW};
    for (int id : weight_ids) VITRS_TRY(op_fill_uniform(ctx, P(m, id), m->param_sizes[id], seed, (uint64_t)id, lo, hi));
    const int gain_ids[] = {P_LN1W, P_LN2W, P_LNFW};
    for (int id : gain_ids) VITRS_TRY(op_fill_const(ctx, P(m, id), m->param_sizes[id], 1.0f));
    VITRS_TRY(refresh_shadow(m));
    if (m->zero1) {
        VITRS_CUDA(ctx, cudaMemsetAsync(m->zm, 0, m->s_total * sizeof(float), ctx->stream));
        VITRS_CUDA(ctx, cudaMemsetAsync(m->zv, 0, m->s_total * sizeof(float), ctx->stream));
        return zero_shard_from_full(m);
    }
    return VITRS_OK;
}

size_t vitrs_model_num_parameters(vitrs_model* m) { return m ? m->num_params : 0; }

int vitrs_model_param_view(vitrs_model* m, int which, int tensor, float** ptr, size_t* count) {
    if (!m) return VITRS_ERR_ARG;
    VITRS_ARG(m->ctx, which >= 0 && which <= 3 && tensor >= 0 && tensor < P_COUNT && ptr);
    float* base = which == 0 ? m->params : which == 1 ? m->grads : which == 2 ? m->m : m->v;
    if (!base)  // ZeRO-1 keeps the moments as shards only
        return vitrs_set_error(m->ctx, VITRS_ERR_UNSUPPORTED, "the AdamW moments are sharded (ZeRO-1): no full view");
    *ptr = base + m->param_off[tensor];
    if (count) *count = m->param_sizes[tensor];
    return VITRS_OK;
}

int vitrs_model_act_view(vitrs_model* m, int which, int tensor, void** ptr, size_t* count, int* elem_size) {
    if (!m) return VITRS_ERR_ARG;
    VITRS_ARG(m->ctx, (which == 0 || which == 1) && tensor >= 0 && tensor < A_COUNT && ptr);
    const Arena& ar = which == 0 ? m->acts : m->gacts;
    *ptr = ar.view[tensor];
    if (count) *count = ar.view[tensor] ? ar.per_image[tensor] * (size_t)m->batch : 0;
    if (elem_size) *elem_size = ar.elem[tensor];
    return VITRS_OK;
}

int vitrs_model_set_dloss_scale(vitrs_model* m, float scale) {
    if (!m) return VITRS_ERR_ARG;
    if (m->dloss_scale != scale) m->graph_epoch++;  // the scale is a launch argument of the loss kernel: recorded graphs are stale
    m->dloss_scale = scale;
    return VITRS_OK;
}

// weights changed behind the model's back (tests writing through param_view): rebuild the bf16 shadow
int vitrs_model_sync_parameters(vitrs_model* m) {
    if (!m) return VITRS_ERR_ARG;
    VITRS_TRY(refresh_shadow(m));
    if (m->zero1) return zero_shard_from_full(m);  // the master shards follow the full view
    return VITRS_OK;
}

int vitrs_model_forward(vitrs_model* m, const float* images, const int* labels, int b) {
    if (!m) return VITRS_ERR_ARG;
    VITRS_TRY(set_batch(m, images, labels, b));
    VITRS_CUDA(m->ctx, cudaSetDevice(m->ctx->device));
    VITRS_TRY(m->mode == VITRS_MODE_F32 ? forward_f32(m) : forward_bf16(m));
    if (!m->has_targets) VITRS_TRY(op_fill_const(m->ctx, m->d_mean_loss, 1, -1.0f));  // rusty_vit.rs:348-350
    return reduce_loss(m);
}

// ---- raw image batches: uint8 samples, normalisation fused into the im2col pass ----
int vitrs_model_set_input_norm(vitrs_model* m, const float* mean, const float* stdev) {
    if (!m) return VITRS_ERR_ARG;
    VITRS_ARG(m->ctx, mean && stdev && stdev[0] > 0.f && stdev[1] > 0.f && stdev[2] > 0.f);
    for (int i = 0; i < 3; ++i) { m->norm_mean[i] = mean[i]; m->norm_std[i] = stdev[i]; }
    m->graph_epoch++;  // the constants are launch arguments of the im2col pass: recorded graphs are stale
    return VITRS_OK;
}

int vitrs_model_forward_u8(vitrs_model* m, const uint8_t* images, int layout, const int* labels, int b) {
    if (!m) return VITRS_ERR_ARG;
    vitrs_ctx* ctx = m->ctx;
    VITRS_ARG(ctx, images != nullptr && (layout == 0 || layout == 1) && b >= 1 && b <= m->max_batch);
    m->batch = b;
    m->images = nullptr;
    m->images_u8 = images;
    m->u8_layout = layout;
    m->labels = labels;
    m->has_targets = labels != nullptr;
    VITRS_CUDA(ctx, cudaSetDevice(ctx->device));
    VITRS_TRY(m->mode == VITRS_MODE_F32 ? forward_f32(m) : forward_bf16(m));
    if (!m->has_targets) VITRS_TRY(op_fill_const(ctx, m->d_mean_loss, 1, -1.0f));
    return reduce_loss(m);
}

int vitrs_model_zero_grad(vitrs_model* m) {
    if (!m) return VITRS_ERR_ARG;
    vitrs_ctx* ctx = m->ctx;
    VITRS_CUDA(ctx, cudaMemsetAsync(m->grads, 0, m->num_params * sizeof(float), ctx->stream));
    if (m->gacts.base) VITRS_CUDA(ctx, cudaMemsetAsync(m->gacts.base, 0, m->gacts.bytes, ctx->stream));
    if (m->mode == VITRS_MODE_F32) VITRS_CUDA(ctx, cudaMemsetAsync(m->dcls_rows, 0, sizeof(float) * (size_t)m->max_batch * m->cfg.channels, ctx->stream));
    return VITRS_OK;
}

int vitrs_model_backward(vitrs_model* m) {
    if (!m) return VITRS_ERR_ARG;
    vitrs_ctx* ctx = m->ctx;
    VITRS_ARG(ctx, m->batch > 0 && m->has_targets);
    VITRS_CUDA(ctx, cudaSetDevice(ctx->device));
    VITRS_TRY(m->mode == VITRS_MODE_F32 ? backward_f32(m) : backward_bf16(m));
    if (ctx->nccl_comm && m->mode == VITRS_MODE_F32) VITRS_TRY(vitrs_allreduce_f32(ctx, m->grads, m->num_params));
    return comm_join(m);
}

// host-only description of the bucketed exchange (no device needed): slices of bucket `bucket` for `cfg`
int vitrs_grad_bucket(const vitrs_config* cfg_in, int bucket, size_t* offsets, size_t* counts, int* big, int* num_slices) {
    if (!cfg_in || !offsets || !counts || !num_slices) return VITRS_ERR_ARG;
    vitrs_config cfg = *cfg_in;
    cfg.max_seq_len = tokens(cfg);
    if (bucket < 0 || bucket > cfg.num_layers + 1) return VITRS_ERR_ARG;
    size_t sizes[P_COUNT], offs[P_COUNT], off = 0;
    param_sizes_of(cfg, sizes);
    for (int i = 0; i < P_COUNT; ++i) { offs[i] = off; off += sizes[i]; }
    *num_slices = bucket_slices(cfg, sizes, offs, bucket, offsets, counts, big);
    return VITRS_OK;
}

int vitrs_model_allreduce_grads(vitrs_model* m) {
    if (!m) return VITRS_ERR_ARG;
    return vitrs_allreduce_f32(m->ctx, m->grads, m->num_params);
}

int vitrs_model_optimizer_step(vitrs_model* m, float lr) {
    if (!m) return VITRS_ERR_ARG;
    VITRS_ARG(m->ctx, !m->zero1);  // the reference's SGD works on the full replicated buffers
    return op_sgd(m->ctx, m->params, m->grads, m->num_params, lr, m->shadow);
}

int vitrs_model_update(vitrs_model* m, float lr, float beta1, float beta2, float eps, float weight_decay) {
    if (!m) return VITRS_ERR_ARG;
    vitrs_ctx* ctx = m->ctx;
    m->adam_step += 1;
    VITRS_TRY(op_adam_set_hyper(ctx, lr, beta1, beta2, eps, weight_decay, m->adam_step, ctx->stream));
    if (m->zero1) return zero_update(m);
    return op_adamw_apply(ctx, m->params, m->grads, m->m, m->v, m->num_params, m->shadow, ctx->stream);
}

int vitrs_model_mean_loss(vitrs_model* m, float* out) {
    if (!m) return VITRS_ERR_ARG;
    vitrs_ctx* ctx = m->ctx;
    VITRS_ARG(ctx, out != nullptr);
    if (m->batch == 0) { *out = -1.0f; return VITRS_OK; }
    // a pure local read: under data parallel the sum over ranks was taken once, by forward, on the comm stream
    const int reduced = m->loss_reduced && m->has_targets;
    if (reduced) VITRS_CUDA(ctx, cudaStreamSynchronize(ctx->comm_stream));
    VITRS_CUDA(ctx, cudaMemcpyAsync(m->h_mean_loss, m->d_mean_loss + (reduced ? 1 : 0), sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    VITRS_CUDA(ctx, cudaMemcpyAsync(m->h_mean_loss + 1, ctx->dev_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    VITRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = m->h_mean_loss[0];
    if (reinterpret_cast<const int*>(m->h_mean_loss)[1] & 1) {  // raised by the loss kernels
        VITRS_CUDA(ctx, cudaMemsetAsync(ctx->dev_flags, 0, sizeof(int), ctx->stream));
        return vitrs_set_error(ctx, VITRS_ERR_ARG, "a class label of the batch is outside [0, %d)", m->cfg.num_classes);
    }
    return VITRS_OK;
}

// ---- gradient exchange options / ZeRO-1 ---------------------------------------------------------------------------------
int vitrs_model_set_comm_dtype(vitrs_model* m, int dtype) {
    if (!m) return VITRS_ERR_ARG;
    VITRS_ARG(m->ctx, (dtype == 0 || dtype == 1) && m->mode == VITRS_MODE_BF16);
    m->comm_dtype = dtype;
    return VITRS_OK;
}

int vitrs_model_enable_zero1(vitrs_model* m) {
    if (!m) return VITRS_ERR_ARG;
    VITRS_ARG(m->ctx, m->mode == VITRS_MODE_BF16);
    if (m->zero1) return VITRS_OK;
    return zero_shard_from_full(m);
}

// ZeRO-1: all-gather the fp32 master weights into the full parameter view (checkpoints, inspection).  No-op otherwise.
int vitrs_model_gather_parameters(vitrs_model* m) {
    if (!m) return VITRS_ERR_ARG;
    if (!m->zero1) return VITRS_OK;
    return zero_gather_full(m, m->zp, nullptr, m->params);  // the small tensors are current in the view already
}

// bytes of optimiser state (fp32 master weights + both AdamW moments) this rank holds
int vitrs_model_optimizer_state_bytes(vitrs_model* m, size_t* bytes) {
    if (!m || !bytes) return VITRS_ERR_ARG;
    // ZeRO-1: sharded master / m / v of the big tensors + replicated fp32 weights / m / v of the small ones
    *bytes = m->zero1 ? (m->s_total + m->small_total) * 3 * sizeof(float) : m->num_params * 3 * sizeof(float);
    return VITRS_OK;
}

// host-only description of the ZeRO-1 partition (no device needed): region offset / padded length of `bucket` in the exchange
// buffer and the length of each rank's shard, for `world` ranks
int vitrs_zero_partition(const vitrs_config* cfg_in, int world, int bucket, size_t* z_off, size_t* z_len, size_t* z_big, size_t* shard) {
    if (!cfg_in || world < 1) return VITRS_ERR_ARG;
    vitrs_config cfg = *cfg_in;
    cfg.max_seq_len = tokens(cfg);
    if (bucket < 0 || bucket > cfg.num_layers + 1) return VITRS_ERR_ARG;
    size_t sizes[P_COUNT], offs[P_COUNT], off = 0, z = 0, big = 0, small = 0;
    param_sizes_of(cfg, sizes);
    for (int i = 0; i < P_COUNT; ++i) { offs[i] = off; off += sizes[i]; }
    for (int b = 0; b <= bucket; ++b) {
        z += big + small;
        zplan_sizes(cfg, sizes, offs, world, b, &big, &small);
    }
    if (z_off) *z_off = z;
    if (z_len) *z_len = big + small;
    if (z_big) *z_big = big;
    if (shard) *shard = big / world;
    return VITRS_OK;
}

// host-only: the device bytes vitrs_model_create / ensure_zplan / enable_zero1 / ensure_stage allocate for this configuration,
// from the same sizing functions they use (param_sizes_of, arena_layout, zplan_sizes, small_runs_of)
int vitrs_model_footprint(const vitrs_config* cfg_in, int max_batch, int mode, int world, int zero1, vitrs_footprint* out) {
    if (!cfg_in || !out || max_batch < 1 || world < 1 || (mode != VITRS_MODE_F32 && mode != VITRS_MODE_BF16)) return VITRS_ERR_ARG;
    vitrs_config cfg = *cfg_in;
    if (cfg.patch_size <= 0 || cfg.image_size % cfg.patch_size || cfg.patch_size % 4 || cfg.channels <= 0 || cfg.num_heads <= 0 ||
        cfg.channels % cfg.num_heads || cfg.channels % 8 || cfg.num_layers < 1 || cfg.num_layers > 62 || cfg.num_classes < 1)
        return VITRS_ERR_ARG;  // (what vitrs_model_create accepts)
    if (zero1 && mode != VITRS_MODE_BF16) return VITRS_ERR_ARG;
    cfg.max_seq_len = tokens(cfg);
    memset(out, 0, sizeof(*out));
    size_t sizes[P_COUNT], offs[P_COUNT], n = 0;
    param_sizes_of(cfg, sizes);
    for (int i = 0; i < P_COUNT; ++i) { offs[i] = n; n += sizes[i]; }
    const size_t B = max_batch, T = cfg.max_seq_len, C = cfg.channels, L = cfg.num_layers, NH = cfg.num_heads, V = cfg.num_classes;
    const size_t kdim = 3u * cfg.patch_size * cfg.patch_size;
    const size_t esz = mode == VITRS_MODE_F32 ? 4 : 2;
    out->num_parameters = n;
    out->weights_f32 = n * sizeof(float);
    out->grads_f32 = n * sizeof(float);
    out->weights_bf16 = mode == VITRS_MODE_BF16 ? n * sizeof(bf16) : 0;
    size_t z = 0, shards = 0;
    for (int b = 0; b < cfg.num_layers + 2; ++b) {
        size_t big = 0, small = 0;
        zplan_sizes(cfg, sizes, offs, world, b, &big, &small);
        z += big + small;
        shards += big / world;
    }
    if (mode == VITRS_MODE_BF16 && (world > 1 || zero1)) out->exchange_buffer = z * sizeof(bf16);
    if (zero1) {
        SmallRun runs[5];
        small_runs_of(offs, sizes, runs);
        const size_t small_total = runs[4].soff + (runs[4].cnt + 3) / 4 * 4;
        out->zero1_master_shard = (shards + 4) * sizeof(float);
        out->adam_moments = 2 * (shards + 4) * sizeof(float) + 2 * small_total * sizeof(float);
    } else {
        out->adam_moments = 2 * n * sizeof(float);
    }
    size_t per_image[A_COUNT], aoffs[A_COUNT];
    int elems[A_COUNT];
    out->activations = arena_layout(cfg, max_batch, mode, false, per_image, elems, aoffs);
    out->activation_grads = arena_layout(cfg, max_batch, mode, true, per_image, elems, aoffs);
    if (mode == VITRS_MODE_BF16)  // dres, dln, dbig, dlogits, dlnf
        out->activation_grads += sizeof(bf16) * B * T * C * 2 + sizeof(bf16) * B * T * 4 * C + sizeof(float) * B * V + sizeof(float) * B * C;
    out->workspace = sizeof(float) * L * B * NH * T + 2 * sizeof(float) * B * C + esz * B * T * kdim + 2 * sizeof(float);
    out->staging = 2 * (B * 3 * cfg.image_size * cfg.image_size * sizeof(float) + B * sizeof(int));
    out->total = out->weights_f32 + out->grads_f32 + out->weights_bf16 + out->adam_moments + out->zero1_master_shard +
                 out->exchange_buffer + out->activations + out->activation_grads + out->workspace + out->staging;
    const double np = (double)(T - 1), t = (double)T, c = (double)C;
    const double fwd = 2.0 * np * (double)kdim * c + (double)L * (24.0 * t * c * c + 4.0 * t * t * c) + 2.0 * c * (double)V;
    out->train_flops_per_image = 3.0 * fwd;
    return VITRS_OK;
}

// ---- the step as a CUDA graph -------------------------------------------------------------------------------------------
// One ViT-B/16 step is 231 launches in 118 ms and the host is far ahead, but ViT-Ti/16 is the same 231 launches in 6.9 ms: 0.5 ms
// of that is launch gaps.  On one GPU in production mode the launch sequence is a pure function of (batch, input pointers, input
// kind) — the AdamW hyper-parameters live in device memory (op_adam_set_hyper) precisely so that nothing step-dependent is a
// launch argument — so the second time a key comes by the step is captured and from then on replayed.  Anything else (a
// communicator, ZeRO-1, verify mode, event profiling, VITRS_NO_STEP_GRAPH) takes the kernel-by-kernel path.
static int step_body(vitrs_model* m) {  // everything of the step that is capturable: no allocation, no synchronisation
    VITRS_TRY(vitrs_model_zero_grad(m));
    VITRS_TRY(forward_bf16(m));
    VITRS_TRY(backward_bf16(m));
    return op_adamw_apply(m->ctx, m->params, m->grads, m->m, m->v, m->num_params, m->shadow, m->ctx->stream);
}

static int train_step_any(vitrs_model* m, const void* images, int kind, const int* labels, int b, float lr, float beta1, float beta2,
                          float eps, float weight_decay) {
    vitrs_ctx* ctx = m->ctx;
    VITRS_ARG(ctx, images != nullptr && labels != nullptr && b >= 1 && b <= m->max_batch);
    const bool graphable = m->mode == VITRS_MODE_BF16 && !ctx->nccl_comm && !m->zero1 && !ctx->prof_on && !ctx->env_no_step_graph;
    if (!graphable) {
        VITRS_TRY(vitrs_model_zero_grad(m));
        if (kind == 0) VITRS_TRY(vitrs_model_forward(m, reinterpret_cast<const float*>(images), labels, b));
        else VITRS_TRY(vitrs_model_forward_u8(m, reinterpret_cast<const uint8_t*>(images), kind - 1, labels, b));
        VITRS_TRY(vitrs_model_backward(m));
        return vitrs_model_update(m, lr, beta1, beta2, eps, weight_decay);
    }
    VITRS_CUDA(ctx, cudaSetDevice(ctx->device));
    m->batch = b;
    m->images = kind == 0 ? reinterpret_cast<const float*>(images) : nullptr;
    m->images_u8 = kind == 0 ? nullptr : reinterpret_cast<const uint8_t*>(images);
    m->u8_layout = kind == 0 ? 0 : kind - 1;
    m->labels = labels;
    m->has_targets = 1;
    m->loss_reduced = 0;
    m->adam_step += 1;
    VITRS_TRY(op_adam_set_hyper(ctx, lr, beta1, beta2, eps, weight_decay, m->adam_step, ctx->stream));  // outside the graph
    StepGraph* slot = nullptr;
    for (StepGraph& g : m->step_graphs)
        if (g.sightings > 0 && g.images == images && g.labels == labels && g.b == b && g.kind == kind && g.epoch == m->graph_epoch) slot = &g;
    if (slot && slot->exec && slot->scratch_gen != ctx->scratch_gen) {  // recorded against a scratch buffer that has since moved
        cudaGraphExecDestroy(slot->exec);
        slot->exec = nullptr;
        slot->sightings = 1;
    }
    if (slot && slot->exec) {
        VITRS_CUDA(ctx, cudaGraphLaunch(slot->exec, ctx->stream));
        slot->age = ++m->graph_tick;
        ctx->launches += slot->launches;
        m->graph_replays++;
        return VITRS_OK;
    }
    VITRS_TRY(step_body(m));
    if (!slot) {  // first sighting: remember the key (least recently used slot)
        slot = &m->step_graphs[0];
        for (StepGraph& g : m->step_graphs)
            if (g.age < slot->age) slot = &g;
        if (slot->exec) cudaGraphExecDestroy(slot->exec);
        *slot = StepGraph{images, labels, b, kind, 1, m->graph_epoch, ctx->scratch_gen, nullptr, 0, ++m->graph_tick};
        return VITRS_OK;
    }
    // second sighting: record the sequence that has just run (capture executes nothing)
    slot->sightings++;
    slot->age = ++m->graph_tick;
    slot->scratch_gen = ctx->scratch_gen;  // (the eager run above has sized the scratch buffer for this step)
    if (!m->cap_stream) VITRS_CUDA(ctx, cudaStreamCreateWithFlags(&m->cap_stream, cudaStreamNonBlocking));
    const uint64_t before = ctx->launches;
    cudaGraph_t graph = nullptr;
    VITRS_CUDA(ctx, cudaStreamBeginCapture(m->cap_stream, cudaStreamCaptureModeThreadLocal));
    cudaStream_t run_stream = ctx->stream;
    ctx->stream = m->cap_stream;  // (the context may run on the legacy default stream, which cannot be captured)
    const int rc = step_body(m);
    ctx->stream = run_stream;
    const cudaError_t ce = cudaStreamEndCapture(m->cap_stream, &graph);
    slot->launches = ctx->launches - before;
    ctx->launches = before;
    if (rc != VITRS_OK || ce != cudaSuccess || cudaGraphInstantiate(&slot->exec, graph, 0) != cudaSuccess) {
        slot->exec = nullptr;   // stay on the kernel-by-kernel path for this key
        slot->sightings = 1 << 30;
        cudaGetLastError();
    }
    if (graph) cudaGraphDestroy(graph);
    return VITRS_OK;
}

int vitrs_model_train_step(vitrs_model* m, const float* images, const int* labels, int b, float lr, float beta1, float beta2,
                           float eps, float weight_decay) {
    if (!m) return VITRS_ERR_ARG;
    return train_step_any(m, images, 0, labels, b, lr, beta1, beta2, eps, weight_decay);
}

int vitrs_model_train_step_u8(vitrs_model* m, const uint8_t* images, int layout, const int* labels, int b, float lr, float beta1,
                              float beta2, float eps, float weight_decay) {
    if (!m) return VITRS_ERR_ARG;
    VITRS_ARG(m->ctx, layout == 0 || layout == 1);
    return train_step_any(m, images, 1 + layout, labels, b, lr, beta1, beta2, eps, weight_decay);
}

int vitrs_model_step_graph_replays(vitrs_model* m, uint64_t* replays) {
    if (!m || !replays) return VITRS_ERR_ARG;
    *replays = m->graph_replays;
    return VITRS_OK;
}

static int ensure_stage(vitrs_model* m) {
    vitrs_ctx* ctx = m->ctx;
    if (m->stage_images[0]) return VITRS_OK;
    const size_t img_elems = (size_t)m->max_batch * 3 * m->cfg.image_size * m->cfg.image_size;
    for (int i = 0; i < 2; ++i) {
        VITRS_CUDA(ctx, cudaMalloc(&m->stage_images[i], img_elems * sizeof(float)));
        VITRS_CUDA(ctx, cudaMalloc(&m->stage_labels[i], sizeof(int) * (size_t)m->max_batch));
        m->stage_src[i] = nullptr;
    }
    return VITRS_OK;
}

// H2D of a batch on the copy stream into the staging slot the compute stream is not using
int vitrs_model_prefetch_host(vitrs_model* m, const float* h_images, const int* h_labels, int b) {
    if (!m) return VITRS_ERR_ARG;
    vitrs_ctx* ctx = m->ctx;
    VITRS_ARG(ctx, h_images && h_labels && b >= 1 && b <= m->max_batch);
    VITRS_TRY(ensure_stage(m));
    const int s = m->stage_next;
    const size_t img_bytes = (size_t)b * 3 * m->cfg.image_size * m->cfg.image_size * sizeof(float);
    VITRS_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, m->stage_free[s], 0));
    VITRS_CUDA(ctx, cudaMemcpyAsync(m->stage_images[s], h_images, img_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    VITRS_CUDA(ctx, cudaMemcpyAsync(m->stage_labels[s], h_labels, sizeof(int) * (size_t)b, cudaMemcpyHostToDevice, ctx->copy_stream));
    VITRS_CUDA(ctx, cudaEventRecord(m->stage_ready[s], ctx->copy_stream));
    m->stage_src[s] = h_images;
    m->stage_batch[s] = b;
    m->stage_next = s ^ 1;
    return VITRS_OK;
}

int vitrs_model_train_step_host(vitrs_model* m, const float* h_images, const int* h_labels, int b, float lr, float beta1,
                                float beta2, float eps, float weight_decay, float* loss_out) {
    if (!m) return VITRS_ERR_ARG;
    vitrs_ctx* ctx = m->ctx;
    VITRS_ARG(ctx, h_images && h_labels && b >= 1 && b <= m->max_batch);
    VITRS_TRY(ensure_stage(m));
    int s = -1;
    for (int i = 0; i < 2; ++i)
        if (m->stage_src[i] == h_images && m->stage_batch[i] == b) s = i;
    if (s < 0) {  // not prefetched: copy now (still through the copy stream so slot reuse stays ordered)
        VITRS_TRY(vitrs_model_prefetch_host(m, h_images, h_labels, b));
        s = m->stage_next ^ 1;
    }
    VITRS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, m->stage_ready[s], 0));
    m->stage_src[s] = nullptr;
    VITRS_TRY(vitrs_model_train_step(m, m->stage_images[s], m->stage_labels[s], b, lr, beta1, beta2, eps, weight_decay));
    VITRS_CUDA(ctx, cudaEventRecord(m->stage_free[s], ctx->stream));
    if (loss_out) VITRS_TRY(vitrs_model_mean_loss(m, loss_out));
    return VITRS_OK;
}

// the same two calls for uint8 host batches (a quarter of the bytes over PCIe; the staging slots are shared with the fp32 path)
int vitrs_model_prefetch_host_u8(vitrs_model* m, const uint8_t* h_images, const int* h_labels, int b) {
    if (!m) return VITRS_ERR_ARG;
    vitrs_ctx* ctx = m->ctx;
    VITRS_ARG(ctx, h_images && h_labels && b >= 1 && b <= m->max_batch);
    VITRS_TRY(ensure_stage(m));
    const int s = m->stage_next;
    const size_t img_bytes = (size_t)b * 3 * m->cfg.image_size * m->cfg.image_size;
    VITRS_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, m->stage_free[s], 0));
    VITRS_CUDA(ctx, cudaMemcpyAsync(m->stage_images[s], h_images, img_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    VITRS_CUDA(ctx, cudaMemcpyAsync(m->stage_labels[s], h_labels, sizeof(int) * (size_t)b, cudaMemcpyHostToDevice, ctx->copy_stream));
    VITRS_CUDA(ctx, cudaEventRecord(m->stage_ready[s], ctx->copy_stream));
    m->stage_src[s] = h_images;
    m->stage_batch[s] = b;
    m->stage_next = s ^ 1;
    return VITRS_OK;
}

int vitrs_model_train_step_host_u8(vitrs_model* m, const uint8_t* h_images, int layout, const int* h_labels, int b, float lr,
                                   float beta1, float beta2, float eps, float weight_decay, float* loss_out) {
    if (!m) return VITRS_ERR_ARG;
    vitrs_ctx* ctx = m->ctx;
    VITRS_ARG(ctx, h_images && h_labels && b >= 1 && b <= m->max_batch);
    VITRS_TRY(ensure_stage(m));
    int s = -1;
    for (int i = 0; i < 2; ++i)
        if (m->stage_src[i] == h_images && m->stage_batch[i] == b) s = i;
    if (s < 0) {
        VITRS_TRY(vitrs_model_prefetch_host_u8(m, h_images, h_labels, b));
        s = m->stage_next ^ 1;
    }
    VITRS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, m->stage_ready[s], 0));
    m->stage_src[s] = nullptr;
    VITRS_TRY(vitrs_model_train_step_u8(m, reinterpret_cast<const uint8_t*>(m->stage_images[s]), layout, m->stage_labels[s], b, lr, beta1,
                                        beta2, eps, weight_decay));
    VITRS_CUDA(ctx, cudaEventRecord(m->stage_free[s], ctx->stream));
    if (loss_out) VITRS_TRY(vitrs_model_mean_loss(m, loss_out));
    return VITRS_OK;
}

// ---- checkpoint: llm.c-style file (rusty_vit.rs:81-129): 256 x i32 header, fp32 params from byte 1024.
// header[0] magic, [1] version, [2..6] max_seq_len, vocab(=classes), layers, heads, channels (the
// reference's slots), [7..10] image, patch, classes, causal, [11] adam step, [12] has m/v.
#define VITRS_CKPT_MAGIC 20261018
#define VITRS_CKPT_VERSION 1
int vitrs_model_save_checkpoint(vitrs_model* m, const char* path) {
    if (!m) return VITRS_ERR_ARG;
    vitrs_ctx* ctx = m->ctx;
    VITRS_ARG(ctx, path != nullptr);
    VITRS_CUDA(ctx, cudaSetDevice(ctx->device));
    VITRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // ZeRO-1: every rank takes part in gathering the shards; the full moments live in temporaries for the duration of the write
    float *full_m = m->m, *full_v = m->v;
    if (m->zero1) {
        VITRS_TRY(zero_gather_full(m, m->zp, nullptr, m->params));
        VITRS_CUDA(ctx, cudaMalloc(&full_m, m->num_params * sizeof(float)));
        VITRS_CUDA(ctx, cudaMalloc(&full_v, m->num_params * sizeof(float)));
        VITRS_TRY(zero_gather_full(m, m->zm, m->m_small, full_m));
        VITRS_TRY(zero_gather_full(m, m->zv, m->v_small, full_v));
    }
    bool ok = true;
    float* host = (float*)malloc(m->num_params * sizeof(float));
    FILE* f = host ? fopen(path, "wb") : nullptr;
    if (f) {
        int32_t header[256] = {0};
        header[0] = VITRS_CKPT_MAGIC; header[1] = VITRS_CKPT_VERSION;
        header[2] = m->cfg.max_seq_len; header[3] = m->cfg.num_classes; header[4] = m->cfg.num_layers;
        header[5] = m->cfg.num_heads; header[6] = m->cfg.channels; header[7] = m->cfg.image_size;
        header[8] = m->cfg.patch_size; header[9] = m->cfg.num_classes; header[10] = m->cfg.causal;
        header[11] = m->adam_step; header[12] = 1;
        ok = fwrite(header, sizeof(int32_t), 256, f) == 256;
        float* srcs[3] = {m->params, full_m, full_v};
        for (int k = 0; k < 3 && ok; ++k) {
            if (cudaMemcpy(host, srcs[k], m->num_params * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) ok = false;
            else ok = fwrite(host, sizeof(float), m->num_params, f) == m->num_params;
        }
        ok = fclose(f) == 0 && ok;
    }
    free(host);
    if (m->zero1) { cudaFree(full_m); cudaFree(full_v); }
    if (!host) return vitrs_set_error(ctx, VITRS_ERR_ARG, "out of host memory for %s", path);
    if (!f) return vitrs_set_error(ctx, VITRS_ERR_ARG, "cannot open %s for writing", path);
    return ok ? VITRS_OK : vitrs_set_error(ctx, VITRS_ERR_CUDA, "short write to %s", path);
}

// Everything is validated and read into host memory BEFORE the first byte of device state changes: a truncated or foreign file
// leaves the model exactly as it was.  Files written by the reference's layout (header slots 2..6, fp32 parameters from byte
// 1024 in tensor order, nothing else) load with zeroed moments.
int vitrs_model_load_checkpoint(vitrs_model* m, const char* path) {
    if (!m) return VITRS_ERR_ARG;
    vitrs_ctx* ctx = m->ctx;
    VITRS_ARG(ctx, path != nullptr);
    VITRS_CUDA(ctx, cudaSetDevice(ctx->device));
    FILE* f = fopen(path, "rb");
    if (!f) return vitrs_set_error(ctx, VITRS_ERR_ARG, "cannot open %s", path);
    int32_t header[256];
    if (fread(header, sizeof(int32_t), 256, f) != 256) { fclose(f); return vitrs_set_error(ctx, VITRS_ERR_ARG, "%s: short header", path); }
    if (header[0] != VITRS_CKPT_MAGIC || header[1] != VITRS_CKPT_VERSION) {
        fclose(f);
        return vitrs_set_error(ctx, VITRS_ERR_ARG, "%s: magic %d / version %d, expected %d / %d", path, header[0], header[1],
                               VITRS_CKPT_MAGIC, VITRS_CKPT_VERSION);
    }
    if (header[2] != m->cfg.max_seq_len || header[4] != m->cfg.num_layers || header[5] != m->cfg.num_heads ||
        header[6] != m->cfg.channels || header[7] != m->cfg.image_size || header[8] != m->cfg.patch_size ||
        header[9] != m->cfg.num_classes || header[10] != m->cfg.causal || header[11] < 0) {
        fclose(f);
        return vitrs_set_error(ctx, VITRS_ERR_ARG, "%s: header does not match the model configuration", path);
    }
    const int sections = header[12] ? 3 : 1;
    const long want = 1024 + (long)sections * (long)m->num_params * 4;
    fseek(f, 0, SEEK_END);
    const long have = ftell(f);
    fseek(f, 1024, SEEK_SET);
    if (have != want) {
        fclose(f);
        return vitrs_set_error(ctx, VITRS_ERR_ARG, "%s: %ld bytes, expected %ld (%d section(s) of %zu fp32)", path, have, want, sections,
                               m->num_params);
    }
    float* host = (float*)malloc((size_t)sections * m->num_params * sizeof(float));
    if (!host) { fclose(f); return vitrs_set_error(ctx, VITRS_ERR_ARG, "out of host memory for %s", path); }
    const bool ok = fread(host, sizeof(float), (size_t)sections * m->num_params, f) == (size_t)sections * m->num_params;
    fclose(f);
    if (!ok) { free(host); return vitrs_set_error(ctx, VITRS_ERR_ARG, "%s: short read", path); }
    VITRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const size_t bytes = m->num_params * sizeof(float);
    if (m->zero1) {  // the shards are re-cut from full buffers
        if (!m->m) VITRS_CUDA(ctx, cudaMalloc(&m->m, bytes));
        if (!m->v) VITRS_CUDA(ctx, cudaMalloc(&m->v, bytes));
    }
    cudaError_t e = cudaMemcpy(m->params, host, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = sections == 3 ? cudaMemcpy(m->m, host + m->num_params, bytes, cudaMemcpyHostToDevice) : cudaMemset(m->m, 0, bytes);
    if (e == cudaSuccess) e = sections == 3 ? cudaMemcpy(m->v, host + 2 * m->num_params, bytes, cudaMemcpyHostToDevice) : cudaMemset(m->v, 0, bytes);
    free(host);
    if (e != cudaSuccess) return vitrs_set_error(ctx, VITRS_ERR_CUDA, "%s: copy to the device failed: %s", path, cudaGetErrorString(e));
    m->adam_step = header[11];
    VITRS_TRY(refresh_shadow(m));
    if (m->zero1) return zero_shard_from_full(m);
    return VITRS_OK;
}

}  // extern "C"
