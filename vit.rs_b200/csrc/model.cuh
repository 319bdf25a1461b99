// model.cuh — the model object's storage (shared by model.cu: training step, and infer.cu: the inference engine).
// Mirrors `struct ViT` of rusty_vit.rs:63-76: parameter / activation tables in the reference's order plus the ViT pieces (D7).
#pragma once
#include "common.cuh"

enum ParamId {
    P_PATCHW, P_PATCHB, P_CLS, P_WPE, P_LN1W, P_LN1B, P_QKVW, P_QKVB, P_ATTPROJW, P_ATTPROJB,
    P_LN2W, P_LN2B, P_FCW, P_FCB, P_FCPROJW, P_FCPROJB, P_LNFW, P_LNFB, P_HEADW, P_HEADB, P_COUNT
};
enum ActId {
    A_ENCODED, A_LN1, A_LN1_MEAN, A_LN1_RSTD, A_QKV, A_ATTY, A_PREATT, A_ATT, A_ATTPROJ, A_RESIDUAL2,
    A_LN2, A_LN2_MEAN, A_LN2_RSTD, A_FCH, A_FCH_GELU, A_FCPROJ, A_RESIDUAL3, A_LNF, A_LNF_MEAN, A_LNF_RSTD,
    A_LOGITS, A_PROBS, A_LOSSES, A_COUNT
};
static_assert(P_COUNT == VITRS_NUM_PARAMETER_TENSORS, "parameter table");
static_assert(A_COUNT == VITRS_NUM_ACTIVATION_TENSORS, "activation table");

struct Arena {
    char* base;
    size_t bytes;
    char* view[A_COUNT];
    size_t per_image[A_COUNT];  // elements per image (all L layers)
    int elem[A_COUNT];          // bytes per element, 0 = not materialised
};

// one captured training step (zero_grad + forward + backward + AdamW apply) for a given (batch, input pointers, input kind)
struct StepGraph {
    const void* images;
    const int* labels;
    int b, kind;            // kind 0 = fp32 NCHW, 1 = uint8 NCHW, 2 = uint8 NHWC
    int sightings;          // the step is captured the second time the same key comes by
    uint64_t epoch;         // vitrs_model::graph_epoch at capture: launch arguments baked into the graph are still current
    uint64_t scratch_gen;   // vitrs_ctx::scratch_gen at capture: the context's scratch buffer (D of the attention backward) has not moved
    cudaGraphExec_t exec;
    uint64_t launches, age;
};

struct vitrs_model {
    vitrs_ctx* ctx;
    vitrs_config cfg;
    int mode, max_batch;
    size_t param_sizes[P_COUNT], param_off[P_COUNT], num_params;
    float *params, *grads, *m, *v;
    bf16* shadow;
    Arena acts, gacts;
    // extras outside the reference's 23 views
    float* lse;          // [L, B, NH, T]
    float* cls_rows;     // [B, C] gathered CLS rows (input of the final LayerNorm)
    float* dcls_rows;    // [B, C]
    void* patches;       // [B*T, 3*p*p] im2col rows
    // bf16 mode: one layer of gradient scratch + fp32 head gradients
    bf16 *dres, *dln, *dbig;
    float *dlogits, *dlnf;
    float* d_mean_loss;
    float* h_mean_loss;  // pinned
    // input staging (double-buffered) for the host-buffer step
    float* stage_images[2];
    int* stage_labels[2];
    cudaEvent_t stage_ready[2], stage_free[2];
    const void* stage_src[2];
    int stage_batch[2];
    int stage_next;
    cudaEvent_t ev_bucket, ev_comm_done;
    int batch, has_targets;
    const float* images;  // borrowed device pointer of the current batch
    const uint8_t* images_u8;  // ... or raw uint8 images (layout u8_layout), normalised inside im2col
    int u8_layout;
    float norm_mean[3], norm_std[3];
    const int* labels;
    float dloss_scale;
    int adam_step;
    // gradient exchange (production mode): bucket-major bf16 buffer ("Z order"), built on first use for the context's world size
    int comm_dtype;        // 0: fp32 slices all-reduced in place (exact sums), 1: packed bf16 buckets (default)
    int loss_reduced;      // d_mean_loss[1] holds (or will hold, on the comm stream) the sum over ranks of the local losses
    int z_world;           // world size the plan below was built for (0 = not built)
    int z_buckets;
    size_t *z_off, *z_len, *z_big, *z_shard, *s_off;  // per bucket: region start / length in the exchange buffer, length of its big (sharded)
                                                      // part, shard length, shard start in zp / zm / zv
    size_t z_size, s_total;
    bf16* comm_buf;        // [z_size]
    // ZeRO-1: fp32 master weights and AdamW moments of this rank's shard of every bucket, in Z order
    unsigned long long unpack_pending;  // buckets whose summed gradients still sit in the exchange buffer (bit per bucket)
    int zero1;
    float *zp, *zm, *zv;   // [s_total]
    float *m_small, *v_small;  // replicated moments of the small tensors, compact (small_runs)
    size_t small_total;
    // CUDA-graph replay of the single-GPU production step
    StepGraph step_graphs[4];
    cudaStream_t cap_stream;
    uint64_t graph_tick, graph_replays;
    uint64_t graph_epoch;   // bumped whenever a value that kernels receive BY VALUE changes (input normalisation, loss scale):
                            // graphs recorded before that (here and in the inference engines) are stale
};

static inline int tokens(const vitrs_config& c) { return (c.image_size / c.patch_size) * (c.image_size / c.patch_size) + 1; }
static inline float* P(const vitrs_model* m, int i) { return m->params + m->param_off[i]; }
static inline float* G(const vitrs_model* m, int i) { return m->grads + m->param_off[i]; }
static inline const bf16* S(const vitrs_model* m, int i) { return m->shadow + m->param_off[i]; }

// matmul_forward call shape (train_vit.rs:384): out[rows, oc] = inp[rows, c] . w[oc, c]^T (+ epilogue `kind`)
template <typename T>
int gemm_fwd(vitrs_ctx* ctx, T* out, const T* inp, const T* w, const float* bias, long rows, int c, int oc, int kind,
             const T* aux, T* out2) {
    GemmDesc g = {};
    g.A = inp; g.a_rs = c; g.a_ks = 1;
    g.B = w; g.b_rs = c; g.b_ks = 1;
    g.M = (int)rows; g.N = oc; g.K = c;
    g.epi.kind = kind; g.epi.bias = bias; g.epi.aux = aux; g.epi.out = out; g.epi.out2 = out2; g.epi.ldo = oc;
    return gemm_dispatch<T>(ctx, g);
}
