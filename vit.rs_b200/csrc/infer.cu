// infer.cu — the inference engine: ViT::forward without targets (rusty_vit.rs:339-350, "logits only, mean_loss = -1").
//
// The training forward (model.cu) keeps every activation of every layer for backward: 59.5 GB at ViT-B/16 batch 1024.  Nothing
// of that is needed when no backward follows, so the engine owns a small ping-pong workspace instead — two [M,C] residual
// streams, one [M,C] LayerNorm output, [M,3C] qkv, [M,C] attention output and one [M,4C] MLP buffer (which the im2col rows share,
// they are dead before the first block) — 3.7 GB at the same size, independent of the layer count.  It borrows the model's
// parameters (bf16 weight shadows for the GEMMs, fp32 gains / biases / embeddings / class head), so a training job can evaluate
// with its live weights and a serving job creates its model with max_batch 1.
//
// The launch sequence of one forward is a pure function of (batch, input pointer, input kind), so after one eager run it is
// captured into a CUDA graph and replayed: ~100 kernel launches become one graph launch, which is what bounds small-batch latency.
// The kernels are the training forward's own (tcgen05 GEMMs with fused bias / bias+GELU / bias+residual epilogues, fused
// attention, LayerNorm), so the logits are bit-identical to vitrs_model_forward(labels = NULL).
#include <stdlib.h>

#include "model.cuh"

struct InferGraph {
    int b, kind;          // kind 0 = fp32 NCHW, 1 = uint8 NCHW, 2 = uint8 NHWC
    const void* src;      // device pointer the graph reads its images from
    uint64_t epoch;       // the model's graph_epoch at capture (input normalisation constants are launch arguments)
    cudaGraphExec_t exec;
    uint64_t launches;    // kernels inside the graph
    uint64_t age;
};

struct vitrs_infer {
    vitrs_model* m;
    vitrs_ctx* ctx;
    int max_batch, use_graph;
    size_t workspace_bytes;
    char* ws;  // one allocation
    bf16 *x0, *x1, *ln, *qkv, *atty, *big;
    float *lse, *mean, *rstd, *cls_rows, *lnf, *lnf_mean, *lnf_rstd, *logits, *probs;
    void* stage;          // device staging of the host entry points: [max_batch, 3, H, W] fp32 (uint8 batches use its first quarter)
    float* h_logits;      // pinned [max_batch, V]
    cudaStream_t cap_stream;  // capture happens here: the context may run on the legacy default stream, which cannot be captured
    InferGraph graphs[8];
    uint64_t tick, graph_replays;
};

namespace {

size_t align256(size_t x) { return (x + 255) / 256 * 256; }

// one forward pass on ctx->stream; every launch below is capturable (no allocation, no synchronisation)
int forward_once(vitrs_infer* e, const void* images, int kind, int b) {
    vitrs_model* m = e->m;
    vitrs_ctx* ctx = e->ctx;
    const vitrs_config& cfg = m->cfg;
    const int T = cfg.max_seq_len, C = cfg.channels, L = cfg.num_layers, NH = cfg.num_heads, V = cfg.num_classes;
    const int kdim = 3 * cfg.patch_size * cfg.patch_size;
    const long rows = (long)b * T;
    bf16* patches = e->big;  // dead before the first block's MLP
    if (kind == 0) VITRS_TRY(op_im2col<bf16>(ctx, patches, reinterpret_cast<const float*>(images), b, cfg.image_size, cfg.patch_size));
    else VITRS_TRY(op_im2col_u8<bf16>(ctx, patches, reinterpret_cast<const uint8_t*>(images), kind - 1, m->norm_mean, m->norm_std, b,
                                      cfg.image_size, cfg.patch_size));
    {
        GemmDesc g = {};
        g.A = patches; g.a_rs = kdim; g.a_ks = 1;
        g.B = S(m, P_PATCHW); g.b_rs = kdim; g.b_ks = 1;
        g.M = (int)rows; g.N = C; g.K = kdim;
        g.epi.kind = EPI_PATCH; g.epi.bias = P(m, P_PATCHB); g.epi.cls = P(m, P_CLS); g.epi.pos = P(m, P_WPE); g.epi.np = T;
        g.epi.out = e->x0; g.epi.ldo = C;
        VITRS_TRY(gemm_tc_bf16(ctx, g));
    }
    for (int l = 0; l < L; ++l) {
        // the block of rusty_vit.rs:300-334 with the fusions of the production forward; x0 -> x1 -> x0
        VITRS_TRY(op_layernorm_forward<bf16>(ctx, e->ln, e->mean, e->rstd, e->x0, P(m, P_LN1W) + l * C, P(m, P_LN1B) + l * C, rows, C));
        VITRS_TRY((gemm_fwd<bf16>(ctx, e->qkv, e->ln, S(m, P_QKVW) + (long)l * 3 * C * C, P(m, P_QKVB) + l * 3 * C, rows, C, 3 * C, EPI_BIAS,
                                  nullptr, nullptr)));
        int r = op_attention_forward_tc(ctx, e->atty, e->lse, e->qkv, b, T, C, NH, cfg.causal);
        if (r == VITRS_ERR_UNSUPPORTED) r = op_attention_forward<bf16>(ctx, e->atty, nullptr, nullptr, e->lse, e->qkv, b, T, C, NH, cfg.causal);
        VITRS_TRY(r);
        VITRS_TRY((gemm_fwd<bf16>(ctx, e->x1, e->atty, S(m, P_ATTPROJW) + (long)l * C * C, P(m, P_ATTPROJB) + l * C, rows, C, C,
                                  EPI_BIAS_RESIDUAL, e->x0, nullptr)));
        VITRS_TRY(op_layernorm_forward<bf16>(ctx, e->ln, e->mean, e->rstd, e->x1, P(m, P_LN2W) + l * C, P(m, P_LN2B) + l * C, rows, C));
        VITRS_TRY((gemm_fwd<bf16>(ctx, e->big, e->ln, S(m, P_FCW) + (long)l * 4 * C * C, P(m, P_FCB) + l * 4 * C, rows, C, 4 * C,
                                  EPI_BIAS_GELU_ONLY, nullptr, nullptr)));
        VITRS_TRY((gemm_fwd<bf16>(ctx, e->x0, e->big, S(m, P_FCPROJW) + (long)l * C * 4 * C, P(m, P_FCPROJB) + l * C, rows, 4 * C, C,
                                  EPI_BIAS_RESIDUAL, e->x1, nullptr)));
    }
    // head (rusty_vit.rs:335-347 on the CLS rows, D7): final LayerNorm, class logits, probabilities
    VITRS_TRY(op_cls_gather<bf16>(ctx, e->cls_rows, e->x0, b, T, C));
    VITRS_TRY(op_layernorm_forward<float>(ctx, e->lnf, e->lnf_mean, e->lnf_rstd, e->cls_rows, P(m, P_LNFW), P(m, P_LNFB), b, C));
    VITRS_TRY((gemm_fwd<float>(ctx, e->logits, e->lnf, P(m, P_HEADW), P(m, P_HEADB), b, C, V, EPI_BIAS, nullptr, nullptr)));
    VITRS_TRY(op_softmax_forward(ctx, e->probs, e->logits, b, V));
    return VITRS_OK;
}

InferGraph* find_graph(vitrs_infer* e, const void* src, int kind, int b) {
    for (InferGraph& g : e->graphs)
        if (g.exec && g.b == b && g.kind == kind && g.src == src && g.epoch == e->m->graph_epoch) return &g;
    return nullptr;
}

int capture_graph(vitrs_infer* e, const void* src, int kind, int b) {
    vitrs_ctx* ctx = e->ctx;
    InferGraph* slot = &e->graphs[0];
    for (InferGraph& g : e->graphs) {
        if (!g.exec) { slot = &g; break; }
        if (g.age < slot->age) slot = &g;  // evict the least recently used
    }
    if (slot->exec) { cudaGraphExecDestroy(slot->exec); slot->exec = nullptr; }
    const uint64_t before = ctx->launches;
    cudaGraph_t graph = nullptr;
    VITRS_CUDA(ctx, cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
    cudaStream_t run_stream = ctx->stream;
    ctx->stream = e->cap_stream;  // the launchers read the context's stream; the graph itself replays on any stream
    const int rc = forward_once(e, src, kind, b);
    ctx->stream = run_stream;
    const cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &graph);
    if (rc != VITRS_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return vitrs_set_error(ctx, VITRS_ERR_CUDA, "stream capture of the inference forward failed: %s", cudaGetErrorString(ce));
    slot->launches = ctx->launches - before;
    ctx->launches = before;  // nothing ran
    const cudaError_t ie = cudaGraphInstantiate(&slot->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { slot->exec = nullptr; return vitrs_set_error(ctx, VITRS_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ie)); }
    slot->b = b; slot->kind = kind; slot->src = src; slot->epoch = e->m->graph_epoch; slot->age = ++e->tick;
    return VITRS_OK;
}

int run(vitrs_infer* e, const void* images, int kind, int b) {
    vitrs_ctx* ctx = e->ctx;
    VITRS_ARG(ctx, images != nullptr && b >= 1 && b <= e->max_batch && kind >= 0 && kind <= 2);
    VITRS_ARG(ctx, e->m->mode == VITRS_MODE_BF16);
    VITRS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (e->use_graph) {
        if (InferGraph* g = find_graph(e, images, kind, b)) {
            VITRS_CUDA(ctx, cudaGraphLaunch(g->exec, ctx->stream));
            g->age = ++e->tick;
            ctx->launches += g->launches;
            e->graph_replays++;
            return VITRS_OK;
        }
    }
    // first sight of this (batch, input): run eagerly (this also opts the kernels in to their shared memory and fills the
    // tensor-map cache), then record the same sequence for every later call
    VITRS_TRY(forward_once(e, images, kind, b));
    if (e->use_graph) VITRS_TRY(capture_graph(e, images, kind, b));
    return VITRS_OK;
}

}  // namespace

// byte offsets of the 15 workspace views (off[15] = the workspace size) for `max_batch` images; cfg.max_seq_len must be set
static void workspace_layout(const vitrs_config& cfg, int max_batch, size_t* off) {
    const size_t B = max_batch, T = cfg.max_seq_len, C = cfg.channels, NH = cfg.num_heads, V = cfg.num_classes, M = B * T;
    const size_t kdim = 3u * cfg.patch_size * cfg.patch_size;
    const size_t big_elems = M * (4 * C > kdim ? 4 * C : kdim);
    const size_t sizes[15] = {M * C * 2, M * C * 2, M * C * 2, M * 3 * C * 2, M * C * 2, big_elems * 2, B * NH * T * 4, M * 4, M * 4,
                              B * C * 4, B * C * 4, B * 4, B * 4, B * V * 4, B * V * 4};
    off[0] = 0;
    for (int i = 0; i < 15; ++i) off[i + 1] = off[i] + align256(sizes[i]);
}

extern "C" {

// host-only (vitrs.h, planning): what vitrs_infer_create(model of cfg, max_batch) allocates on the device
int vitrs_infer_footprint(const vitrs_config* cfg_in, int max_batch, uint64_t* workspace_bytes, uint64_t* staging_bytes) {
    if (!cfg_in || max_batch < 1 || cfg_in->patch_size <= 0 || cfg_in->image_size % cfg_in->patch_size) return VITRS_ERR_ARG;
    vitrs_config cfg = *cfg_in;
    cfg.max_seq_len = tokens(cfg);
    size_t off[16];
    workspace_layout(cfg, max_batch, off);
    if (workspace_bytes) *workspace_bytes = off[15];
    if (staging_bytes) *staging_bytes = (uint64_t)max_batch * 3 * cfg.image_size * cfg.image_size * sizeof(float);
    return VITRS_OK;
}

int vitrs_infer_create(vitrs_model* m, int max_batch, vitrs_infer** out) {
    if (!m) return VITRS_ERR_ARG;
    vitrs_ctx* ctx = m->ctx;
    VITRS_ARG(ctx, out != nullptr && max_batch >= 1);
    if (m->mode != VITRS_MODE_BF16) return vitrs_set_error(ctx, VITRS_ERR_UNSUPPORTED, "the inference engine runs the bf16 production kernels");
    *out = nullptr;
    VITRS_CUDA(ctx, cudaSetDevice(ctx->device));
    vitrs_infer* e = (vitrs_infer*)calloc(1, sizeof(vitrs_infer));
    e->m = m; e->ctx = ctx; e->max_batch = max_batch; e->use_graph = 1;
    const vitrs_config& cfg = m->cfg;
    const size_t B = max_batch, V = cfg.num_classes;
    size_t off[16];
    workspace_layout(cfg, max_batch, off);
    e->workspace_bytes = off[15];
    const size_t img_bytes = B * 3 * cfg.image_size * cfg.image_size * sizeof(float);
    if (cudaMalloc(&e->ws, e->workspace_bytes) != cudaSuccess || cudaMalloc(&e->stage, img_bytes) != cudaSuccess ||
        cudaMallocHost(&e->h_logits, B * V * sizeof(float)) != cudaSuccess ||
        cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
        const int rc = vitrs_set_error(ctx, VITRS_ERR_CUDA, "inference workspace (%zu + %zu bytes): %s", e->workspace_bytes, img_bytes,
                                       cudaGetErrorString(cudaGetLastError()));
        cudaFree(e->ws); cudaFree(e->stage);
        free(e);
        return rc;
    }
    char* w = e->ws;
    e->x0 = (bf16*)(w + off[0]); e->x1 = (bf16*)(w + off[1]); e->ln = (bf16*)(w + off[2]); e->qkv = (bf16*)(w + off[3]);
    e->atty = (bf16*)(w + off[4]); e->big = (bf16*)(w + off[5]); e->lse = (float*)(w + off[6]); e->mean = (float*)(w + off[7]);
    e->rstd = (float*)(w + off[8]); e->cls_rows = (float*)(w + off[9]); e->lnf = (float*)(w + off[10]); e->lnf_mean = (float*)(w + off[11]);
    e->lnf_rstd = (float*)(w + off[12]); e->logits = (float*)(w + off[13]); e->probs = (float*)(w + off[14]);
    *out = e;
    return VITRS_OK;
}

int vitrs_infer_destroy(vitrs_infer* e) {
    if (!e) return VITRS_OK;
    cudaSetDevice(e->ctx->device);
    cudaStreamSynchronize(e->ctx->stream);
    for (InferGraph& g : e->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    cudaFree(e->ws);
    cudaFree(e->stage);
    if (e->h_logits) cudaFreeHost(e->h_logits);
    if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
    free(e);
    return VITRS_OK;
}

int vitrs_infer_set_graph(vitrs_infer* e, int enabled) {
    if (!e) return VITRS_ERR_ARG;
    e->use_graph = enabled != 0;
    return VITRS_OK;
}

int vitrs_infer_forward(vitrs_infer* e, const float* images, int b) {
    if (!e) return VITRS_ERR_ARG;
    return run(e, images, 0, b);
}

int vitrs_infer_forward_u8(vitrs_infer* e, const uint8_t* images, int layout, int b) {
    if (!e) return VITRS_ERR_ARG;
    VITRS_ARG(e->ctx, layout == 0 || layout == 1);
    return run(e, images, 1 + layout, b);
}

int vitrs_infer_outputs(vitrs_infer* e, float** logits, float** probs) {
    if (!e) return VITRS_ERR_ARG;
    if (logits) *logits = e->logits;
    if (probs) *probs = e->probs;
    return VITRS_OK;
}

static int host_call(vitrs_infer* e, const void* h_images, size_t bytes, int kind, int b, float* h_logits) {
    vitrs_ctx* ctx = e->ctx;
    VITRS_ARG(ctx, h_images != nullptr && h_logits != nullptr && b >= 1 && b <= e->max_batch);
    VITRS_CUDA(ctx, cudaSetDevice(ctx->device));
    VITRS_CUDA(ctx, cudaMemcpyAsync(e->stage, h_images, bytes, cudaMemcpyHostToDevice, ctx->stream));
    VITRS_TRY(run(e, e->stage, kind, b));
    const size_t out_bytes = (size_t)b * e->m->cfg.num_classes * sizeof(float);
    VITRS_CUDA(ctx, cudaMemcpyAsync(e->h_logits, e->logits, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    VITRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(h_logits, e->h_logits, out_bytes);
    return VITRS_OK;
}

int vitrs_infer_forward_host(vitrs_infer* e, const float* h_images, int b, float* h_logits) {
    if (!e) return VITRS_ERR_ARG;
    const size_t img = (size_t)e->m->cfg.image_size;
    return host_call(e, h_images, (size_t)b * 3 * img * img * sizeof(float), 0, b, h_logits);
}

int vitrs_infer_forward_host_u8(vitrs_infer* e, const uint8_t* h_images, int layout, int b, float* h_logits) {
    if (!e) return VITRS_ERR_ARG;
    VITRS_ARG(e->ctx, layout == 0 || layout == 1);
    const size_t img = (size_t)e->m->cfg.image_size;
    return host_call(e, h_images, (size_t)b * 3 * img * img, 1 + layout, b, h_logits);
}

int vitrs_infer_stats(vitrs_infer* e, size_t* workspace_bytes, uint64_t* graph_replays) {
    if (!e) return VITRS_ERR_ARG;
    if (workspace_bytes) *workspace_bytes = e->workspace_bytes;
    if (graph_replays) *graph_replays = e->graph_replays;
    return VITRS_OK;
}

}  // extern "C"
