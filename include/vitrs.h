/*
 * vitrs.h — C ABI of libvitrs.so: the B200 (sm_100a) replacement for the ViT.rs hot path.
 *
 * This is the drop-in boundary.  The reference has no plugin system; its
 * boundary is the llm.c-style free-function set of /root/reference/train_vit.rs:376-670
 * (raw f32 pointers, c_int dims, no return value) plus the model entry points
 * ViT::build_from_checkpoint / forward / backward (rusty_vit.rs:79,269,354) and
 * optimizer_step (rusty_vit.rs:949).  Every entry point below names the reference item it
 * replaces.  Differences, all forced by the device:
 *   - pointers are DEVICE pointers; calls are asynchronous on the context's stream;
 *   - an int status is returned (0 = ok, <0 = error; vitrs_last_error() gives the text)
 *     because a CUDA launch can fail where a CPU loop cannot;
 *   - every op exists as _f32 (verify mode, fp32 multiply + fp32 accumulate) and _bf16
 *     (production: bf16 activations/weights, fp32 statistics, accumulators and gradients
 *     of parameters);
 *   - attention takes an explicit `causal` flag (reference: causal, train_vit.rs:413; every
 *     ViT config: 0) and its preatt/att buffers may be NULL (the fused kernel keeps only
 *     lse[B,NH,T]).
 * Semantics kept from the reference: forward ops overwrite their outputs; backward ops
 * ACCUMULATE (+=) into dinp/dweight/dbias (train_vit.rs:538,549,552,626-633,650), so the
 * caller zeroes gradients once per step; NULL bias / dbias are legal (train_vit.rs:388,548);
 * ops never allocate, free or retain caller pointers.
 *
 * There is no CPU fallback: every entry point fails with VITRS_ERR_CUDA when no sm_100
 * device is present.  Plain C types only — no torch / C++ types cross this boundary.
 */
#ifndef VITRS_H
#define VITRS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VITRS_OK 0
#define VITRS_ERR_CUDA -1
#define VITRS_ERR_ARG -2
#define VITRS_ERR_UNSUPPORTED -3
#define VITRS_ERR_NCCL -4

typedef struct vitrs_ctx vitrs_ctx;
typedef struct vitrs_model vitrs_model;
typedef uint16_t vitrs_bf16; /* raw bfloat16 bits */

/* ---- context ------------------------------------------------------------------------- */
int vitrs_ctx_create(vitrs_ctx** out, int device);
int vitrs_ctx_destroy(vitrs_ctx* ctx);
/* run on a caller-owned CUDA stream (cudaStream_t passed as void*); NULL is CUDA's default stream,
 * as everywhere in CUDA.  reset_stream returns to the context's own non-blocking stream. */
int vitrs_ctx_set_stream(vitrs_ctx* ctx, void* cuda_stream);
int vitrs_ctx_reset_stream(vitrs_ctx* ctx);
void* vitrs_ctx_stream(vitrs_ctx* ctx);
int vitrs_ctx_synchronize(vitrs_ctx* ctx);
const char* vitrs_last_error(vitrs_ctx* ctx);
/* device-side error flags raised by kernels since the last call, then cleared (bit 0: a class label outside [0, classes));
 * synchronises.  vitrs_model_mean_loss checks the same flags and fails with VITRS_ERR_ARG. */
int vitrs_ctx_error_flags(vitrs_ctx* ctx, int* flags);
/* kernels launched by this library on this context since creation (bench.py: gpu_launches) */
uint64_t vitrs_launch_count(vitrs_ctx* ctx);
const char* vitrs_version(void);
/* measurement aid (bench.py roofline): between begin and end every tcgen05 GEMM launch is bracketed
 * by CUDA events on the launching stream; end synchronises and returns their summed duration, the
 * summed algorithmic flops (2*M*N*K) and the launch count. */
int vitrs_profile_begin(vitrs_ctx* ctx);
int vitrs_profile_end(vitrs_ctx* ctx, double* gemm_ms, double* gemm_flops, int* gemm_launches);

/* device memory for hosts that have no allocator of their own (the Rust crate uses these) */
int vitrs_malloc(vitrs_ctx* ctx, void** ptr, size_t bytes);
int vitrs_free(vitrs_ctx* ctx, void* ptr);
int vitrs_malloc_host(vitrs_ctx* ctx, void** ptr, size_t bytes); /* pinned */
int vitrs_free_host(vitrs_ctx* ctx, void* ptr);
int vitrs_memcpy_h2d(vitrs_ctx* ctx, void* dst, const void* src, size_t bytes);
int vitrs_memcpy_d2h(vitrs_ctx* ctx, void* dst, const void* src, size_t bytes); /* synchronises */
int vitrs_memset(vitrs_ctx* ctx, void* dst, int value, size_t bytes);
/* conversions between the two modes' storage */
int vitrs_cast_f32_to_bf16(vitrs_ctx* ctx, vitrs_bf16* dst, const float* src, size_t n);
int vitrs_cast_bf16_to_f32(vitrs_ctx* ctx, float* dst, const vitrs_bf16* src, size_t n);

/* ---- L1 operators, fp32 verify mode (train_vit.rs line of the op each one replaces) ---- */
int vitrs_residual_forward_f32(vitrs_ctx*, float* out, const float* inp1, const float* inp2, int n);           /* :376 */
int vitrs_matmul_forward_f32(vitrs_ctx*, float* out, const float* inp, const float* weight,
                             const float* bias, int b, int t, int c, int oc);                                   /* :384 */
int vitrs_attention_forward_f32(vitrs_ctx*, float* out, float* preatt, float* att, const float* inp,
                                int b, int t, int c, int nh, int causal);                                       /* :400 */
int vitrs_layernorm_forward_f32(vitrs_ctx*, float* out, float* mean, float* rstd, const float* inp,
                                const float* weight, const float* bias, int b, int t, int c);                   /* :453 */
int vitrs_gelu_forward_f32(vitrs_ctx*, float* out, const float* inp, int n);                                    /* :482 */
int vitrs_softmax_forward_f32(vitrs_ctx*, float* probs, const float* logits, int b, int t, int v);              /* :493 */
int vitrs_residual_backward_f32(vitrs_ctx*, float* dinp1, float* dinp2, const float* dout, int n);              /* :521 */
int vitrs_matmul_backward_f32(vitrs_ctx*, float* dinp, float* dweight, float* dbias, const float* dout,
                              const float* inp, const float* weight, int b, int t, int c, int oc);              /* :530 */
int vitrs_attention_backward_f32(vitrs_ctx*, float* dinp, float* dpreatt, float* datt, const float* dout,
                                 const float* inp, const float* att, int b, int t, int c, int nh, int causal);  /* :559 */
int vitrs_layernorm_backward_f32(vitrs_ctx*, float* dinp, float* dweight, float* dbias, const float* dout,
                                 const float* inp, const float* weight, const float* mean, const float* rstd,
                                 int b, int t, int c);                                                          /* :603 */
int vitrs_gelu_backward_f32(vitrs_ctx*, float* dinp, const float* inp, const float* dout, int n);               /* :639 */
/* rusty_vit.rs:836 (loss = -ln p[target], DEVIATIONS D5) and the fused backward called at rusty_vit.rs:371 */
int vitrs_crossentropy_forward_f32(vitrs_ctx*, float* losses, const float* probs, const int* targets, int b, int t, int v);
int vitrs_crossentropy_softmax_backward_f32(vitrs_ctx*, float* dlogits, const float* dlosses, const float* probs,
                                            const int* targets, int b, int t, int v);
/* encoder_forward/backward as called at rusty_vit.rs:282,448 (token + position embedding) */
int vitrs_encoder_forward_f32(vitrs_ctx*, float* encoded, const int* inputs, const float* wte, const float* wpe, int b, int t, int c);
int vitrs_encoder_backward_f32(vitrs_ctx*, float* dwte, float* dwpe, const float* dencoded, const int* inputs, int b, int t, int c);
/* ViT replacement of the encoder (DEVIATIONS D7): images [B,3,H,W] fp32 NCHW */
int vitrs_patch_embed_forward_f32(vitrs_ctx*, float* encoded, const float* images, const float* patchw, const float* patchb,
                                  const float* cls, const float* wpe, int b, int img, int patch, int c);
int vitrs_patch_embed_backward_f32(vitrs_ctx*, float* dpatchw, float* dpatchb, float* dcls, float* dwpe,
                                   const float* dencoded, const float* images, int b, int img, int patch, int c);

/* ---- L1 operators, bf16 production mode ------------------------------------------------
 * activations and matmul weights are bf16; biases, LayerNorm gains/biases, statistics and
 * every parameter gradient are fp32.  attention_backward recomputes probabilities from
 * lse (written by attention_forward_bf16), so att may be NULL. */
int vitrs_residual_forward_bf16(vitrs_ctx*, vitrs_bf16* out, const vitrs_bf16* inp1, const vitrs_bf16* inp2, int n);
int vitrs_matmul_forward_bf16(vitrs_ctx*, vitrs_bf16* out, const vitrs_bf16* inp, const vitrs_bf16* weight,
                              const float* bias, int b, int t, int c, int oc);
int vitrs_attention_forward_bf16(vitrs_ctx*, vitrs_bf16* out, float* lse, const vitrs_bf16* inp,
                                 int b, int t, int c, int nh, int causal);
int vitrs_layernorm_forward_bf16(vitrs_ctx*, vitrs_bf16* out, float* mean, float* rstd, const vitrs_bf16* inp,
                                 const float* weight, const float* bias, int b, int t, int c);
int vitrs_gelu_forward_bf16(vitrs_ctx*, vitrs_bf16* out, const vitrs_bf16* inp, int n);
int vitrs_residual_backward_bf16(vitrs_ctx*, vitrs_bf16* dinp1, vitrs_bf16* dinp2, const vitrs_bf16* dout, int n);
int vitrs_matmul_backward_bf16(vitrs_ctx*, vitrs_bf16* dinp, float* dweight, float* dbias, const vitrs_bf16* dout,
                               const vitrs_bf16* inp, const vitrs_bf16* weight, int b, int t, int c, int oc);
int vitrs_attention_backward_bf16(vitrs_ctx*, vitrs_bf16* dinp, const vitrs_bf16* dout, const vitrs_bf16* out,
                                  const float* lse, const vitrs_bf16* inp, int b, int t, int c, int nh, int causal);
int vitrs_layernorm_backward_bf16(vitrs_ctx*, vitrs_bf16* dinp, float* dweight, float* dbias, const vitrs_bf16* dout,
                                  const vitrs_bf16* inp, const float* weight, const float* mean, const float* rstd,
                                  int b, int t, int c);
int vitrs_gelu_backward_bf16(vitrs_ctx*, vitrs_bf16* dinp, const vitrs_bf16* inp, const vitrs_bf16* dout, int n);

/* raw GEMM entry (tests, tuning): D[M,N] = sum_k A(m,k) * B(n,k), bf16 in, fp32 accumulate on
 * tcgen05.  a_mn_major / b_mn_major = 1 when the contraction index is the SLOW index of the
 * operand in memory (A stored [K,M] / B stored [K,N]).  out_f32_accumulate: D is fp32 and is
 * added into (split-K allowed); otherwise D is bf16 and overwritten. */
int vitrs_gemm_bf16(vitrs_ctx*, void* D, const vitrs_bf16* A, const vitrs_bf16* B, int M, int N, int K,
                    int lda, int ldb, int ldd, int a_mn_major, int b_mn_major, int out_f32_accumulate);

/* the same GEMM with one of the fused epilogues of the training step (tests, tuning; bf16 D, overwritten):
 *   1 BIAS           D = acc + bias[n]                              matmul_forward (train_vit.rs:384)
 *   2 BIAS_GELU      D = acc + bias[n]; D2 = gelu(D)                 matmul_forward + gelu_forward (:482)
 *   3 BIAS_RESIDUAL  D = acc + bias[n] + aux[m,n]                    matmul_forward + residual_forward (:376)
 *   4 GELU_BWD       D = acc * gelu'(aux[m,n])                       matmul_backward dinp + gelu_backward (:639)
 *   8 BIAS_GELU_ONLY D = gelu(acc + bias[n])                         the same pair when nothing keeps the pre-activation (inference)
 * D, D2, aux are [M, ldd] bf16; bias may be NULL (kinds 1-3); a_colsum (nullable, MN-major A only) receives
 * a_colsum[m] += sum_k A(m,k), the fused bias gradient of matmul_backward (:548-550). */
int vitrs_gemm_bf16_fused(vitrs_ctx*, vitrs_bf16* D, vitrs_bf16* D2, const vitrs_bf16* aux, const float* bias,
                          float* a_colsum, const vitrs_bf16* A, const vitrs_bf16* B, int M, int N, int K, int lda, int ldb,
                          int ldd, int a_mn_major, int b_mn_major, int epilogue);

/* ---- optimiser ---------------------------------------------------------------------------
 * optimizer_step (train_vit.rs:737): SGD over the flat buffer.  adamw_step: DEVIATIONS D8.
 * shadow (nullable): bf16 copy of the updated parameters written in the same pass. */
int vitrs_sgd_step(vitrs_ctx*, float* params, const float* grads, size_t n, float lr, vitrs_bf16* shadow);
int vitrs_adamw_step(vitrs_ctx*, float* params, const float* grads, float* m, float* v, size_t n,
                     float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                     vitrs_bf16* shadow);
/* init_parameters (train_vit.rs:674): U[lo,hi) from the counter generator shared with the oracle */
int vitrs_fill_uniform(vitrs_ctx*, float* dst, size_t n, uint64_t seed, uint64_t stream, float lo, float hi);

/* ---- L2 model (rusty_vit.rs:63-450) ------------------------------------------------------ */
typedef struct {
    int max_seq_len; /* (image/patch)^2 + 1 */
    int vocab_size;  /* reference slot (rusty_vit.rs:12); carries num_classes */
    int num_layers;
    int num_heads;
    int channels;
    int image_size;
    int patch_size;
    int num_classes;
    int causal;
} vitrs_config;

#define VITRS_MODE_F32 0  /* verify */
#define VITRS_MODE_BF16 1 /* production */
#define VITRS_NUM_PARAMETER_TENSORS 20
#define VITRS_NUM_ACTIVATION_TENSORS 23

/* ViT::build_from_checkpoint's allocation half (rusty_vit.rs:105-259): six flat device
 * allocations carved into named views.  max_batch fixes the activation arena. */
int vitrs_model_create(vitrs_ctx* ctx, const vitrs_config* cfg, int max_batch, int mode, vitrs_model** out);
int vitrs_model_destroy(vitrs_model* m);
/* init_parameters (rusty_vit.rs:864): init_mode 0 = U[0,1)*0.02 (reference), 1 = U[-1,1)*0.02 */
int vitrs_model_init_parameters(vitrs_model* m, uint64_t seed, int init_mode);
/* llm.c-style checkpoint (rusty_vit.rs:81-129): 256 x i32 header, fp32 params from byte 1024 */
int vitrs_model_save_checkpoint(vitrs_model* m, const char* path);
int vitrs_model_load_checkpoint(vitrs_model* m, const char* path);
size_t vitrs_model_num_parameters(vitrs_model* m);
/* after writing parameters through param_view: rebuild the bf16 weight shadows (no-op in f32 mode) */
int vitrs_model_sync_parameters(vitrs_model* m);
/* named views: tensor index in the reference's order (see VITRS_PARAM_NAMES in the Python /
 * Rust mirrors).  which: 0 params (fp32), 1 grads (fp32), 2 adam m, 3 adam v */
int vitrs_model_param_view(vitrs_model* m, int which, int tensor, float** ptr, size_t* count);
/* activations: which 0 = acts, 1 = grads_acts; elem_size 4 (f32 mode) or 2 (bf16 mode; the
 * head tensors lnf..losses are fp32 in both modes).  ptr is NULL for tensors the fused path
 * does not materialise (preatt, att, attproj, fcproj in bf16 mode). */
int vitrs_model_act_view(vitrs_model* m, int which, int tensor, void** ptr, size_t* count, int* elem_size);
/* 1/B_global for data parallel (generalises 1/(b*t), rusty_vit.rs:366); 0 => 1/b */
int vitrs_model_set_dloss_scale(vitrs_model* m, float scale);
/* ViT::forward (rusty_vit.rs:269): images [b,3,H,W] fp32 and labels [b] on the device;
 * labels NULL => logits only, mean_loss = -1 (rusty_vit.rs:348-350) */
int vitrs_model_forward(vitrs_model* m, const float* images, const int* labels, int b);
int vitrs_model_zero_grad(vitrs_model* m);
/* ViT::backward (rusty_vit.rs:354) */
int vitrs_model_backward(vitrs_model* m);
/* optimizer_step(model, lr) (rusty_vit.rs:949) and its AdamW form */
int vitrs_model_optimizer_step(vitrs_model* m, float lr);
int vitrs_model_update(vitrs_model* m, float lr, float beta1, float beta2, float eps, float weight_decay);
/* mean_loss field (rusty_vit.rs:75): synchronises the stream and reads it back.  A pure local read: under data parallel
 * forward() has already summed the ranks' losses once (on the comm stream), so it may be read twice, or on one rank only.
 * Fails with VITRS_ERR_ARG when a label of the batch was outside [0, classes). */
int vitrs_model_mean_loss(vitrs_model* m, float* out);
/* one whole training step from HOST buffers (pinned or pageable): H2D copy of images/labels,
 * zero_grad, forward, backward, [all-reduce], AdamW, D2H of the loss.  prefetch stages the
 * NEXT step's batch on the copy stream so it overlaps this step's compute. */
int vitrs_model_prefetch_host(vitrs_model* m, const float* h_images, const int* h_labels, int b);
int vitrs_model_train_step_host(vitrs_model* m, const float* h_images, const int* h_labels, int b,
                                float lr, float beta1, float beta2, float eps, float weight_decay,
                                float* loss_out);

/* ---- raw image batches: the data path in front of the step (the reference takes in-memory
 * buffers, rusty_vit.rs:269, and has no loader).  Images as datasets store them: uint8, layout 0 = NCHW
 * [B,3,H,W], layout 1 = NHWC [B,H,W,3] (CIFAR-10 records, decoded JPEGs).  (x / 255 - mean[c]) / std[c] is
 * applied inside the im2col pass of the patch embedding, so the host sends one byte per sample value.
 * Default normalisation: mean = std = 0.5 (-> [-1, 1]).  Everything else is as in the fp32 entry points. */
int vitrs_model_set_input_norm(vitrs_model*, const float* mean3, const float* std3);
int vitrs_model_forward_u8(vitrs_model*, const uint8_t* images, int layout, const int* labels, int b);
int vitrs_model_train_step_u8(vitrs_model*, const uint8_t* images, int layout, const int* labels, int b, float lr,
                              float beta1, float beta2, float eps, float weight_decay);
int vitrs_model_prefetch_host_u8(vitrs_model*, const uint8_t* h_images, const int* h_labels, int b);
int vitrs_model_train_step_host_u8(vitrs_model*, const uint8_t* h_images, int layout, const int* h_labels, int b,
                                   float lr, float beta1, float beta2, float eps, float weight_decay, float* loss_out);
/* same step with the batch already resident on the device.  On one GPU in production mode the launch sequence of the step is
 * captured the second time the same (batch, images, labels) comes by and replayed as a CUDA graph from then on (the AdamW
 * hyper-parameters live in device memory); vitrs_model_step_graph_replays counts the replays.  VITRS_NO_STEP_GRAPH=1 disables. */
int vitrs_model_train_step(vitrs_model* m, const float* images, const int* labels, int b,
                           float lr, float beta1, float beta2, float eps, float weight_decay);
int vitrs_model_step_graph_replays(vitrs_model* m, uint64_t* replays);

/* ---- inference engine: ViT::forward without targets (rusty_vit.rs:339-350, logits only) ---------------------
 * Borrows the parameters of a production-mode model (which must outlive it; create that model with max_batch 1 when it only
 * serves) and owns a ping-pong workspace instead of the [L, ...] activation arena of the training forward: two [M,C] residual
 * streams, [M,C] LayerNorm output, [M,3C] qkv, [M,C] attention output, [M,4C] MLP buffer — independent of the layer count
 * (ViT-B/16, batch 1024: 3.7 GB against 59.5 GB).  After one eager run per (batch, input pointer, input kind) the launch
 * sequence is replayed as a CUDA graph.  Same kernels as the training forward: logits are bit-identical to
 * vitrs_model_forward(labels = NULL).  Images: fp32 NCHW, or uint8 (layout 0 = NCHW, 1 = NHWC, normalised on the device with
 * the model's vitrs_model_set_input_norm constants). */
typedef struct vitrs_infer vitrs_infer;
int vitrs_infer_create(vitrs_model* m, int max_batch, vitrs_infer** out);
int vitrs_infer_destroy(vitrs_infer* e);
/* device images -> device logits / probabilities [b, classes] fp32 (vitrs_infer_outputs), asynchronous on the context's stream */
int vitrs_infer_forward(vitrs_infer* e, const float* images, int b);
int vitrs_infer_forward_u8(vitrs_infer* e, const uint8_t* images, int layout, int b);
int vitrs_infer_outputs(vitrs_infer* e, float** logits, float** probs);
/* host images -> host logits: H2D copy, forward, D2H copy, synchronised (the serving call; latency figures of bench.py) */
int vitrs_infer_forward_host(vitrs_infer* e, const float* h_images, int b, float* h_logits);
int vitrs_infer_forward_host_u8(vitrs_infer* e, const uint8_t* h_images, int layout, int b, float* h_logits);
/* 0 = launch kernel by kernel every time (A/B aid); default 1 */
int vitrs_infer_set_graph(vitrs_infer* e, int enabled);
int vitrs_infer_stats(vitrs_infer* e, size_t* workspace_bytes, uint64_t* graph_replays);

/* ---- record loader: the reader + loader thread in front of vitrs_model_train_step_host_u8 ------------------
 * The reference's loop takes in-memory buffers (rusty_vit.rs:269) and has no reader.  Files hold fixed-size records in the
 * CIFAR binary layout: `label_bytes` label bytes (CIFAR-10: 1; CIFAR-100: 2, the last one is used), then 3 x H x W uint8 samples,
 * channel-major.  A loader thread assembles (optionally shuffled: a function of seed and epoch only) batches into a ring of four
 * pinned host slots; ctx may be NULL for pageable slots (tools, CPU tests).  The files are read whole at open. */
typedef struct vitrs_loader vitrs_loader;
int vitrs_loader_open(vitrs_ctx* ctx, const char* const* paths, int num_paths, int image_size, int label_bytes, int batch,
                      int shuffle, uint64_t seed, int drop_last, vitrs_loader** out);
/* data parallel: every rank opens the same files with the same seed; the (seed, epoch) order is cut into batches and rank r takes
 * batches r, r + world, ... of each whole round of `world` batches (all ranks see the same number of batches per epoch) */
int vitrs_loader_open_sharded(vitrs_ctx* ctx, const char* const* paths, int num_paths, int image_size, int label_bytes, int batch,
                              int shuffle, uint64_t seed, int drop_last, int rank, int world, vitrs_loader** out);
/* the same with a transform between the stored record and the delivered image, done on the host by `workers` threads (the loader
 * thread and workers - 1 helpers share the images of a batch): random crop of the record zero-padded by crop_pad pixels on every
 * side and random horizontal flip (the standard CIFAR augmentations; a function of seed, epoch and record index only, so a batch
 * is the same bytes whatever the thread count), then a bilinear resize with half-pixel centres from image_size to out_size — 32 x 32
 * CIFAR records feeding a 224 x 224 model.  Zero / absent fields mean: no resize, no augmentation, one thread. */
typedef struct {
    int image_size;  /* side of the stored records */
    int out_size;    /* side of the delivered images (0: image_size) */
    int label_bytes;
    int batch;
    int shuffle;
    int drop_last;
    uint64_t seed;
    int rank, world; /* data parallel sharding as in vitrs_loader_open_sharded (world 0 is not valid: use 1) */
    int workers;     /* host threads transforming images (0 or 1: the loader thread alone) */
    int random_flip; /* horizontal flip with probability 1/2 */
    int crop_pad;    /* random crop after padding by this many zero pixels per side (0: none) */
} vitrs_loader_options;
int vitrs_loader_open_transform(vitrs_ctx* ctx, const char* const* paths, int num_paths, const vitrs_loader_options* options,
                                vitrs_loader** out);
/* side of the images the loader delivers */
int vitrs_loader_image_size(vitrs_loader* loader);
int vitrs_loader_close(vitrs_loader* loader);
int vitrs_loader_info(vitrs_loader* loader, size_t* num_records, int* batches_per_epoch, int* num_classes_seen);
/* blocks until the next batch is assembled: uint8 NCHW images [b,3,H,W] and int labels [b] in host memory that stays valid
 * until the call after the next one; epoch (nullable) counts passes over the files */
int vitrs_loader_next(vitrs_loader* loader, const uint8_t** h_images, const int** h_labels, int* b, uint64_t* epoch);
/* one AdamW training step on the loader's next batch; the batch after it, when already assembled, is staged on the copy stream so
 * its H2D transfer overlaps this step */
int vitrs_model_train_step_loader(vitrs_model* m, vitrs_loader* loader, float lr, float beta1, float beta2, float eps,
                                  float weight_decay, float* loss_out, int* batch_out);

/* ---- data parallel (not in the reference) ------------------------------------
 * NCCL is resolved at run time from the already-loaded libnccl.so.2 (dlopen), so the library
 * has no link-time dependency on it.  unique_id is the 128-byte ncclUniqueId made by rank 0. */
int vitrs_comm_unique_id(vitrs_ctx* ctx, void* id128);
int vitrs_comm_init(vitrs_ctx* ctx, const void* id128, int rank, int world);
/* same with an explicit cap on the thread blocks NCCL may use per collective (ncclConfig_t.maxCTAs; 0 = NCCL's default).
 * vitrs_comm_init uses 4 (env VITRS_NCCL_MAX_CTAS overrides): the persistent GEMM / attention kernels occupy every SM with
 * one large-shared-memory CTA, so each SM NCCL takes delays a CTA pair of the next GEMM. */
int vitrs_comm_init_config(vitrs_ctx* ctx, const void* id128, int rank, int world, int max_ctas);
/* ncclCommGetAsyncError: *nccl_result != 0 (and a VITRS_ERR_NCCL return) when a peer or the fabric failed */
int vitrs_comm_async_error(vitrs_ctx* ctx, int* nccl_result);
int vitrs_comm_destroy(vitrs_ctx* ctx);
int vitrs_comm_world(vitrs_ctx* ctx, int* rank, int* world);
/* sum all-reduce of the fp32 gradient buffer in reverse-layer buckets on the comm stream;
 * called by train_step when a communicator exists, exported for tests */
int vitrs_model_allreduce_grads(vitrs_model* m);
/* the bucket schedule itself, computable on the host: bucket 0 = final LayerNorm + head, 1..L = blocks
 * L-1..0 (12 slices each), L+1 = embeddings; offsets/counts (elements, capacity 12) index the flat buffer; big (nullable,
 * capacity 12): 1 for the GEMM weight matrices (read through the bf16 shadow; sharded by ZeRO-1), 0 for the tensors the
 * kernels read in fp32 (gains, biases, embeddings, class head; always replicated) */
int vitrs_grad_bucket(const vitrs_config* cfg, int bucket, size_t* offsets, size_t* counts, int* big, int* num_slices);
int vitrs_allreduce_f32(vitrs_ctx* ctx, float* buf, size_t n);
/* what the gradient exchange of the production mode puts on the wire: 1 (default) = one contiguous bf16 message per bucket
 * (the bucket's slices are packed into a bucket-major exchange buffer, summed, and unpacked into the fp32 gradient views);
 * 0 = the fp32 slices themselves, in place (exact sums; twice the bytes, 12 small messages per block) */
int vitrs_model_set_comm_dtype(vitrs_model* m, int dtype);

/* ---- ZeRO-1 sharded optimiser (the reference's optimizer_step rusty_vit.rs:949-955 and its unused moment
 * buffers :67-68 generalised): after enable, the GEMM weight matrices of each bucket (98.7 % of ViT-B/16's parameters) are
 * reduce-scattered (bf16) instead of all-reduced, every rank keeps their fp32 master weights and both AdamW moments for its
 * 1/world shard only, update() runs AdamW on the shard and all-gathers the bf16 weights into the shadow the GEMMs read.  The
 * tensors the kernels read in fp32 (gains, biases, embeddings, class head) stay replicated: all-reduced and updated everywhere.  The full fp32 parameter view goes stale until
 * vitrs_model_gather_parameters (save_checkpoint gathers by itself; every rank must call it).  grads views then hold the
 * rank-local gradients.  Works without a communicator (world 1) as the same code path. */
int vitrs_model_enable_zero1(vitrs_model* m);
int vitrs_model_gather_parameters(vitrs_model* m);
/* bytes of optimiser state (fp32 master weights + AdamW m and v) held by this rank */
int vitrs_model_optimizer_state_bytes(vitrs_model* m, size_t* bytes);
/* the partition itself, computable on the host: region [z_off, z_off + z_len) of `bucket` in the bucket-major exchange
 * buffer; its first z_big elements are the bucket's big slices padded to a multiple of 8 * world and cut into world shards
 * of `shard` elements, the rest are its small slices (padded to 8), which stay replicated */
int vitrs_zero_partition(const vitrs_config* cfg, int world, int bucket, size_t* z_off, size_t* z_len, size_t* z_big, size_t* shard);

/* ---- planning on the host (no context, no device, no CUDA call) ---------------------------------------------------------
 * The arithmetic the library itself uses to size its allocations and to route its GEMMs, exported so that a deployment can
 * be sized for the 180 GB of a B200 and the kernel routing of every model shape can be checked without a GPU. */

/* device bytes of a model created with vitrs_model_create(cfg, max_batch, mode), by what they hold (the six flat
 * allocations of rusty_vit.rs:105-259 and what the fused step adds).  world / zero1 describe the data-parallel setup the
 * model will run in (1 / 0: one GPU); exchange_buffer assumes the default bf16 gradient exchange. */
typedef struct {
    uint64_t num_parameters;
    uint64_t weights_f32;        /* the flat fp32 parameter buffer (master weights; always whole) */
    uint64_t grads_f32;          /* the flat fp32 gradient buffer */
    uint64_t weights_bf16;       /* bf16 shadow the GEMMs read (0 in verify mode) */
    uint64_t adam_moments;       /* m + v: whole, or under ZeRO-1 the rank's shards + the replicated small tensors */
    uint64_t zero1_master_shard; /* ZeRO-1: fp32 master weights of the rank's shard of the GEMM weight matrices */
    uint64_t exchange_buffer;    /* bucket-major bf16 gradient exchange buffer (world > 1 or ZeRO-1, production mode) */
    uint64_t activations;        /* the [L, ...] activation arena (ActivationTensors, rusty_vit.rs:37-61) */
    uint64_t activation_grads;   /* verify mode: the second arena; production: one block's worth + the head */
    uint64_t workspace;          /* im2col patches, lse, CLS rows, loss scalars */
    uint64_t staging;            /* the two device batch slots train_step_host copies into (allocated at first use) */
    uint64_t total;
    double train_flops_per_image; /* algorithmic: 3 x forward matmul flops */
} vitrs_footprint;
int vitrs_model_footprint(const vitrs_config* cfg, int max_batch, int mode, int world, int zero1, vitrs_footprint* out);
/* device bytes of vitrs_infer_create(model, max_batch): the ping-pong workspace and the image staging slot */
int vitrs_infer_footprint(const vitrs_config* cfg, int max_batch, uint64_t* workspace_bytes, uint64_t* staging_bytes);

/* what vitrs_gemm_bf16 / vitrs_gemm_bf16_fused (and the model's own calls) launch for a dense, 16-byte-aligned problem of
 * these extents on a device with sm_count SMs: kernel family, tile, CTAs per tile, shared-memory ring depth, split-K factor
 * and grid.  epilogue: 0 none, 1-4 / 8 as for vitrs_gemm_bf16_fused, 5 fp32 accumulate (weight gradients), 6 patch
 * embedding, 7 row dot.  flags: the context switches that change the choice. */
#define VITRS_PLAN_SIMT 0
#define VITRS_PLAN_TCGEN05 1
#define VITRS_PLAN_NO_SMALL 1   /* VITRS_GEMM_NO_SMALL: keep the large tiles on small problems */
#define VITRS_PLAN_SINGLE_CTA 2 /* VITRS_GEMM_CG=1: no CTA pairs */
#define VITRS_PLAN_PATCH_TC 4   /* VITRS_GEMM_PATCH_TC: tensor-core patch-embedding epilogue on every tile shape */
typedef struct {
    int kernel;    /* VITRS_PLAN_SIMT | VITRS_PLAN_TCGEN05 */
    int tile_m, tile_n;
    int cta_group; /* 2 = a CTA pair (tcgen05 cta_group::2) per tile */
    int stages;    /* TMA ring depth (0 for the SIMT kernel) */
    int splits;    /* split-K factor (weight gradients) */
    int k_blocks_per_split; /* 64-element K blocks per unit */
    int tiles;     /* output tiles (units = tiles * splits) */
    int grid;      /* CTAs launched (persistent: at most one per SM) */
} vitrs_gemm_plan_t;
int vitrs_gemm_plan(int M, int N, int K, int a_mn_major, int b_mn_major, int epilogue, int sm_count, int flags,
                    vitrs_gemm_plan_t* out);

#ifdef __cplusplus
}
#endif
#endif
