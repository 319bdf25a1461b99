// common.cuh — context, error plumbing and small device helpers shared by every kernel file.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/vitrs.h"

typedef __nv_bfloat16 bf16;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct AdamHyper;
struct vitrs_ctx {
    int device;
    int sm_count;
    cudaStream_t stream;      // compute stream (own or caller's)
    cudaStream_t own_stream;
    cudaStream_t copy_stream; // H2D staging
    cudaStream_t comm_stream; // NCCL
    uint64_t launches;
    char err[512];
    PFN_encodeTiled encode_tiled;
    // scratch for two-stage reductions (LayerNorm backward partials, column sums)
    float* scratch;
    size_t scratch_floats;
    // per-launch GEMM timing (vitrs_profile_begin / _end)
    uint64_t scratch_gen;  // bumped when the scratch buffer is reallocated: a recorded graph holding the old pointer is stale
    int prof_on, prof_count, prof_cap;
    cudaEvent_t* prof_ev;  // 2 per launch
    double* prof_flops;
    // NCCL (resolved with dlopen)
    void* nccl_lib;
    void* nccl_comm;
    int rank, world, nccl_max_ctas;
    // tensor maps are pure functions of (base, extents, strides, box): encoded once, then served from this table
    struct vitrs_map_entry* map_cache;
    int map_cache_used;
    // diagnostic switches, read once at context creation
    int env_gemm_cg1, env_gemm_splits, env_dp_defer, env_attn_fwd_stream, env_attn_bwd_stream, env_attn_bwd_overwrite,
        env_no_map_cache, env_attn_fwd_legacy, env_attn_fwd_nostagger, env_gemm_static, env_no_step_graph, env_gemm_no_small, env_attn_fwd_nosplit, env_gemm_patch_tc;
    // device-side error flags raised by kernels (bit 0: class label out of range), reported by vitrs_model_mean_loss
    int* dev_flags;
    AdamHyper* d_hyper;  // AdamW hyper-parameters of the current step (same allocation as dev_flags)
    unsigned int* gemm_sched;  // {next unit, drained clusters} of the GEMM's dynamic tile scheduler (same allocation)
};

// collectives on the comm stream (ctx.cu; dtype 0 = fp32, 1 = bf16); no-ops without a communicator
int vitrs_nccl_allreduce(vitrs_ctx* ctx, const void* send, void* recv, size_t count, int dtype);
int vitrs_nccl_reduce_scatter(vitrs_ctx* ctx, const void* send, void* recv, size_t recv_count, int dtype);
int vitrs_nccl_all_gather(vitrs_ctx* ctx, const void* send, void* recv, size_t send_count, int dtype);
int vitrs_nccl_group(vitrs_ctx* ctx, int begin);
// opt a kernel in to `bytes` of dynamic shared memory on this context's device (no-op when already done)
int vitrs_func_smem(vitrs_ctx* ctx, const void* fn, size_t bytes);
// cached cuTensorMapEncodeTiled: bf16 elements, 128B swizzle, rank 2 or 3; strides in BYTES (rank - 1 of them)
int vitrs_tensor_map(vitrs_ctx* ctx, CUtensorMap* out, int rank, const void* base, const uint64_t* dims, const uint64_t* strides,
                     const uint32_t* box);

int vitrs_set_error(vitrs_ctx* ctx, int code, const char* fmt, ...);
int vitrs_ensure_scratch(vitrs_ctx* ctx, size_t floats);
// event bracket for one GEMM launch while profiling is on (no-ops otherwise)
void vitrs_prof_before(vitrs_ctx* ctx, double flops);
void vitrs_prof_after(vitrs_ctx* ctx);

#define VITRS_CUDA(ctx, expr)                                                                           \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return vitrs_set_error(ctx, VITRS_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,    \
                                   cudaGetErrorString(e__));                                            \
    } while (0)

#define VITRS_TRY(expr)           \
    do {                          \
        int r__ = (expr);         \
        if (r__ != VITRS_OK) return r__; \
    } while (0)

// every kernel launch goes through this so launches are counted and errors surface at once
#define VITRS_LAUNCHED(ctx)                                                                             \
    do {                                                                                                \
        (ctx)->launches++;                                                                              \
        cudaError_t e__ = cudaGetLastError();                                                           \
        if (e__ != cudaSuccess)                                                                         \
            return vitrs_set_error(ctx, VITRS_ERR_CUDA, "%s:%d launch -> %s", __FILE__, __LINE__,       \
                                   cudaGetErrorString(e__));                                            \
    } while (0)

#define VITRS_ARG(ctx, cond)                                                                            \
    do {                                                                                                \
        if (!(cond)) return vitrs_set_error(ctx, VITRS_ERR_ARG, "%s:%d bad argument: %s", __FILE__, __LINE__, #cond); \
    } while (0)

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

// ---- device helpers -----------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float x) { return x; }
__device__ __forceinline__ float to_f32(bf16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f32(float x);
template <> __device__ __forceinline__ float from_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float x) { return __float2bfloat16_rn(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// 16-byte vector of T: 4 floats or 8 bf16
template <typename T> struct Vec16;
template <> struct Vec16<float> {
    static constexpr int N = 4;
    float4 raw;
    __device__ __forceinline__ void load(const float* p) { raw = *reinterpret_cast<const float4*>(p); }
    __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = raw; }
    __device__ __forceinline__ float get(int i) const { return reinterpret_cast<const float*>(&raw)[i]; }
    __device__ __forceinline__ void set(int i, float v) { reinterpret_cast<float*>(&raw)[i] = v; }
};
template <> struct Vec16<bf16> {
    static constexpr int N = 8;
    uint4 raw;
    __device__ __forceinline__ void load(const bf16* p) { raw = *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ void store(bf16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
    __device__ __forceinline__ float get(int i) const {
        return __bfloat162float(reinterpret_cast<const bf16*>(&raw)[i]);
    }
    __device__ __forceinline__ void set(int i, float v) { reinterpret_cast<bf16*>(&raw)[i] = __float2bfloat16_rn(v); }
};

// tanh-GELU (train_vit.rs:482-491) and its derivative (DEVIATIONS D4)
#define VITRS_GELU_K 0.044715f
#define VITRS_GELU_S 0.7978845608028654f /* sqrt(2/pi) */
template <bool FAST> __device__ __forceinline__ float tanh_sel(float u) {
    if (FAST) {
        float r;
        asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(u));
        return r;
    }
    return tanhf(u);
}
template <bool FAST> __device__ __forceinline__ float gelu_fwd(float x) {
    float u = VITRS_GELU_S * (x + VITRS_GELU_K * x * x * x);
    return 0.5f * x * (1.0f + tanh_sel<FAST>(u));
}
template <bool FAST> __device__ __forceinline__ float gelu_grad(float x) {
    float u = VITRS_GELU_S * (x + VITRS_GELU_K * x * x * x);
    float th = tanh_sel<FAST>(u);
    float sech2 = 1.0f - th * th;
    return 0.5f * (1.0f + th) + x * 0.5f * sech2 * VITRS_GELU_S * (1.0f + 3.0f * VITRS_GELU_K * x * x);
}

// Two-wide fp32 versions on sm_100's packed f32x2 FMA / MUL (one issue slot for two lanes of a thread): the GEMM
// epilogues that apply GELU / GELU' are instruction-issue-bound, and this halves their fp32 ALU instructions.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 gelu_fwd2(float2 x) {
    const float2 x2 = mul2(x, x);
    const float2 u = mul2(x, fma2(x2, splat2(VITRS_GELU_S * VITRS_GELU_K), splat2(VITRS_GELU_S)));
    const float2 th = make_float2(tanh_sel<true>(u.x), tanh_sel<true>(u.y));
    const float2 hx = mul2(x, splat2(0.5f));
    return fma2(hx, th, hx);
}
__device__ __forceinline__ float2 gelu_grad2(float2 x) {
    const float2 x2 = mul2(x, x);
    const float2 u = mul2(x, fma2(x2, splat2(VITRS_GELU_S * VITRS_GELU_K), splat2(VITRS_GELU_S)));
    const float2 th = make_float2(tanh_sel<true>(u.x), tanh_sel<true>(u.y));
    const float2 nth = make_float2(-th.x, -th.y);
    const float2 sech2 = fma2(nth, th, splat2(1.0f));
    const float2 a = fma2(th, splat2(0.5f), splat2(0.5f));
    const float2 p = fma2(x2, splat2(3.0f * VITRS_GELU_K * VITRS_GELU_S), splat2(VITRS_GELU_S));
    const float2 q = mul2(mul2(x, splat2(0.5f)), sech2);
    return fma2(q, p, a);
}

// ---- GEMM epilogues shared by the SIMT (fp32) and tcgen05 (bf16) kernels -------------------
enum EpiKind {
    EPI_NONE = 0,          // out = acc
    EPI_BIAS = 1,          // out = acc + bias[n]                         (matmul_forward)
    EPI_BIAS_GELU = 2,     // out = acc + bias[n]; out2 = gelu(out)       (fc + gelu_forward)
    EPI_BIAS_RESIDUAL = 3, // out = acc + bias[n] + aux[m,n]              (proj + residual_forward)
    EPI_GELU_BWD = 4,      // out = acc * gelu'(aux[m,n])                 (fcproj dX + gelu_backward)
    EPI_ACCUM_F32 = 5,     // out(fp32) += acc, atomically (split-K)      (dweight)
    EPI_ROWDOT = 7,        // out = acc; rowdot[(m / np * N/64 + n/64) * np + m % np] = sum over the 64-column head slice of out * aux
                           //                                             (attproj dX + D = rowsum(dO * O) of attention_backward)
    EPI_BIAS_GELU_ONLY = 8, // out = gelu(acc + bias[n]) (pre-activation rounded as if stored)   (fc + gelu_forward, inference: no backward
                           //                                             will ask for the pre-activation)
    EPI_PATCH = 6,         // tok = m % np: out = tok ? acc + bias[n] + pos[tok,n] : cls[n] + pos[0,n]   (patch embedding)
};

struct Epilogue {
    int kind;
    int accumulate;     // out += result (read-modify-write; backward ops' += contract)
    const float* bias;  // [N] fp32 or null
    const void* aux;    // residual / pre-GELU activations, same dtype and ld as out
    void* out;
    void* out2;         // second bf16 output (EPI_BIAS_GELU) or the fp32 row-dot vector (EPI_ROWDOT)
    long ldo;           // leading dimension of out / out2 / aux
    const float* pos;   // EPI_PATCH: wpe [np, N]
    const float* cls;   // EPI_PATCH: class token [N]
    int np;             // EPI_PATCH, EPI_ROWDOT: tokens per image (patches + 1)
};

// gemm operand description: element (row, k) of an operand is at base[row*rs + k*ks]
struct GemmDesc {
    const void* A;
    const void* B;
    long a_rs, a_ks, b_rs, b_ks;
    int M, N, K;
    Epilogue epi;
    float* a_colsum;  // optional, MN-major A only (a_rs == 1): a_colsum[m] += sum_k A(m, k), fused into the tensor-core GEMM
};

int gemm_simt_f32(vitrs_ctx* ctx, const GemmDesc& g);
int gemm_tc_bf16(vitrs_ctx* ctx, const GemmDesc& g);
template <typename T> inline int gemm_dispatch(vitrs_ctx* ctx, const GemmDesc& g);
template <> inline int gemm_dispatch<float>(vitrs_ctx* ctx, const GemmDesc& g) { return gemm_simt_f32(ctx, g); }
template <> inline int gemm_dispatch<bf16>(vitrs_ctx* ctx, const GemmDesc& g) { return gemm_tc_bf16(ctx, g); }
// fp32 GEMM on bf16 data is never wanted; bf16 GEMMs too small / unaligned for the tcgen05
// kernel go to gemm_simt_bf16 (same SIMT kernel instantiated on bf16 storage)
int gemm_simt_bf16(vitrs_ctx* ctx, const GemmDesc& g);

// ---- typed op launchers (elementwise.cu, attention.cu, patch_embed.cu) ----------------------
template <typename T> int op_residual_forward(vitrs_ctx*, T* out, const T* a, const T* b, long n);
template <typename T> int op_residual_backward(vitrs_ctx*, T* d1, T* d2, const T* dout, long n);
template <typename T> int op_gelu_forward(vitrs_ctx*, T* out, const T* inp, long n);
template <typename T> int op_gelu_backward(vitrs_ctx*, T* dinp, const T* inp, const T* dout, long n);
template <typename T> int op_layernorm_forward(vitrs_ctx*, T* out, float* mean, float* rstd, const T* inp,
                                               const float* w, const float* b, long rows, int c);
// dinp += LN-backward(dout); dweight/dbias += column sums; colsum_out (nullable) += column
// sums of the UPDATED dinp (the bias gradient of the GEMM that produced this residual stream)
template <typename T> int op_layernorm_backward(vitrs_ctx*, T* dinp, float* dweight, float* dbias, const T* dout,
                                                const T* inp, const float* w, const float* mean, const float* rstd,
                                                long rows, int c, float* colsum_out);
template <typename T> int op_colsum(vitrs_ctx*, float* out_accum, const T* inp, long rows, int cols, long ld);
int op_softmax_forward(vitrs_ctx*, float* probs, const float* logits, long rows, int v);
int op_crossentropy_forward(vitrs_ctx*, float* losses, const float* probs, const int* targets, long rows, int v);
int op_crossentropy_softmax_backward(vitrs_ctx*, float* dlogits, const float* dlosses, const float* probs,
                                     const int* targets, long rows, int v);
// fused head loss: probs, losses, mean loss and dlogits (+=) in one pass
int op_head_loss(vitrs_ctx*, float* probs, float* losses, float* mean_loss, float* dlogits, const float* logits,
                 const int* targets, int rows, int v, float dloss);
template <typename T> int op_cls_gather(vitrs_ctx*, float* out, const T* inp, int b, int t, int c);
template <typename T> int op_cls_scatter_add(vitrs_ctx*, T* dinp, const float* dout, int b, int t, int c);
int op_adamw(vitrs_ctx*, float* p, const float* g, float* m, float* v, size_t n, float lr, float b1, float b2,
             float eps, float wd, int step, bf16* shadow);
int op_sgd(vitrs_ctx*, float* p, const float* g, size_t n, float lr, bf16* shadow);
// AdamW split for CUDA-graph replay and ZeRO-1: hyper-parameters go to device memory once per step, the update reads them
struct AdamHyper { float lr, b1, b2, eps, wd, step_size, bc2_sqrt, pad; };
int op_adam_set_hyper(vitrs_ctx*, float lr, float b1, float b2, float eps, float wd, int step, cudaStream_t stream);
int op_adamw_apply(vitrs_ctx*, float* p, const float* g, float* m, float* v, size_t n, bf16* shadow, cudaStream_t stream);
int op_adamw_apply_shard(vitrs_ctx*, float* p, bf16* gz, float* m, float* v, size_t n, cudaStream_t stream);
// slices of one gradient bucket: flat (tensor-major) offsets, lengths, and offsets in the bucket-major exchange buffer
struct SliceTable { size_t src_off[12], cnt[12], z_off[12]; int n; };
int op_pack_f32_to_bf16(vitrs_ctx*, bf16* z, const float* flat, const SliceTable&, cudaStream_t);
int op_unpack_bf16_to_f32(vitrs_ctx*, float* flat, const bf16* z, const SliceTable&, cudaStream_t);
int op_unpack_bf16_to_bf16(vitrs_ctx*, bf16* flat, const bf16* z, const SliceTable&, cudaStream_t);
int op_pack_f32_to_f32(vitrs_ctx*, float* z, const float* flat, const SliceTable&, cudaStream_t);
int op_unpack_f32_to_f32(vitrs_ctx*, float* flat, const float* z, const SliceTable&, cudaStream_t);
int op_fill_uniform(vitrs_ctx*, float* dst, size_t n, uint64_t seed, uint64_t stream, float lo, float hi);
int op_fill_const(vitrs_ctx*, float* dst, size_t n, float v);
int op_scaled_sum(vitrs_ctx*, float* out_accum, const float* inp, long n, float scale);
int op_cast_f32_bf16(vitrs_ctx*, bf16* dst, const float* src, size_t n);
int op_cast_bf16_f32(vitrs_ctx*, float* dst, const bf16* src, size_t n);

template <typename T> int op_attention_forward(vitrs_ctx*, T* out, float* preatt, float* att, float* lse, const T* qkv,
                                               int b, int t, int c, int nh, int causal);
// dqkv += ...; probabilities are read from att when given, else recomputed from lse.
// dpreatt/datt (nullable, fp32) are the reference's materialised buffers, filled (+=) only when given.
template <typename T> int op_attention_backward(vitrs_ctx*, T* dqkv, float* dpreatt, float* datt, const T* dout,
                                                const T* qkv, const float* att, const float* lse, int b, int t, int c,
                                                int nh, int causal);
// production attention (tensor cores); same contracts
int op_attention_forward_tc(vitrs_ctx*, bf16* out, float* lse, const bf16* qkv, int b, int t, int c, int nh, int causal);
int op_attention_bwd_prep(vitrs_ctx*, float* dsum, const bf16* dout, const bf16* out, int b, int t, int c, int nh);
int op_attention_backward_tc(vitrs_ctx*, bf16* dqkv, const bf16* dout, const bf16* out, const bf16* qkv, const float* lse,
                             int b, int t, int c, int nh, int causal, int accumulate, const float* dsum_ready = nullptr);

// patch embedding pieces (patch_embed.cu): im2col rows are tokens, [B*T, 3*p*p], CLS rows zero
template <typename T> int op_im2col(vitrs_ctx*, T* patches, const float* images, int b, int img, int patch);
template <typename T>
int op_im2col_u8(vitrs_ctx*, T* patches, const uint8_t* images, int layout, const float* mean, const float* stdev, int b, int img, int patch);
// dwpe += sum_b denc; dcls += sum_b denc[b,0]; dpatchb += sum over patch tokens
template <typename T> int op_patch_backward_reduce(vitrs_ctx*, float* dwpe, float* dcls, float* dpatchb, const T* denc,
                                                   int b, int t, int c);
int op_encoder_forward(vitrs_ctx*, float* enc, const int* inputs, const float* wte, const float* wpe, int b, int t, int c);
int op_encoder_backward(vitrs_ctx*, float* dwte, float* dwpe, const float* denc, const int* inputs, int b, int t, int c);
