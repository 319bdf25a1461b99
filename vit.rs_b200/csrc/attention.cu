// attention.cu — SIMT multi-head attention forward / backward.
//
// This is verify mode's attention (fp32 multiply and accumulate, able to materialise the
// reference's preatt / att / dpreatt / datt buffers) and the route for head sizes the
// tensor-core kernel (attention_tc.cu) does not take.  Same contracts as the reference ops:
//   attention_forward   train_vit.rs:400-451 / rusty_vit.rs:512-563 / attention.rs:1-58
//   attention_backward  train_vit.rs:559-601
// with the deviations recorded in DEVIATIONS.md: rows are (b*T + t) and score buffers are
// [B,NH,T,T] (D2), every weight incl. the diagonal is normalised (D3), running max starts at
// -inf (D6), the mask is a flag (D1).  inp is packed qkv [B,T,3C]: Q at column h*hs, K at
// C + h*hs, V at 2C + h*hs; scale = 1/sqrt(hs).
//
// The backward pass never runs the reference's O(T^3) softmax-Jacobian loop (tv:583-589):
// dpreatt = att * (datt - rowsum(att * datt)) is the same quantity in O(T^2).
// It is split in two kernels so that no atomics are needed: one owns query rows (dQ, and the
// optional datt / dpreatt outputs), one owns key rows (dK, dV).  Probabilities come from the
// materialised att when given, otherwise they are recomputed from the saved log-sum-exp.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kTile = 32;       // keys (or queries) per shared-memory tile
constexpr int kMaxDimRegs = 4;  // head size <= 128

template <typename T>
__device__ __forceinline__ void load_tile(float* dst, int dst_ld, const T* __restrict__ src, long row0_off, int c3,
                                          int first, int limit, int hs) {
    // rows [first, first+kTile) of one head's Q, K or V slab -> dst[r*dst_ld + i]; rows >= limit are zero
    for (int idx = threadIdx.x; idx < kTile * hs; idx += kThreads) {
        const int r = idx / hs, i = idx - r * hs;
        const int t = first + r;
        dst[r * dst_ld + i] = t < limit ? to_f32(src[row0_off + (long)t * c3 + i]) : 0.f;
    }
}

// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
attn_fwd_kernel(T* __restrict__ out, float* __restrict__ preatt, float* __restrict__ att, float* __restrict__ lse,
                const T* __restrict__ qkv, int Tn, int C, int NH, int causal, int rpw) {
    extern __shared__ float sm[];
    const int hs = C / NH, c3 = 3 * C;
    const int rows = kWarps * rpw;
    float* qs = sm;                       // [rows][hs]
    float* kvs = qs + rows * hs;          // [kTile][hs+1]
    float* sc = kvs + kTile * (hs + 1);   // [rows][Tn]
    const int bh = blockIdx.y, b = bh / NH, h = bh - b * NH;
    const int q0 = blockIdx.x * rows;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float scale = 1.0f / sqrtf((float)hs);
    const long base = (long)b * Tn * c3 + h * hs;

    for (int idx = threadIdx.x; idx < rows * hs; idx += kThreads) {
        const int r = idx / hs, i = idx - r * hs;
        const int tq = q0 + r;
        qs[idx] = tq < Tn ? to_f32(qkv[base + (long)tq * c3 + i]) : 0.f;
    }
    const int kend_block = causal ? min(Tn, q0 + rows) : Tn;

    for (int k0 = 0; k0 < kend_block; k0 += kTile) {
        __syncthreads();
        load_tile(kvs, hs + 1, qkv, base + C, c3, k0, Tn, hs);
        __syncthreads();
        const int tk = k0 + lane;
        for (int j = 0; j < rpw; ++j) {
            const int r = warp * rpw + j;
            float dot = 0.f;
            for (int i = 0; i < hs; ++i) dot = fmaf(qs[r * hs + i], kvs[lane * (hs + 1) + i], dot);
            if (tk < Tn) sc[r * Tn + tk] = dot * scale;
        }
    }
    __syncthreads();

    for (int j = 0; j < rpw; ++j) {
        const int r = warp * rpw + j, tq = q0 + r;
        if (tq >= Tn) continue;
        const int kend = causal ? tq + 1 : Tn;
        const long srow = ((long)bh * Tn + tq) * Tn;
        float mx = -INFINITY;
        for (int tk = lane; tk < kend; tk += 32) mx = fmaxf(mx, sc[r * Tn + tk]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int tk = lane; tk < kend; tk += 32) {
            const float s = sc[r * Tn + tk];
            if (preatt) preatt[srow + tk] = s;
            const float e = expf(s - mx);
            sc[r * Tn + tk] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = sum == 0.f ? 0.f : 1.0f / sum;
        for (int tk = lane; tk < kend; tk += 32) {
            const float p = sc[r * Tn + tk] * inv;
            sc[r * Tn + tk] = p;
            if (att) att[srow + tk] = p;
        }
        for (int tk = kend + lane; tk < Tn; tk += 32) {
            sc[r * Tn + tk] = 0.f;
            if (preatt) preatt[srow + tk] = 0.f;
            if (att) att[srow + tk] = 0.f;
        }
        if (lse && lane == 0) lse[(long)bh * Tn + tq] = mx + logf(sum);
    }

    float acc[4][kMaxDimRegs];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int ii = 0; ii < kMaxDimRegs; ++ii) acc[j][ii] = 0.f;
    for (int k0 = 0; k0 < kend_block; k0 += kTile) {
        __syncthreads();
        load_tile(kvs, hs + 1, qkv, base + 2 * C, c3, k0, Tn, hs);
        __syncthreads();
        const int kmax = min(kTile, kend_block - k0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j >= rpw) break;
            const int r = warp * rpw + j;
            if (q0 + r >= Tn) continue;
            for (int kk = 0; kk < kmax; ++kk) {
                const float p = sc[r * Tn + k0 + kk];
#pragma unroll
                for (int ii = 0; ii < kMaxDimRegs; ++ii) {
                    const int i = lane + 32 * ii;
                    if (i < hs) acc[j][ii] = fmaf(p, kvs[kk * (hs + 1) + i], acc[j][ii]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j >= rpw) break;
        const int tq = q0 + warp * rpw + j;
        if (tq >= Tn) continue;
#pragma unroll
        for (int ii = 0; ii < kMaxDimRegs; ++ii) {
            const int i = lane + 32 * ii;
            if (i < hs) out[((long)b * Tn + tq) * C + h * hs + i] = from_f32<T>(acc[j][ii]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// backward, query-row owner: dQ (+=), D = rowsum(P * dP) -> dsum, optional datt / dpreatt (+=)
template <typename T, bool USE_ATT>
__global__ void __launch_bounds__(kThreads)
attn_bwd_q_kernel(T* __restrict__ dqkv, float* __restrict__ dpreatt, float* __restrict__ datt, float* __restrict__ dsum,
                  const T* __restrict__ dout, const T* __restrict__ qkv, const float* __restrict__ att,
                  const float* __restrict__ lse, int Tn, int C, int NH, int causal, int rpw) {
    extern __shared__ float sm[];
    const int hs = C / NH, c3 = 3 * C;
    const int rows = kWarps * rpw;
    float* qs = sm;                        // [rows][hs]
    float* dos = qs + rows * hs;           // [rows][hs]
    float* kvs = dos + rows * hs;          // [kTile][hs+1]
    float* P = kvs + kTile * (hs + 1);     // [rows][Tn]
    float* dP = P + rows * Tn;             // [rows][Tn]
    const int bh = blockIdx.y, b = bh / NH, h = bh - b * NH;
    const int q0 = blockIdx.x * rows;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float scale = 1.0f / sqrtf((float)hs);
    const long base = (long)b * Tn * c3 + h * hs;

    for (int idx = threadIdx.x; idx < rows * hs; idx += kThreads) {
        const int r = idx / hs, i = idx - r * hs;
        const int tq = q0 + r;
        qs[idx] = tq < Tn ? to_f32(qkv[base + (long)tq * c3 + i]) : 0.f;
        dos[idx] = tq < Tn ? to_f32(dout[((long)b * Tn + tq) * C + h * hs + i]) : 0.f;
    }
    const int kend_block = causal ? min(Tn, q0 + rows) : Tn;

    // P
    if (USE_ATT) {
        __syncthreads();
        for (int j = 0; j < rpw; ++j) {
            const int r = warp * rpw + j, tq = q0 + r;
            for (int tk = lane; tk < kend_block; tk += 32)
                P[r * Tn + tk] = tq < Tn ? att[((long)bh * Tn + tq) * Tn + tk] : 0.f;
        }
    } else {
        for (int k0 = 0; k0 < kend_block; k0 += kTile) {
            __syncthreads();
            load_tile(kvs, hs + 1, qkv, base + C, c3, k0, Tn, hs);
            __syncthreads();
            const int tk = k0 + lane;
            for (int j = 0; j < rpw; ++j) {
                const int r = warp * rpw + j, tq = q0 + r;
                float dot = 0.f;
                for (int i = 0; i < hs; ++i) dot = fmaf(qs[r * hs + i], kvs[lane * (hs + 1) + i], dot);
                if (tk < Tn) {
                    const bool live = tq < Tn && (!causal || tk <= tq);
                    P[r * Tn + tk] = live ? expf(dot * scale - lse[(long)bh * Tn + tq]) : 0.f;
                }
            }
        }
    }
    // dP = dO . V
    for (int k0 = 0; k0 < kend_block; k0 += kTile) {
        __syncthreads();
        load_tile(kvs, hs + 1, qkv, base + 2 * C, c3, k0, Tn, hs);
        __syncthreads();
        const int tk = k0 + lane;
        for (int j = 0; j < rpw; ++j) {
            const int r = warp * rpw + j, tq = q0 + r;
            float dot = 0.f;
            for (int i = 0; i < hs; ++i) dot = fmaf(dos[r * hs + i], kvs[lane * (hs + 1) + i], dot);
            if (tk < Tn) dP[r * Tn + tk] = (tq < Tn && (!causal || tk <= tq)) ? dot : 0.f;
        }
    }
    __syncthreads();
    // D, dS
    for (int j = 0; j < rpw; ++j) {
        const int r = warp * rpw + j, tq = q0 + r;
        if (tq >= Tn) continue;
        const int kend = causal ? tq + 1 : Tn;
        float d = 0.f;
        for (int tk = lane; tk < kend; tk += 32) d = fmaf(P[r * Tn + tk], dP[r * Tn + tk], d);
        d = warp_sum(d);
        if (lane == 0) dsum[(long)bh * Tn + tq] = d;
        const long srow = ((long)bh * Tn + tq) * Tn;
        for (int tk = lane; tk < kend_block; tk += 32) {
            const float p = P[r * Tn + tk], g = dP[r * Tn + tk];
            const float ds = tk < kend ? p * (g - d) : 0.f;
            if (tk < kend) {
                if (datt) datt[srow + tk] += g;
                if (dpreatt) dpreatt[srow + tk] += ds;
            }
            dP[r * Tn + tk] = ds;
        }
    }
    // dQ = scale * dS . K
    float acc[4][kMaxDimRegs];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int ii = 0; ii < kMaxDimRegs; ++ii) acc[j][ii] = 0.f;
    for (int k0 = 0; k0 < kend_block; k0 += kTile) {
        __syncthreads();
        load_tile(kvs, hs + 1, qkv, base + C, c3, k0, Tn, hs);
        __syncthreads();
        const int kmax = min(kTile, kend_block - k0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j >= rpw) break;
            const int r = warp * rpw + j;
            if (q0 + r >= Tn) continue;
            for (int kk = 0; kk < kmax; ++kk) {
                const float ds = dP[r * Tn + k0 + kk];
#pragma unroll
                for (int ii = 0; ii < kMaxDimRegs; ++ii) {
                    const int i = lane + 32 * ii;
                    if (i < hs) acc[j][ii] = fmaf(ds, kvs[kk * (hs + 1) + i], acc[j][ii]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j >= rpw) break;
        const int tq = q0 + warp * rpw + j;
        if (tq >= Tn) continue;
#pragma unroll
        for (int ii = 0; ii < kMaxDimRegs; ++ii) {
            const int i = lane + 32 * ii;
            if (i < hs) {
                T* d = dqkv + base + (long)tq * c3 + i;
                *d = from_f32<T>(to_f32(*d) + acc[j][ii] * scale);
            }
        }
    }
}

// backward, key-row owner: dK (+=), dV (+=) for kTile keys, looping over query tiles
template <typename T, bool USE_ATT>
__global__ void __launch_bounds__(kThreads)
attn_bwd_kv_kernel(T* __restrict__ dqkv, const float* __restrict__ dsum, const T* __restrict__ dout,
                   const T* __restrict__ qkv, const float* __restrict__ att, const float* __restrict__ lse, int Tn, int C,
                   int NH, int causal) {
    extern __shared__ float sm[];
    const int hs = C / NH, c3 = 3 * C;
    float* ks = sm;                          // [kTile][hs+1]
    float* vs = ks + kTile * (hs + 1);       // [kTile][hs+1]
    float* qs = vs + kTile * (hs + 1);       // [kTile][hs]
    float* dos = qs + kTile * hs;            // [kTile][hs]
    float* Pt = dos + kTile * hs;            // [kTile q][33]
    float* dSt = Pt + kTile * 33;            // [kTile q][33]
    float* lrow = dSt + kTile * 33;          // [kTile] lse of the query tile
    float* drow = lrow + kTile;              // [kTile] D of the query tile
    const int bh = blockIdx.y, b = bh / NH, h = bh - b * NH;
    const int k0 = blockIdx.x * kTile;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float scale = 1.0f / sqrtf((float)hs);
    const long base = (long)b * Tn * c3 + h * hs;
    constexpr int RPW = kTile / kWarps;  // 4

    load_tile(ks, hs + 1, qkv, base + C, c3, k0, Tn, hs);
    load_tile(vs, hs + 1, qkv, base + 2 * C, c3, k0, Tn, hs);

    float dk[RPW][kMaxDimRegs], dv[RPW][kMaxDimRegs];
#pragma unroll
    for (int j = 0; j < RPW; ++j)
#pragma unroll
        for (int ii = 0; ii < kMaxDimRegs; ++ii) dk[j][ii] = dv[j][ii] = 0.f;

    const int qstart = causal ? k0 : 0;
    for (int q0 = qstart; q0 < Tn; q0 += kTile) {
        __syncthreads();
        load_tile(qs, hs, qkv, base, c3, q0, Tn, hs);
        for (int idx = threadIdx.x; idx < kTile * hs; idx += kThreads) {
            const int r = idx / hs, i = idx - r * hs;
            const int tq = q0 + r;
            dos[idx] = tq < Tn ? to_f32(dout[((long)b * Tn + tq) * C + h * hs + i]) : 0.f;
        }
        if (threadIdx.x < kTile) {
            const int tq = q0 + threadIdx.x;
            lrow[threadIdx.x] = (!USE_ATT && tq < Tn) ? lse[(long)bh * Tn + tq] : 0.f;
            drow[threadIdx.x] = tq < Tn ? dsum[(long)bh * Tn + tq] : 0.f;
        }
        __syncthreads();
        // P and dS for (tq = q0 + warp*4 + j, tk = k0 + lane)
        const int tk = k0 + lane;
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            const int r = warp * RPW + j, tq = q0 + r;
            const bool live = tq < Tn && tk < Tn && (!causal || tk <= tq);
            float p = 0.f, g = 0.f;
            if (USE_ATT) {
                if (live) p = att[((long)bh * Tn + tq) * Tn + tk];
            } else {
                float dot = 0.f;
                for (int i = 0; i < hs; ++i) dot = fmaf(qs[r * hs + i], ks[lane * (hs + 1) + i], dot);
                if (live) p = expf(dot * scale - lrow[r]);
            }
            for (int i = 0; i < hs; ++i) g = fmaf(dos[r * hs + i], vs[lane * (hs + 1) + i], g);
            Pt[r * 33 + lane] = p;
            dSt[r * 33 + lane] = live ? p * (g - drow[r]) : 0.f;
        }
        __syncthreads();
        // this warp's keys kk = warp*4 + j accumulate over the 32 queries of the tile
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            const int kk = warp * RPW + j;
            for (int r = 0; r < kTile; ++r) {
                const float p = Pt[r * 33 + kk], ds = dSt[r * 33 + kk];
#pragma unroll
                for (int ii = 0; ii < kMaxDimRegs; ++ii) {
                    const int i = lane + 32 * ii;
                    if (i < hs) {
                        dv[j][ii] = fmaf(p, dos[r * hs + i], dv[j][ii]);
                        dk[j][ii] = fmaf(ds, qs[r * hs + i], dk[j][ii]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
        const int tk = k0 + warp * RPW + j;
        if (tk >= Tn) continue;
#pragma unroll
        for (int ii = 0; ii < kMaxDimRegs; ++ii) {
            const int i = lane + 32 * ii;
            if (i < hs) {
                T* pk = dqkv + base + C + (long)tk * c3 + i;
                T* pv = dqkv + base + 2 * C + (long)tk * c3 + i;
                *pk = from_f32<T>(to_f32(*pk) + dk[j][ii] * scale);
                *pv = from_f32<T>(to_f32(*pv) + dv[j][ii]);
            }
        }
    }
}

constexpr size_t kMaxSmem = 227 * 1024;

inline size_t fwd_smem(int rows, int hs, int Tn) { return sizeof(float) * ((size_t)rows * hs + kTile * (hs + 1) + (size_t)rows * Tn); }
inline size_t bwdq_smem(int rows, int hs, int Tn) {
    return sizeof(float) * (2 * (size_t)rows * hs + kTile * (hs + 1) + 2 * (size_t)rows * Tn);
}
inline size_t bwdkv_smem(int hs) { return sizeof(float) * (2 * kTile * (hs + 1) + 2 * kTile * hs + 2 * kTile * 33 + 2 * kTile); }

}  // namespace

template <typename T>
int op_attention_forward(vitrs_ctx* ctx, T* out, float* preatt, float* att, float* lse, const T* qkv, int b, int t, int c,
                         int nh, int causal) {
    if (b <= 0 || t <= 0) return VITRS_OK;
    VITRS_ARG(ctx, nh > 0 && c % nh == 0 && c / nh <= 32 * kMaxDimRegs);
    const int hs = c / nh;
    int rpw = 4;
    while (rpw > 1 && fwd_smem(kWarps * rpw, hs, t) > kMaxSmem) rpw >>= 1;
    const size_t smem = fwd_smem(kWarps * rpw, hs, t);
    if (smem > kMaxSmem) return vitrs_set_error(ctx, VITRS_ERR_UNSUPPORTED, "attention_forward: sequence length %d too long for the SIMT kernel", t);
    auto k = attn_fwd_kernel<T>;
    VITRS_TRY(vitrs_func_smem(ctx, (const void*)k, smem));
    dim3 grid(ceil_div(t, kWarps * rpw), b * nh);
    k<<<grid, kThreads, smem, ctx->stream>>>(out, preatt, att, lse, qkv, t, c, nh, causal, rpw);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}

template <typename T>
int op_attention_backward(vitrs_ctx* ctx, T* dqkv, float* dpreatt, float* datt, const T* dout, const T* qkv, const float* att,
                          const float* lse, int b, int t, int c, int nh, int causal) {
    if (b <= 0 || t <= 0) return VITRS_OK;
    VITRS_ARG(ctx, nh > 0 && c % nh == 0 && c / nh <= 32 * kMaxDimRegs);
    VITRS_ARG(ctx, att != nullptr || lse != nullptr);
    const int hs = c / nh;
    int rpw = 4;
    while (rpw > 1 && bwdq_smem(kWarps * rpw, hs, t) > kMaxSmem) rpw >>= 1;
    const size_t smem_q = bwdq_smem(kWarps * rpw, hs, t);
    if (smem_q > kMaxSmem) return vitrs_set_error(ctx, VITRS_ERR_UNSUPPORTED, "attention_backward: sequence length %d too long for the SIMT kernel", t);
    VITRS_TRY(vitrs_ensure_scratch(ctx, (size_t)b * nh * t));
    float* dsum = ctx->scratch;
    const size_t smem_kv = bwdkv_smem(hs);
    dim3 grid_q(ceil_div(t, kWarps * rpw), b * nh), grid_kv(ceil_div(t, kTile), b * nh);
    if (att) {
        auto kq = attn_bwd_q_kernel<T, true>;
        auto kkv = attn_bwd_kv_kernel<T, true>;
        VITRS_TRY(vitrs_func_smem(ctx, (const void*)kq, smem_q));
        VITRS_TRY(vitrs_func_smem(ctx, (const void*)kkv, smem_kv));
        kq<<<grid_q, kThreads, smem_q, ctx->stream>>>(dqkv, dpreatt, datt, dsum, dout, qkv, att, lse, t, c, nh, causal, rpw);
        VITRS_LAUNCHED(ctx);
        kkv<<<grid_kv, kThreads, smem_kv, ctx->stream>>>(dqkv, dsum, dout, qkv, att, lse, t, c, nh, causal);
        VITRS_LAUNCHED(ctx);
    } else {
        auto kq = attn_bwd_q_kernel<T, false>;
        auto kkv = attn_bwd_kv_kernel<T, false>;
        VITRS_TRY(vitrs_func_smem(ctx, (const void*)kq, smem_q));
        VITRS_TRY(vitrs_func_smem(ctx, (const void*)kkv, smem_kv));
        kq<<<grid_q, kThreads, smem_q, ctx->stream>>>(dqkv, dpreatt, datt, dsum, dout, qkv, att, lse, t, c, nh, causal, rpw);
        VITRS_LAUNCHED(ctx);
        kkv<<<grid_kv, kThreads, smem_kv, ctx->stream>>>(dqkv, dsum, dout, qkv, att, lse, t, c, nh, causal);
        VITRS_LAUNCHED(ctx);
    }
    return VITRS_OK;
}

template int op_attention_forward<float>(vitrs_ctx*, float*, float*, float*, float*, const float*, int, int, int, int, int);
template int op_attention_forward<bf16>(vitrs_ctx*, bf16*, float*, float*, float*, const bf16*, int, int, int, int, int);
template int op_attention_backward<float>(vitrs_ctx*, float*, float*, float*, const float*, const float*, const float*,
                                          const float*, int, int, int, int, int);
template int op_attention_backward<bf16>(vitrs_ctx*, bf16*, float*, float*, const bf16*, const bf16*, const float*, const float*,
                                         int, int, int, int, int);
