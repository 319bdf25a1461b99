// elementwise.cu — the bandwidth-bound operators: residual, GELU, LayerNorm forward/backward,
// column sums, softmax / cross-entropy, AdamW / SGD, init and casts.
// All are vectorised (16-byte accesses), coalesced, warp-shuffle kernels; statistics and
// accumulators are fp32 in both modes.  Reference lines: train_vit.rs (tv), rusty_vit.rs (rv).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

inline int grid_for(long work_items, int per_block, int sm_count, int waves = 8) {
    long g = (work_items + per_block - 1) / per_block;
    long cap = (long)sm_count * waves;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ---- residual (tv:376-382, tv:521-528) -------------------------------------------------------
template <typename T>
__global__ void residual_fwd_kernel(T* __restrict__ out, const T* __restrict__ a, const T* __restrict__ b, long n) {
    constexpr int VN = Vec16<T>::N;
    const long nv = n / VN;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        Vec16<T> x, y, o;
        x.load(a + i * VN);
        y.load(b + i * VN);
#pragma unroll
        for (int j = 0; j < VN; ++j) o.set(j, x.get(j) + y.get(j));
        o.store(out + i * VN);
    }
    for (long i = nv * VN + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = from_f32<T>(to_f32(a[i]) + to_f32(b[i]));
}

template <typename T>
__global__ void residual_bwd_kernel(T* __restrict__ d1, T* __restrict__ d2, const T* __restrict__ dout, long n) {
    constexpr int VN = Vec16<T>::N;
    const long nv = n / VN;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        Vec16<T> g, x, y;
        g.load(dout + i * VN);
        x.load(d1 + i * VN);
        y.load(d2 + i * VN);
#pragma unroll
        for (int j = 0; j < VN; ++j) {
            x.set(j, x.get(j) + g.get(j));
            y.set(j, y.get(j) + g.get(j));
        }
        x.store(d1 + i * VN);
        y.store(d2 + i * VN);
    }
    for (long i = nv * VN + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float g = to_f32(dout[i]);
        d1[i] = from_f32<T>(to_f32(d1[i]) + g);
        d2[i] = from_f32<T>(to_f32(d2[i]) + g);
    }
}

// ---- GELU (tv:482-491, tv:639-653 with D4) ---------------------------------------------------
template <typename T, bool FAST>
__global__ void gelu_fwd_kernel(T* __restrict__ out, const T* __restrict__ inp, long n) {
    constexpr int VN = Vec16<T>::N;
    const long nv = n / VN;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        Vec16<T> x, o;
        x.load(inp + i * VN);
#pragma unroll
        for (int j = 0; j < VN; ++j) o.set(j, gelu_fwd<FAST>(x.get(j)));
        o.store(out + i * VN);
    }
    for (long i = nv * VN + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = from_f32<T>(gelu_fwd<FAST>(to_f32(inp[i])));
}

template <typename T, bool FAST>
__global__ void gelu_bwd_kernel(T* __restrict__ dinp, const T* __restrict__ inp, const T* __restrict__ dout, long n) {
    constexpr int VN = Vec16<T>::N;
    const long nv = n / VN;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        Vec16<T> x, g, d;
        x.load(inp + i * VN);
        g.load(dout + i * VN);
        d.load(dinp + i * VN);
#pragma unroll
        for (int j = 0; j < VN; ++j) d.set(j, d.get(j) + gelu_grad<FAST>(x.get(j)) * g.get(j));
        d.store(dinp + i * VN);
    }
    for (long i = nv * VN + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dinp[i] = from_f32<T>(to_f32(dinp[i]) + gelu_grad<FAST>(to_f32(inp[i])) * to_f32(dout[i]));
}

// ---- LayerNorm forward (tv:453-480): warps stride over rows, the row cached in registers -------
// Gains and biases sit in shared memory (16-byte reads, conflict-free) so that the kernel fits three blocks per SM, and
// each warp requests its next row before it reduces the current one: per SM ~24 warps x 2 rows x 48 B per lane are in
// flight, which is what 6.5 TB/s x HBM latency asks for (one row per warp at two blocks per SM reached 68 % of it).
template <typename T, int MAXNV>
__global__ void __launch_bounds__(kThreads, MAXNV <= 4 ? 3 : 1)
ln_fwd_kernel(T* __restrict__ out, float* __restrict__ mean, float* __restrict__ rstd, const T* __restrict__ inp,
              const float* __restrict__ w, const float* __restrict__ bias, long rows, int c) {
    constexpr int VN = Vec16<T>::N;
    extern __shared__ float ln_wb[];  // [c] gains, [c] biases
    float* wsm = ln_wb;
    float* bsm = ln_wb + c;
    for (int i = threadIdx.x; i < c; i += kThreads) {
        wsm[i] = w[i];
        bsm[i] = bias[i];
    }
    if (threadIdx.x == 0) bsm[c] = 1.0f;
    __syncthreads();
    const float2 one2 = splat2(bsm[c]);
    const int lane = threadIdx.x & 31;
    const long nwarps = (long)gridDim.x * (kThreads / 32);
    long row = (long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    Vec16<T> v[MAXNV], nx[MAXNV];
    if (row < rows) {
#pragma unroll
        for (int i = 0; i < MAXNV; ++i) {
            const int idx = (i * 32 + lane) * VN;
            if (idx < c) v[i].load(inp + row * c + idx);
        }
    }
    for (; row < rows; row += nwarps) {
        const long next = row + nwarps;
        if (next < rows) {
#pragma unroll
            for (int i = 0; i < MAXNV; ++i) {
                const int idx = (i * 32 + lane) * VN;
                if (idx < c) nx[i].load(inp + next * c + idx);
            }
        }
        // packed f32x2 arithmetic: at the SM clock of a power-capped training step this kernel is as much instruction- as
        // bandwidth-bound (x + y is x * one + y with a multiplier ptxas cannot fold, see ln_bwd_kernel)
        constexpr int NP = MAXNV * VN / 2;
        float2 xf[NP];  // the row in fp32, converted once
        float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < MAXNV; ++i) {
            const bool live = (i * 32 + lane) * VN < c;
#pragma unroll
            for (int j = 0; j < VN; j += 2) {
                xf[(i * VN + j) >> 1] = live ? make_float2(v[i].get(j), v[i].get(j + 1)) : make_float2(0.f, 0.f);
                s2 = fma2(xf[(i * VN + j) >> 1], one2, s2);
            }
        }
        const float m = warp_sum(s2.x + s2.y) / (float)c;
        const float2 nm2 = splat2(-m);
        float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < MAXNV; ++i) {
            const bool live = (i * 32 + lane) * VN < c;
#pragma unroll
            for (int j = 0; j < VN; j += 2) {
                const float2 d = fma2(xf[(i * VN + j) >> 1], one2, nm2);
                xf[(i * VN + j) >> 1] = d;
                if (live) q2 = fma2(d, d, q2);
            }
        }
        const float var = warp_sum(q2.x + q2.y) / (float)c;
        const float rs = 1.0f / sqrtf(var + 1e-5f);
        const float2 rs2 = splat2(rs);
        T* y = out + row * c;
#pragma unroll
        for (int i = 0; i < MAXNV; ++i) {
            const int idx = (i * 32 + lane) * VN;
            if (idx < c) {
                Vec16<T> o;
#pragma unroll
                for (int j4 = 0; j4 < VN / 4; ++j4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wsm + idx + j4 * 4);
                    const float4 b4 = *reinterpret_cast<const float4*>(bsm + idx + j4 * 4);
                    const float2 o0 = fma2(mul2(rs2, xf[(i * VN + j4 * 4) >> 1]), make_float2(w4.x, w4.y), make_float2(b4.x, b4.y));
                    const float2 o1 = fma2(mul2(rs2, xf[(i * VN + j4 * 4 + 2) >> 1]), make_float2(w4.z, w4.w), make_float2(b4.z, b4.w));
                    o.set(j4 * 4 + 0, o0.x);
                    o.set(j4 * 4 + 1, o0.y);
                    o.set(j4 * 4 + 2, o1.x);
                    o.set(j4 * 4 + 3, o1.y);
                }
                o.store(y + idx);
            }
        }
        if (lane == 0) {
            mean[row] = m;
            rstd[row] = rs;
        }
#pragma unroll
        for (int i = 0; i < MAXNV; ++i) v[i] = nx[i];
    }
}

// any c (unaligned / very wide rows): scalar accesses, three passes through L1
template <typename T>
__global__ void ln_fwd_generic_kernel(T* __restrict__ out, float* __restrict__ mean, float* __restrict__ rstd,
                                      const T* __restrict__ inp, const float* __restrict__ w,
                                      const float* __restrict__ bias, long rows, int c) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const T* x = inp + row * c;
    float s = 0.f;
    for (int i = lane; i < c; i += 32) s += to_f32(x[i]);
    const float m = warp_sum(s) / (float)c;
    float q = 0.f;
    for (int i = lane; i < c; i += 32) {
        float d = to_f32(x[i]) - m;
        q += d * d;
    }
    const float rs = 1.0f / sqrtf(warp_sum(q) / (float)c + 1e-5f);
    for (int i = lane; i < c; i += 32) out[row * c + i] = from_f32<T>((rs * (to_f32(x[i]) - m)) * w[i] + bias[i]);
    if (lane == 0) {
        mean[row] = m;
        rstd[row] = rs;
    }
}

// ---- LayerNorm backward (tv:603-637) ---------------------------------------------------------
// Warps stride over rows; each lane keeps fp32 partial column sums for the columns it owns
// (dweight, dbias and, optionally, the column sum of the updated dinp), reduced across the
// block's warps in shared memory and added to global with one atomic per column per block.
// Gains live in shared memory (broadcast-free 16-byte reads) to keep the kernel at two blocks per SM;
// all three row operands (dout, inp, dinp) are requested before the first reduction.
template <typename T, int MAXNV, bool COLSUM>
__global__ void __launch_bounds__(kThreads, MAXNV <= 4 ? 2 : 1)
ln_bwd_kernel(T* __restrict__ dinp, float* __restrict__ dweight, float* __restrict__ dbias, const T* __restrict__ dout,
              const T* __restrict__ inp, const float* __restrict__ w, const float* __restrict__ mean,
              const float* __restrict__ rstd, long rows, int c, float* __restrict__ colsum_out) {
    constexpr int VN = Vec16<T>::N;
    constexpr int NW = kThreads / 32;
    extern __shared__ float red[];  // [NW][c] reduction scratch, then [c] gains
    float* wsm = red + NW * c;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < c; i += kThreads) wsm[i] = w[i];
    if (threadIdx.x == 0) wsm[c] = 1.0f;
    __syncthreads();
    // x + y as one FFMA2 (x * one + y): ptxas splits add.f32x2, and a literal 1.0 multiplier, into two FADDs; a value it cannot
    // fold keeps the packed form
    const float2 one2 = splat2(wsm[c]);
    // packed f32x2 arithmetic throughout (the kernel is co-limited by instruction issue: ~27 -> ~17 instructions per element)
    constexpr int NP = MAXNV * VN / 2;
    float2 dw_acc[NP], db_acc[NP], cs_acc[COLSUM ? NP : 1];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        dw_acc[i] = make_float2(0.f, 0.f);
        db_acc[i] = make_float2(0.f, 0.f);
        if (COLSUM) cs_acc[i] = make_float2(0.f, 0.f);
    }
    for (long row = (long)blockIdx.x * NW + warp; row < rows; row += (long)gridDim.x * NW) {
        Vec16<T> gy[MAXNV], xv[MAXNV], dv[MAXNV];
#pragma unroll
        for (int i = 0; i < MAXNV; ++i) {
            const int idx = (i * 32 + lane) * VN;
            if (idx < c) {
                gy[i].load(dout + row * c + idx);
                xv[i].load(inp + row * c + idx);
                dv[i].load(dinp + row * c + idx);
            }
        }
        const float m = mean[row], rs = rstd[row];
        const float2 rs2 = splat2(rs), mrs2 = splat2(-m * rs);  // nrm = x * rs - m * rs
        float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < MAXNV; ++i) {
            const int idx = (i * 32 + lane) * VN;
            if (idx < c) {
#pragma unroll
                for (int j4 = 0; j4 < VN / 4; ++j4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wsm + idx + j4 * 4);
                    const float2 wv[2] = {make_float2(w4.x, w4.y), make_float2(w4.z, w4.w)};
#pragma unroll
                    for (int pp = 0; pp < 2; ++pp) {
                        const int j = j4 * 4 + pp * 2;
                        const float2 g = make_float2(gy[i].get(j), gy[i].get(j + 1));
                        const float2 nrm = fma2(make_float2(xv[i].get(j), xv[i].get(j + 1)), rs2, mrs2);
                        const float2 dn = mul2(wv[pp], g);
                        s1 = fma2(wv[pp], g, s1);
                        s2 = fma2(dn, nrm, s2);
                        db_acc[(i * VN + j) >> 1] = fma2(g, one2, db_acc[(i * VN + j) >> 1]);
                        dw_acc[(i * VN + j) >> 1] = fma2(nrm, g, dw_acc[(i * VN + j) >> 1]);
                    }
                }
            }
        }
        const float dn_mean = warp_sum(s1.x + s1.y) / (float)c;
        const float dnn_mean = warp_sum(s2.x + s2.y) / (float)c;
        const float2 ndn2 = splat2(-dn_mean), ndnn2 = splat2(-dnn_mean);
#pragma unroll
        for (int i = 0; i < MAXNV; ++i) {
            const int idx = (i * 32 + lane) * VN;
            if (idx < c) {
#pragma unroll
                for (int j4 = 0; j4 < VN / 4; ++j4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wsm + idx + j4 * 4);
                    const float2 wv[2] = {make_float2(w4.x, w4.y), make_float2(w4.z, w4.w)};
#pragma unroll
                    for (int pp = 0; pp < 2; ++pp) {
                        const int j = j4 * 4 + pp * 2;
                        const float2 nrm = fma2(make_float2(xv[i].get(j), xv[i].get(j + 1)), rs2, mrs2);
                        const float2 dnc = fma2(wv[pp], make_float2(gy[i].get(j), gy[i].get(j + 1)), ndn2);  // dn - mean(dn)
                        // dinp += (dn - mean(dn) - nrm * mean(dn * nrm)) * rstd
                        const float2 t = fma2(nrm, ndnn2, dnc);
                        const float2 upd = fma2(t, rs2, make_float2(dv[i].get(j), dv[i].get(j + 1)));
                        dv[i].set(j, upd.x);
                        dv[i].set(j + 1, upd.y);
                        if (COLSUM)  // the value as stored (rounded)
                            cs_acc[(i * VN + j) >> 1] = fma2(make_float2(dv[i].get(j), dv[i].get(j + 1)), one2, cs_acc[(i * VN + j) >> 1]);
                    }
                }
                dv[i].store(dinp + row * c + idx);
            }
        }
    }
    // block reduction, one quantity at a time through the same [NW][c] buffer
    for (int qn = 0; qn < (COLSUM ? 3 : 2); ++qn) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < MAXNV; ++i) {
            const int idx = (i * 32 + lane) * VN;
            if (idx < c) {
#pragma unroll
                for (int j = 0; j < VN; ++j)
                {
                    const float2 v2 = qn == 0 ? dw_acc[(i * VN + j) >> 1] : (qn == 1 ? db_acc[(i * VN + j) >> 1] : cs_acc[COLSUM ? (i * VN + j) >> 1 : 0]);
                    red[warp * c + idx + j] = (j & 1) ? v2.y : v2.x;
                }
            }
        }
        __syncthreads();
        float* dst = qn == 0 ? dweight : (qn == 1 ? dbias : colsum_out);
        for (int col = threadIdx.x; col < c; col += kThreads) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < NW; ++k) s += red[k * c + col];
            atomicAdd(dst + col, s);
        }
    }
}

template <typename T>
__global__ void ln_bwd_generic_kernel(T* __restrict__ dinp, float* __restrict__ dweight, float* __restrict__ dbias,
                                      const T* __restrict__ dout, const T* __restrict__ inp, const float* __restrict__ w,
                                      const float* __restrict__ mean, const float* __restrict__ rstd, long rows, int c,
                                      float* __restrict__ colsum_out) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float m = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < c; i += 32) {
        float nrm = (to_f32(inp[row * c + i]) - m) * rs;
        float dn = w[i] * to_f32(dout[row * c + i]);
        s1 += dn;
        s2 += dn * nrm;
    }
    const float dn_mean = warp_sum(s1) / (float)c, dnn_mean = warp_sum(s2) / (float)c;
    for (int i = lane; i < c; i += 32) {
        float g = to_f32(dout[row * c + i]);
        float nrm = (to_f32(inp[row * c + i]) - m) * rs;
        float dn = w[i] * g;
        atomicAdd(dbias + i, g);
        atomicAdd(dweight + i, nrm * g);
        T upd = from_f32<T>(to_f32(dinp[row * c + i]) + (dn - dn_mean - nrm * dnn_mean) * rs);
        dinp[row * c + i] = upd;
        if (colsum_out) atomicAdd(colsum_out + i, to_f32(upd));
    }
}

// ---- column sums: out[col] += sum_rows inp[row, col] (dbias of matmul_backward, tv:548-550) ----
template <typename T>
__global__ void colsum_kernel(float* __restrict__ out, const T* __restrict__ inp, long rows, int cols, long ld) {
    constexpr int VN = Vec16<T>::N;
    __shared__ float red[8][32 * VN + 1];
    const int col0 = (blockIdx.x * 32 + threadIdx.x) * VN;
    float acc[VN];
#pragma unroll
    for (int j = 0; j < VN; ++j) acc[j] = 0.f;
    if (col0 < cols) {
        for (long r = (long)blockIdx.y * 8 + threadIdx.y; r < rows; r += (long)gridDim.y * 8) {
            Vec16<T> v;
            v.load(inp + r * ld + col0);
#pragma unroll
            for (int j = 0; j < VN; ++j) acc[j] += v.get(j);
        }
    }
#pragma unroll
    for (int j = 0; j < VN; ++j) red[threadIdx.y][threadIdx.x * VN + j] = acc[j];
    __syncthreads();
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int cidx = tid; cidx < 32 * VN; cidx += 256) {
        const int col = blockIdx.x * 32 * VN + cidx;
        if (col < cols) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) s += red[k][cidx];
            atomicAdd(out + col, s);
        }
    }
}

template <typename T>
__global__ void colsum_generic_kernel(float* __restrict__ out, const T* __restrict__ inp, long rows, int cols, long ld) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= cols) return;
    float s = 0.f;
    for (long r = blockIdx.y; r < rows; r += gridDim.y) s += to_f32(inp[r * ld + col]);
    atomicAdd(out + col, s);
}

// ---- softmax / cross-entropy (tv:493-517, rv:836-843 with D5, rv:371) --------------------------
__global__ void softmax_fwd_kernel(float* __restrict__ probs, const float* __restrict__ logits, long rows, int v) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* z = logits + row * v;
    float mx = -INFINITY;
    for (int i = lane; i < v; i += 32) mx = fmaxf(mx, z[i]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int i = lane; i < v; i += 32) s += expf(z[i] - mx);
    s = warp_sum(s);
    for (int i = lane; i < v; i += 32) probs[row * v + i] = expf(z[i] - mx) / s;
}

// A class label outside [0, v) raises bit 0 of *flags (reported by vitrs_ctx_error_flags / vitrs_model_mean_loss) and the row
// contributes neither a loss nor a target term: a bad dataset label must not read or write out of bounds.
__global__ void ce_fwd_kernel(float* __restrict__ losses, const float* __restrict__ probs, const int* __restrict__ targets,
                              long rows, int v, int* __restrict__ flags) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int tgt = targets[r];
    if (tgt < 0 || tgt >= v) {
        atomicOr(flags, 1);
        losses[r] = 0.f;
        return;
    }
    losses[r] = -logf(probs[r * v + tgt]);
}

__global__ void ce_softmax_bwd_kernel(float* __restrict__ dlogits, const float* __restrict__ dlosses,
                                      const float* __restrict__ probs, const int* __restrict__ targets, long rows, int v,
                                      int* __restrict__ flags) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * v) return;
    const long r = i / v;
    const int j = (int)(i - r * v);
    const int tgt = targets[r];
    if (j == 0 && (tgt < 0 || tgt >= v)) atomicOr(flags, 1);
    const float ind = (j == tgt) ? 1.f : 0.f;
    dlogits[i] += (probs[i] - ind) * dlosses[r];
}

// fused head: softmax_forward + crossentropy_forward + mean (rv:337-347) + the fused backward
// (rv:366-371) in one pass, one warp per image.  mean_loss (+= loss * dloss, i.e. the mean when
// dloss = 1/B) must be zeroed by the caller.
__global__ void head_loss_kernel(float* __restrict__ probs, float* __restrict__ losses, float* __restrict__ mean_loss,
                                 float* __restrict__ dlogits, const float* __restrict__ logits,
                                 const int* __restrict__ targets, int rows, int v, float dloss, int* __restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* z = logits + (long)row * v;
    float mx = -INFINITY;
    for (int i = lane; i < v; i += 32) mx = fmaxf(mx, z[i]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int i = lane; i < v; i += 32) s += expf(z[i] - mx);
    s = warp_sum(s);
    const int tgt = targets ? targets[row] : -1;
    if (targets && lane == 0 && (tgt < 0 || tgt >= v)) {
        atomicOr(flags, 1);
        losses[row] = 0.f;
    }
    for (int i = lane; i < v; i += 32) {
        float p = expf(z[i] - mx) / s;
        probs[(long)row * v + i] = p;
        if (targets) {
            if (dlogits) dlogits[(long)row * v + i] += (p - (i == tgt ? 1.f : 0.f)) * dloss;
            if (i == tgt) {
                float l = -logf(p);
                losses[row] = l;
                atomicAdd(mean_loss, l * dloss);
            }
        }
    }
}

// ---- CLS row gather / scatter (D7) ------------------------------------------------------------
template <typename T>
__global__ void cls_gather_kernel(float* __restrict__ out, const T* __restrict__ inp, int b, int t, int c) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)b * c) return;
    const long bi = i / c, ci = i - bi * c;
    out[i] = to_f32(inp[bi * t * c + ci]);
}
template <typename T>
__global__ void cls_scatter_kernel(T* __restrict__ dinp, const float* __restrict__ dout, int b, int t, int c) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)b * c) return;
    const long bi = i / c, ci = i - bi * c;
    const long o = bi * t * c + ci;
    dinp[o] = from_f32<T>(to_f32(dinp[o]) + dout[i]);
}

// ---- optimiser (tv:737-743; AdamW per D8, torch.optim.AdamW operation order) -----------------
// Hyper-parameters live in device memory (written by a one-thread kernel per step) so that a captured CUDA graph of the
// step replays with the current learning rate and bias corrections.  GZ: gradients arrive as bf16 (the reduce-scattered
// shard of the gradient exchange buffer, ZeRO-1) and the updated bf16 weight goes back to the same place.
__global__ void adam_set_hyper_kernel(AdamHyper* dst, AdamHyper h) { *dst = h; }

template <bool GZ>
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, size_t n, const AdamHyper* __restrict__ hp, bf16* shadow) {
    const AdamHyper h = *hp;
    const float lr = h.lr, b1 = h.b1, b2 = h.b2, eps = h.eps, wd = h.wd, step_size = h.step_size, bc2_sqrt = h.bc2_sqrt;
    const size_t nv = n / 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        float4 gg;
        if (GZ) {
            const uint2 raw = reinterpret_cast<const uint2*>(shadow)[i];
            gg.x = __uint_as_float(raw.x << 16); gg.y = __uint_as_float(raw.x & 0xFFFF0000u);
            gg.z = __uint_as_float(raw.y << 16); gg.w = __uint_as_float(raw.y & 0xFFFF0000u);
        } else {
            gg = reinterpret_cast<const float4*>(g)[i];
        }
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float pj = pa[j] * (1.0f - lr * wd);
            float mj = b1 * ma[j] + (1.0f - b1) * ga[j];
            float vj = b2 * va[j] + (1.0f - b2) * ga[j] * ga[j];
            ma[j] = mj;
            va[j] = vj;
            pa[j] = pj - step_size * (mj / (sqrtf(vj) / bc2_sqrt + eps));
        }
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
        if (shadow) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(pp.x, pp.y), hi = __floats2bfloat162_rn(pp.z, pp.w);
            uint2 packed;
            packed.x = *reinterpret_cast<uint32_t*>(&lo);
            packed.y = *reinterpret_cast<uint32_t*>(&hi);
            reinterpret_cast<uint2*>(shadow)[i] = packed;
        }
    }
    for (size_t i = nv * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gi = GZ ? __bfloat162float(shadow[i]) : g[i];
        float pj = p[i] * (1.0f - lr * wd);
        float mj = b1 * m[i] + (1.0f - b1) * gi;
        float vj = b2 * v[i] + (1.0f - b2) * gi * gi;
        m[i] = mj;
        v[i] = vj;
        pj = pj - step_size * (mj / (sqrtf(vj) / bc2_sqrt + eps));
        p[i] = pj;
        if (shadow) shadow[i] = __float2bfloat16_rn(pj);
    }
}

__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// ---- gradient exchange buffer ("Z order": bucket-major, slices of a bucket back to back) ---------------------------
// One launch moves every slice of a bucket between the tensor-major flat buffers and its contiguous region of the
// exchange buffer (blockIdx.y = slice).  PACK: z[z_off + i] = cvt(flat[src_off + i]); else flat[src_off + i] = cvt(z[z_off + i]).
template <typename TF, typename TZ, bool PACK, int VEC>
__global__ void slice_copy_kernel(TF* __restrict__ flat, TZ* __restrict__ z, const SliceTable tab, size_t total_units) {
    // the slices form one virtual array of units (VEC elements each: 4 when every slice is 4-aligned, else 1); a thread walks
    // it with a grid stride and finds the slice of a unit in the (<= 12 entry) table
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x; u < total_units; u += stride) {
        size_t local = u;
        int sl = 0;
        while (sl + 1 < tab.n && local >= tab.cnt[sl] / VEC) { local -= tab.cnt[sl] / VEC; ++sl; }
        TF* f = flat + tab.src_off[sl] + local * VEC;
        TZ* zz = z + tab.z_off[sl] + local * VEC;
        if (VEC == 4) {
            float x[4];
            if (PACK) {
                if (sizeof(TF) == 4) { const float4 a = *reinterpret_cast<const float4*>(f); x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; }
                else { const uint2 a = *reinterpret_cast<const uint2*>(f); x[0] = __uint_as_float(a.x << 16); x[1] = __uint_as_float(a.x & 0xFFFF0000u); x[2] = __uint_as_float(a.y << 16); x[3] = __uint_as_float(a.y & 0xFFFF0000u); }
                if (sizeof(TZ) == 4) *reinterpret_cast<float4*>(zz) = make_float4(x[0], x[1], x[2], x[3]);
                else { uint2 o; o.x = pack2_bf16(x[0], x[1]); o.y = pack2_bf16(x[2], x[3]); *reinterpret_cast<uint2*>(zz) = o; }
            } else {
                if (sizeof(TZ) == 4) { const float4 a = *reinterpret_cast<const float4*>(zz); x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; }
                else { const uint2 a = *reinterpret_cast<const uint2*>(zz); x[0] = __uint_as_float(a.x << 16); x[1] = __uint_as_float(a.x & 0xFFFF0000u); x[2] = __uint_as_float(a.y << 16); x[3] = __uint_as_float(a.y & 0xFFFF0000u); }
                if (sizeof(TF) == 4) *reinterpret_cast<float4*>(f) = make_float4(x[0], x[1], x[2], x[3]);
                else { uint2 o; o.x = pack2_bf16(x[0], x[1]); o.y = pack2_bf16(x[2], x[3]); *reinterpret_cast<uint2*>(f) = o; }
            }
        } else {
            if (PACK) *zz = from_f32<TZ>(to_f32(*f));
            else *f = from_f32<TF>(to_f32(*zz));
        }
    }
}

__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, size_t n, float lr, bf16* __restrict__ shadow) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float pj = p[i] - lr * g[i];
        p[i] = pj;
        if (shadow) shadow[i] = __float2bfloat16_rn(pj);
    }
}

// counter-based generator shared with oracle/vit_oracle.c::vit_rand_u01 (D9)
__global__ void fill_uniform_kernel(float* __restrict__ dst, size_t n, uint64_t seed, uint64_t stream, float lo, float hi) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const uint64_t base = seed * 0x9E3779B97F4A7C15ull + stream * 0xD1B54A32D192ED03ull;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t x = base + i;
        x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
        x ^= x >> 27; x *= 0x94D049BB133111EBull;
        x ^= x >> 31;
        float u = (float)(x >> 40) * (1.0f / 16777216.0f);
        dst[i] = __fadd_rn(lo, __fmul_rn(hi - lo, u));  // no FMA contraction: bit-equal to the C oracle
    }
}

__global__ void fill_const_kernel(float* __restrict__ dst, size_t n, float v) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = v;
}

// out += scale * sum(inp[0..n)), one warp, ascending strides (mean loss, rv:342-347)
__global__ void scaled_sum_kernel(float* __restrict__ out, const float* __restrict__ inp, long n, float scale) {
    float s = 0.f;
    for (long i = threadIdx.x; i < n; i += 32) s += inp[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) *out += s * scale;
}

__global__ void cast_f32_bf16_kernel(bf16* __restrict__ dst, const float* __restrict__ src, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = __float2bfloat16_rn(src[i]);
}
__global__ void cast_bf16_f32_kernel(float* __restrict__ dst, const bf16* __restrict__ src, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = __bfloat162float(src[i]);
}

template <typename T> constexpr bool is_bf16() { return sizeof(T) == 2; }
inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

// ================================ launchers ====================================================
template <typename T> int op_residual_forward(vitrs_ctx* ctx, T* out, const T* a, const T* b, long n) {
    if (n <= 0) return VITRS_OK;
    VITRS_ARG(ctx, aligned16(out) && aligned16(a) && aligned16(b));
    residual_fwd_kernel<T><<<grid_for(n / Vec16<T>::N + 1, kThreads, ctx->sm_count), kThreads, 0, ctx->stream>>>(out, a, b, n);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}
template <typename T> int op_residual_backward(vitrs_ctx* ctx, T* d1, T* d2, const T* dout, long n) {
    if (n <= 0) return VITRS_OK;
    VITRS_ARG(ctx, aligned16(d1) && aligned16(d2) && aligned16(dout));
    residual_bwd_kernel<T><<<grid_for(n / Vec16<T>::N + 1, kThreads, ctx->sm_count), kThreads, 0, ctx->stream>>>(d1, d2, dout, n);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}
template <typename T> int op_gelu_forward(vitrs_ctx* ctx, T* out, const T* inp, long n) {
    if (n <= 0) return VITRS_OK;
    VITRS_ARG(ctx, aligned16(out) && aligned16(inp));
    gelu_fwd_kernel<T, is_bf16<T>()><<<grid_for(n / Vec16<T>::N + 1, kThreads, ctx->sm_count), kThreads, 0, ctx->stream>>>(out, inp, n);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}
template <typename T> int op_gelu_backward(vitrs_ctx* ctx, T* dinp, const T* inp, const T* dout, long n) {
    if (n <= 0) return VITRS_OK;
    VITRS_ARG(ctx, aligned16(dinp) && aligned16(inp) && aligned16(dout));
    gelu_bwd_kernel<T, is_bf16<T>()><<<grid_for(n / Vec16<T>::N + 1, kThreads, ctx->sm_count), kThreads, 0, ctx->stream>>>(dinp, inp, dout, n);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}

template <typename T>
int op_layernorm_forward(vitrs_ctx* ctx, T* out, float* mean, float* rstd, const T* inp, const float* w, const float* b,
                         long rows, int c) {
    if (rows <= 0) return VITRS_OK;
    constexpr int VN = Vec16<T>::N;
    const int nv = (c + 32 * VN - 1) / (32 * VN);
    int grid = ceil_div(rows, kThreads / 32);
    const bool vec_ok = (c % VN == 0) && aligned16(out) && aligned16(inp) && aligned16(w) && aligned16(b) && nv <= 8;
    if (!vec_ok) {
        ln_fwd_generic_kernel<T><<<grid, kThreads, 0, ctx->stream>>>(out, mean, rstd, inp, w, b, rows, c);
        VITRS_LAUNCHED(ctx);
        return VITRS_OK;
    }
    const int per_sm = nv <= 4 ? 3 : 1;
    if (grid > per_sm * ctx->sm_count) grid = per_sm * ctx->sm_count;  // persistent: warps stride over rows
    const size_t smem = ((size_t)2 * c + 4) * sizeof(float);  // gains, biases, the constant 1 (nv <= 8: at most 16 KB)
    if (false) {
    } else if (nv <= 1) {
        ln_fwd_kernel<T, 1><<<grid, kThreads, smem, ctx->stream>>>(out, mean, rstd, inp, w, b, rows, c);
    } else if (nv <= 2) {
        ln_fwd_kernel<T, 2><<<grid, kThreads, smem, ctx->stream>>>(out, mean, rstd, inp, w, b, rows, c);
    } else if (nv <= 3) {
        ln_fwd_kernel<T, 3><<<grid, kThreads, smem, ctx->stream>>>(out, mean, rstd, inp, w, b, rows, c);
    } else if (nv <= 4) {
        ln_fwd_kernel<T, 4><<<grid, kThreads, smem, ctx->stream>>>(out, mean, rstd, inp, w, b, rows, c);
    } else {
        ln_fwd_kernel<T, 8><<<grid, kThreads, smem, ctx->stream>>>(out, mean, rstd, inp, w, b, rows, c);
    }
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}

template <typename T, int MAXNV>
static int launch_ln_bwd(vitrs_ctx* ctx, T* dinp, float* dw, float* db, const T* dout, const T* inp, const float* w,
                         const float* mean, const float* rstd, long rows, int c, float* colsum_out) {
    int grid = ceil_div(rows, kThreads / 32);
    if (grid > 2 * ctx->sm_count) grid = 2 * ctx->sm_count;
    const size_t smem = ((size_t)(kThreads / 32 + 1) * c + 4) * sizeof(float);  // reduction scratch, gains, the constant 1
    if (colsum_out) {
        auto k = ln_bwd_kernel<T, MAXNV, true>;
        if (smem > 48 * 1024) VITRS_TRY(vitrs_func_smem(ctx, (const void*)k, smem));
        k<<<grid, kThreads, smem, ctx->stream>>>(dinp, dw, db, dout, inp, w, mean, rstd, rows, c, colsum_out);
    } else {
        auto k = ln_bwd_kernel<T, MAXNV, false>;
        if (smem > 48 * 1024) VITRS_TRY(vitrs_func_smem(ctx, (const void*)k, smem));
        k<<<grid, kThreads, smem, ctx->stream>>>(dinp, dw, db, dout, inp, w, mean, rstd, rows, c, colsum_out);
    }
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}

template <typename T>
int op_layernorm_backward(vitrs_ctx* ctx, T* dinp, float* dw, float* db, const T* dout, const T* inp, const float* w,
                          const float* mean, const float* rstd, long rows, int c, float* colsum_out) {
    if (rows <= 0) return VITRS_OK;
    constexpr int VN = Vec16<T>::N;
    const int nv = (c + 32 * VN - 1) / (32 * VN);
    const bool vec_ok = (c % VN == 0) && aligned16(dinp) && aligned16(dout) && aligned16(inp) && nv <= 8;
    if (!vec_ok) {
        ln_bwd_generic_kernel<T><<<ceil_div(rows, kThreads / 32), kThreads, 0, ctx->stream>>>(dinp, dw, db, dout, inp, w, mean,
                                                                                            rstd, rows, c, colsum_out);
        VITRS_LAUNCHED(ctx);
        return VITRS_OK;
    }
    if (nv <= 1) return launch_ln_bwd<T, 1>(ctx, dinp, dw, db, dout, inp, w, mean, rstd, rows, c, colsum_out);
    if (nv <= 2) return launch_ln_bwd<T, 2>(ctx, dinp, dw, db, dout, inp, w, mean, rstd, rows, c, colsum_out);
    if (nv <= 3) return launch_ln_bwd<T, 3>(ctx, dinp, dw, db, dout, inp, w, mean, rstd, rows, c, colsum_out);
    if (nv <= 4) return launch_ln_bwd<T, 4>(ctx, dinp, dw, db, dout, inp, w, mean, rstd, rows, c, colsum_out);
    return launch_ln_bwd<T, 8>(ctx, dinp, dw, db, dout, inp, w, mean, rstd, rows, c, colsum_out);
}

template <typename T> int op_colsum(vitrs_ctx* ctx, float* out, const T* inp, long rows, int cols, long ld) {
    if (rows <= 0 || cols <= 0) return VITRS_OK;
    constexpr int VN = Vec16<T>::N;
    if (cols % VN == 0 && ld % VN == 0 && aligned16(inp)) {
        dim3 block(32, 8);
        const int gx = ceil_div(cols, 32 * VN);
        int gy = ceil_div(rows, 8 * 16);
        const int cap = (4 * ctx->sm_count + gx - 1) / gx;
        if (gy > cap) gy = cap;
        if (gy < 1) gy = 1;
        colsum_kernel<T><<<dim3(gx, gy), block, 0, ctx->stream>>>(out, inp, rows, cols, ld);
    } else {
        int gy = ceil_div(rows, 64);
        if (gy > 256) gy = 256;
        colsum_generic_kernel<T><<<dim3(ceil_div(cols, 128), gy), 128, 0, ctx->stream>>>(out, inp, rows, cols, ld);
    }
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}

int op_softmax_forward(vitrs_ctx* ctx, float* probs, const float* logits, long rows, int v) {
    if (rows <= 0) return VITRS_OK;
    softmax_fwd_kernel<<<ceil_div(rows, 4), 128, 0, ctx->stream>>>(probs, logits, rows, v);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}
int op_crossentropy_forward(vitrs_ctx* ctx, float* losses, const float* probs, const int* targets, long rows, int v) {
    if (rows <= 0) return VITRS_OK;
    ce_fwd_kernel<<<ceil_div(rows, 128), 128, 0, ctx->stream>>>(losses, probs, targets, rows, v, ctx->dev_flags);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}
int op_crossentropy_softmax_backward(vitrs_ctx* ctx, float* dlogits, const float* dlosses, const float* probs,
                                     const int* targets, long rows, int v) {
    if (rows <= 0) return VITRS_OK;
    ce_softmax_bwd_kernel<<<ceil_div(rows * v, 256), 256, 0, ctx->stream>>>(dlogits, dlosses, probs, targets, rows, v, ctx->dev_flags);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}
int op_head_loss(vitrs_ctx* ctx, float* probs, float* losses, float* mean_loss, float* dlogits, const float* logits,
                 const int* targets, int rows, int v, float dloss) {
    if (rows <= 0) return VITRS_OK;
    head_loss_kernel<<<ceil_div(rows, 4), 128, 0, ctx->stream>>>(probs, losses, mean_loss, dlogits, logits, targets, rows, v, dloss, ctx->dev_flags);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}

template <typename T> int op_cls_gather(vitrs_ctx* ctx, float* out, const T* inp, int b, int t, int c) {
    cls_gather_kernel<T><<<ceil_div((long)b * c, 256), 256, 0, ctx->stream>>>(out, inp, b, t, c);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}
template <typename T> int op_cls_scatter_add(vitrs_ctx* ctx, T* dinp, const float* dout, int b, int t, int c) {
    cls_scatter_kernel<T><<<ceil_div((long)b * c, 256), 256, 0, ctx->stream>>>(dinp, dout, b, t, c);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}

int op_adam_set_hyper(vitrs_ctx* ctx, float lr, float b1, float b2, float eps, float wd, int step, cudaStream_t stream) {
    AdamHyper h;
    const float bc1 = 1.0f - powf(b1, (float)step);
    const float bc2 = 1.0f - powf(b2, (float)step);
    h.lr = lr; h.b1 = b1; h.b2 = b2; h.eps = eps; h.wd = wd; h.step_size = lr / bc1; h.bc2_sqrt = sqrtf(bc2); h.pad = 0.f;
    adam_set_hyper_kernel<<<1, 1, 0, stream>>>(ctx->d_hyper, h);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}
// the update with the hyper-parameters last set on this context (graph-capturable: no step-dependent launch argument)
int op_adamw_apply(vitrs_ctx* ctx, float* p, const float* g, float* m, float* v, size_t n, bf16* shadow, cudaStream_t stream) {
    if (n == 0) return VITRS_OK;
    VITRS_ARG(ctx, aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v) && (!shadow || ((uintptr_t)shadow & 7) == 0));
    adamw_kernel<false><<<grid_for((long)(n / 4 + 1), kThreads, ctx->sm_count), kThreads, 0, stream>>>(p, g, m, v, n, ctx->d_hyper, shadow);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}
// ZeRO-1 shard: gradients are the bf16 values at gz (reduce-scattered), which receives the updated bf16 weights
int op_adamw_apply_shard(vitrs_ctx* ctx, float* p, bf16* gz, float* m, float* v, size_t n, cudaStream_t stream) {
    if (n == 0) return VITRS_OK;
    VITRS_ARG(ctx, aligned16(p) && aligned16(m) && aligned16(v) && ((uintptr_t)gz & 7) == 0);
    adamw_kernel<true><<<grid_for((long)(n / 4 + 1), kThreads, ctx->sm_count), kThreads, 0, stream>>>(p, nullptr, m, v, n, ctx->d_hyper, gz);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}
int op_adamw(vitrs_ctx* ctx, float* p, const float* g, float* m, float* v, size_t n, float lr, float b1, float b2,
             float eps, float wd, int step, bf16* shadow) {
    if (n == 0) return VITRS_OK;
    VITRS_TRY(op_adam_set_hyper(ctx, lr, b1, b2, eps, wd, step, ctx->stream));
    return op_adamw_apply(ctx, p, g, m, v, n, shadow, ctx->stream);
}

template <typename TF, typename TZ, bool PACK>
static int slice_copy(vitrs_ctx* ctx, TF* flat, TZ* z, const SliceTable& tab, cudaStream_t stream) {
    if (tab.n <= 0) return VITRS_OK;
    size_t total = 0;
    bool vec = true;
    for (int i = 0; i < tab.n; ++i) {
        total += tab.cnt[i];
        vec = vec && ((tab.cnt[i] | tab.src_off[i] | tab.z_off[i]) & 3) == 0;
    }
    if (total == 0) return VITRS_OK;
    const size_t units = vec ? total / 4 : total;
    int gx = (int)((units + kThreads - 1) / kThreads);
    if (gx > 8 * ctx->sm_count) gx = 8 * ctx->sm_count;
    if (vec) slice_copy_kernel<TF, TZ, PACK, 4><<<gx, kThreads, 0, stream>>>(flat, z, tab, units);
    else slice_copy_kernel<TF, TZ, PACK, 1><<<gx, kThreads, 0, stream>>>(flat, z, tab, units);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}
int op_pack_f32_to_bf16(vitrs_ctx* ctx, bf16* z, const float* flat, const SliceTable& tab, cudaStream_t s) {
    return slice_copy<float, bf16, true>(ctx, const_cast<float*>(flat), z, tab, s);
}
int op_unpack_bf16_to_f32(vitrs_ctx* ctx, float* flat, const bf16* z, const SliceTable& tab, cudaStream_t s) {
    return slice_copy<float, bf16, false>(ctx, flat, const_cast<bf16*>(z), tab, s);
}
int op_unpack_bf16_to_bf16(vitrs_ctx* ctx, bf16* flat, const bf16* z, const SliceTable& tab, cudaStream_t s) {
    return slice_copy<bf16, bf16, false>(ctx, flat, const_cast<bf16*>(z), tab, s);
}
int op_pack_f32_to_f32(vitrs_ctx* ctx, float* z, const float* flat, const SliceTable& tab, cudaStream_t s) {
    return slice_copy<float, float, true>(ctx, const_cast<float*>(flat), z, tab, s);
}
int op_unpack_f32_to_f32(vitrs_ctx* ctx, float* flat, const float* z, const SliceTable& tab, cudaStream_t s) {
    return slice_copy<float, float, false>(ctx, flat, const_cast<float*>(z), tab, s);
}
int op_sgd(vitrs_ctx* ctx, float* p, const float* g, size_t n, float lr, bf16* shadow) {
    if (n == 0) return VITRS_OK;
    sgd_kernel<<<grid_for((long)n, kThreads, ctx->sm_count), kThreads, 0, ctx->stream>>>(p, g, n, lr, shadow);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}
int op_fill_uniform(vitrs_ctx* ctx, float* dst, size_t n, uint64_t seed, uint64_t stream, float lo, float hi) {
    if (n == 0) return VITRS_OK;
    fill_uniform_kernel<<<grid_for((long)n, kThreads, ctx->sm_count), kThreads, 0, ctx->stream>>>(dst, n, seed, stream, lo, hi);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}
int op_fill_const(vitrs_ctx* ctx, float* dst, size_t n, float v) {
    if (n == 0) return VITRS_OK;
    fill_const_kernel<<<grid_for((long)n, kThreads, ctx->sm_count), kThreads, 0, ctx->stream>>>(dst, n, v);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}
int op_scaled_sum(vitrs_ctx* ctx, float* out, const float* inp, long n, float scale) {
    scaled_sum_kernel<<<1, 32, 0, ctx->stream>>>(out, inp, n, scale);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}
int op_cast_f32_bf16(vitrs_ctx* ctx, bf16* dst, const float* src, size_t n) {
    if (n == 0) return VITRS_OK;
    cast_f32_bf16_kernel<<<grid_for((long)n, kThreads, ctx->sm_count), kThreads, 0, ctx->stream>>>(dst, src, n);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}
int op_cast_bf16_f32(vitrs_ctx* ctx, float* dst, const bf16* src, size_t n) {
    if (n == 0) return VITRS_OK;
    cast_bf16_f32_kernel<<<grid_for((long)n, kThreads, ctx->sm_count), kThreads, 0, ctx->stream>>>(dst, src, n);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}

#define INSTANTIATE(T)                                                                                          \
    template int op_residual_forward<T>(vitrs_ctx*, T*, const T*, const T*, long);                              \
    template int op_residual_backward<T>(vitrs_ctx*, T*, T*, const T*, long);                                   \
    template int op_gelu_forward<T>(vitrs_ctx*, T*, const T*, long);                                            \
    template int op_gelu_backward<T>(vitrs_ctx*, T*, const T*, const T*, long);                                 \
    template int op_layernorm_forward<T>(vitrs_ctx*, T*, float*, float*, const T*, const float*, const float*, long, int); \
    template int op_layernorm_backward<T>(vitrs_ctx*, T*, float*, float*, const T*, const T*, const float*,     \
                                          const float*, const float*, long, int, float*);                       \
    template int op_colsum<T>(vitrs_ctx*, float*, const T*, long, int, long);                                   \
    template int op_cls_gather<T>(vitrs_ctx*, float*, const T*, int, int, int);                                 \
    template int op_cls_scatter_add<T>(vitrs_ctx*, T*, const float*, int, int, int);
INSTANTIATE(float)
INSTANTIATE(bf16)
