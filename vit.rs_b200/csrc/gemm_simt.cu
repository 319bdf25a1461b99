// gemm_simt.cu — plain SIMT GEMM with fp32 multiply and fp32 accumulate.
// This is verify mode's matmul (TF32's 10-bit mantissa cannot meet the 1e-4 gate) and the
// route for the few GEMMs too small or too ragged for the tcgen05 kernel (the class head).
// D[m,n] = sum_k A(m,k) * B(n,k) with arbitrary element strides, then the shared epilogue.
// Replaces the three-loop matmul_forward / matmul_backward of train_vit.rs:384-398, 530-557.
//
// In fp32 the reference's arithmetic is reproduced exactly, not just its result: the
// accumulator starts at the bias (train_vit.rs:389) or at the value being accumulated into
// (`+=`, :538,552), k runs ascending, and each step is a rounded multiply followed by a rounded
// add (Rust never contracts to FMA).  With the reference's all-positive init the sums are long
// same-sign chains whose rounding alone is ~1e-4 of the result, so matching the order is what
// makes the 1e-4 verify gate meaningful.
#include "epilogue.cuh"

namespace {

constexpr int BK = 16;

// TM x TM outputs per thread, 16 x 16 threads: [64 x 64] tiles (TM = 4), or [32 x 32] (TM = 2) when the larger tile would leave
// most of the chip idle (the class head of the small models: 12 CTAs for ViT-Ti/16's dX).  The order of additions per output
// (k ascending) does not depend on the tile: results are bit-identical.
template <typename T, bool FAST, bool EXACT, int TM>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmDesc g) {
    constexpr int BM = 16 * TM, BN = 16 * TM;
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const T* __restrict__ A = reinterpret_cast<const T*>(g.A);
    const T* __restrict__ B = reinterpret_cast<const T*>(g.B);
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long m0 = (long)blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;
    const bool a_kfast = g.a_ks == 1, b_kfast = g.b_ks == 1;
    const Epilogue& e = g.epi;
    float acc[TM][TM];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TM; ++j) {
            const long m = m0 + ty * TM + i;
            const int n = n0 + tx * TM + j;
            float init = 0.f;
            if (m < g.M && n < g.N) {
                if (e.kind == EPI_ACCUM_F32) init = reinterpret_cast<const float*>(e.out)[m * e.ldo + n];
                else if (e.accumulate) init = to_f32(reinterpret_cast<const T*>(e.out)[m * e.ldo + n]);
                else if (e.bias && (e.kind == EPI_BIAS || e.kind == EPI_BIAS_GELU || e.kind == EPI_BIAS_RESIDUAL || e.kind == EPI_PATCH))
                    init = e.bias[n];
            }
            acc[i][j] = init;
        }

    for (int k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int e = tid + i * 256;
            // consecutive threads follow the operand's unit-stride index
            const int am = a_kfast ? e / BK : e % BM, ak = a_kfast ? e % BK : e / BM;
            const long gm = m0 + am;
            const int gk = k0 + ak;
            As[ak][am] = (gm < g.M && gk < g.K) ? to_f32(A[gm * g.a_rs + gk * g.a_ks]) : 0.f;
            const int bn = b_kfast ? e / BK : e % BN, bk = b_kfast ? e % BK : e / BN;
            const int gn = n0 + bn;
            const int gk2 = k0 + bk;
            Bs[bk][bn] = (gn < g.N && gk2 < g.K) ? to_f32(B[(long)gn * g.b_rs + gk2 * g.b_ks]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TM];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TM; ++j) b[j] = Bs[k][tx * TM + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TM; ++j)
                    acc[i][j] = EXACT ? __fadd_rn(acc[i][j], __fmul_rn(a[i], b[j])) : fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const long m = m0 + ty * TM + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TM; ++j) {
            const int n = n0 + tx * TM + j;
            if (n >= g.N) continue;
            if (e.kind == EPI_ACCUM_F32) {
                reinterpret_cast<float*>(e.out)[m * e.ldo + n] = acc[i][j];  // the old value is already in acc
            } else {
                Epilogue e2 = e;  // bias / old value were the accumulator's starting point
                e2.bias = nullptr;
                e2.accumulate = 0;
                epi_store_scalar<T, FAST>(e2, m, n, acc[i][j]);
            }
        }
    }
}

template <typename T, bool FAST, bool EXACT> int launch(vitrs_ctx* ctx, const GemmDesc& g) {
    if (g.M <= 0 || g.N <= 0) return VITRS_OK;
    if ((long)ceil_div(g.N, 64) * ceil_div(g.M, 64) >= ctx->sm_count) {
        dim3 grid(ceil_div(g.N, 64), ceil_div(g.M, 64));
        gemm_simt_kernel<T, FAST, EXACT, 4><<<grid, 256, 0, ctx->stream>>>(g);
    } else {
        dim3 grid(ceil_div(g.N, 32), ceil_div(g.M, 32));
        gemm_simt_kernel<T, FAST, EXACT, 2><<<grid, 256, 0, ctx->stream>>>(g);
    }
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}

}  // namespace

int gemm_simt_f32(vitrs_ctx* ctx, const GemmDesc& g) { return launch<float, false, true>(ctx, g); }
int gemm_simt_bf16(vitrs_ctx* ctx, const GemmDesc& g) { return launch<bf16, true, false>(ctx, g); }
