// epilogue.cuh — the fused GEMM epilogues, one scalar definition used by both the SIMT fp32
// kernel and the tcgen05 bf16 kernel (which calls it on register tiles).
#pragma once
#include "common.cuh"

// result for output element (m, n) given the fp32 accumulator; TO = storage type of out/aux.
// Returns the value to store in `out`; *second receives the value for out2 (EPI_BIAS_GELU).
template <typename TO, bool FAST>
__device__ __forceinline__ float epi_value(const Epilogue& e, long m, int n, float acc, float aux, float* second) {
    float v = acc;
    switch (e.kind) {
        case EPI_BIAS:
            if (e.bias) v += __ldg(e.bias + n);
            break;
        case EPI_BIAS_GELU: {
            if (e.bias) v += __ldg(e.bias + n);
            // gelu_forward consumes the stored (rounded) pre-activation, as the unfused op would
            float x = to_f32(from_f32<TO>(v));
            *second = gelu_fwd<FAST>(x);
            break;
        }
        case EPI_BIAS_GELU_ONLY:
            if (e.bias) v += __ldg(e.bias + n);
            v = gelu_fwd<FAST>(to_f32(from_f32<TO>(v)));
            break;
        case EPI_BIAS_RESIDUAL:
            if (e.bias) v += __ldg(e.bias + n);
            v += aux;
            break;
        case EPI_GELU_BWD:
            v *= gelu_grad<FAST>(aux);
            break;
        case EPI_PATCH: {
            const int tok = (int)(m % e.np);
            if (tok == 0) v = __ldg(e.cls + n);
            else if (e.bias) v += __ldg(e.bias + n);
            v += __ldg(e.pos + (long)tok * e.ldo + n);
            break;
        }
        default:
            break;
    }
    return v;
}

__device__ __forceinline__ long epi_out_row(const Epilogue&, long m) { return m; }

__device__ __forceinline__ bool epi_needs_aux(int kind) { return kind == EPI_BIAS_RESIDUAL || kind == EPI_GELU_BWD; }

template <typename TO, bool FAST>
__device__ __forceinline__ void epi_store_scalar(const Epilogue& e, long m, int n, float acc) {
    if (e.kind == EPI_ACCUM_F32) {
        atomicAdd(reinterpret_cast<float*>(e.out) + m * e.ldo + n, acc);
        return;
    }
    const long orow = epi_out_row(e, m);
    TO* out = reinterpret_cast<TO*>(e.out) + orow * e.ldo + n;
    float aux = 0.f;
    if (epi_needs_aux(e.kind)) aux = to_f32(reinterpret_cast<const TO*>(e.aux)[m * e.ldo + n]);
    float second = 0.f;
    float v = epi_value<TO, FAST>(e, m, n, acc, aux, &second);
    if (e.accumulate) v += to_f32(*out);
    *out = from_f32<TO>(v);
    if (e.kind == EPI_BIAS_GELU) reinterpret_cast<TO*>(e.out2)[orow * e.ldo + n] = from_f32<TO>(second);
}
