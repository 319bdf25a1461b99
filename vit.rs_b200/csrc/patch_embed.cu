// patch_embed.cu — the pieces either side of the patch-projection GEMM (DEVIATIONS D7), and the
// reference's token encoder (called at rusty_vit.rs:282,448; restated from its signature).
//
// Patch embedding = im2col + GEMM.  The im2col matrix has one row per TOKEN, [B*T, 3*p*p], with
// the CLS row of every image left zero: the projection GEMM then runs over the same row space as
// `encoded` (no row remapping), its epilogue (EPI_PATCH) writes cls + wpe[0] into token 0 and
// acc + patchb + wpe[tok] elsewhere, and the weight gradient dpatchw += dencoded^T . patches
// needs no gather because the zero rows contribute nothing.  Column order of a patch vector is
// (channel, row, col) = the Conv2d weight order used by the oracle (vit_oracle.c patch_embed_forward).
#include "common.cuh"

namespace {

__device__ __forceinline__ void store4(float* dst, float4 v) { *reinterpret_cast<float4*>(dst) = v; }
__device__ __forceinline__ void store4(bf16* dst, float4 v) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 packed;
    packed.x = *reinterpret_cast<uint32_t*>(&lo);
    packed.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(dst) = packed;
}

template <typename T>
__global__ void im2col_kernel(T* __restrict__ patches, const float* __restrict__ images, int b, int img, int patch) {
    const int g = img / patch, np = g * g, t = np + 1, kdim = 3 * patch * patch;
    const int k4 = kdim / 4;
    const long total = (long)b * t * k4;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int kq = (int)(idx % k4);
        const long row = idx / k4;
        const int tok = (int)(row % t);
        const long bi = row / t;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tok > 0) {
            const int n = tok - 1, py = n / g, px = n - py * g;
            const int k = kq * 4;
            const int ch = k / (patch * patch), rem = k - ch * patch * patch;
            const int i = rem / patch, j = rem - i * patch;
            v = *reinterpret_cast<const float4*>(images + ((bi * 3 + ch) * img + (py * patch + i)) * (long)img + px * patch + j);
        }
        store4(patches + row * kdim + kq * 4, v);
    }
}

// The same pass over raw dataset images: uint8 samples, NCHW [B,3,H,W] (layout 0) or NHWC [B,H,W,3] (layout 1, the order of
// CIFAR-10 records and decoded JPEGs), normalised on the fly: (x / 255 - mean[ch]) / std[ch].  scale[ch] = 1 / (255 std),
// shift[ch] = -mean / std: one FMA per sample; the host sends one byte per sample instead of four.
struct NormParams {
    float scale[3], shift[3];
};
template <typename T>
__global__ void im2col_u8_kernel(T* __restrict__ patches, const uint8_t* __restrict__ images, int layout, NormParams nrm, int b, int img,
                                 int patch) {
    const int g = img / patch, np = g * g, t = np + 1, kdim = 3 * patch * patch;
    const int k4 = kdim / 4;
    const long total = (long)b * t * k4;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int kq = (int)(idx % k4);
        const long row = idx / k4;
        const int tok = (int)(row % t);
        const long bi = row / t;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tok > 0) {
            const int n = tok - 1, py = n / g, px = n - py * g;
            const int k = kq * 4;
            const int ch = k / (patch * patch), rem = k - ch * patch * patch;
            const int i = rem / patch, j = rem - i * patch;
            const int y = py * patch + i, x = px * patch + j;
            uchar4 u;
            if (layout == 0) {
                u = *reinterpret_cast<const uchar4*>(images + ((bi * 3 + ch) * img + y) * (long)img + x);
            } else {
                const uint8_t* px0 = images + ((bi * img + y) * (long)img + x) * 3 + ch;
                u = make_uchar4(px0[0], px0[3], px0[6], px0[9]);
            }
            const float sc = ch == 0 ? nrm.scale[0] : (ch == 1 ? nrm.scale[1] : nrm.scale[2]);
            const float sh = ch == 0 ? nrm.shift[0] : (ch == 1 ? nrm.shift[1] : nrm.shift[2]);
            v = make_float4(fmaf((float)u.x, sc, sh), fmaf((float)u.y, sc, sh), fmaf((float)u.z, sc, sh), fmaf((float)u.w, sc, sh));
        }
        store4(patches + row * kdim + kq * 4, v);
    }
}

// sums over the batch: dwpe[t,c] += s, dcls[c] += s (t == 0), dpatchb[c] += s (t > 0)
template <typename T>
__global__ void patch_bwd_reduce_kernel(float* __restrict__ dwpe, float* __restrict__ dcls, float* __restrict__ dpatchb,
                                        const T* __restrict__ denc, int b, int t, int c) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= t * c) return;
    const int tok = idx / c, col = idx - tok * c;
    float s = 0.f;
    for (int bi = blockIdx.y; bi < b; bi += gridDim.y) s += to_f32(denc[((long)bi * t + tok) * c + col]);
    atomicAdd(dwpe + idx, s);
    if (tok == 0) atomicAdd(dcls + col, s);
    else atomicAdd(dpatchb + col, s);
}

__global__ void encoder_fwd_kernel(float* __restrict__ enc, const int* __restrict__ inputs, const float* __restrict__ wte,
                                   const float* __restrict__ wpe, int b, int t, int c) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)b * t * c) return;
    const int i = (int)(idx % c);
    const long bt = idx / c;
    const int ti = (int)(bt % t);
    enc[idx] = wte[(long)inputs[bt] * c + i] + wpe[(long)ti * c + i];
}

__global__ void encoder_bwd_kernel(float* __restrict__ dwte, float* __restrict__ dwpe, const float* __restrict__ denc,
                                   const int* __restrict__ inputs, int b, int t, int c) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)b * t * c) return;
    const int i = (int)(idx % c);
    const long bt = idx / c;
    const int ti = (int)(bt % t);
    const float g = denc[idx];
    atomicAdd(dwte + (long)inputs[bt] * c + i, g);
    atomicAdd(dwpe + (long)ti * c + i, g);
}

}  // namespace

template <typename T> int op_im2col(vitrs_ctx* ctx, T* patches, const float* images, int b, int img, int patch) {
    if (b <= 0) return VITRS_OK;
    VITRS_ARG(ctx, patch > 0 && img % patch == 0 && patch % 4 == 0 && ((uintptr_t)images & 15) == 0);
    const int g = img / patch, t = g * g + 1, kdim = 3 * patch * patch;
    const long total = (long)b * t * (kdim / 4);
    long grid = (total + 255) / 256;
    if (grid > (long)ctx->sm_count * 16) grid = (long)ctx->sm_count * 16;
    im2col_kernel<T><<<(int)grid, 256, 0, ctx->stream>>>(patches, images, b, img, patch);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}

template <typename T>
int op_im2col_u8(vitrs_ctx* ctx, T* patches, const uint8_t* images, int layout, const float* mean, const float* stdev, int b, int img,
                 int patch) {
    if (b <= 0) return VITRS_OK;
    VITRS_ARG(ctx, patch > 0 && img % patch == 0 && patch % 4 == 0 && (layout == 0 || layout == 1));
    VITRS_ARG(ctx, layout == 1 || ((uintptr_t)images & 3) == 0);
    NormParams nrm;
    for (int ch = 0; ch < 3; ++ch) {
        VITRS_ARG(ctx, stdev[ch] > 0.f);
        nrm.scale[ch] = 1.0f / (255.0f * stdev[ch]);
        nrm.shift[ch] = -mean[ch] / stdev[ch];
    }
    const int g = img / patch, t = g * g + 1, kdim = 3 * patch * patch;
    const long total = (long)b * t * (kdim / 4);
    long grid = (total + 255) / 256;
    if (grid > (long)ctx->sm_count * 16) grid = (long)ctx->sm_count * 16;
    im2col_u8_kernel<T><<<(int)grid, 256, 0, ctx->stream>>>(patches, images, layout, nrm, b, img, patch);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}

template <typename T>
int op_patch_backward_reduce(vitrs_ctx* ctx, float* dwpe, float* dcls, float* dpatchb, const T* denc, int b, int t, int c) {
    if (b <= 0) return VITRS_OK;
    const int gx = ceil_div((long)t * c, 256);
    int gy = (4 * ctx->sm_count + gx - 1) / gx;
    if (gy > b) gy = b;
    if (gy < 1) gy = 1;
    patch_bwd_reduce_kernel<T><<<dim3(gx, gy), 256, 0, ctx->stream>>>(dwpe, dcls, dpatchb, denc, b, t, c);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}

int op_encoder_forward(vitrs_ctx* ctx, float* enc, const int* inputs, const float* wte, const float* wpe, int b, int t, int c) {
    const long n = (long)b * t * c;
    if (n <= 0) return VITRS_OK;
    encoder_fwd_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(enc, inputs, wte, wpe, b, t, c);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}

int op_encoder_backward(vitrs_ctx* ctx, float* dwte, float* dwpe, const float* denc, const int* inputs, int b, int t, int c) {
    const long n = (long)b * t * c;
    if (n <= 0) return VITRS_OK;
    encoder_bwd_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(dwte, dwpe, denc, inputs, b, t, c);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}

template int op_im2col<float>(vitrs_ctx*, float*, const float*, int, int, int);
template int op_im2col<bf16>(vitrs_ctx*, bf16*, const float*, int, int, int);
template int op_im2col_u8<float>(vitrs_ctx*, float*, const uint8_t*, int, const float*, const float*, int, int, int);
template int op_im2col_u8<bf16>(vitrs_ctx*, bf16*, const uint8_t*, int, const float*, const float*, int, int, int);
template int op_patch_backward_reduce<float>(vitrs_ctx*, float*, float*, float*, const float*, int, int, int);
template int op_patch_backward_reduce<bf16>(vitrs_ctx*, float*, float*, float*, const bf16*, int, int, int);
