// Microbenchmark: sustained cycles per tcgen05.mma (cta_group::1, M = 128, K = 16, bf16) for the shapes the attention kernels
// issue — what a stream of small MMAs costs when nothing else runs on the SM.  Build + run: scripts/exp_mma_rate.sh
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../vit.rs_b200/csrc/tc_ptx.cuh"

__device__ __forceinline__ uint32_t idesc_of(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// mode 0: SS, A and B K-major   1: SS, A and B MN-major   2: TS (A from TMEM), B MN-major   3: SS K-major, same A descriptor re-used but
// B walks over 4 tiles (as mode 0; placeholder)   `busy` > 0: the other three warps hammer shared memory with loads meanwhile
template <int NACC, int STYLE>
__global__ void __launch_bounds__(256, 1) mma_rate(int mode, int N, int reps, int busy, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sA = base, sB = base + 4 * 16384, bar = base + 8 * 16384 + 64;
    volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(gen + 8 * 16384);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    for (int i = threadIdx.x; i < 8 * 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gen)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *slot, 0);
    if (warp == 0) {
        const uint32_t idesc = idesc_of(128, N, mode == 1, mode >= 1);
        const uint64_t dA = mode == 1 ? make_desc(sA, 16384, 1024) : make_desc(sA, 0, 1024);
        const uint64_t dB = mode >= 1 ? make_desc(sB, 16384, 1024) : make_desc(sB, 0, 1024);
        const long long t0 = clock64();
        // STYLE 0: one elect per MMA (converged loop)   1: one elected region around the whole loop   2: lane 0 by threadIdx
        if (STYLE == 0) {
            for (int r = 0; r < reps; ++r) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t acc = tmem + 256 + (uint32_t)((k % NACC) * 64);
                    if (elect_one()) {
                        if (mode == 2) umma_bf16_ts(acc, tmem + 8 * k, dB + 128 * k, idesc, 1);
                        else umma_bf16(acc, dA + 2 * k, dB + 2 * k, idesc, 1);
                    }
                }
            }
        } else if (STYLE == 1 ? elect_one() : (threadIdx.x == 0)) {
            for (int r = 0; r < reps; ++r) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t acc = tmem + 256 + (uint32_t)((k % NACC) * 64);
                    if (mode == 2) umma_bf16_ts(acc, tmem + 8 * k, dB + 128 * k, idesc, 1);
                    else umma_bf16(acc, dA + 2 * k, dB + 2 * k, idesc, 1);
                }
            }
        }
        __syncwarp();
        const long long t1 = clock64();
        if (elect_one()) umma_commit(bar);
        mbar_wait(bar, 0);
        const long long t2 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    } else if (busy && warp >= 4) {
        // shared-memory traffic from SIMT warps: 16-byte loads over the operand tiles
        uint32_t acc = 0;
        const uint32_t lane_off = (threadIdx.x & 127) * 16;
        for (int it = 0; it < busy; ++it) {
            uint32_t a, b, c, d;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(sA + ((lane_off + it * 2048) & 65535)));
            acc += a + b + c + d;
        }
        if (acc == 0x12345678) out[2] = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

template <int NACC, int STYLE>
void run(int mode, int N, long long* d_out, size_t smem) {
    cudaFuncSetAttribute(mma_rate<NACC, STYLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int reps = 500;
    long long h[2] = {0, 0};
    for (int rep = 0; rep < 2; ++rep) {
        mma_rate<NACC, STYLE><<<148, 256, smem>>>(mode, N, reps, 0, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    }
    cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
    printf("mode %d (0 SS K-major, 2 TS) N=%3d nacc=%d style=%d : issue %6.1f clk/MMA, complete %6.1f clk/MMA (tensor floor %d)\n", mode, N, NACC, STYLE,
           (double)h[0] / (4 * reps), (double)h[1] / (4 * reps), 128 * N / 256);
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 64);
    const size_t smem = 8 * 16384 + 2048;
    for (int mode = 0; mode < 3; mode += 2)
        for (int N : {16, 64, 128, 256}) {
            run<1, 0>(mode, N, d_out, smem);
            run<1, 1>(mode, N, d_out, smem);
            run<1, 2>(mode, N, d_out, smem);
            if (N <= 64) { run<2, 0>(mode, N, d_out, smem); run<4, 1>(mode, N, d_out, smem); }
        }
    return 0;
}
