// gemm_tc.cu — the bf16 tensor-core GEMM: TMA -> shared-memory ring -> tcgen05.mma -> TMEM ->
// tcgen05.ld epilogue, persistent over the 148 SMs, hand-written for sm_100a.
//
//   D[m,n] = sum_k A(m,k) * B(n,k)      bf16 operands, fp32 accumulation in tensor memory
//
// One kernel covers the three operand layouts the training step needs
//   forward   Y  = X . W^T        A K-major,  B K-major    (matmul_forward, train_vit.rs:384)
//   dinp      dX = dY . W         A K-major,  B MN-major   (matmul_backward pass 1, :532-541)
//   dweight   dW = dY^T . X       A MN-major, B MN-major   (matmul_backward pass 2, :543-555)
// "MN-major" = the contraction index is the slow index in memory; the 128B-swizzled
// shared-memory tile is then [k rows][64 mn elements] per TMA box and the UMMA descriptor
// carries the transpose, so no operand is ever transposed in HBM.
//
// Tiles are 256 x 256 and belong to a CTA PAIR (cluster of two = the two SMs of a TPC, tcgen05 cta_group::2): each CTA
// stages its own 128 rows of A and 128 of the 256 rows of B, one thread of the leader CTA issues the 256 x 256 x 16 MMAs for
// both, each CTA holds and drains its own 128 accumulator rows.  Per CTA the tensor core then reads a third less shared
// memory per flop than with 128 x 256 single-CTA tiles (kept for M <= 128 or N <= 128): +11 % on the K = 768 shapes.
//
// CTA = 12 warps: warp 0 lane 0 issues TMA, warp 1 lane 0 issues tcgen05.mma and commits,
// warp 2 owns the TMEM allocation, warps 4-11 are the epilogue (two warps per TMEM lane quarter,
// each taking half of the tile's 64-column chunks).  Two accumulators live in TMEM (2 x BN
// columns) so tile i's epilogue overlaps tile i+1's main loop.  The epilogue moves whole
// [32 rows][64 columns] bf16 boxes with TMA in both directions (residual / pre-GELU in, result
// out) through a warp-private 128B-swizzled staging tile, so no thread computes a global address
// and edge tiles are clipped by the tensor map.  dweight runs split-K with vectorised fp32 reductions (red.global.add.v4.f32),
// which is exactly the += contract of the reference's backward ops.
#include <stdlib.h>

#include "epilogue.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int kThreads = 384;
constexpr int kEpilogueWarp0 = 4;
constexpr int kEpilogueWarps = 8;
constexpr int kColsumWarps = 4;  // epilogue warps that also sum the A tiles (fused bias gradient)
constexpr int kSchedSlots = 4;   // depth of the per-CTA tile queue (the producer runs up to two tiles ahead of the epilogue)

// Tile scheduler.  Work units are handed out by the TMA producer thread of the cluster's leader CTA through a small queue in
// the shared memory of every CTA of the cluster: the unit number travels with st.async and completes the transaction count of
// that CTA's `full` barrier, every other role (the peer's producer, the MMA issuer, the epilogue warps of both CTAs) arrives
// on the leader's `empty` barrier once it has read the slot.  With a global counter (p.sched) the units go to whichever cluster
// asks first, so a cluster that starts late — its SMs were held by a collective's CTAs when the grid launched — simply takes
// fewer tiles instead of becoming the tail of the whole launch (measured at 2 ranks: GEMMs that overlap an NCCL kernel ran
// +40 % with the static stride).  Without a counter the same queue carries the static stride.
__device__ __forceinline__ void sched_publish(uint32_t full_cluster_addr, uint32_t slot_cluster_addr, uint32_t unit) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], 4;" ::"r"(full_cluster_addr) : "memory");
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];" ::"r"(slot_cluster_addr), "r"(unit),
                 "r"(full_cluster_addr)
                 : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

struct TcParams {
    int M, N, K;
    int m_tiles, n_tiles, kb_total, kb_per_split, splits;
    float* a_colsum;  // MN-major A only: a_colsum[m] += sum_k A(m, k)  (the bias gradient of matmul_backward, tv:548-550)
    unsigned int* sched;  // [0] next unit, [1] clusters that have drained the queue (dynamic tile scheduler); null = static stride
    Epilogue epi;
};

// CG = CTAs per tile: 1, or 2 = a CTA pair (cta_group::2) that computes a [256 x BN] tile with each CTA staging its own 128
// rows of A and BN/2 rows of B — per CTA the tensor core then reads a third less shared memory per flop
template <int BN, int STAGES, int CG>
struct SmemLayout {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN / CG * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGING_OFFSET = STAGES * STAGE_BYTES;  // 8 epilogue warps x [32 rows][128 B]
    static constexpr int STAGING_BYTES = kEpilogueWarps * 32 * 128;
    static constexpr int BAR_OFFSET = STAGING_OFFSET + STAGING_BYTES;
    static constexpr int SCHED_OFFSET = BAR_OFFSET + 256;  // tile queue: kSchedSlots x {full, empty} barriers + kSchedSlots unit numbers
    static constexpr int TOTAL = BAR_OFFSET + 512 + 1024;  // barriers + tile queue + alignment slack
};

// ---- epilogue: one warp, its 32 accumulator rows, one 64-column chunk ------------------------------
// The accumulator row of a thread (TMEM lane) is contiguous along n, but a warp's 32 rows are ld
// elements apart in global memory.  Both directions therefore go through a warp-private
// [32 rows][128 B] shared tile in the TMA 128B-swizzle layout: TMA moves the box, threads touch only
// their own row (16-byte chunks xor-swizzled by row -> conflict-free).
__device__ __forceinline__ uint32_t stage_addr(uint32_t stage, int row, int c8) { return stage + row * 128 + ((c8 ^ (row & 7)) << 4); }

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// this thread's row (= lane) of the staging tile, 8 bf16 at a time
__device__ __forceinline__ void stage_read8(uint32_t stage, int lane, int c8, float (&f)[8]) {
    uint32_t x[4];
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]) : "r"(stage_addr(stage, lane, c8)) : "memory");
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        f[2 * j] = __uint_as_float(x[j] << 16);
        f[2 * j + 1] = __uint_as_float(x[j] & 0xFFFF0000u);
    }
}
__device__ __forceinline__ void stage_write_row(uint32_t stage, int lane, const float (&f)[64]) {
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
        uint32_t x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 v = __floats2bfloat162_rn(f[c8 * 8 + 2 * j], f[c8 * 8 + 2 * j + 1]);
            x[j] = *reinterpret_cast<uint32_t*>(&v);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_addr(stage, lane, c8)), "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]) : "memory");
    }
}

struct EpiMaps {
    const CUtensorMap* out;
    const CUtensorMap* out2;
    const CUtensorMap* aux;
};

// box [m_base, +32) x [n_base, +64) of `map` -> staging tile; every lane returns once the bytes have landed
__device__ __forceinline__ void stage_fetch(const CUtensorMap* map, uint32_t stage, uint32_t bar, uint32_t& phase, int n_base, int m_base,
                                            int lane) {
    if (lane == 0) {
        mbar_expect_tx(bar, 32 * 128);
        tma_load_2d(stage, map, bar, n_base, m_base);
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
}
// staging tile (fully written by the warp) -> box of `map`
__device__ __forceinline__ void stage_flush(const CUtensorMap* map, uint32_t stage, int n_base, int m_base, int lane) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) tma_store_2d(map, stage, n_base, m_base);
}

// KIND is a compile-time constant here: the per-element code is straight-line (a runtime switch inside the
// element loop is if-converted by the compiler and every output then pays for every epilogue kind)
template <int KIND>
__device__ __forceinline__ void epilogue_chunk(const Epilogue& e, const EpiMaps& maps, uint32_t stage, uint32_t bar, uint32_t& bar_phase,
                                               uint32_t taddr, int m_base, int n_base, int M, int N, int lane, float2 bias2,
                                               bool aux_issued) {
    constexpr bool kAux = KIND == EPI_BIAS_RESIDUAL || KIND == EPI_GELU_BWD || KIND == EPI_ROWDOT;
    // the staging tile is reusable once the previous TMA store has finished reading it; the residual / pre-GELU box is
    // requested before the accumulator is read so that its latency overlaps the TMEM loads and the bias adds (the first
    // chunk of a tile was requested before the tile's main loop finished: aux_issued)
    if (kAux && !aux_issued) {
        if (lane == 0) {
            bulk_wait_read0();
            mbar_expect_tx(bar, 32 * 128);
            tma_load_2d(stage, maps.aux, bar, n_base, m_base);
        }
        __syncwarp();
    }
    float acc[64];
    {
        uint32_t v[32];
        tmem_ld32(taddr, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]);
        tmem_ld32(taddr + 32, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[32 + j] = __uint_as_float(v[j]);
    }
    if (KIND == EPI_BIAS || KIND == EPI_BIAS_GELU || KIND == EPI_BIAS_GELU_ONLY || KIND == EPI_BIAS_RESIDUAL) {
        // bias[n_base + 2*lane .. +1] was fetched by this lane before the tile's accumulator was ready (this kernel leaves the
        // L1 no capacity, a global load here costs an L2 round trip per chunk); every lane needs all 64 values: shuffle broadcast
#pragma unroll
        for (int j = 0; j < 64; j += 2) {
            acc[j] += __shfl_sync(0xffffffffu, bias2.x, j >> 1);
            acc[j + 1] += __shfl_sync(0xffffffffu, bias2.y, j >> 1);
        }
    }
    if (KIND == EPI_BIAS_GELU_ONLY) {
#pragma unroll
        for (int j = 0; j < 64; j += 2) {
            const __nv_bfloat162 r2 = __floats2bfloat162_rn(acc[j], acc[j + 1]);
            const uint32_t rw = *reinterpret_cast<const uint32_t*>(&r2);
            const float2 g = gelu_fwd2(make_float2(__uint_as_float(rw << 16), __uint_as_float(rw & 0xFFFF0000u)));
            acc[j] = g.x;
            acc[j + 1] = g.y;
        }
    }
    float rowdot = 0.f;
    if (kAux) {
        mbar_wait(bar, bar_phase);
        bar_phase ^= 1u;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
            float a[8];
            stage_read8(stage, lane, c8, a);
            if (KIND == EPI_BIAS_RESIDUAL) {
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[c8 * 8 + j] += a[j];
            } else if (KIND == EPI_ROWDOT) {
                // the chunk is one head's 64 columns of dO: D = sum dO * O with dO as it is stored (rounded to bf16)
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const __nv_bfloat162 r2 = __floats2bfloat162_rn(acc[c8 * 8 + j], acc[c8 * 8 + j + 1]);
                    const uint32_t rw = *reinterpret_cast<const uint32_t*>(&r2);
                    rowdot = fmaf(__uint_as_float(rw << 16), a[j], rowdot);
                    rowdot = fmaf(__uint_as_float(rw & 0xFFFF0000u), a[j + 1], rowdot);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; j += 2) {  // packed f32x2 math: this epilogue is instruction-issue-bound
                    const float2 g = mul2(make_float2(acc[c8 * 8 + j], acc[c8 * 8 + j + 1]), gelu_grad2(make_float2(a[j], a[j + 1])));
                    acc[c8 * 8 + j] = g.x;
                    acc[c8 * 8 + j + 1] = g.y;
                }
            }
        }
        __syncwarp();  // every lane has read its aux row before the tile is overwritten
        if (KIND == EPI_ROWDOT) {
            const long m = (long)m_base + lane;
            if (m < M) {
                const long img = m / e.np;
                reinterpret_cast<float*>(e.out2)[(img * (N >> 6) + (n_base >> 6)) * e.np + (m - img * e.np)] = rowdot;
            }
        }
    }
    if (KIND == EPI_PATCH) {
        const long m = (long)m_base + lane;
        const int tok = (int)(m % e.np);
        if (n_base + 64 <= N && ((e.ldo | n_base) & 3) == 0 &&
            ((reinterpret_cast<uintptr_t>(e.pos) | reinterpret_cast<uintptr_t>(e.cls) | reinterpret_cast<uintptr_t>(e.bias)) & 15) == 0) {
            // this thread's 64 columns of its position row and of the bias (the class token for token 0, whose im2col row is
            // zero) by 16-byte loads; same order of additions as the scalar form below: (acc + bias) + pos, cls + pos.
            // (One scalar load per operand and element made this GEMM 9x slower than the same shape with a plain bias.)
            const float4* prow = reinterpret_cast<const float4*>(e.pos + (long)tok * e.ldo + n_base);
            const float4* brow = tok == 0 ? reinterpret_cast<const float4*>(e.cls + n_base)
                                          : (e.bias ? reinterpret_cast<const float4*>(e.bias + n_base) : nullptr);
#pragma unroll
            for (int j4 = 0; j4 < 16; ++j4) {
                if ((j4 & 3) == 0) asm volatile("" ::: "memory");  // at most 8 loads in flight: hoisting all 32 would spill
                const float4 pz = __ldg(prow + j4);
                const float4 bz = brow ? __ldg(brow + j4) : make_float4(0.f, 0.f, 0.f, 0.f);
                acc[j4 * 4 + 0] = ((tok == 0 ? 0.f : acc[j4 * 4 + 0]) + bz.x) + pz.x;
                acc[j4 * 4 + 1] = ((tok == 0 ? 0.f : acc[j4 * 4 + 1]) + bz.y) + pz.y;
                acc[j4 * 4 + 2] = ((tok == 0 ? 0.f : acc[j4 * 4 + 2]) + bz.z) + pz.z;
                acc[j4 * 4 + 3] = ((tok == 0 ? 0.f : acc[j4 * 4 + 3]) + bz.w) + pz.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                const int n = n_base + j;
                if (m < M && n < N) {
                    float v = tok == 0 ? __ldg(e.cls + n) : acc[j] + (e.bias ? __ldg(e.bias + n) : 0.f);
                    acc[j] = v + __ldg(e.pos + (long)tok * e.ldo + n);
                }
            }
        }
    }
    if (e.accumulate) {  // the reference's `+=` contract at the ABI; the fused model path never takes it
        if (!kAux) {
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
        }
        stage_fetch(maps.out, stage, bar, bar_phase, n_base, m_base, lane);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
            float a[8];
            stage_read8(stage, lane, c8, a);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[c8 * 8 + j] += a[j];
        }
        __syncwarp();
    }
    if (!kAux && !e.accumulate) {  // (the aux / accumulate paths already waited before fetching their box)
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
    }
    stage_write_row(stage, lane, acc);
    stage_flush(maps.out, stage, n_base, m_base, lane);
    if (KIND == EPI_BIAS_GELU) {
        // gelu_forward consumes the stored (bf16-rounded) pre-activation, as the unfused op would
#pragma unroll
        for (int j = 0; j < 64; j += 2) {
            // (rounded through the packed conversion: the scalar cvt shares the quarter-rate unit with tanh)
            const __nv_bfloat162 r2 = __floats2bfloat162_rn(acc[j], acc[j + 1]);
            const uint32_t rw = *reinterpret_cast<const uint32_t*>(&r2);
            const float2 g = gelu_fwd2(make_float2(__uint_as_float(rw << 16), __uint_as_float(rw & 0xFFFF0000u)));
            acc[j] = g.x;
            acc[j + 1] = g.y;
        }
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
        stage_write_row(stage, lane, acc);
        stage_flush(maps.out2, stage, n_base, m_base, lane);
    }
}

// PATCH: the instantiation that carries the patch-embedding epilogue (and only that one), so that its 16-byte table loads do not
// weigh on the register allocation of the kernels the training step spends its time in
template <int BN, int STAGES, bool A_MN, bool B_MN, int CG, bool PATCH = false>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
               const __grid_constant__ CUtensorMap tmOut2, const __grid_constant__ CUtensorMap tmAux, const TcParams p) {
    using L = SmemLayout<BN, STAGES, CG>;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + L::BAR_OFFSET;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto mdone_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };  // CTA pair + column sums only: MMAs of stage s retired
    auto tfull_bar = [&](int a) { return bar_base + 8u * (3 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (3 * STAGES + 2 + a); };
    auto aux_bar = [&](int w) { return bar_base + 8u * (3 * STAGES + 4 + w); };  // one per epilogue warp
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + L::BAR_OFFSET + 8 * (3 * STAGES + 4 + kEpilogueWarps));
    static_assert(8 * (3 * STAGES + 4 + kEpilogueWarps) + 4 <= 256, "barrier block");
    const uint32_t sched_base = smem_base + L::SCHED_OFFSET;
    auto sfull_bar = [&](int s) { return sched_base + 8u * s; };
    auto sempty_bar = [&](int s) { return sched_base + 8u * (kSchedSlots + s); };  // used in the leader CTA only
    auto sunit_slot = [&](int s) { return sched_base + 16u * kSchedSlots + 4u * s; };
    // CTA pair: rank 0 (the leader) issues the MMAs and owns the full / tempty barriers; both CTAs load, both run epilogues
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    const bool colsum = A_MN && p.a_colsum != nullptr;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t TMEM_COLS = 2 * BN;  // two accumulators; power of two (256 or 512)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmOut);
        tma_prefetch_desc(&tmOut2);
        tma_prefetch_desc(&tmAux);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            // freed by the MMA commit and, when the column sums of A ride along, by the four epilogue warps that read the tile
            // (in a CTA pair the column-sum warps of the peer cannot see the leader's full barrier: there they read the tile after
            // its MMAs have retired — mdone — and the slot is freed by them alone)
            mbar_init(empty_bar(s), colsum ? (CG == 2 ? kColsumWarps : 1 + kColsumWarps) : 1);
            mbar_init(mdone_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), CG * kEpilogueWarps);  // one arrive per epilogue warp (of both CTAs of a pair)
        }
        for (int w = 0; w < kEpilogueWarps; ++w) mbar_init(aux_bar(w), 1);
        for (int q = 0; q < kSchedSlots; ++q) {
            mbar_init(sfull_bar(q), 1);
            // consumers: 8 epilogue warps per CTA, the MMA issuer of the leader, the TMA producer of the peer
            mbar_init(sempty_bar(q), CG == 2 ? 2 * kEpilogueWarps + 2 : kEpilogueWarps + 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                         "r"(TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                         "r"(TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all();  // the peer's barriers are initialised before anything is signalled across the pair
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles = p.m_tiles * p.n_tiles;  // tiles of CG*128 rows
    const int total_units = tiles * p.splits;
    const int unit0 = blockIdx.x / CG, unit_step = gridDim.x / CG;
    // consumer side of the tile queue: the it-th unit of this CTA (>= total_units: the queue is drained)
    const uint32_t sempty_leader0 = CG == 2 ? mapa_shared(sempty_bar(0), 0) : sempty_bar(0);
    auto next_unit = [&](int it) -> int {
        const int q = it % kSchedSlots;
        mbar_wait(sfull_bar(q), (uint32_t)((it / kSchedSlots) & 1));
        const int unit = (int)ld_shared_u32(sunit_slot(q));
        if (CG == 2) mbar_arrive_cluster(sempty_leader0 + 8u * q);
        else mbar_arrive(sempty_bar(q));
        return unit;
    };
    constexpr int TM = BM * CG;
    const int m_rank = (int)rank * BM, n_rank = (int)rank * (BN / CG);  // this CTA's share of the operand tiles

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        int stage = 0;
        uint32_t phase = 0;
        auto load = [&](uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
            if (CG == 2) tma_load_2d_pair(dst, map, bar, c0, c1);
            else tma_load_2d(dst, map, bar, c0, c1);
        };
        // The leader's producer is the cluster's tile scheduler.  With many units per cluster it asks for unit i + 1 when it starts
        // loading unit i, so the counter's round trip hides under the loads (asking only when unit i is loaded cost the forward
        // GEMMs 16 %: this thread is the one that must never bubble).  With few units per cluster — the split-K weight-gradient
        // GEMMs have about one — any look-ahead would hoard: the clusters that start first would take two units each and leave
        // the others idle; there it asks on demand.
        const bool lookahead = total_units >= 4 * unit_step;
        auto fetch = [&](int it) -> int { return p.sched ? (int)atomicAdd(p.sched, 1u) : unit0 + it * unit_step; };
        int prefetched = (rank == 0 && lookahead) ? fetch(0) : 0;
        for (int it = 0;; ++it) {
            int unit;
            if (rank == 0) {
                const int q = it % kSchedSlots;
                if (it >= kSchedSlots) mbar_wait(sempty_bar(q), (uint32_t)(((it / kSchedSlots) - 1) & 1));
                unit = lookahead ? prefetched : fetch(it);
                if (lookahead && unit < total_units) prefetched = fetch(it + 1);
#pragma unroll
                for (int r = 0; r < CG; ++r)
                    sched_publish(CG == 2 ? mapa_shared(sfull_bar(q), (uint32_t)r) : sfull_bar(q),
                                  CG == 2 ? mapa_shared(sunit_slot(q), (uint32_t)r) : sunit_slot(q), (uint32_t)unit);
                if (unit >= total_units) {
                    // this cluster is done asking; the last cluster to get here re-arms the counter for the next launch
                    if (p.sched && atomicAdd(p.sched + 1, 1u) == (unsigned int)unit_step - 1u) {
                        p.sched[1] = 0u;
                        __threadfence();
                        p.sched[0] = 0u;
                    }
                    break;
                }
            } else {
                unit = next_unit(it);
                if (unit >= total_units) break;
            }
            // split-major: consecutive units (= concurrently running CTAs) walk the same K range, so every
            // operand slab is fetched from HBM once and shared through L2
            const int split = unit / tiles, tile = unit - split * tiles;
            const int m0 = (tile / p.n_tiles) * TM + m_rank, n0 = (tile % p.n_tiles) * BN + n_rank;
            const int kb0 = split * p.kb_per_split;
            const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                // pair: both CTAs' bytes are counted on the leader's barrier, which the leader arms for both
                if (rank == 0) mbar_expect_tx(full_bar(stage), CG * L::STAGE_BYTES);
                const uint32_t fb = CG == 2 ? mapa_shared(full_bar(stage), 0) : full_bar(stage);
                const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
                const uint32_t sb = sa + L::A_BYTES;
                const int k0 = kb * BK;
                if (!A_MN) {
                    load(sa, &tmA, fb, k0, m0);  // box {64 k, 128 m}
                } else {
#pragma unroll
                    for (int j = 0; j < BM / 64; ++j)  // boxes {64 m, 64 k}
                        load(sa + j * (BK * 128), &tmA, fb, m0 + j * 64, k0);
                }
                if (!B_MN) {
                    load(sb, &tmB, fb, k0, n0);  // box {64 k, BN/CG n}
                } else {
#pragma unroll
                    for (int j = 0; j < BN / CG / 64; ++j)
                        load(sb + j * (BK * 128), &tmB, fb, n0 + j * 64, k0);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1 && lane == 0 && rank == 0) {
        // ===== MMA issuer (single thread; of the leader CTA in a pair) =====
        // instruction descriptor: D fp32 [4,6)=1, A bf16 [7,10)=1, B bf16 [10,13)=1,
        // A major bit 15, B major bit 16 (1 = MN-major), N>>3 [17,23), M>>4 [24,29)
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int it = 0;; ++it) {
            const int unit = next_unit(it);
            if (unit >= total_units) break;
            const int split = unit / tiles;
            const int kb0 = split * p.kb_per_split;
            const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
                const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    // K-major: 8-row groups 1024 B apart (SBO), +32 B per 16-element K step.
                    // MN-major: 64-element MN atoms BK*128 B apart (LBO), 8-k groups 1024 B apart
                    // (SBO), +2048 B per 16-element K step.
                    const uint64_t adesc = A_MN ? make_desc(sa + k * 2048, BK * 128, 1024) : make_desc(sa + k * 32, 0, 1024);
                    const uint64_t bdesc = B_MN ? make_desc(sb + k * 2048, BK * 128, 1024) : make_desc(sb + k * 32, 0, 1024);
                    if (CG == 2) umma_bf16_pair(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    else umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                }
                // frees the smem slot when these MMAs retire
                if (CG == 2) umma_commit_pair(colsum ? mdone_bar(stage) : empty_bar(stage));
                else umma_commit(empty_bar(stage));
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
            // accumulator complete -> epilogue (of both CTAs)
            if (CG == 2) umma_commit_pair(tfull_bar(acc));
            else umma_commit(tfull_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    } else if (warp >= kEpilogueWarp0) {
        // ===== epilogue: TMEM -> registers -> global =====
        const int ew = warp - kEpilogueWarp0;       // 0..7
        const int quarter = ew & 3;                 // == warp % 4: the TMEM lanes this warp may read
        const int chalf = ew >> 2;                  // which half of the tile's 64-column chunks
        constexpr int kChunksPerWarp = BN / 64 / 2;
        const uint32_t stage_smem = smem_base + L::STAGING_OFFSET + ew * (32 * 128);
        const uint32_t my_bar = aux_bar(ew);
        uint32_t bar_phase = 0;
        const EpiMaps maps = {&tmOut, &tmOut2, &tmAux};
        int acc = 0;
        uint32_t acc_phase = 0;
        int cs_stage = 0;        // column-sum consumer: walks the stage ring in step with the MMA warp
        uint32_t cs_phase = 0;
        const uint32_t tempty_remote0 = CG == 2 ? mapa_shared(tempty_bar(0), 0) : 0u;
        for (int it = 0;; ++it) {
            int unit = 0;
            if (lane == 0) unit = next_unit(it);
            unit = __shfl_sync(0xffffffffu, unit, 0);
            if (unit >= total_units) break;
            const int tile = unit % tiles;
            const int m0 = (tile / p.n_tiles) * TM + m_rank, n0 = (tile % p.n_tiles) * BN;
            if (colsum && (ew & 2)) {
                // dbias fused into the weight-gradient GEMM: the A tiles ([64 k][64 m] boxes of dout) pass through shared
                // memory anyway; four of the otherwise idle epilogue warps (those on the schedulers that host neither the TMA
                // nor the MMA warp) sum them over k.  The n_tiles units that see the same A tile share the work: the unit with
                // n-tile t sums the 4-row groups rg with rg % n_tiles == t.
                const int split = unit / tiles;
                const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                const int box = ew & 1, c8 = lane & 7;           // 64-column box, 16-byte chunk
                const int rg0 = (ew >> 2) * 8 + (lane >> 3) * 2;  // this thread's two 4-row groups of the 64 k-rows
                const int nt_ = tile % p.n_tiles;
                bool own[2];
#pragma unroll
                for (int g2 = 0; g2 < 2; ++g2) own[g2] = p.n_tiles <= 16 ? ((rg0 + g2) % p.n_tiles) == nt_ : (rg0 + g2) == nt_;
                float part[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) part[i] = 0.f;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(CG == 2 ? mdone_bar(cs_stage) : full_bar(cs_stage), cs_phase);
                    const uint32_t abox = smem_base + cs_stage * L::STAGE_BYTES + box * (BK * 128);
#pragma unroll
                    for (int g2 = 0; g2 < 2; ++g2) {
                        if (!own[g2]) continue;
#pragma unroll
                        for (int rr = 0; rr < 4; ++rr) {
                            const int row = (rg0 + g2) * 4 + rr;
                            uint32_t x[4];
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3])
                                         : "r"(abox + row * 128 + ((c8 ^ (row & 7)) << 4)) : "memory");
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                part[2 * j] += __uint_as_float(x[j] << 16);
                                part[2 * j + 1] += __uint_as_float(x[j] & 0xFFFF0000u);
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty_bar(cs_stage));
                    if (++cs_stage == STAGES) { cs_stage = 0; cs_phase ^= 1u; }
                }
                // 4 row pairs -> one sum per column (lanes with equal c8 hold the same 8 columns)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    part[i] += __shfl_xor_sync(0xffffffffu, part[i], 8);
                    part[i] += __shfl_xor_sync(0xffffffffu, part[i], 16);
                }
                if ((lane >> 3) == 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int col = m0 + box * 64 + c8 * 8 + i;
                        if (col < p.M) atomicAdd(p.a_colsum + col, part[i]);
                    }
                }
            }
            const int m_base = m0 + quarter * 32;
            const long m = (long)m_base + lane;
            const bool has_aux = p.epi.kind == EPI_BIAS_RESIDUAL || p.epi.kind == EPI_GELU_BWD || p.epi.kind == EPI_ROWDOT;
            const int nb_first = n0 + chalf * kChunksPerWarp * 64;
            const bool first_live = nb_first < p.N && m_base < p.M;
            float2 bias2[kChunksPerWarp];  // this lane's two bias values of each of the warp's chunks, requested ahead of the accumulator
#pragma unroll
            for (int c = 0; c < kChunksPerWarp; ++c) {
                const int n = nb_first + c * 64 + 2 * lane;
                bias2[c] = make_float2(0.f, 0.f);
                if (p.epi.bias && p.epi.kind != EPI_PATCH && p.epi.kind != EPI_ACCUM_F32) {
                    if (n < p.N) bias2[c].x = __ldg(p.epi.bias + n);
                    if (n + 1 < p.N) bias2[c].y = __ldg(p.epi.bias + n + 1);
                }
            }
            if (has_aux && first_live) {  // this tile's first residual / pre-GELU box, while the main loop is still running
                if (lane == 0) {
                    bulk_wait_read0();
                    mbar_expect_tx(my_bar, 32 * 128);
                    tma_load_2d(stage_smem, &tmAux, my_bar, nb_first, m_base);
                }
                __syncwarp();
            }
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
            if (p.epi.kind == EPI_ACCUM_F32) {
                // dweight: fp32 vector reductions straight from the accumulator rows (split-K partial sums)
#pragma unroll 1
                for (int ch = chalf * (BN / 64); ch < (chalf + 1) * (BN / 64); ++ch) {
                    uint32_t v[32];
                    tmem_ld32(trow + ch * 32, v);  // warp-collective: no per-lane predicate around it
                    const int nb = n0 + ch * 32;
                    if (m < p.M && nb < p.N) {
                        float* dst = reinterpret_cast<float*>(p.epi.out) + m * p.epi.ldo + nb;
#pragma unroll
                        for (int g = 0; g < 8; ++g)
                            if (nb + g * 4 < p.N)
                                red_add_v4(dst + g * 4, __uint_as_float(v[g * 4]), __uint_as_float(v[g * 4 + 1]),
                                           __uint_as_float(v[g * 4 + 2]), __uint_as_float(v[g * 4 + 3]));
                    }
                }
            } else {
#pragma unroll 1
                for (int ch = chalf * kChunksPerWarp; ch < (chalf + 1) * kChunksPerWarp; ++ch) {
                    const uint32_t taddr = trow + ch * 64;
                    const int nb = n0 + ch * 64;
                    if (nb >= p.N || m_base >= p.M) continue;  // warp-uniform: nothing of this chunk is inside the matrix
                    const bool pre = has_aux && nb == nb_first;
                    float2 b2 = bias2[0];  // static indices only: a runtime index would put the array in local memory
#pragma unroll
                    for (int c = 1; c < kChunksPerWarp; ++c)
                        if (ch - chalf * kChunksPerWarp == c) b2 = bias2[c];
                    if (PATCH) {
                        epilogue_chunk<EPI_PATCH>(p.epi, maps, stage_smem, my_bar, bar_phase, taddr, m_base, nb, p.M, p.N, lane, b2, pre);
                    } else switch (p.epi.kind) {  // warp-uniform branch to straight-line per-kind code
                        case EPI_BIAS: epilogue_chunk<EPI_BIAS>(p.epi, maps, stage_smem, my_bar, bar_phase, taddr, m_base, nb, p.M, p.N, lane, b2, pre); break;
                        case EPI_BIAS_GELU: epilogue_chunk<EPI_BIAS_GELU>(p.epi, maps, stage_smem, my_bar, bar_phase, taddr, m_base, nb, p.M, p.N, lane, b2, pre); break;
                        case EPI_BIAS_RESIDUAL: epilogue_chunk<EPI_BIAS_RESIDUAL>(p.epi, maps, stage_smem, my_bar, bar_phase, taddr, m_base, nb, p.M, p.N, lane, b2, pre); break;
                        case EPI_GELU_BWD: epilogue_chunk<EPI_GELU_BWD>(p.epi, maps, stage_smem, my_bar, bar_phase, taddr, m_base, nb, p.M, p.N, lane, b2, pre); break;
                        case EPI_BIAS_GELU_ONLY: epilogue_chunk<EPI_BIAS_GELU_ONLY>(p.epi, maps, stage_smem, my_bar, bar_phase, taddr, m_base, nb, p.M, p.N, lane, b2, pre); break;
                        case EPI_ROWDOT: epilogue_chunk<EPI_ROWDOT>(p.epi, maps, stage_smem, my_bar, bar_phase, taddr, m_base, nb, p.M, p.N, lane, b2, pre); break;
                        default: epilogue_chunk<EPI_NONE>(p.epi, maps, stage_smem, my_bar, bar_phase, taddr, m_base, nb, p.M, p.N, lane, b2, pre); break;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster(tempty_remote0 + 8u * acc);
                else mbar_arrive(tempty_bar(acc));
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (lane == 0) bulk_wait_all();  // the last TMA stores must be complete before the CTA retires its shared memory
    }

    tc_fence_before();
    if (CG == 2) cluster_sync_all();  // neither CTA retires shared / tensor memory the other may still signal or write
    else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- host side --------------------------------------------------------------------------------
int encode_map(vitrs_ctx* ctx, CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t outer_stride_elems,
               uint32_t box_inner, uint32_t box_outer) {
    const uint64_t dims[2] = {inner, outer};
    const uint64_t strides[1] = {outer_stride_elems * 2};
    const uint32_t box[2] = {box_inner, box_outer};
    return vitrs_tensor_map(ctx, map, 2, base, dims, strides, box);
}

template <int BN, int STAGES, bool A_MN, bool B_MN, int CG, bool PATCH = false>
int launch_tc(vitrs_ctx* ctx, const CUtensorMap* maps, const TcParams& p) {
    using L = SmemLayout<BN, STAGES, CG>;
    auto kern = gemm_tc_kernel<BN, STAGES, A_MN, B_MN, CG, PATCH>;
    VITRS_TRY(vitrs_func_smem(ctx, (const void*)kern, L::TOTAL));
    const int units = p.m_tiles * p.n_tiles * p.splits;
    const int slots = ctx->sm_count / CG;
    const int grid = CG * (units < slots ? units : slots);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = L::TOTAL;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;  // a pair = the two SMs of one TPC
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    vitrs_prof_before(ctx, 2.0 * p.M * p.N * p.K);
    VITRS_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], maps[4], p));
    vitrs_prof_after(ctx);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}

template <int BN, int STAGES, int CG>
int launch_tc_major(vitrs_ctx* ctx, bool a_mn, bool b_mn, const CUtensorMap* maps, const TcParams& p) {
    if (p.epi.kind == EPI_PATCH) {  // (patches and weights are both K-major: tc_eligible)
        if (a_mn || b_mn) return VITRS_ERR_UNSUPPORTED;
        return launch_tc<BN, STAGES, false, false, CG, true>(ctx, maps, p);
    }
    if (!a_mn && !b_mn) return launch_tc<BN, STAGES, false, false, CG>(ctx, maps, p);
    if (!a_mn && b_mn) return launch_tc<BN, STAGES, false, true, CG>(ctx, maps, p);
    if (a_mn && b_mn) return launch_tc<BN, STAGES, true, true, CG>(ctx, maps, p);
    return launch_tc<BN, STAGES, true, false, CG>(ctx, maps, p);
}

inline bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

// true when the tcgen05 kernel can take this problem as described (alignment / layout rules of
// TMA and of the vectorised epilogue); everything else is routed to the SIMT kernel by the caller
static bool tc_eligible(const GemmDesc& g) {
    const bool a_k = g.a_ks == 1, a_mn = g.a_rs == 1 && g.a_ks != 1;
    const bool b_k = g.b_ks == 1, b_mn = g.b_rs == 1 && g.b_ks != 1;
    if (!(a_k || a_mn) || !(b_k || b_mn)) return false;
    if (!al16(g.A) || !al16(g.B) || !al16(g.epi.out)) return false;
    if ((a_k ? g.a_rs : g.a_ks) % 8 || (b_k ? g.b_rs : g.b_ks) % 8) return false;
    if (a_mn && g.M % 8) return false;  // TMA inner extent bytes must be a multiple of 16
    if (b_mn && g.N % 8) return false;
    if (a_k && g.K % 8) return false;
    if (g.N % 8 || g.epi.ldo % 8) return false;
    if (g.epi.aux && !al16(g.epi.aux)) return false;
    if (g.epi.out2 && !al16(g.epi.out2)) return false;
    if (g.M < 1 || g.N < 8 || g.K < 8) return false;
    if (g.epi.kind == EPI_ROWDOT && (g.N % 64 || g.epi.np < 1 || g.M % g.epi.np || !g.epi.aux || !g.epi.out2)) return false;
    return true;
}

// Every host-side decision of one bf16 GEMM call as a pure function of the problem and the context's switches: kernel family,
// tile, CTAs per tile, ring depth, split-K factor and grid.  gemm_tc_bf16 launches what this returns; vitrs_gemm_plan exports
// it, so the routing of every BASELINE shape is pinned by tests that need no device (tests/test_plan.py).
struct TcSwitches {
    int sm_count;
    int no_small;   // VITRS_GEMM_NO_SMALL
    int cg1;        // VITRS_GEMM_CG=1
    int patch_tc;   // VITRS_GEMM_PATCH_TC
    int splits;     // VITRS_GEMM_SPLITS (0 = heuristic)
};
struct TcPlan {
    bool simt;  // the SIMT kernel takes the call (gemm_simt.cu)
    int BN, CG, stages;
    int m_tiles, n_tiles, kb_total, splits, kb_per_split, grid;
};

static TcPlan plan_tc(int M, int N, int K, int kind, bool a_mn, bool b_mn, const TcSwitches& sw) {
    TcPlan pl = {};
    if (kind == EPI_PATCH && (a_mn || b_mn)) {  // (only the K-major instantiation carries this epilogue)
        pl.simt = true;
        return pl;
    }
    int BN = N > 128 ? 256 : 128;
    // a CTA pair per [256 x 256] tile whenever the problem has that many rows
    int CG = (BN == 256 && M > BM) ? 2 : 1;
    // Small problems (the inference engine at batch 1-64: M = 197 .. 12 608 rows) leave most SMs without a tile at that size and
    // the launch is a stream of weights through a handful of CTAs: when the large tiles fill less than half of the chip, take the
    // smaller ones — [256 x 256] pairs -> [128 x 256] -> [128 x 128] single CTAs — as far as that adds CTAs.  (Not for the
    // split-K weight gradients, which fill the chip by splitting K.)
    if (kind != EPI_ACCUM_F32 && !sw.no_small) {
        const long slots = sw.sm_count;
        auto ctas = [&](int bn, int cg) { return (long)ceil_div(M, BM * cg) * ceil_div(N, bn) * cg; };
        if (CG == 2 && ctas(256, 2) < slots / 2) CG = 1;
        if (BN == 256 && CG == 1 && ctas(256, 1) < slots / 2) BN = 128;
    }
    if (sw.cg1) CG = 1;  // VITRS_GEMM_CG=1: tuning aid (scripts/bench_gemm.py)
    // The patch-embedding instantiation has run on a GPU with K = 768 on CTA pairs (ViT-B/16, S/16, Ti/16: tests and bench) and
    // with K = 192 on CTA pairs at test batch sizes (tests/test_gpu_parity_configs.py, a few dozen tiles) and on the [128 x 128]
    // single-CTA tiles.  The one bench run of ViT-B/8 at batch 256 (K = 192 on CTA pairs, 2 355 tiles) taken after it went in did
    // not finish within its time limit, with no GPU time left to find out why: until that combination has
    // been examined it takes the SIMT kernel, whose EPI_PATCH path every verify-mode test exercises.  VITRS_GEMM_PATCH_TC=1
    // lifts the restriction.
    if (kind == EPI_PATCH && BN == 256 && K < 768 && !sw.patch_tc) {
        pl.simt = true;
        return pl;
    }
    pl.BN = BN;
    pl.CG = CG;
    pl.stages = (CG == 1 && BN == 256) ? 4 : 6;
    pl.m_tiles = ceil_div(M, BM * CG);
    pl.n_tiles = ceil_div(N, BN);
    pl.kb_total = ceil_div(K, BK);
    int splits = 1;
    const int tiles = pl.m_tiles * pl.n_tiles;
    const int slots = sw.sm_count / CG;  // tiles in flight
    if (kind == EPI_ACCUM_F32 && tiles < slots) {
        // Split K so that one wave of CTAs covers the SMs: the smallest split count that fills >= 92 % of
        // whole waves, else the best fill.  More splits than needed cost fp32 reductions and, worse, break
        // the sharing of operand slabs between concurrently running CTAs (measured: 37 splits -> 6x DRAM traffic).
        const int max_s = pl.kb_total / 16 > 1 ? (pl.kb_total / 16 < 32 ? pl.kb_total / 16 : 32) : 1;
        double best = 0.0;
        for (int s = 1; s <= max_s; ++s) {
            const long units = (long)tiles * s;
            const long waves = (units + slots - 1) / slots;
            const double eff = (double)units / (double)(waves * slots);
            if (eff > best + 1e-9) { best = eff; splits = s; }
            if (eff >= 0.92) break;
        }
    }
    if (kind == EPI_ACCUM_F32 && sw.splits > 0) splits = sw.splits;  // VITRS_GEMM_SPLITS: tuning aid
    pl.kb_per_split = ceil_div(pl.kb_total, splits);
    pl.splits = ceil_div(pl.kb_total, pl.kb_per_split);
    const int units = tiles * pl.splits;
    pl.grid = CG * (units < slots ? units : slots);  // persistent: one CTA (pair) per SM (pair), fewer when the work is smaller
    return pl;
}

int gemm_tc_bf16(vitrs_ctx* ctx, const GemmDesc& g) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return VITRS_OK;
    if (!tc_eligible(g)) {
        if (g.a_colsum) VITRS_TRY(op_colsum<bf16>(ctx, g.a_colsum, reinterpret_cast<const bf16*>(g.A), g.K, g.M, g.a_ks));
        if (g.epi.kind == EPI_ROWDOT) {  // plain GEMM, then the row dots of every 64-column head slice as a separate pass
            GemmDesc plain = g;
            plain.epi.kind = EPI_NONE;
            plain.epi.aux = nullptr;
            plain.epi.out2 = nullptr;
            VITRS_TRY(gemm_simt_bf16(ctx, plain));
            if (g.N % 64 || g.epi.np < 1 || g.M % g.epi.np || g.epi.ldo != g.N)
                return vitrs_set_error(ctx, VITRS_ERR_ARG, "EPI_ROWDOT needs N %% 64 == 0, M %% tokens == 0 and a dense output");
            return op_attention_bwd_prep(ctx, reinterpret_cast<float*>(g.epi.out2), reinterpret_cast<const bf16*>(g.epi.out),
                                         reinterpret_cast<const bf16*>(g.epi.aux), g.M / g.epi.np, g.epi.np, g.N, g.N / 64);
        }
        return gemm_simt_bf16(ctx, g);
    }
    const bool a_mn = g.a_ks != 1, b_mn = g.b_ks != 1;
    const TcSwitches sw = {ctx->sm_count, ctx->env_gemm_no_small, ctx->env_gemm_cg1, ctx->env_gemm_patch_tc, ctx->env_gemm_splits};
    const TcPlan pl = plan_tc(g.M, g.N, g.K, g.epi.kind, a_mn, b_mn, sw);
    if (pl.simt) return gemm_simt_bf16(ctx, g);
    const int BN = pl.BN, CG = pl.CG;
    CUtensorMap maps[5];  // A, B, out, out2, aux
    CUtensorMap &tmA = maps[0], &tmB = maps[1];
    if (!a_mn) VITRS_TRY(encode_map(ctx, &tmA, g.A, g.K, g.M, g.a_rs, BK, BM));
    else VITRS_TRY(encode_map(ctx, &tmA, g.A, g.M, g.K, g.a_ks, 64, BK));
    if (!b_mn) VITRS_TRY(encode_map(ctx, &tmB, g.B, g.K, g.N, g.b_rs, BK, BN / CG));
    else VITRS_TRY(encode_map(ctx, &tmB, g.B, g.N, g.K, g.b_ks, 64, BK));
    if (g.epi.kind == EPI_ACCUM_F32) {
        maps[2] = maps[3] = maps[4] = tmA;  // unused by the reduction epilogue
    } else {
        // [M, N] bf16 matrices with leading dimension ldo, moved as boxes of 32 rows x 64 columns
        VITRS_TRY(encode_map(ctx, &maps[2], g.epi.out, g.N, g.M, g.epi.ldo, 64, 32));
        if (g.epi.out2 && g.epi.kind != EPI_ROWDOT) VITRS_TRY(encode_map(ctx, &maps[3], g.epi.out2, g.N, g.M, g.epi.ldo, 64, 32));
        else maps[3] = maps[2];
        if (g.epi.aux) VITRS_TRY(encode_map(ctx, &maps[4], g.epi.aux, g.N, g.M, g.epi.ldo, 64, 32));
        else maps[4] = maps[2];
    }

    TcParams p;
    p.M = g.M; p.N = g.N; p.K = g.K;
    p.m_tiles = pl.m_tiles;
    p.n_tiles = pl.n_tiles;
    p.kb_total = pl.kb_total;
    p.kb_per_split = pl.kb_per_split;
    p.splits = pl.splits;
    p.epi = g.epi;
    p.a_colsum = g.a_colsum;
    p.sched = ctx->env_gemm_static ? nullptr : ctx->gemm_sched;  // VITRS_GEMM_STATIC: the static stride (A/B aid)
    if (CG == 2) return launch_tc_major<256, 6, 2>(ctx, a_mn, b_mn, maps, p);
    if (BN == 256) return launch_tc_major<256, 4, 1>(ctx, a_mn, b_mn, maps, p);
    return launch_tc_major<128, 6, 1>(ctx, a_mn, b_mn, maps, p);
}

// The plan of a dense, 16-byte-aligned call of these extents (vitrs.h): host arithmetic only, no context, no CUDA call.
extern "C" int vitrs_gemm_plan(int M, int N, int K, int a_mn_major, int b_mn_major, int epilogue, int sm_count, int flags,
                               vitrs_gemm_plan_t* out) {
    if (!out || M < 1 || N < 1 || K < 1 || sm_count < 2) return VITRS_ERR_ARG;
    memset(out, 0, sizeof(*out));
    GemmDesc g = {};
    void* const aligned = reinterpret_cast<void*>(uintptr_t(4096));  // never dereferenced: tc_eligible only looks at alignment
    g.A = aligned; g.B = aligned;
    g.a_rs = a_mn_major ? 1 : K; g.a_ks = a_mn_major ? M : 1;
    g.b_rs = b_mn_major ? 1 : K; g.b_ks = b_mn_major ? N : 1;
    g.M = M; g.N = N; g.K = K;
    g.epi.kind = epilogue; g.epi.out = aligned; g.epi.ldo = N;
    if (epilogue == EPI_ROWDOT) { g.epi.aux = aligned; g.epi.out2 = aligned; g.epi.np = 1; }
    TcPlan pl = {};
    pl.simt = !tc_eligible(g);
    if (!pl.simt) {
        const TcSwitches sw = {sm_count, flags & VITRS_PLAN_NO_SMALL ? 1 : 0, flags & VITRS_PLAN_SINGLE_CTA ? 1 : 0,
                               flags & VITRS_PLAN_PATCH_TC ? 1 : 0, 0};
        pl = plan_tc(M, N, K, epilogue, a_mn_major != 0, b_mn_major != 0, sw);
    }
    if (pl.simt) {
        // gemm_simt.cu: [64 x 64] tiles, [32 x 32] when those would not give every SM a CTA
        const bool big = (long)ceil_div(N, 64) * ceil_div(M, 64) >= sm_count;
        out->kernel = VITRS_PLAN_SIMT;
        out->tile_m = out->tile_n = big ? 64 : 32;
        out->cta_group = 1;
        out->splits = 1;
        out->grid = ceil_div(N, out->tile_n) * ceil_div(M, out->tile_m);
        return VITRS_OK;
    }
    out->kernel = VITRS_PLAN_TCGEN05;
    out->tile_m = BM * pl.CG;
    out->tile_n = pl.BN;
    out->cta_group = pl.CG;
    out->stages = pl.stages;
    out->splits = pl.splits;
    out->k_blocks_per_split = pl.kb_per_split;
    out->tiles = pl.m_tiles * pl.n_tiles;
    out->grid = pl.grid;
    return VITRS_OK;
}
