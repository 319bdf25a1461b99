// attention_tc.cu — fused multi-head attention on the sm_100a tensor path (tcgen05.mma, TMEM
// accumulators, TMA loads) for head size 64 and sequences of up to 256 tokens (ViT-Ti/S/B at
// patch 16: T = 197).  Replaces attention_forward / attention_backward of train_vit.rs:400-451,
// 559-601: no T x T buffer ever reaches HBM; forward keeps lse[B,NH,T], backward recomputes the
// probabilities from it and uses D = rowsum(dO * O) in place of the reference's O(T^3)
// softmax-Jacobian loop (tv:583-589).
//
// Kernels in this file (the launchers at the end pick): attn_fwd_persist_kernel (forward, 128 < T <= 256: persistent, two
// softmax groups, operands of the next head prefetched), attn_fwd_tc2_kernel (forward, T <= 128 and causal masks: one CTA per
// 128-query tile, P in tensor memory), attn_bwd_persist_kernel (backward, persistent and software-pipelined, non-causal
// T <= 256), attn_bwd_tc_kernel (backward with a causal mask, T <= 256) and the streaming kernels for any sequence length.
// The notes below describe the common data flow.
//
// One CTA owns one (batch, head) at a time.  Q, K, V (and dO) tiles of 128 tokens x 64 are read straight
// out of the packed qkv[B,T,3C] activation by 3-D TMA boxes (row pitch 3C, column offset
// {0,C,2C} + h*64, rows >= T zero-filled), so the reference layout needs no permute kernels.
//
// forward   S_i = Q_i K^T  (M=128 queries, N=T keys, fp32 in TMEM)  ->  one thread per query row:
//           exp2 / sum from TMEM in one pass, P_i packed to bf16 in place in TMEM (the A operand of the next MMA)  ->
//           O_i = P_i V  ->  O_i / sum to out[B,T,C] at column h*64, lse.
//           The two 128-row query tiles run on two 4-warp groups concurrently.
// backward  transposed orientation, keys on TMEM lanes: for every (key tile j, query sub-tile s)
//           S^T = K_j Q_s^T and dP^T = V_j dO_s^T  ->  P^T = exp2(S^T*c - lse), dS^T = P^T o (dP^T - D) * scale, both packed
//           in place in TMEM as the A operands of dV_j += P^T dO_s and dK_j += dS^T Q_s; dS^T is also written to shared
//           memory, where it is the MN-major A operand of dQ_i += dS K_j.
//           dV_j, dK_j and both dQ_i accumulate in TMEM; nothing is reduced through global memory.
// Every tcgen05 instruction is issued by a CONVERGED warp from one elected region per batch (see tc_ptx.cuh: elect_one).
#include <stdlib.h>

#include "tc_ptx.cuh"

// Diagnostic build only (make trace -> libvitrs_trace.so, scripts/attn_trace.py): CTA 0 stamps its SM clock at the pipeline's
// hand-over points so the critical path of the persistent kernels can be read off a timeline.  Compiled out otherwise.
#ifdef VITRS_ATTN_TRACE
// per-warp rows of 2048 stamps, written with plain stores (an atomic cursor would cost the stamping warp a round trip to L2)
__device__ unsigned long long g_attn_trace[20 * 2048];  // one row of stamps per warp (the widest kernel has 18)
#define TR_DECL unsigned int tr_i__ = 0;
#define TR(ev, a)                                                                                                              \
    do {                                                                                                                       \
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && tr_i__ < 2048u)                                                       \
            g_attn_trace[((threadIdx.x >> 5) << 11) + tr_i__++] = ((unsigned long long)clock64() << 24) |                        \
                                                                  ((unsigned long long)(threadIdx.x >> 5) << 16) |               \
                                                                  ((unsigned long long)(ev) << 8) | (unsigned long long)((a) & 255); \
    } while (0)
extern "C" int vitrs_debug_trace_read(unsigned long long* out, unsigned int* n) {
    static unsigned long long host[20 * 2048];
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    cudaMemcpyFromSymbol(host, g_attn_trace, sizeof(host));
    unsigned int cnt = 0;
    for (int i = 0; i < 20 * 2048; ++i)
        if (host[i]) { if (out) out[cnt] = host[i]; ++cnt; }
    memset(host, 0, sizeof(host));
    cudaMemcpyToSymbol(g_attn_trace, host, sizeof(host));
    *n = cnt;
    return 0;
}
#else
#define TR_DECL
#define TR(ev, a)
#endif

namespace {

constexpr int kThreads = 256;
constexpr int TILE = 128;             // tokens per tile
constexpr int HS = 64;                // head size
constexpr int TILE_BYTES = TILE * HS * 2;  // 16 KB: one TMA box, 128 rows of 128 bytes
constexpr float kLog2e = 1.4426950408889634f;
constexpr int SUB = 64;                       // 64-row sub-tile (streamed tiles, pipelined backward)
constexpr int SUB_BYTES = SUB * HS * 2;       // 8 KB

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void mbar_arrive_cnt(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 16-byte chunk `c8` (8 bf16) of row `row` inside a [rows][64] bf16 tile with the 128-byte swizzle
// that TMA and the UMMA descriptors use (chunk index xor row mod 8); tile base is 1024-aligned
__device__ __forceinline__ uint32_t sw128(uint32_t tile_base, int row, int c8) {
    return tile_base + row * 128 + ((c8 ^ (row & 7)) << 4);
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot_addr, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_addr), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
// 32 fp32 values -> bf16 -> 64 contiguous bytes of global memory (optionally added to what is there)
__device__ __forceinline__ void store_row32(bf16* dst, const uint32_t* v, float mul, bool accumulate) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g * 8 + j]) * mul;
        if (accumulate) {
            Vec16<bf16> old;
            old.load(dst + g * 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] += old.get(j);
        }
        uint4 o;
        o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
        *reinterpret_cast<uint4*>(dst + g * 8) = o;
    }
}

// this thread's 32 fp32 accumulator columns -> (+ old bf16 values) -> bf16 -> its half-row of a swizzled [128][64] staging tile
__device__ __forceinline__ void stage_half_row(uint32_t tile, int r, int half, const uint32_t (&v)[32], const bf16* old) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g * 8 + j]);
        if (old) {
            Vec16<bf16> o;
            o.load(old + g * 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] += o.get(j);
        }
        st_shared_v4(sw128(tile, r, half * 4 + g), pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
    }
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1),
                 "r"(c2)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// =============================================== forward ==========================================
// ---- forward, one CTA per (batch, head, 128-query tile), two CTAs per SM ------------------------------
// P never leaves tensor memory: the softmax threads overwrite their S row in place with packed bf16
// (tcgen05.st) and the P.V MMA takes its A operand from TMEM, so a CTA needs only Q_i, K, V in shared
// memory (80 KB at T = 197) and 256 TMEM columns; the second resident CTA's loads and MMAs overlap this
// CTA's exponentials.  O is staged through the dead Q tile and leaves with one TMA store (rows >= T clipped).
__global__ void __launch_bounds__(128, 2)
attn_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_out, float* __restrict__ lse, int T,
                    int C, int NH, int causal, int NT, int NK, uint32_t tmem_cols) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sQ = base, sK = sQ + TILE_BYTES, sV = sK + NT * TILE_BYTES;
    const uint32_t bar0 = sV + NT * TILE_BYTES;
    const uint32_t bar_qk = bar0, bar_v = bar0 + 8, bar_s = bar0 + 16, bar_o = bar0 + 24;
    volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(gen + (bar0 - base) + 32);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int qt = blockIdx.x % NT, bh = blockIdx.x / NT, b = bh / NH, h = bh - b * NH;
    if (tid == 0) {
        tma_prefetch_desc(&tm_qkv);
        tma_prefetch_desc(&tm_out);
        mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the loads are in flight while the CTA allocates tensor memory and synchronises
        mbar_expect_tx(bar_qk, (uint32_t)((1 + NT) * TILE_BYTES));
        tma_load_3d(sQ, &tm_qkv, bar_qk, h * HS, qt * TILE, b);
        for (int i = 0; i < NT; ++i) tma_load_3d(sK + i * TILE_BYTES, &tm_qkv, bar_qk, C + h * HS, i * TILE, b);
        mbar_expect_tx(bar_v, (uint32_t)(NT * TILE_BYTES));
        for (int i = 0; i < NT; ++i) tma_load_3d(sV + i * TILE_BYTES, &tm_qkv, bar_v, 2 * C + h * HS, i * TILE, b);
    }
    if (warp == 1) tmem_alloc(smem_u32((const void*)slot), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *slot;
    const uint32_t o_col = tmem_cols - HS;  // O accumulator: the last 64 columns (dead part of S once P is packed)

    if (tid == 0) {
        const uint32_t idesc = make_idesc(TILE, NK, 0, 0);
        const uint64_t dq = make_desc(sQ, 0, 1024), dk = make_desc(sK, 0, 1024);
        mbar_wait(bar_qk, 0);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HS / 16; ++k) umma_bf16(tmem_base, dq + 2 * k, dk + 2 * k, idesc, k > 0);
        umma_commit(bar_s);
    }
    mbar_wait(bar_s, 0);
    tc_fence_after();
    const int r = tid;
    const int q = qt * TILE + r;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int kend = causal ? min(T, q + 1) : T;
    const int nchunks = (NK + 31) >> 5;
    const float scale = 1.0f / sqrtf((float)HS);
    const float sl2 = kLog2e * scale;
    float mx = -INFINITY;
    for (int ch = 0; ch < nchunks; ++ch) {
        uint32_t v[32];
        tmem_ld32(lane_addr + ch * 32, v);
        if (ch * 32 + 32 <= kend) {  // whole chunk visible: no per-element masks (6 of the 7 chunks at T = 197)
            float m0 = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1])), m1 = fmaxf(__uint_as_float(v[2]), __uint_as_float(v[3]));
#pragma unroll
            for (int c = 4; c < 32; c += 4) {
                m0 = fmaxf(m0, fmaxf(__uint_as_float(v[c]), __uint_as_float(v[c + 1])));
                m1 = fmaxf(m1, fmaxf(__uint_as_float(v[c + 2]), __uint_as_float(v[c + 3])));
            }
            mx = fmaxf(mx, fmaxf(m0, m1));
        } else {
#pragma unroll
            for (int c = 0; c < 32; ++c)
                if (ch * 32 + c < kend) mx = fmaxf(mx, __uint_as_float(v[c]));
        }
    }
    float sum = 0.f;
    const float mxs = mx * sl2;
    for (int ch = 0; ch < nchunks; ++ch) {
        uint32_t v[32], pk[16];
        tmem_ld32(lane_addr + ch * 32, v);
        if (ch * 32 + 32 <= kend) {
            float s0 = 0.f, s1 = 0.f;
            const float2 sl22 = splat2(sl2), nmx2 = splat2(-mxs);
            for (int c = 0; c < 16; ++c) {  // one packed FMA per pair of scores
                const float2 a = fma2(make_float2(__uint_as_float(v[2 * c]), __uint_as_float(v[2 * c + 1])), sl22, nmx2);
                const float p0 = ex2(a.x), p1 = ex2(a.y);
                s0 += p0;
                s1 += p1;
                pk[c] = pack_bf16(p0, p1);
            }
            sum += s0 + s1;
        } else {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const int k = ch * 32 + 2 * c;
                const float p0 = k < kend ? ex2(__uint_as_float(v[2 * c]) * sl2 - mxs) : 0.f;
                const float p1 = k + 1 < kend ? ex2(__uint_as_float(v[2 * c + 1]) * sl2 - mxs) : 0.f;
                sum += p0 + p1;
                pk[c] = pack_bf16(p0, p1);
            }
        }
        tmem_st16(lane_addr + ch * 16, pk);  // in place: columns [16ch, 16ch+16) were consumed by chunk ch/2 <= ch
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        const uint32_t idesc = make_idesc(TILE, HS, 0, 1);  // A = P from TMEM (K-major by construction), B = V MN-major
        const uint64_t dv = make_desc(sV, TILE_BYTES, 1024);
        mbar_wait(bar_v, 0);
        for (int k16 = 0; k16 < NK / 16; ++k16) umma_bf16_ts(tmem_base + o_col, tmem_base + k16 * 8, dv + 128 * k16, idesc, k16 > 0);
        umma_commit(bar_o);
    }
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float inv = 1.0f / sum;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld32(lane_addr + o_col + half * 32, v);
#pragma unroll
        for (int g = 0; g < 4; ++g)
            st_shared_v4(sw128(sQ, r, half * 4 + g), pack_bf16(__uint_as_float(v[g * 8]) * inv, __uint_as_float(v[g * 8 + 1]) * inv),
                         pack_bf16(__uint_as_float(v[g * 8 + 2]) * inv, __uint_as_float(v[g * 8 + 3]) * inv),
                         pack_bf16(__uint_as_float(v[g * 8 + 4]) * inv, __uint_as_float(v[g * 8 + 5]) * inv),
                         pack_bf16(__uint_as_float(v[g * 8 + 6]) * inv, __uint_as_float(v[g * 8 + 7]) * inv));
    }
    if (q < T) lse[(long)bh * T + q] = mx * scale + logf(sum);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tma_store_3d(&tm_out, sQ, h * HS, qt * TILE, b);
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ---- forward, persistent (128 < T <= 256, non-causal): one CTA per SM walks over its (batch, head) pairs ----------------
// The kernel above pays a launch, a tensor-memory allocation and an un-overlapped ~2 us operand fetch per 128-query tile, and
// fetches K and V of a head once per query tile.  Here the CTA stays resident:
//   loader (warp 0, one thread)  Q_0 Q_1 K_0 K_1 V_0 V_1 of the NEXT head stream into the other 96 KB stage while this head runs
//                                (each head's operands cross HBM -> shared memory exactly once);
//   issuer (warp 1, one thread)  a polling state machine over the two query tiles: O_g = P_g V as soon as P_g is packed, with the
//                                next head's S_g = Q_g K^T queued right behind it (in-order tensor pipe), except for the few
//                                score columns that alias the O accumulator, which follow when the group has read O out;
//   softmax group g (8 warps)    two threads per query row of tile g (4 warps, one thread per row, in the NOSPLIT variant): one
//                                pass over the scores in TMEM — exp2 / sum against a lazily moved exponent reference, P packed to
//                                bf16 in place (A operand of P.V) — then O scaled and staged through the dead Q_g tile, one TMA store.
// Group 1 starts half a period late, so one group's pass runs under the other group's MMA / read-out phases.  Warps whose 32
// rows are all beyond T (rows 224..255 at T = 197) skip the arithmetic and only keep the barriers moving.
// TMEM: two 256-column regions {S [0,NK) -> P [0,NK/2); O [192,256)}.
// What bounds it (profiles/r2_attn_fwd_trace.txt, r2_softmax_chunk_rate.txt): tcgen05.ld moves 64 B / clock / SM, so reading the
// scores (224 live rows x 208 columns x 4 B) and O takes 3.8 k cycles per head whatever the arithmetic costs — 14 softmax warps
// each asking for a 4 KB chunk keep that port ~85 % busy while both groups are in their passes, which is why a 32-column chunk
// takes ~1 050 cycles in the kernel against 610 for the same instruction mix without the loads.  The MUFU pipe (16 ex2 / clock /
// SM: 3.3 k cycles per head) is the second floor; evaluating a share of the exponentials on the FMA pipe does not pay here
// (7 issue slots per element against 3: scripts/exp_mufu_rate.cu).  The remaining time is each group's serial chain
// S -> pass -> P.V -> O out, during which only the other group reads.
constexpr int kMainCols = 192;  // score columns that do not overlap the O accumulator at [192, 256)
constexpr float kTau = 12.f;  // NOSPLIT: probabilities may exceed 1 by up to 2^kTau before the exponent reference is moved
constexpr float kPBig = 1.8446744e19f;  // 2^64: SPLIT / streaming forward move a row's exponent reference only when a chunk of probabilities sums past this
// SPLIT: first score column of a row's second thread — half of NK rounded up to a multiple of 16 (the P.V MMAs consume 16 keys each).
// For 128 < T <= 256 both parts are at least 32 columns wide and each part's first 32 columns lie below T.
__host__ __device__ __forceinline__ int fwd_split_point(int NK) { return ((NK >> 1) + 15) & ~15; }
constexpr int kFwdThreads = 320;       // warp 0 loader, warp 1 issuer / TMEM owner, warps 2-5 group 0, warps 6-9 group 1
constexpr int kFwdSplitThreads = 576;  // SPLIT: two threads per query row — warps 2-9 group 0, warps 10-17 group 1; within a group the
                                       // first four warps take keys [0, 128), the other four keys [128, T)

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

// the same for a converged warp: one answer for all lanes (a completed phase stays completed, so "any lane saw it" is exact)
__device__ __forceinline__ bool mbar_test_warp(uint32_t bar, uint32_t parity) { return __any_sync(0xffffffffu, mbar_test(bar, parity)); }

template <bool STAGGER, bool SPLIT>
__global__ void __launch_bounds__(SPLIT ? kFwdSplitThreads : kFwdThreads, 1)
attn_fwd_persist_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_out, float* __restrict__ lse,
                        int T, int C, int NH, int NK, int total_heads) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    constexpr uint32_t STAGE = 6 * TILE_BYTES;  // Q_0 Q_1 K_0 K_1 V_0 V_1
    const uint32_t bar0 = base + 2 * STAGE;
    const uint32_t full_qk = bar0, full_v = bar0 + 16, s_main = bar0 + 32, p_ready = bar0 + 48, o_ready = bar0 + 64,
                   tmem_free = bar0 + 80, stage_free = bar0 + 96, s_tail = bar0 + 112;  // two barriers each
    volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(gen + (bar0 - base) + 128);
    float* xch = reinterpret_cast<float*>(gen + (bar0 - base) + 256);  // SPLIT: [group][half][row][3] floats exchanged between a row's two threads
    constexpr uint32_t TMEM_COLS = 512, REGION = 256, cO = REGION - HS;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform: role branches stay converged
    TR_DECL
    const int nheads = (total_heads - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // heads of this CTA
    if (tid == 0) {
        tma_prefetch_desc(&tm_qkv);
        tma_prefetch_desc(&tm_out);
        for (int i = 0; i < 2; ++i) {
            mbar_init(full_qk + 8 * i, 1);
            mbar_init(full_v + 8 * i, 1);
            mbar_init(s_main + 8 * i, 1);
            mbar_init(s_tail + 8 * i, 1);
            mbar_init(p_ready + 8 * i, SPLIT ? 256 : 128);
            mbar_init(o_ready + 8 * i, 1);
            mbar_init(tmem_free + 8 * i, SPLIT ? 256 : 128);
            mbar_init(stage_free + 8 * i, 2);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32((const void*)slot), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *slot, 0);

    if (warp == 0) {
        if (lane == 0) {
            // ================================ loader ================================
            for (int G = 0; G < nheads; ++G) {
                const int bh = (int)blockIdx.x + G * (int)gridDim.x, b = bh / NH, h = bh - b * NH;
                const int st = G & 1;
                const uint32_t sQ = base + st * STAGE, sK = sQ + 2 * TILE_BYTES, sV = sK + 2 * TILE_BYTES;
                if (G >= 2) mbar_wait(stage_free + 8 * st, (uint32_t)(((G >> 1) - 1) & 1));  // both groups are done with head G - 2
                mbar_expect_tx(full_qk + 8 * st, (uint32_t)(4 * TILE_BYTES));
                tma_load_3d(sK, &tm_qkv, full_qk + 8 * st, C + h * HS, 0, b);
                tma_load_3d(sQ, &tm_qkv, full_qk + 8 * st, h * HS, 0, b);
                tma_load_3d(sK + TILE_BYTES, &tm_qkv, full_qk + 8 * st, C + h * HS, TILE, b);
                tma_load_3d(sQ + TILE_BYTES, &tm_qkv, full_qk + 8 * st, h * HS, TILE, b);
                mbar_expect_tx(full_v + 8 * st, (uint32_t)(2 * TILE_BYTES));
                tma_load_3d(sV, &tm_qkv, full_v + 8 * st, 2 * C + h * HS, 0, b);
                tma_load_3d(sV + TILE_BYTES, &tm_qkv, full_v + 8 * st, 2 * C + h * HS, TILE, b);
            }
        }
    } else if (warp == 1) {
        {
            // ================================ issuer (whole warp converged, tcgen05 instructions on the elected lane) ================
            // S = Q K^T (both operands K-major) is issued in two pieces.  Columns [0, kMainCols) of the next head go out right behind
            // this head's P.V MMAs: tcgen05.mma executes in issue order, so overwriting the probabilities the P.V MMAs are still
            // reading needs no barrier, and the O accumulator at [192, 256) is untouched.  The few columns that overlap O
            // (keys >= 192) follow once the group has read O out.  The group thus finds its next scores waiting when it comes back.
            const int n_main = NK < kMainCols ? NK : kMainCols, n_tail = NK - n_main;
            const uint32_t idesc_main = make_idesc(TILE, n_main, 0, 0);
            const uint32_t idesc_tail = make_idesc(TILE, n_tail > 0 ? n_tail : 16, 0, 0);
            const uint32_t idesc_o = make_idesc(TILE, HS, 0, 1);   // O = P V: A = P from TMEM, B = V MN-major
            int nMain[2] = {0, 0}, nTail[2] = {0, 0}, nPV[2] = {0, 0};
            auto issue_main = [&](int g, int G) {
                const uint32_t sQ = base + (G & 1) * STAGE;
                const uint64_t dq = make_desc(sQ + g * TILE_BYTES, 0, 1024), dk = make_desc(sQ + 2 * TILE_BYTES, 0, 1024);
                if (elect_one()) {  // one elected region per batch: the MMAs go out back to back (scripts/exp_mma_rate.cu)
#pragma unroll
                    for (int k = 0; k < HS / 16; ++k) umma_bf16(tmem_base + (uint32_t)g * REGION, dq + 2 * k, dk + 2 * k, idesc_main, k > 0);
                    umma_commit(s_main + 8 * g);
                }
                __syncwarp();
                nMain[g] = G + 1;
                TR(52 + g, G);
            };
            while (nPV[0] < nheads || nPV[1] < nheads) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    if (nPV[g] < nTail[g]) {
                        const int G = nPV[g], st = G & 1;
                        if (mbar_test_warp(p_ready + 8 * g, (uint32_t)(G & 1)) && mbar_test_warp(full_v + 8 * st, (uint32_t)((G >> 1) & 1))) {
                            tc_fence_after();
                            const uint32_t region = tmem_base + (uint32_t)g * REGION;
                            const uint64_t dv = make_desc(base + st * STAGE + 4 * TILE_BYTES, TILE_BYTES, 1024);
                            if (elect_one()) {
                                // (SPLIT: the probabilities of keys >= SP were packed in place over THEIR OWN scores, from column SP on)
                                const int sp16 = fwd_split_point(NK) / 16;
                                for (int k16 = 0; k16 < NK / 16; ++k16)
                                    umma_bf16_ts(region + cO, region + (SPLIT && k16 >= sp16 ? sp16 * 16 + (k16 - sp16) * 8 : k16 * 8), dv + 128 * k16, idesc_o, k16 > 0);
                                umma_commit(o_ready + 8 * g);
                            }
                            __syncwarp();
                            nPV[g] = G + 1;
                            TR(50 + g, G);
                            // the next head's main scores right behind, if its operands have landed
                            if (G + 1 < nheads && mbar_test_warp(full_qk + 8 * ((G + 1) & 1), (uint32_t)(((G + 1) >> 1) & 1))) {
                                tc_fence_after();
                                issue_main(g, G + 1);
                            }
                        }
                    }
                    if (nMain[g] < nheads && nMain[g] == nPV[g]) {  // (not issued behind P.V: first head, or operands were late)
                        const int G = nMain[g];
                        // group 1 enters half a period behind group 0: its first scores wait for group 0's first probabilities
                        const bool held = STAGGER && g == 1 && G == 0 && nPV[0] == 0;
                        if (!held && mbar_test_warp(full_qk + 8 * (G & 1), (uint32_t)((G >> 1) & 1))) {
                            tc_fence_after();
                            issue_main(g, G);
                        }
                    }
                    if (nTail[g] < nMain[g]) {
                        const int G = nTail[g];
                        if (G == 0 || mbar_test_warp(tmem_free + 8 * g, (uint32_t)((G - 1) & 1))) {  // O of the previous head has been read out
                            tc_fence_after();
                            const uint32_t sQ = base + (G & 1) * STAGE;
                            const uint64_t dq = make_desc(sQ + g * TILE_BYTES, 0, 1024);
                            const uint64_t dk = make_desc(sQ + 2 * TILE_BYTES + (kMainCols / 8) * 1024, 0, 1024);  // key rows >= kMainCols
                            if (elect_one()) {
                                if (n_tail > 0) {
#pragma unroll
                                    for (int k = 0; k < HS / 16; ++k)
                                        umma_bf16(tmem_base + (uint32_t)g * REGION + kMainCols, dq + 2 * k, dk + 2 * k, idesc_tail, k > 0);
                                }
                                umma_commit(s_tail + 8 * g);
                            }
                            __syncwarp();
                            nTail[g] = G + 1;
                            TR(54 + g, G);
                        }
                    }
                }
            }
        }
    } else if (SPLIT) {
        // ================================ softmax groups, two threads per query row ================================
        // With one thread per row a group is one warp per scheduler, and its pass over the scores is a chain of dependent MUFU /
        // FMA / convert instructions (810 cycles per 32-column chunk, measured).  Here a row's columns are split between two
        // threads of the same TMEM lane quarter at SP = NK/2 rounded up to 16 (T = 197: keys [0, 112) | [112, 208)); each packs its
        // probabilities in place over its own (already read) scores — half 1 from column SP on, which the P.V MMAs are told about.
        // The two halves agree on the exponent reference before they exponentiate (maximum of their first chunks, through shared
        // memory).  After that NO maximum is taken: a chunk goes fma -> ex2 -> sum -> pack with the next chunk's tensor-memory
        // load in flight, and only a chunk whose probabilities sum past 2^64 (a key 44 nats above the first 64: never, in
        // practice) takes the slow path that moves the reference and rescales what was written by an exact power of two.
        // bf16 and fp32 share their exponent range, so P <= 2^64 is as exact as P <= 1; the result is the softmax itself.
        const int gi = warp - 2;
        const int g = gi >> 3, half = (gi >> 2) & 1;
        const int r = (warp & 3) * 32 + lane;  // TMEM lane = query row within the tile
        const int q = g * TILE + r;
        const bool warp_live = g * TILE + (warp & 3) * 32 < T;
        const bool store_leader = (gi & 7) == 0 && lane == 0;
        const uint32_t lane_addr = tmem_base + (uint32_t)g * REGION + ((uint32_t)((warp & 3) * 32) << 16);
        const int SP = fwd_split_point(NK);
        const int c_lo = half ? SP : 0, c_hi = half ? NK : SP;  // this thread's score columns (both extents are multiples of 16, >= 32)
        const float scale = 1.0f / sqrtf((float)HS);
        const float sl2 = kLog2e * scale;
        float* mine = xch + ((g * 2 + half) * TILE + r) * 3;
        const float* partner = xch + ((g * 2 + (half ^ 1)) * TILE + r) * 3;
        int pending_stage = -1;  // leader: stage whose O store still has to be confirmed read
        for (int G = 0; G < nheads; ++G) {
            const int bh = (int)blockIdx.x + G * (int)gridDim.x, b = bh / NH, h = bh - b * NH;
            const int st = G & 1;
            const uint32_t par = (uint32_t)(G & 1);
            const uint32_t sO = base + st * STAGE + g * TILE_BYTES;  // the dead Q_g tile of this stage
            TR(41, G);
            mbar_wait(s_main + 8 * g, par);
            tc_fence_after();
            TR(42, G);
            if (store_leader && pending_stage >= 0) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                mbar_arrive_cnt(stage_free + 8 * pending_stage);
                pending_stage = -1;
            }
            float m_ref = 0.f, sum = 0.f;
            bool tail_ready = false;
            uint32_t v[32];
            auto pcol = [&](int col) { return c_lo + ((col - c_lo) >> 1); };  // where the packed probabilities of score column `col` go
            auto issue_load = [&](int col) {  // (no wait) the chunk at `col`: 32 columns, or the 16 that are left
                const int w = c_hi - col < 32 ? 16 : 32;
                if (!tail_ready && col + w > kMainCols) {  // the columns from kMainCols on arrive with the second piece of the score MMA
                    mbar_wait(s_tail + 8 * g, par);
                    tc_fence_after();
                    tail_ready = true;
                }
                if (w == 32) tmem_ld32_issue(lane_addr + col, v);
                else tmem_ld16_issue(lane_addr + col, v);
            };
            // one chunk of W score columns, already in v: exponentiate against m_ref, pack, store in place; the next chunk's load is
            // issued as soon as the scores have gone through the FMA (v is dead from there on)
            auto step = [&](auto Wc, int col) {
                constexpr int W = decltype(Wc)::value;
                const bool has_next = col + W < c_hi;
                const bool full = col + W <= T;
                float2 a[W / 2];
                uint32_t pk[16];
                float csum;
                auto scale_scores = [&]() {
                    const float2 sl22 = splat2(sl2), nref2 = splat2(-m_ref);
#pragma unroll
                    for (int c = 0; c < W / 2; ++c) a[c] = fma2(make_float2(__uint_as_float(v[2 * c]), __uint_as_float(v[2 * c + 1])), sl22, nref2);
                };
                auto exponentiate = [&]() {
                    float s0 = 0.f, s1 = 0.f;
                    if (full) {
#pragma unroll
                        for (int c = 0; c < W / 2; ++c) {
                            const float p0 = ex2(a[c].x), p1 = ex2(a[c].y);
                            s0 += p0;
                            s1 += p1;
                            pk[c] = pack_bf16(p0, p1);
                        }
                    } else {  // (a branch of its own, warp-uniform: only the chunk that contains key T pays for the masks)
                        const int live = T - col;
#pragma unroll
                        for (int c = 0; c < W / 2; ++c) {
                            const float p0 = 2 * c < live ? ex2(a[c].x) : 0.f, p1 = 2 * c + 1 < live ? ex2(a[c].y) : 0.f;
                            s0 += p0;
                            s1 += p1;
                            pk[c] = pack_bf16(p0, p1);
                        }
                    }
                    csum = s0 + s1;
                };
                scale_scores();
                if (has_next) issue_load(col + W);
                exponentiate();
                if (__any_sync(0xffffffffu, !(csum < kPBig))) {
                    // never in practice: a probability of this chunk left the comfortable range.  Its scores are still intact in
                    // tensor memory (the chunk's probabilities have not been stored yet): read them again, move this row's
                    // reference to their maximum and rescale what the thread has written so far.
                    if (has_next) tmem_ld_wait32(v);  // retire the load in flight: its registers are needed
                    if (W == 32) tmem_ld32_issue(lane_addr + col, v);
                    else tmem_ld16_issue(lane_addr + col, v);
                    tmem_ld_wait32(v);
                    float cm = -INFINITY;
#pragma unroll
                    for (int c = 0; c < W; ++c)
                        if (full || col + c < T) cm = fmaxf(cm, __uint_as_float(v[c]));
                    const float new_ref = !(csum < kPBig) ? ceilf(cm * sl2) : m_ref;
                    const float f = ex2(m_ref - new_ref);  // 2^(integer <= 0): exact
                    tmem_st_wait();
                    for (int pc = pcol(c_lo); pc < pcol(col); pc += 16) {  // (every chunk before this one was 32 columns wide)
                        uint32_t old[16];
                        tmem_ld16(lane_addr + pc, old);
#pragma unroll
                        for (int c = 0; c < 16; ++c)
                            old[c] = pack_bf16(__uint_as_float(old[c] << 16) * f, __uint_as_float(old[c] & 0xFFFF0000u) * f);
                        tmem_st16(lane_addr + pc, old);
                    }
                    sum *= f;
                    m_ref = new_ref;
                    scale_scores();
                    if (has_next) issue_load(col + W);
                    exponentiate();
                }
                if (W == 32) tmem_st16(lane_addr + pcol(col), pk);  // in place: these columns held scores this thread has already read
                else tmem_st8(lane_addr + pcol(col), pk);
                sum += csum;
                if (has_next) tmem_ld_wait32(v);
            };
            // ---- the common exponent reference: the larger of the two halves' first-chunk maxima, rounded up ----
            float cm2 = -INFINITY;
            if (warp_live) {
                issue_load(c_lo);
                tmem_ld_wait32(v);
                TR(60, G);
                float m0 = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1])), m1 = fmaxf(__uint_as_float(v[2]), __uint_as_float(v[3]));
#pragma unroll
                for (int c = 4; c < 32; c += 4) {  // (the first chunk of either half lies below key 128 + 32 <= T... see fwd_split_point)
                    m0 = fmaxf(m0, fmaxf(__uint_as_float(v[c]), __uint_as_float(v[c + 1])));
                    m1 = fmaxf(m1, fmaxf(__uint_as_float(v[c + 2]), __uint_as_float(v[c + 3])));
                }
                cm2 = fmaxf(m0, m1) * sl2;
            }
            mine[0] = cm2;
            named_bar_sync(1 + g, 256);
            TR(61, G);
            if (warp_live) {
                m_ref = ceilf(fmaxf(cm2, partner[0]));
                for (int col = c_lo; col < c_hi; col += 32) {
                    if (c_hi - col >= 32) step(std::integral_constant<int, 32>{}, col);
                    else step(std::integral_constant<int, 16>{}, col);
                    TR(62, col >> 4);
                }
                tmem_st_wait();
                TR(63, G);
            }
            if (!tail_ready) mbar_wait(s_tail + 8 * g, par);  // (keeps the barrier's phase in step when no column needed it)
            // ---- the two halves of a row meet: common reference, total sum ----
            mine[1] = m_ref;
            mine[2] = sum;
            named_bar_sync(1 + g, 256);
            TR(64, G);
            float total = 1.f, m_all = m_ref;
            if (warp_live) {
                const float pm = partner[1], ps = partner[2];
                m_all = fmaxf(m_ref, pm);
                if (__any_sync(0xffffffffu, m_ref < m_all)) {  // never in practice: the other half moved its reference
                    const float f = ex2(m_ref - m_all);
                    for (int pc = pcol(c_lo); pc < pcol(c_hi); pc += 8) {
                        uint32_t old[16];
                        tmem_ld16(lane_addr + pc, old);  // (reads 8 columns more than it rewrites)
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            old[c] = pack_bf16(__uint_as_float(old[c] << 16) * f, __uint_as_float(old[c] & 0xFFFF0000u) * f);
                        tmem_st8(lane_addr + pc, old);
                    }
                    tmem_st_wait();
                    sum *= f;
                }
                total = sum + ps * ex2(pm - m_all);
            }
            tc_fence_before();
            mbar_arrive_cnt(p_ready + 8 * g);
            TR(43, G);
            mbar_wait(o_ready + 8 * g, par);
            tc_fence_after();
            TR(44, G);
            if (warp_live) {
                const float inv = 1.0f / total;
                uint32_t o[32];
                tmem_ld32(lane_addr + cO + half * 32, o);  // each half takes 32 of the 64 output columns
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8)
                    st_shared_v4(sw128(sO, r, half * 4 + g8), pack_bf16(__uint_as_float(o[g8 * 8]) * inv, __uint_as_float(o[g8 * 8 + 1]) * inv),
                                 pack_bf16(__uint_as_float(o[g8 * 8 + 2]) * inv, __uint_as_float(o[g8 * 8 + 3]) * inv),
                                 pack_bf16(__uint_as_float(o[g8 * 8 + 4]) * inv, __uint_as_float(o[g8 * 8 + 5]) * inv),
                                 pack_bf16(__uint_as_float(o[g8 * 8 + 6]) * inv, __uint_as_float(o[g8 * 8 + 7]) * inv));
                if (half == 0 && q < T) lse[(long)bh * T + q] = (m_all + __log2f(total)) * (1.0f / kLog2e);
            }
            tc_fence_before();
            mbar_arrive_cnt(tmem_free + 8 * g);  // S_g of the next head may overwrite this region
            TR(45, G);
            fence_proxy_async();
            named_bar_sync(1 + g, 256);
            TR(46, G);
            if (store_leader) {
                tma_store_3d(&tm_out, sO, h * HS, g * TILE, b);  // rows >= T are clipped by the tensor map
                pending_stage = st;
            }
        }
        if (store_leader) {
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the last stores are complete before shared memory is retired
            if (pending_stage >= 0) mbar_arrive_cnt(stage_free + 8 * pending_stage);
        }
    } else {
        // ================================ softmax groups ================================
        const int g = (warp - 2) >> 2;
        const int r = (warp & 3) * 32 + lane;  // TMEM lane = query row within the tile
        const int q = g * TILE + r;
        const bool warp_live = g * TILE + (warp & 3) * 32 < T;
        const bool store_leader = ((warp - 2) & 3) == 0 && lane == 0;
        const uint32_t lane_addr = tmem_base + (uint32_t)g * REGION + ((uint32_t)((warp & 3) * 32) << 16);
        const int nchunks = (NK + 31) >> 5;
        const float scale = 1.0f / sqrtf((float)HS);
        const float sl2 = kLog2e * scale;
        int pending_stage = -1;  // leader: stage whose O store still has to be confirmed read
        for (int G = 0; G < nheads; ++G) {
            const int bh = (int)blockIdx.x + G * (int)gridDim.x, b = bh / NH, h = bh - b * NH;
            const int st = G & 1;
            const uint32_t par = (uint32_t)(G & 1);
            const uint32_t sO = base + st * STAGE + g * TILE_BYTES;  // the dead Q_g tile of this stage
            TR(41, G);
            mbar_wait(s_main + 8 * g, par);
            tc_fence_after();
            TR(42, G);
            if (store_leader && pending_stage >= 0) {
                // the previous head's O store was issued a whole S MMA ago and has read its staging tile by now: tell the loader that
                // the stage may be refilled (it then has more than a head's time to fetch head G + 1)
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                mbar_arrive_cnt(stage_free + 8 * pending_stage);
                pending_stage = -1;
            }
            // One pass over the scores.  Tensor-memory reads (64 B / clock / SM) are the floor of this kernel, so the row maximum is
            // not taken in a pass of its own: the exponent reference m_ref (an integer, log2 domain) is the first chunk's maximum
            // rounded up, every probability is 2^(s - m_ref), and a later chunk whose maximum exceeds m_ref by more than kTau
            // (rare: a key more than 8 nats above the first 32) moves the reference and rescales the probabilities already
            // written by an exact power of two.  P <= 2^kTau keeps every sum far inside fp32 / bf16 range; the result is the
            // softmax itself, whatever the reference.
            float m_ref = 0.f, sum = 0.f;
            if (warp_live) {
                uint32_t va[32], vb[32];
                auto process = [&](int ch, uint32_t (&v)[32]) {
                    const int k0 = ch * 32;
                    const bool full = k0 + 32 <= T;
                    float cm;
                    if (full) {
                        float m0 = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1])), m1 = fmaxf(__uint_as_float(v[2]), __uint_as_float(v[3]));
#pragma unroll
                        for (int c = 4; c < 32; c += 4) {
                            m0 = fmaxf(m0, fmaxf(__uint_as_float(v[c]), __uint_as_float(v[c + 1])));
                            m1 = fmaxf(m1, fmaxf(__uint_as_float(v[c + 2]), __uint_as_float(v[c + 3])));
                        }
                        cm = fmaxf(m0, m1);
                    } else {
                        cm = -INFINITY;
#pragma unroll
                        for (int c = 0; c < 32; ++c)
                            if (k0 + c < T) cm = fmaxf(cm, __uint_as_float(v[c]));
                    }
                    const float cm2 = cm * sl2;
                    if (ch == 0) {
                        m_ref = ceilf(cm2);
                    } else if (__any_sync(0xffffffffu, cm2 > m_ref + kTau)) {
                        // rare: move the reference up and rescale what was written (warp-uniform branch, per-lane factor)
                        const float new_ref = cm2 > m_ref + kTau ? ceilf(cm2) : m_ref;
                        const float f = ex2(m_ref - new_ref);  // 2^(integer <= 0): exact
                        tmem_st_wait();
                        for (int blk = 0; blk < ch; ++blk) {
                            uint32_t pk[16];
                            tmem_ld16(lane_addr + blk * 16, pk);
#pragma unroll
                            for (int c = 0; c < 16; ++c)
                                pk[c] = pack_bf16(__uint_as_float(pk[c] << 16) * f, __uint_as_float(pk[c] & 0xFFFF0000u) * f);
                            tmem_st16(lane_addr + blk * 16, pk);
                        }
                        sum *= f;
                        m_ref = new_ref;
                    }
                    uint32_t pk[16];
                    if (full) {
                        const float2 sl22 = splat2(sl2), nref2 = splat2(-m_ref);
                        float s0 = 0.f, s1 = 0.f;
#pragma unroll
                        for (int c = 0; c < 16; ++c) {  // one packed FMA per pair of scores
                            const float2 a = fma2(make_float2(__uint_as_float(v[2 * c]), __uint_as_float(v[2 * c + 1])), sl22, nref2);
                            const float p0 = ex2(a.x), p1 = ex2(a.y);
                            s0 += p0;
                            s1 += p1;
                            pk[c] = pack_bf16(p0, p1);
                        }
                        sum += s0 + s1;
                    } else {
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            const int k = k0 + 2 * c;
                            const float p0 = k < T ? ex2(__uint_as_float(v[2 * c]) * sl2 - m_ref) : 0.f;
                            const float p1 = k + 1 < T ? ex2(__uint_as_float(v[2 * c + 1]) * sl2 - m_ref) : 0.f;
                            sum += p0 + p1;
                            pk[c] = pack_bf16(p0, p1);
                        }
                    }
                    tmem_st16(lane_addr + ch * 16, pk);  // in place: columns [16ch, 16ch+16) hold scores of chunk ch/2 <= ch, consumed
                };
                // the next chunk's load is in flight while this chunk is exponentiated; the columns from kMainCols on arrive with
                // the second piece of the score MMA (see the issuer)
                bool tail_ready = false;
                auto issue_load = [&](int ch, uint32_t (&v)[32]) {
                    if (!tail_ready && ch * 32 + 32 > kMainCols) {
                        mbar_wait(s_tail + 8 * g, par);
                        tc_fence_after();
                        tail_ready = true;
                    }
                    tmem_ld32_issue(lane_addr + ch * 32, v);
                };
                issue_load(0, va);
                tmem_ld_wait32(va);
                for (int ch = 0; ch < nchunks; ch += 2) {
                    if (ch + 1 < nchunks) issue_load(ch + 1, vb);
                    process(ch, va);
                    if (ch + 1 < nchunks) {
                        tmem_ld_wait32(vb);
                        if (ch + 2 < nchunks) issue_load(ch + 2, va);
                        process(ch + 1, vb);
                        if (ch + 2 < nchunks) tmem_ld_wait32(va);
                    }
                }
                tmem_st_wait();
                if (!tail_ready) mbar_wait(s_tail + 8 * g, par);  // (keeps the barrier's phase in step when no column needed it)
            } else {
                mbar_wait(s_tail + 8 * g, par);
            }
            tc_fence_before();
            mbar_arrive_cnt(p_ready + 8 * g);
            TR(43, G);
            mbar_wait(o_ready + 8 * g, par);
            tc_fence_after();
            TR(44, G);
            if (warp_live) {
                const float inv = 1.0f / sum;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t v[32];
                    tmem_ld32(lane_addr + cO + half * 32, v);
#pragma unroll
                    for (int g8 = 0; g8 < 4; ++g8)
                        st_shared_v4(sw128(sO, r, half * 4 + g8), pack_bf16(__uint_as_float(v[g8 * 8]) * inv, __uint_as_float(v[g8 * 8 + 1]) * inv),
                                     pack_bf16(__uint_as_float(v[g8 * 8 + 2]) * inv, __uint_as_float(v[g8 * 8 + 3]) * inv),
                                     pack_bf16(__uint_as_float(v[g8 * 8 + 4]) * inv, __uint_as_float(v[g8 * 8 + 5]) * inv),
                                     pack_bf16(__uint_as_float(v[g8 * 8 + 6]) * inv, __uint_as_float(v[g8 * 8 + 7]) * inv));
                }
                if (q < T) lse[(long)bh * T + q] = (m_ref + __log2f(sum)) * (1.0f / kLog2e);
            }
            tc_fence_before();
            mbar_arrive_cnt(tmem_free + 8 * g);  // S_g of the next head may overwrite this region
            TR(45, G);
            fence_proxy_async();
            named_bar_sync(1 + g, 128);
            TR(46, G);
            if (store_leader) {
                tma_store_3d(&tm_out, sO, h * HS, g * TILE, b);  // rows >= T are clipped by the tensor map
                pending_stage = st;
            }
        }
        if (store_leader) {
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the last stores are complete before shared memory is retired
            if (pending_stage >= 0) mbar_arrive_cnt(stage_free + 8 * pending_stage);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// D[b,h,q] = sum_i dO[b,q,h*64+i] * O[b,q,h*64+i]: one warp per token, 8 lanes per head
__global__ void attn_bwd_prep_kernel(float* __restrict__ dsum, const bf16* __restrict__ dout, const bf16* __restrict__ out, int B, int T,
                                     int C, int NH) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= (long)B * T) return;
    const int b = (int)(row / T), q = (int)(row - (long)b * T);
    for (int c8 = lane; c8 < C / 8; c8 += 32) {
        Vec16<bf16> x, y;
        x.load(dout + row * C + c8 * 8);
        y.load(out + row * C + c8 * 8);
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s = fmaf(x.get(j), y.get(j), s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if ((lane & 7) == 0) dsum[((long)b * NH + (c8 >> 3)) * T + q] = s;
    }
}

// =============================================== backward =========================================
__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do, bf16* __restrict__ dqkv,
                   const float* __restrict__ lse, const float* __restrict__ dsum, int T, int C, int NH, int causal, int NT,
                   int accumulate) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sQ = base, sK = sQ + NT * TILE_BYTES, sV = sK + NT * TILE_BYTES, sdO = sV + NT * TILE_BYTES;
    const uint32_t sPT = sdO + NT * TILE_BYTES;   // [2 blocks of 64 queries][128 keys][128 B]
    const uint32_t sdST = sPT + 2 * TILE_BYTES;
    const uint32_t sStat = sdST + 2 * TILE_BYTES;  // lse*log2e [256], D [256]
    const uint32_t bar0 = sStat + 2 * 256 * 4;
    const uint32_t bar_load = bar0, bar_s = bar0 + 8, bar_acc = bar0 + 16;
    float* stat = reinterpret_cast<float*>(gen + (sStat - base));
    volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(gen + (bar0 - base) + 24);
    constexpr uint32_t TMEM_COLS = 512;
    constexpr uint32_t cST = 0, cDPT = 128, cDV = 256, cDK = 320, cDQ = 384;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int bh = blockIdx.x, b = bh / NH, h = bh - b * NH;
    if (tid == 0) {
        tma_prefetch_desc(&tm_qkv);
        tma_prefetch_desc(&tm_do);
        mbar_init(bar_load, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(smem_u32((const void*)slot), TMEM_COLS);
    for (int i = tid; i < 256; i += kThreads) {
        stat[i] = i < T ? lse[(long)bh * T + i] * kLog2e : 0.f;
        stat[256 + i] = i < T ? dsum[(long)bh * T + i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *slot;
    const bool leader = tid == 0;
    if (leader) {
        mbar_expect_tx(bar_load, (uint32_t)(4 * NT * TILE_BYTES));
        for (int i = 0; i < NT; ++i) {
            tma_load_3d(sQ + i * TILE_BYTES, &tm_qkv, bar_load, h * HS, i * TILE, b);
            tma_load_3d(sK + i * TILE_BYTES, &tm_qkv, bar_load, C + h * HS, i * TILE, b);
            tma_load_3d(sV + i * TILE_BYTES, &tm_qkv, bar_load, 2 * C + h * HS, i * TILE, b);
            tma_load_3d(sdO + i * TILE_BYTES, &tm_do, bar_load, h * HS, i * TILE, b);
        }
        mbar_wait(bar_load, 0);
    }
    const int r = tid & 127;                 // TMEM lane = key (S^T phase) or output row (epilogues)
    const int half = warp >> 2;              // which half of the columns this thread handles
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const float scale = 1.0f / sqrtf((float)HS);
    const float sl2 = kLog2e * scale;
    uint32_t s_phase = 0, acc_phase = 0;
    uint32_t dq_started = 0;  // bit i: dQ_i already holds a partial sum

    for (int j = 0; j < NT; ++j) {           // key tile
        const int key = j * TILE + r;
        const int nk16 = (min(TILE, T - j * TILE) + 15) & ~15;  // keys of this tile, padded to the MMA K step
        bool first_i = true;
        for (int i = causal ? j : 0; i < NT; ++i) {  // query tile (causal: tiles left of the diagonal are empty)
            const int nq = min(TILE, T - i * TILE);  // valid queries in the tile
            const int nq16 = (nq + 15) & ~15;
            if (leader) {
                tc_fence_after();
                const uint32_t idesc = make_idesc(TILE, nq16, 0, 0);
#pragma unroll
                for (int k = 0; k < HS / 16; ++k)  // S^T = K_j Q_i^T
                    umma_bf16(tmem_base + cST, make_desc(sK + j * TILE_BYTES + k * 32, 0, 1024),
                              make_desc(sQ + i * TILE_BYTES + k * 32, 0, 1024), idesc, k > 0);
#pragma unroll
                for (int k = 0; k < HS / 16; ++k)  // dP^T = V_j dO_i^T
                    umma_bf16(tmem_base + cDPT, make_desc(sV + j * TILE_BYTES + k * 32, 0, 1024),
                              make_desc(sdO + i * TILE_BYTES + k * 32, 0, 1024), idesc, k > 0);
                umma_commit(bar_s);  // also covers the dV/dK/dQ MMAs of the previous iteration (they read sPT / sdST)
            }
            mbar_wait(bar_s, s_phase);
            s_phase ^= 1u;
            tc_fence_after();
            // columns (queries) are split between the two threads that share a TMEM lane
            const int nchunks = (nq16 + 31) >> 5;
            const int ch0 = half == 0 ? 0 : (nchunks + 1) >> 1, ch1 = half == 0 ? (nchunks + 1) >> 1 : nchunks;
            for (int ch = ch0; ch < ch1; ++ch) {
                uint32_t s[32], dp[32];
                tmem_ld32(tmem_base + lane_off + cST + ch * 32, s);
                tmem_ld32(tmem_base + lane_off + cDPT + ch * 32, dp);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float p[8], ds[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int qc = ch * 32 + g * 8 + c;  // query within the tile
                        const int q = i * TILE + qc;
                        const bool live = key < T && qc < nq && (!causal || key <= q);
                        const float pv = live ? ex2(__uint_as_float(s[g * 8 + c]) * sl2 - stat[q & 255]) : 0.f;
                        p[c] = pv;
                        ds[c] = pv * (__uint_as_float(dp[g * 8 + c]) - stat[256 + (q & 255)]) * scale;
                    }
                    const int qcol = ch * 32 + g * 8;
                    const uint32_t off = (qcol >> 6) * TILE_BYTES;
                    const int c8 = (qcol & 63) >> 3;
                    st_shared_v4(sw128(sPT + off, r, c8), pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]),
                                 pack_bf16(p[6], p[7]));
                    st_shared_v4(sw128(sdST + off, r, c8), pack_bf16(ds[0], ds[1]), pack_bf16(ds[2], ds[3]), pack_bf16(ds[4], ds[5]),
                                 pack_bf16(ds[6], ds[7]));
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (leader) {
                tc_fence_after();
                const uint32_t idesc_kk = make_idesc(TILE, HS, 0, 1);  // A K-major (P^T / dS^T rows = keys), B MN-major
                for (int k16 = 0; k16 < nq16 / 16; ++k16) {            // contraction over the queries of tile i
                    const uint32_t aoff = (k16 >> 2) * TILE_BYTES + (k16 & 3) * 32;
                    umma_bf16(tmem_base + cDV, make_desc(sPT + aoff, 0, 1024), make_desc(sdO + i * TILE_BYTES + k16 * 2048, TILE_BYTES, 1024),
                              idesc_kk, (!first_i || k16 > 0) ? 1u : 0u);
                    umma_bf16(tmem_base + cDK, make_desc(sdST + aoff, 0, 1024), make_desc(sQ + i * TILE_BYTES + k16 * 2048, TILE_BYTES, 1024),
                              idesc_kk, (!first_i || k16 > 0) ? 1u : 0u);
                }
                // dQ_i += dS K_j: A = dS^T read MN-major (m = query contiguous, 64-query atoms TILE_BYTES apart), contraction over keys
                const uint32_t idesc_mn = make_idesc(TILE, HS, 1, 1);
                for (int k16 = 0; k16 < nk16 / 16; ++k16)
                    umma_bf16(tmem_base + cDQ + i * HS, make_desc(sdST + k16 * 2048, TILE_BYTES, 1024),
                              make_desc(sK + j * TILE_BYTES + k16 * 2048, TILE_BYTES, 1024), idesc_mn, (((dq_started >> i) & 1u) || k16 > 0) ? 1u : 0u);
                dq_started |= 1u << i;
            }
            first_i = false;
        }
        // dV_j, dK_j are complete once every MMA issued so far has retired
        if (leader) umma_commit(bar_acc);
        mbar_wait(bar_acc, acc_phase);
        acc_phase ^= 1u;
        tc_fence_after();
        {
            uint32_t v[32];
            tmem_ld32(tmem_base + lane_off + cDV + half * 32, v);
            if (key < T) store_row32(dqkv + ((long)b * T + key) * 3 * C + 2 * C + h * HS + half * 32, v, 1.0f, accumulate);
            tmem_ld32(tmem_base + lane_off + cDK + half * 32, v);
            if (key < T) store_row32(dqkv + ((long)b * T + key) * 3 * C + C + h * HS + half * 32, v, 1.0f, accumulate);
        }
        tc_fence_before();  // the next key tile's first dV/dK MMA overwrites these columns after the next __syncthreads
    }
    // dQ: all MMAs were covered by the last bar_acc commit
    for (int i = 0; i < NT; ++i) {
        const int q = i * TILE + r;
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_off + cDQ + i * HS + half * 32, v);
        if (q < T) store_row32(dqkv + ((long)b * T + q) * 3 * C + h * HS + half * 32, v, 1.0f, accumulate);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}


// ---- backward, software-pipelined and persistent (non-causal, T <= 256) -----------------------------------
// Same five matmuls as attn_bwd_tc_kernel, but the work is cut into (128-key tile j) x (64-query sub-tile s)
// iterations and four agents run concurrently:
//   issuer (warp 0)      : the whole warp runs the control flow converged and every step's tcgen05 instructions leave from ONE
//                          elected region, back to back (an `if (lane == 0)` region costs 94 cycles per MMA in R2UR / vote
//                          wrappers, an elect per MMA 76; the pipe itself takes 42 cycles for a 128x64x16 MMA whose A operand
//                          is in tensor memory and 74 from shared memory: scripts/exp_mma_rate.cu).  S^T / dP^T of
//                          iteration n+2 are queued right behind dV / dK of iteration n, dQ when a query tile is complete.
//   SIMT group A / B     : 128 threads each (one per key = TMEM lane), alternate iterations; P^T and dS^T are packed to bf16
//                          in place in tensor memory (the A operands of dV and dK); dS^T also goes to one of four shared
//                          tiles, which pairwise are the MN-major A operand of dQ.  The groups do nothing else.
//   read-out group       : 128 threads, takes dV_j / dK_j (per key tile) and dQ_0 / dQ_1 (per head) out of tensor memory,
//                          hands the accumulators back as soon as the values are in registers, then converts, stages
//                          (dV / dK through the dead V_j / K_j tiles, dQ through a tile of its own) and TMA-stores.
//   loader / statistics  : one thread refills operand tiles for the next head as they die, one warp fetches lse / D a head ahead.
// TMEM: 2 x {S^T 64, dP^T 64} + dV 64 + dK 64 + dQ_0 64 + dQ_1 64 = 512 columns.
constexpr int kPipeThreads = 512;  // warps 0-3: control (issuer = warp 0, TMEM owner = warp 1), 4-7: group A, 8-11: group B, 12-15: read-out

// Sub-tile processed at position t of key tile j.  Iterations alternate between the two SIMT groups; with an even number of
// sub-tiles of which the last is short (T = 197: 64, 64, 64, 5 queries) the same group would get the short one in every key
// tile, so odd key tiles swap their last two sub-tiles (0, 1, 3, 2): both groups then see 208 query columns in total.
__device__ __forceinline__ int sub_at(int j, int t, int nsub) {
    return ((j & 1) && !(nsub & 1) && t >= nsub - 2) ? 2 * nsub - 3 - t : t;
}


// ---- the persistent kernel: one CTA per SM walks over its (batch, head) pairs ----------
// A one-CTA-per-head version of this pipeline spent 2.9 of its 14.5 us (ViT-B/16) on being launched, allocating tensor memory
// and waiting for its 128 KB of operands with nothing else resident on the SM (the shared memory it needs excludes a second
// CTA).  Here the CTA stays, and a loader thread refills each operand tile for the next head as soon as the current head is done with it:
// K_0 / V_0 once the first key tile's dV / dK have left through them, rows 0-127 of Q and dO when the last key tile is through
// its first two sub-tiles, the other rows when the head's last MMA has retired, K_1 / V_1 (on a barrier of their own: they are
// not needed before key tile 1) after the read-out group's stores of that tile.  Row statistics are double-buffered.  Barrier
// parities are derived from the running head index G (every barrier completes a fixed number of phases per head).
__global__ void __launch_bounds__(kPipeThreads, 1)
attn_bwd_persist_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                        const __grid_constant__ CUtensorMap tm_dqkv, bf16* __restrict__ dqkv, const float* __restrict__ lse,
                        const float* __restrict__ dsum, int T, int C, int NH, int NT, int accumulate, int total_heads) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sQ = base, sK = sQ + NT * TILE_BYTES, sV = sK + NT * TILE_BYTES, sdO = sV + NT * TILE_BYTES;
    const uint32_t sdS = sdO + NT * TILE_BYTES;            // 4 x [128 keys][64 queries] bf16, 128B-swizzled
    const uint32_t sDQ = sdS + 4 * TILE_BYTES;             // staging tile of the dQ stores
    const uint32_t sStat = sDQ + TILE_BYTES;               // 2 x {lse*log2e [256], D [256]}
    const uint32_t bar0 = sStat + 2 * 2 * 256 * 4;
    const uint32_t load0 = bar0, load1 = bar0 + 8, s_full = bar0 + 16, p_full = bar0 + 32, ds_free = bar0 + 48,  // .., .., [2], [2], [4]
                   acc_full = bar0 + 80, acc_free = bar0 + 88, dq_full = bar0 + 96, dq_free = bar0 + 104, free_kv = bar0 + 112,  // free_kv[2]
                   stat_full = bar0 + 128, stat_free = bar0 + 144, q0_free = bar0 + 160, load2 = bar0 + 168;                       // [2], [2], .., ..
    float* stat_all = reinterpret_cast<float*>(gen + (sStat - base));
    volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(gen + (bar0 - base) + 176);
    constexpr uint32_t TMEM_COLS = 512, cDV = 256, cDK = 320, cDQ = 384;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform: role branches stay converged
    TR_DECL
    const int NSUB = (T + SUB - 1) / SUB;   // 64-query sub-tiles
    const int N = NT * NSUB;                // iterations per head: n = j * NSUB + t
    const int nheads = (total_heads - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // heads of this CTA
    if (tid == 0) {
        tma_prefetch_desc(&tm_qkv);
        tma_prefetch_desc(&tm_do);
        tma_prefetch_desc(&tm_dqkv);
        mbar_init(load0, 1);
        mbar_init(load1, 1);
        mbar_init(load2, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(s_full + 8 * i, 1);
            mbar_init(p_full + 8 * i, 128);
            mbar_init(free_kv + 8 * i, 1);
            mbar_init(stat_full + 8 * i, 1);
            mbar_init(stat_free + 8 * i, 256);
        }
        for (int i = 0; i < 4; ++i) mbar_init(ds_free + 8 * i, 1);
        mbar_init(acc_full, 1);
        mbar_init(acc_free, 128);
        mbar_init(dq_full, 1);
        mbar_init(dq_free, 128);
        mbar_init(q0_free, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32((const void*)slot), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *slot, 0);
    const float scale = 1.0f / sqrtf((float)HS);
    const float sl2 = kLog2e * scale;

    if (warp == 2) {
        if (lane == 0) {
            // ================================ loader ================================
            for (int G = 0; G < nheads; ++G) {
                const int bh = (int)blockIdx.x + G * (int)gridDim.x, b = bh / NH, h = bh - b * NH;
                const uint32_t par = (uint32_t)((G - 1) & 1);  // parity of the previous head's single-phase barriers
                TR(30, G);
                if (G > 0) mbar_wait(free_kv, par);            // dV_0 / dK_0 of the previous head have left through V_0 / K_0
                TR(31, G);
                mbar_expect_tx(load0, (uint32_t)(4 * TILE_BYTES));
                tma_load_3d(sK, &tm_qkv, load0, C + h * HS, 0, b);
                tma_load_3d(sV, &tm_qkv, load0, 2 * C + h * HS, 0, b);
                // the first 128 rows of Q and dO are dead once the last key tile is through its first two sub-tiles, well before
                // the head ends: the next head's first scores can then be issued right behind this head's last MMAs
                if (G > 0) mbar_wait(q0_free, par);
                TR(32, G);
                tma_load_3d(sQ, &tm_qkv, load0, h * HS, 0, b);
                tma_load_3d(sdO, &tm_do, load0, h * HS, 0, b);
                if (NT > 1) {
                    if (G > 0) mbar_wait(dq_full, par);        // every MMA of the previous head has retired: the other rows are dead too
                    // rows >= 128 of Q / dO (first needed by the third sub-tile of key tile 0) and K_1 / V_1 (first needed by key
                    // tile 1) complete on barriers of their own: K_1 / V_1 wait for the read-out group's stores of the previous head
                    mbar_expect_tx(load1, (uint32_t)(2 * TILE_BYTES));
                    tma_load_3d(sQ + TILE_BYTES, &tm_qkv, load1, h * HS, TILE, b);
                    tma_load_3d(sdO + TILE_BYTES, &tm_do, load1, h * HS, TILE, b);
                    TR(33, G);
                    if (G > 0) mbar_wait(free_kv + 8, par);
                    TR(34, G);
                    mbar_expect_tx(load2, (uint32_t)(2 * TILE_BYTES));
                    tma_load_3d(sK + TILE_BYTES, &tm_qkv, load2, C + h * HS, TILE, b);
                    tma_load_3d(sV + TILE_BYTES, &tm_qkv, load2, 2 * C + h * HS, TILE, b);
                }
            }
        }
    } else if (warp == 3) {
        // ================================ row statistics, one head ahead ================================
        for (int G = 0; G < nheads; ++G) {
            const int bh = (int)blockIdx.x + G * (int)gridDim.x, sb = G & 1;
            if (G >= 2) mbar_wait(stat_free + 8 * sb, (uint32_t)(((G >> 1) - 1) & 1));
            float* st = stat_all + sb * 512;
            for (int i = lane; i < 256; i += 32) {
                st[i] = i < T ? -lse[(long)bh * T + i] * kLog2e : 0.f;     // P = exp2(S * scale * log2e + st[i])
                st[256 + i] = i < T ? -dsum[(long)bh * T + i] * scale : 0.f;  // dS = P * (dP * scale + st[256 + i])
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_cnt(stat_full + 8 * sb);
        }
    } else if (warp == 0) {
        {
            // ================================ issuer (whole warp converged, tcgen05 instructions on the elected lane) ================
            const uint32_t idesc_kk = make_idesc(TILE, HS, 0, 1);
            const uint32_t idesc_mn = make_idesc(TILE, HS, 1, 1);
            const uint64_t dK_k = make_desc(sK, 0, 1024), dV_k = make_desc(sV, 0, 1024);          // K-major A operands (rows = keys)
            const uint64_t dQ_k = make_desc(sQ, 0, 1024), ddO_k = make_desc(sdO, 0, 1024);        // K-major B operands (rows = queries)
            const uint64_t dQ_mn = make_desc(sQ, TILE_BYTES, 1024), ddO_mn = make_desc(sdO, TILE_BYTES, 1024);  // MN-major B operands
            const uint64_t dK_mn = make_desc(sK, TILE_BYTES, 1024);
            const uint64_t ddS_mn = make_desc(sdS, TILE_BYTES, 1024);  // (dK takes dS^T from tensor memory)
            auto off = [](uint32_t bytes) { return (uint64_t)(bytes >> 4); };
            // Every wait is taken by the converged warp; the tcgen05 instructions of one step then go out from ONE elected region,
            // back to back (measured, scripts/exp_mma_rate.cu: an elect per MMA costs 76 cycles per MMA whatever its size, a
            // threadIdx test 94; inside one region a 128x64x16 MMA with A in tensor memory takes 42 and with A in shared memory 74).
            for (int G = 0; G < nheads; ++G) {
                const uint32_t gpar = (uint32_t)(G & 1);
                bool q1_ready = false, kv1_ready = false, dq_ready = G == 0;
                uint32_t done_mask = 0u;  // sub-tiles of the current key tile whose dS^T is in shared memory
                auto wait_tile1 = [&](int n) {  // operands iteration n is the first to touch: rows >= 128 of Q / dO, K_1 / V_1
                    if (n >= N) return;
                    const int j = n / NSUB, s_ = sub_at(j, n - j * NSUB, NSUB);
                    if (!q1_ready && s_ >= TILE / SUB) {
                        mbar_wait(load1, gpar);
                        q1_ready = true;
                    }
                    if (!kv1_ready && j > 0) {
                        mbar_wait(load2, gpar);
                        kv1_ready = true;
                    }
                    tc_fence_after();
                };
                auto issue_scores = [&](int n) {  // (elected lane) S^T = K_j Q_s^T and dP^T = V_j dO_s^T of iteration n
                    const int j = n / NSUB, s_ = sub_at(j, n - j * NSUB, NSUB), bx = n & 1;
                    const int nq16 = (min(SUB, T - s_ * SUB) + 15) & ~15;
                    const uint32_t idesc = make_idesc(TILE, nq16, 0, 0);
                    const uint64_t qo = off((s_ >> 1) * TILE_BYTES + (s_ & 1) * SUB_BYTES), ko = off(j * TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < HS / 16; ++k) umma_bf16(tmem_base + bx * 128, dK_k + ko + 2 * k, dQ_k + qo + 2 * k, idesc, k > 0);
#pragma unroll
                    for (int k = 0; k < HS / 16; ++k) umma_bf16(tmem_base + bx * 128 + 64, dV_k + ko + 2 * k, ddO_k + qo + 2 * k, idesc, k > 0);
                    umma_commit(s_full + 8 * bx);
                };
                TR(22, G);
                mbar_wait(load0, gpar);
                tc_fence_after();
                TR(23, G);
                wait_tile1(0);
                wait_tile1(1);
                if (elect_one()) {
                    issue_scores(0);
                    if (N > 1) issue_scores(1);
                }
                __syncwarp();
                for (int m = 0; m < N; ++m) {
                    const int j = m / NSUB, t_ = m - j * NSUB, s_ = sub_at(j, t_, NSUB), bx = m & 1;
                    const int nq16 = (min(SUB, T - s_ * SUB) + 15) & ~15;
                    const int nk16 = (min(TILE, T - j * TILE) + 15) & ~15;
                    const uint64_t qo = off((s_ >> 1) * TILE_BYTES + (s_ & 1) * SUB_BYTES);
                    if (t_ == 0) done_mask = 0u;
                    const int per_buf = (N + 1 - bx) >> 1;  // iterations per head on buffer bx
                    done_mask |= 1u << s_;
                    const int partner = s_ ^ 1;
                    const bool pair_done = partner >= NSUB || ((done_mask >> partner) & 1u);  // query tile i = s/2 complete for this key tile
                    TR(20, m);
                    mbar_wait(p_full + 8 * bx, (uint32_t)((G * per_buf + (m >> 1)) & 1));
                    TR(21, m);
                    if (t_ == 0 && G * NT + j > 0) mbar_wait(acc_free, (uint32_t)((G * NT + j - 1) & 1));  // previous key tile's dV / dK read out
                    wait_tile1(m + 2);
                    if (pair_done && !dq_ready) {  // the previous head's dQ accumulators have been read out
                        mbar_wait(dq_free, (uint32_t)((G - 1) & 1));
                        dq_ready = true;
                    }
                    tc_fence_after();
                    if (elect_one()) {
                        // dV_j += P^T dO_s and dK_j += dS^T Q_s: both A operands are the bf16 tiles the group packed in place in tensor memory
                        for (int k16 = 0; k16 < nq16 / 16; ++k16) {
                            const uint32_t acc = (t_ > 0 || k16 > 0) ? 1u : 0u;
                            umma_bf16_ts(tmem_base + cDV, tmem_base + bx * 128 + k16 * 8, ddO_mn + qo + 128 * k16, idesc_kk, acc);
                            umma_bf16_ts(tmem_base + cDK, tmem_base + bx * 128 + 64 + k16 * 8, dQ_mn + qo + 128 * k16, idesc_kk, acc);
                        }
                        if (m + 2 < N) issue_scores(m + 2);
                        // last key tile, both sub-tiles of query tile 0 issued: no later MMA of this head reads Q_0 / dO_0
                        if (j == NT - 1 && s_ < 2 && pair_done) umma_commit(q0_free);
                        if (pair_done) {  // dQ_i += dS K_j: A = dS^T read MN-major from the groups' shared tiles
                            const int i = s_ >> 1;
                            const uint64_t ao = off((2 * i) * TILE_BYTES), ko = off(j * TILE_BYTES);
                            for (int k16 = 0; k16 < nk16 / 16; ++k16)
                                umma_bf16(tmem_base + cDQ + i * HS, ddS_mn + ao + 128 * k16, dK_mn + ko + 128 * k16, idesc_mn, (j > 0 || k16 > 0) ? 1u : 0u);
                            umma_commit(ds_free + 8 * ((2 * i) & 3));
                            if (2 * i + 1 < NSUB) umma_commit(ds_free + 8 * ((2 * i + 1) & 3));
                        }
                        if (t_ == NSUB - 1) umma_commit(acc_full);
                    }
                    __syncwarp();
                    TR(24, m);
                }
                if (elect_one()) umma_commit(dq_full);
                __syncwarp();
            }
        }
    } else if (warp >= 12) {
        // ================================ read-out group (warps 12-15, one per TMEM lane quarter) ================================
        // Takes every accumulator out of tensor memory so that the SIMT groups never leave their iterations: dV_j / dK_j when a key
        // tile completes (acc_full), dQ_0 / dQ_1 when the head completes (dq_full).  The accumulators are handed back to the
        // issuer (acc_free / dq_free) as soon as their values sit in registers; conversion, staging and the TMA stores follow
        // off the critical path.  dV_j / dK_j leave through the dead V_j / K_j tiles, dQ through a staging tile of its own.
        const int r = (warp & 3) * 32 + lane;
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const bool leader = warp == 12 && lane == 0;
        bool stage_busy = false;  // (group-uniform) a dQ store may still be reading the staging tile
        auto pack32 = [&](uint32_t (&v)[32], uint32_t (&o)[16]) {
#pragma unroll
            for (int c = 0; c < 16; ++c) o[c] = pack_bf16(__uint_as_float(v[2 * c]), __uint_as_float(v[2 * c + 1]));
        };
        // 64 packed bf16 values (this thread's row of a [128][64] tile) -> (+ what global memory holds) -> swizzled staging tile
        auto stage_row = [&](uint32_t tile, const uint32_t (&lo)[16], const uint32_t (&hi)[16], const bf16* old) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4) {
                    uint32_t w[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) w[c] = half == 0 ? lo[g4 * 4 + c] : hi[g4 * 4 + c];
                    if (old) {
                        Vec16<bf16> o;
                        o.load(old + half * 32 + g4 * 8);
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            w[c] = pack_bf16(__uint_as_float(w[c] << 16) + o.get(2 * c), __uint_as_float(w[c] & 0xFFFF0000u) + o.get(2 * c + 1));
                    }
                    st_shared_v4(sw128(tile, r, half * 4 + g4), w[0], w[1], w[2], w[3]);
                }
            }
        };
        auto read64 = [&](uint32_t col, uint32_t (&lo)[16], uint32_t (&hi)[16]) {  // 64 fp32 accumulator columns -> packed bf16
            uint32_t v[32];
            tmem_ld32(tmem_base + lane_off + col, v);
            pack32(v, lo);
            tmem_ld32(tmem_base + lane_off + col + 32, v);
            pack32(v, hi);
        };
        for (int G = 0; G < nheads; ++G) {
            const int bh = (int)blockIdx.x + G * (int)gridDim.x, b = bh / NH, h = bh - b * NH;
            for (int j = 0; j < NT; ++j) {
                const int key = j * TILE + r;
                uint32_t vlo[16], vhi[16], klo[16], khi[16];
                TR(5, j);
                mbar_wait(acc_full, (uint32_t)((G * NT + j) & 1));
                tc_fence_after();
                TR(6, j);
                read64(cDV, vlo, vhi);
                read64(cDK, klo, khi);
                tc_fence_before();
                mbar_arrive_cnt(acc_free);  // the next key tile's dV / dK MMAs may start
                const bf16* old = (accumulate && key < T) ? dqkv + ((long)b * T + key) * 3 * C + h * HS : nullptr;
                // every MMA that read K_j / V_j has retired (acc_full covers the key tile's scores and its dQ MMAs)
                stage_row(sV + j * TILE_BYTES, vlo, vhi, old ? old + 2 * C : nullptr);
                stage_row(sK + j * TILE_BYTES, klo, khi, old ? old + C : nullptr);
                fence_proxy_async();
                asm volatile("bar.sync 2, 128;" ::: "memory");
                if (leader) {
                    tma_store_3d(&tm_dqkv, sV + j * TILE_BYTES, 2 * C + h * HS, j * TILE, b);
                    tma_store_3d(&tm_dqkv, sK + j * TILE_BYTES, C + h * HS, j * TILE, b);
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // (also covers an earlier dQ store)
                    mbar_arrive_cnt(free_kv + 8 * j);  // the loader may refill K_j / V_j for the next head
                }
                stage_busy = false;  // (the leader has just waited for every earlier store's reads; the others pass bar 2 before staging again)
                TR(7, j);
            }
            // ---- dQ_0, dQ_1 ----
            {
                uint32_t alo[16], ahi[16], blo[16], bhi[16];
                TR(8, G);
                mbar_wait(dq_full, (uint32_t)(G & 1));
                tc_fence_after();
                TR(9, G);
                read64(cDQ, alo, ahi);
                if (NT > 1) read64(cDQ + HS, blo, bhi);
                tc_fence_before();
                mbar_arrive_cnt(dq_free);  // the next head's dQ MMAs may start
                auto store_dq = [&](int i, const uint32_t (&lo)[16], const uint32_t (&hi)[16]) {
                    const int q = i * TILE + r;
                    if (stage_busy) {
                        if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        asm volatile("bar.sync 2, 128;" ::: "memory");
                    }
                    stage_row(sDQ, lo, hi, (accumulate && q < T) ? dqkv + ((long)b * T + q) * 3 * C + h * HS : nullptr);
                    fence_proxy_async();
                    asm volatile("bar.sync 2, 128;" ::: "memory");
                    if (leader) tma_store_3d(&tm_dqkv, sDQ, h * HS, i * TILE, b);
                    stage_busy = true;
                };
                store_dq(0, alo, ahi);
                if (NT > 1) store_dq(1, blo, bhi);
                TR(10, G);
            }
        }
        if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // every store is complete before shared memory is retired
    } else if (warp >= 4) {
        // ================================ SIMT groups ================================
        const int g = (warp - 4) >> 2;          // group 0 / 1
        const int r = (warp & 3) * 32 + lane;   // TMEM lane = key within the tile
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const int per_buf = (N + 1 - g) >> 1;   // iterations per head of this group
        for (int G = 0; G < nheads; ++G) {
            const float* stat = stat_all + (G & 1) * 512;
            TR(1, G);
            mbar_wait(stat_full + 8 * (G & 1), (uint32_t)((G >> 1) & 1));
            for (int j = 0; j < NT; ++j) {
                const int key = j * TILE + r;
                for (int t_ = 0; t_ < NSUB; ++t_) {
                    const int n = j * NSUB + t_;
                    if ((n & 1) != g) continue;
                    const int s_ = sub_at(j, t_, NSUB);
                    const int nq = min(SUB, T - s_ * SUB), nq16 = (nq + 15) & ~15;
                    const uint32_t xb = tmem_base + lane_off + (uint32_t)(g * 128);
                    TR(2, n);
                    mbar_wait(s_full + 8 * g, (uint32_t)((G * per_buf + (n >> 1)) & 1));
                    tc_fence_after();
                    const int bs = s_ & 3;
                    const uint32_t sbuf = sdS + bs * TILE_BYTES;
                    // the MMAs that read this dS tile last time (previous key tile, or the previous head) have retired
                    if (G * NT + j > 0) mbar_wait(ds_free + 8 * bs, (uint32_t)((G * NT + j - 1) & 1));
                    TR(3, n);
                    const int nch = (nq16 + 31) >> 5;
                    for (int ch = 0; ch < nch; ++ch) {
                        uint32_t sv[32], dp[32], pk[16], dk[16];
                        tmem_ld32(xb + ch * 32, sv);
                        tmem_ld32(xb + 64 + ch * 32, dp);
                        const int q0 = s_ * SUB + ch * 32;             // first query of the chunk (q0 + 31 < 256)
                        const bool full = key < T && ch * 32 + 32 <= nq;  // warp-uniform except for the key tail: no per-element masks
#pragma unroll
                        for (int g8 = 0; g8 < 4; ++g8) {
                            const float4 l0 = *reinterpret_cast<const float4*>(stat + q0 + g8 * 8), l1 = *reinterpret_cast<const float4*>(stat + q0 + g8 * 8 + 4);
                            const float4 d0 = *reinterpret_cast<const float4*>(stat + 256 + q0 + g8 * 8), d1 = *reinterpret_cast<const float4*>(stat + 256 + q0 + g8 * 8 + 4);
                            const float lq[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
                            const float dq[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
                            float ds[8], pv[8];
                            // packed f32x2 FMAs: three arithmetic instructions per pair of scores instead of eight
#pragma unroll
                            for (int c = 0; c < 8; c += 2) {
                                const int qc = ch * 32 + g8 * 8 + c;
                                const float2 a = fma2(make_float2(__uint_as_float(sv[g8 * 8 + c]), __uint_as_float(sv[g8 * 8 + c + 1])), splat2(sl2),
                                                      make_float2(lq[c], lq[c + 1]));
                                const float e0 = ex2(a.x), e1 = ex2(a.y);
                                pv[c] = (full || (key < T && qc < nq)) ? e0 : 0.f;
                                pv[c + 1] = (full || (key < T && qc + 1 < nq)) ? e1 : 0.f;
                                const float2 t = fma2(make_float2(__uint_as_float(dp[g8 * 8 + c]), __uint_as_float(dp[g8 * 8 + c + 1])), splat2(scale),
                                                      make_float2(dq[c], dq[c + 1]));
                                const float2 d2 = mul2(make_float2(pv[c], pv[c + 1]), t);
                                ds[c] = d2.x;
                                ds[c + 1] = d2.y;
                            }
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                pk[g8 * 4 + c] = pack_bf16(pv[2 * c], pv[2 * c + 1]);
                                dk[g8 * 4 + c] = pack_bf16(ds[2 * c], ds[2 * c + 1]);
                            }
                            st_shared_v4(sw128(sbuf, r, ch * 4 + g8), dk[g8 * 4], dk[g8 * 4 + 1], dk[g8 * 4 + 2], dk[g8 * 4 + 3]);
                        }
                        tmem_st16(xb + ch * 16, pk);       // P^T in place: columns [16ch, 16ch+16) were consumed by chunk <= ch
                        tmem_st16(xb + 64 + ch * 16, dk);  // dS^T likewise over dP^T: the A operand of dK (shared memory keeps the copy dQ reads)
                    }
                    tmem_st_wait();
                    fence_proxy_async();
                    tc_fence_before();
                    mbar_arrive_cnt(p_full + 8 * g);
                    TR(4, n);
                }
            }
            mbar_arrive_cnt(stat_free + 8 * (G & 1));  // this head's row statistics are no longer read
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// =====================================================================================================
// Streaming kernels: any sequence length (ViT-B/8: T = 785), two CTAs per SM, every probability operand in
// tensor memory.  Forward streams 128-key K/V tiles with a single-pass softmax.  Backward is split in two
// kernels so that nothing is reduced through global memory and each CTA needs only 256 TMEM columns:
//   dKV: one CTA per (batch, head, 128-key tile), keys on TMEM lanes, streams 64-query tiles of Q and dO;
//   dQ : one CTA per (batch, head, 128-query tile), queries on TMEM lanes, streams 64-key tiles of K and V.
// S and dP are recomputed in both (7 matmuls instead of 5); P / dS are packed to bf16 in place in TMEM and
// consumed as the A operand of the next MMA, so no probability ever touches shared or global memory.
// =====================================================================================================

__global__ void __launch_bounds__(128, 2)
attn_fwd_stream_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_out, float* __restrict__ lse,
                       int T, int C, int NH, int causal) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sQ = base, sK = sQ + TILE_BYTES, sV = sK + 2 * TILE_BYTES;
    const uint32_t bar0 = sV + 2 * TILE_BYTES;
    const uint32_t bar_q = bar0, bar_k = bar0 + 8, bar_v = bar0 + 24, bar_s = bar0 + 40, bar_o = bar0 + 48;  // bar_k[2], bar_v[2]
    volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(gen + (bar0 - base) + 56);
    constexpr uint32_t TMEM_COLS = 256, cO = 192;  // S [0,128) -> P packed [0,64); O [192,256)

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform: warp 0 issues the tcgen05 instructions converged
    const int NT = (T + TILE - 1) / TILE;
    const int qt = blockIdx.x % NT, bh = blockIdx.x / NT, b = bh / NH, h = bh - b * NH;
    const int NJ = causal ? qt + 1 : NT;
    if (tid == 0) {
        tma_prefetch_desc(&tm_qkv);
        tma_prefetch_desc(&tm_out);
        mbar_init(bar_q, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(bar_k + 8 * i, 1); mbar_init(bar_v + 8 * i, 1); }
        mbar_init(bar_s, 1);
        mbar_init(bar_o, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(smem_u32((const void*)slot), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *slot, 0);
    if (tid == 0) {
        mbar_expect_tx(bar_q, TILE_BYTES);
        tma_load_3d(sQ, &tm_qkv, bar_q, h * HS, qt * TILE, b);
        for (int j = 0; j < 2 && j < NJ; ++j) {
            mbar_expect_tx(bar_k + 8 * j, TILE_BYTES);
            tma_load_3d(sK + j * TILE_BYTES, &tm_qkv, bar_k + 8 * j, C + h * HS, j * TILE, b);
            mbar_expect_tx(bar_v + 8 * j, TILE_BYTES);
            tma_load_3d(sV + j * TILE_BYTES, &tm_qkv, bar_v + 8 * j, 2 * C + h * HS, j * TILE, b);
        }
    }
    const int r = tid;
    const int q = qt * TILE + r;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int kend = causal ? min(T, q + 1) : T;
    const float scale = 1.0f / sqrtf((float)HS);
    const float sl2 = kLog2e * scale;
    // ONE pass over each score tile: tensor-memory reads (64 B / clock / SM) are what this kernel is made of, so neither a maximum
    // pass nor a per-tile rescale of O is taken.  The exponent reference m_ref (an integer, log2 domain) is the maximum of the
    // row's first 32 keys rounded up; every probability is 2^(s - m_ref), P and O simply grow with it (bf16 and fp32 share
    // their exponent range), and only a chunk whose probabilities sum past 2^64 moves the reference: what the row has
    // written so far — this tile's P, the running sum and the O accumulator — is then rescaled by an exact power of two.
    float m_ref = 0.f, l_run = 0.f;
    // keys below this bound are unmasked for every row of the warp: chunks entirely below it run without per-element masks
    const int kfull = causal ? min(T, qt * TILE + warp * 32 + 1) : T;

    for (int j = 0; j < NJ; ++j) {
        const int buf = j & 1;
        const uint32_t ph = (uint32_t)((j >> 1) & 1);
        const int k0 = j * TILE;
        const int nk16 = (min(TILE, T - k0) + 15) & ~15;
        if (warp == 0) {  // waits by the converged warp, the MMAs back to back from one elected region (scripts/exp_mma_rate.cu)
            if (j == 0) mbar_wait(bar_q, 0);
            mbar_wait(bar_k + 8 * buf, ph);
            tc_fence_after();
            const uint32_t idesc = make_idesc(TILE, nk16, 0, 0);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < HS / 16; ++k)
                    umma_bf16(tmem_base, make_desc(sQ + k * 32, 0, 1024), make_desc(sK + buf * TILE_BYTES + k * 32, 0, 1024), idesc, k > 0);
                umma_commit(bar_s);
            }
            __syncwarp();
        }
        mbar_wait(bar_s, (uint32_t)(j & 1));
        tc_fence_after();
        if (tid == 0 && j + 2 < NJ) {  // the K buffer is free again
            mbar_expect_tx(bar_k + 8 * buf, TILE_BYTES);
            tma_load_3d(sK + buf * TILE_BYTES, &tm_qkv, bar_k + 8 * buf, C + h * HS, (j + 2) * TILE, b);
        }
        const int nchunks = (nk16 + 31) >> 5;  // (a last chunk of 16 columns is read as 32: the stale columns are masked, they lie beyond T)
        bool o_settled = j == 0;  // P_{j-1} V_{j-1} has landed in O (known only once bar_o has been waited for)
        uint32_t v[32];
        tmem_ld32_issue(lane_addr, v);
        tmem_ld_wait32(v);
        if (j == 0) {
            float cm = -INFINITY;
#pragma unroll
            for (int c = 0; c < 32; ++c)
                if (c < kend) cm = fmaxf(cm, __uint_as_float(v[c]));  // (key 0 is never masked)
            m_ref = ceilf(cm * sl2);
        }
        for (int ch = 0; ch < nchunks; ++ch) {
            const int kc = k0 + ch * 32;
            const bool full = kc + 32 <= kfull;  // warp-uniform
            const bool has_next = ch + 1 < nchunks;
            float2 a[16];
            uint32_t pk[16];
            float csum;
            auto scale_scores = [&]() {
                const float2 sl22 = splat2(sl2), nref2 = splat2(-m_ref);
#pragma unroll
                for (int c = 0; c < 16; ++c) a[c] = fma2(make_float2(__uint_as_float(v[2 * c]), __uint_as_float(v[2 * c + 1])), sl22, nref2);
            };
            auto exponentiate = [&]() {
                float s0 = 0.f, s1 = 0.f;
                if (full) {
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const float p0 = ex2(a[c].x), p1 = ex2(a[c].y);
                        s0 += p0;
                        s1 += p1;
                        pk[c] = pack_bf16(p0, p1);
                    }
                } else {
                    const int live = kend - kc;  // per thread (causal: the diagonal)
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const float p0 = 2 * c < live ? ex2(a[c].x) : 0.f, p1 = 2 * c + 1 < live ? ex2(a[c].y) : 0.f;
                        s0 += p0;
                        s1 += p1;
                        pk[c] = pack_bf16(p0, p1);
                    }
                }
                csum = s0 + s1;
            };
            scale_scores();
            if (has_next) tmem_ld32_issue(lane_addr + (ch + 1) * 32, v);  // v is dead: the next chunk's load overlaps the exponentials
            exponentiate();
            if (__any_sync(0xffffffffu, !(csum < kPBig))) {
                // never in practice: move this row's reference to the chunk's maximum (its scores are still intact in tensor memory)
                if (has_next) tmem_ld_wait32(v);
                tmem_ld32_issue(lane_addr + ch * 32, v);
                tmem_ld_wait32(v);
                float cm = -INFINITY;
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    if (kc + c < kend) cm = fmaxf(cm, __uint_as_float(v[c]));
                const float new_ref = !(csum < kPBig) ? ceilf(cm * sl2) : m_ref;
                const float f = ex2(m_ref - new_ref);  // 2^(integer <= 0): exact
                tmem_st_wait();
                for (int blk = 0; blk < ch; ++blk) {  // this tile's probabilities written so far
                    uint32_t old[16];
                    tmem_ld16(lane_addr + blk * 16, old);
#pragma unroll
                    for (int c = 0; c < 16; ++c)
                        old[c] = pack_bf16(__uint_as_float(old[c] << 16) * f, __uint_as_float(old[c] & 0xFFFF0000u) * f);
                    tmem_st16(lane_addr + blk * 16, old);
                }
                if (j > 0) {  // the accumulator of the earlier tiles
                    if (!o_settled) {
                        mbar_wait(bar_o, (uint32_t)((j - 1) & 1));
                        tc_fence_after();
                        o_settled = true;
                    }
                    for (int blk = 0; blk < HS / 16; ++blk) {
                        uint32_t o16[16];
                        tmem_ld16(lane_addr + cO + blk * 16, o16);
#pragma unroll
                        for (int c = 0; c < 16; ++c) o16[c] = __float_as_uint(__uint_as_float(o16[c]) * f);
                        tmem_st16(lane_addr + cO + blk * 16, o16);
                    }
                }
                l_run *= f;
                m_ref = new_ref;
                scale_scores();
                if (has_next) tmem_ld32_issue(lane_addr + (ch + 1) * 32, v);
                exponentiate();
            }
            tmem_st16(lane_addr + ch * 16, pk);  // in place: these columns held scores this thread has already read
            l_run += csum;
            if (has_next) tmem_ld_wait32(v);
        }
        if (j > 0) {
            if (!o_settled) mbar_wait(bar_o, (uint32_t)((j - 1) & 1));  // P_{j-1} V_{j-1} has landed in O
            tc_fence_after();
            if (tid == 0 && j + 1 < NJ) {  // its V buffer is free again
                mbar_expect_tx(bar_v + 8 * ((j + 1) & 1), TILE_BYTES);
                tma_load_3d(sV + ((j + 1) & 1) * TILE_BYTES, &tm_qkv, bar_v + 8 * ((j + 1) & 1), 2 * C + h * HS, (j + 1) * TILE, b);
            }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            mbar_wait(bar_v + 8 * buf, j == 1 ? 0u : ph);
            const uint32_t idesc = make_idesc(TILE, HS, 0, 1);
            if (elect_one()) {
                for (int k16 = 0; k16 < nk16 / 16; ++k16)
                    umma_bf16_ts(tmem_base + cO, tmem_base + k16 * 8, make_desc(sV + buf * TILE_BYTES + k16 * 2048, TILE_BYTES, 1024), idesc,
                                 (j > 0 || k16 > 0) ? 1u : 0u);
                umma_commit(bar_o);
            }
            __syncwarp();
        }
    }
    mbar_wait(bar_o, (uint32_t)((NJ - 1) & 1));
    tc_fence_after();
    const float inv = 1.0f / l_run;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld32(lane_addr + cO + half * 32, v);
#pragma unroll
        for (int g = 0; g < 4; ++g)
            st_shared_v4(sw128(sQ, r, half * 4 + g), pack_bf16(__uint_as_float(v[g * 8]) * inv, __uint_as_float(v[g * 8 + 1]) * inv),
                         pack_bf16(__uint_as_float(v[g * 8 + 2]) * inv, __uint_as_float(v[g * 8 + 3]) * inv),
                         pack_bf16(__uint_as_float(v[g * 8 + 4]) * inv, __uint_as_float(v[g * 8 + 5]) * inv),
                         pack_bf16(__uint_as_float(v[g * 8 + 6]) * inv, __uint_as_float(v[g * 8 + 7]) * inv));
    }
    if (q < T) lse[(long)bh * T + q] = (m_ref + __log2f(l_run)) * (1.0f / kLog2e);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tma_store_3d(&tm_out, sQ, h * HS, qt * TILE, b);
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

__global__ void __launch_bounds__(256, 2)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tm_kv, const __grid_constant__ CUtensorMap tm_q64, const __grid_constant__ CUtensorMap tm_do64,
                    const __grid_constant__ CUtensorMap tm_dqkv, bf16* __restrict__ dqkv, const float* __restrict__ lse,
                    const float* __restrict__ dsum, int T, int C, int NH, int causal, int accumulate) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sK = base, sV = sK + TILE_BYTES, sQ = sV + TILE_BYTES, sdO = sQ + 2 * SUB_BYTES;
    const uint32_t bar0 = sdO + 2 * SUB_BYTES;
    const uint32_t bar_kv = bar0, bar_q = bar0 + 8, bar_s = bar0 + 24, bar_acc = bar0 + 32;  // bar_q[2]: Q_s and dO_s of one buffer
    volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(gen + (bar0 - base) + 40);
    float* stat = reinterpret_cast<float*>(gen + (bar0 - base) + 64);  // lse*log2e [T], D [T]
    constexpr uint32_t TMEM_COLS = 256, cST = 0, cDPT = 64, cDV = 128, cDK = 192;

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform: warp 0 issues the tcgen05 instructions converged
    const int NT = (T + TILE - 1) / TILE, NS = (T + SUB - 1) / SUB;
    const int jt = blockIdx.x % NT, bh = blockIdx.x / NT, b = bh / NH, h = bh - b * NH;
    const int s0 = causal ? (jt * TILE) / SUB : 0;
    if (tid == 0) {
        tma_prefetch_desc(&tm_kv); tma_prefetch_desc(&tm_q64); tma_prefetch_desc(&tm_do64); tma_prefetch_desc(&tm_dqkv);
        mbar_init(bar_kv, 1); mbar_init(bar_q, 1); mbar_init(bar_q + 8, 1); mbar_init(bar_s, 1); mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(smem_u32((const void*)slot), TMEM_COLS);
    // row statistics of the whole head, negated and pre-scaled so that each is the addend of one FMA:
    // P = exp2(S * sl2 + stat[q]), dS = P * (dP * scale + stat[TP + q])
    const int TP = (T + SUB - 1) & ~(SUB - 1);
    for (int i = tid; i < TP; i += 256) {
        stat[i] = i < T ? -lse[(long)bh * T + i] * kLog2e : 0.f;
        stat[TP + i] = i < T ? -dsum[(long)bh * T + i] * (1.0f / sqrtf((float)HS)) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *slot, 0);
    const bool leader = tid == 0;
    auto load_sub = [&](int s, int buf) {
        mbar_expect_tx(bar_q + 8 * buf, 2 * SUB_BYTES);
        tma_load_3d(sQ + buf * SUB_BYTES, &tm_q64, bar_q + 8 * buf, h * HS, s * SUB, b);
        tma_load_3d(sdO + buf * SUB_BYTES, &tm_do64, bar_q + 8 * buf, h * HS, s * SUB, b);
    };
    if (leader) {
        mbar_expect_tx(bar_kv, 2 * TILE_BYTES);
        tma_load_3d(sK, &tm_kv, bar_kv, C + h * HS, jt * TILE, b);
        tma_load_3d(sV, &tm_kv, bar_kv, 2 * C + h * HS, jt * TILE, b);
        load_sub(s0, 0);
        if (s0 + 1 < NS) load_sub(s0 + 1, 1);
    }
    const int r = tid & 127, half = warp >> 2;
    const int key = jt * TILE + r;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const float scale = 1.0f / sqrtf((float)HS);
    const float sl2 = kLog2e * scale;
    const int warp_key_max = jt * TILE + (warp & 3) * 32 + 31;  // last key of this warp's 32 TMEM lanes
    const bool warp_keys_live = warp_key_max < T;

    for (int s = s0; s < NS; ++s) {
        const int it = s - s0, buf = it & 1;
        const uint32_t ph = (uint32_t)((it >> 1) & 1);
        const int nq = min(SUB, T - s * SUB), nq16 = (nq + 15) & ~15;
        if (warp == 0) {
            if (it == 0) mbar_wait(bar_kv, 0);
            mbar_wait(bar_q + 8 * buf, ph);
            tc_fence_after();
            const uint32_t idesc = make_idesc(TILE, nq16, 0, 0);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < HS / 16; ++k)  // S^T = K_j Q_s^T
                    umma_bf16(tmem_base + cST, make_desc(sK + k * 32, 0, 1024), make_desc(sQ + buf * SUB_BYTES + k * 32, 0, 1024), idesc, k > 0);
#pragma unroll
                for (int k = 0; k < HS / 16; ++k)  // dP^T = V_j dO_s^T
                    umma_bf16(tmem_base + cDPT, make_desc(sV + k * 32, 0, 1024), make_desc(sdO + buf * SUB_BYTES + k * 32, 0, 1024), idesc, k > 0);
                umma_commit(bar_s);  // also covers the dV / dK MMAs of the previous sub-tile
            }
            __syncwarp();
        }
        mbar_wait(bar_s, (uint32_t)(it & 1));
        tc_fence_after();
        if (leader && it >= 1 && s + 1 < NS) load_sub(s + 1, buf ^ 1);  // the previous sub-tile's buffer is free
        const bool mine = half * 32 < nq16;  // warp-uniform
        uint32_t sv[32], dpv[32];
        if (mine) {
            tmem_ld32(tmem_base + lane_off + cST + half * 32, sv);
            tmem_ld32(tmem_base + lane_off + cDPT + half * 32, dpv);
        }
        // packing in place needs no barrier: each thread overwrites only columns it has read itself (the second half of a row
        // packs from column 32 on, which the dV / dK MMAs are told about)
        if (mine) {
            uint32_t pp[16], dd[16];
            const int qbase = s * SUB + half * 32;  // first query of this thread's 32 columns
            // warp-uniform: every key of the warp and every query of the chunk is live and (causal) on or below the diagonal
            const bool fast = warp_keys_live && half * 32 + 32 <= nq && (!causal || warp_key_max <= qbase);
            if (fast) {
                // packed f32x2 arithmetic, statistics by 16-byte shared loads: ~5 instructions per element instead of ~25
                const float2 sl22 = splat2(sl2), sc2 = splat2(scale);
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                    const float4 l0 = *reinterpret_cast<const float4*>(stat + qbase + g8 * 8), l1 = *reinterpret_cast<const float4*>(stat + qbase + g8 * 8 + 4);
                    const float4 d0 = *reinterpret_cast<const float4*>(stat + TP + qbase + g8 * 8), d1 = *reinterpret_cast<const float4*>(stat + TP + qbase + g8 * 8 + 4);
                    const float lq[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
                    const float dq[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
                    for (int c = 0; c < 8; c += 2) {
                        const float2 a = fma2(make_float2(__uint_as_float(sv[g8 * 8 + c]), __uint_as_float(sv[g8 * 8 + c + 1])), sl22, make_float2(lq[c], lq[c + 1]));
                        const float2 pr = make_float2(ex2(a.x), ex2(a.y));
                        const float2 t = fma2(make_float2(__uint_as_float(dpv[g8 * 8 + c]), __uint_as_float(dpv[g8 * 8 + c + 1])), sc2, make_float2(dq[c], dq[c + 1]));
                        const float2 d2 = mul2(pr, t);
                        pp[g8 * 4 + (c >> 1)] = pack_bf16(pr.x, pr.y);
                        dd[g8 * 4 + (c >> 1)] = pack_bf16(d2.x, d2.y);
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    float pv[2], dv[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int qc = half * 32 + 2 * c + e;
                        const int q = s * SUB + qc;
                        const bool live = key < T && qc < nq && (!causal || key <= q);
                        const float p = live ? ex2(__uint_as_float(sv[2 * c + e]) * sl2 + stat[min(q, T - 1)]) : 0.f;
                        pv[e] = p;
                        dv[e] = p * (__uint_as_float(dpv[2 * c + e]) * scale + stat[TP + min(q, T - 1)]);
                    }
                    pp[c] = pack_bf16(pv[0], pv[1]);
                    dd[c] = pack_bf16(dv[0], dv[1]);
                }
            }
            tmem_st16(tmem_base + lane_off + cST + half * 32, pp);
            tmem_st16(tmem_base + lane_off + cDPT + half * 32, dd);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            const uint32_t idesc = make_idesc(TILE, HS, 0, 1);  // A from TMEM, B MN-major
            if (elect_one()) {
                for (int k16 = 0; k16 < nq16 / 16; ++k16) {
                    const uint32_t acc = (it > 0 || k16 > 0) ? 1u : 0u;
                    const uint32_t acol = (uint32_t)((k16 >> 1) * 32 + (k16 & 1) * 8);  // queries [32, 64) were packed from column 32 on
                    umma_bf16_ts(tmem_base + cDV, tmem_base + cST + acol, make_desc(sdO + buf * SUB_BYTES + k16 * 2048, SUB_BYTES, 1024), idesc, acc);
                    umma_bf16_ts(tmem_base + cDK, tmem_base + cDPT + acol, make_desc(sQ + buf * SUB_BYTES + k16 * 2048, SUB_BYTES, 1024), idesc, acc);
                }
                if (s + 1 == NS) umma_commit(bar_acc);
            }
            __syncwarp();
        }
    }
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    {
        uint32_t v[32];
        const bf16* oldv = (accumulate && key < T) ? dqkv + ((long)b * T + key) * 3 * C + 2 * C + h * HS + half * 32 : nullptr;
        const bf16* oldk = (accumulate && key < T) ? dqkv + ((long)b * T + key) * 3 * C + C + h * HS + half * 32 : nullptr;
        tmem_ld32(tmem_base + lane_off + cDV + half * 32, v);
        stage_half_row(sV, r, half, v, oldv);
        tmem_ld32(tmem_base + lane_off + cDK + half * 32, v);
        stage_half_row(sK, r, half, v, oldk);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (leader) {
        tma_store_3d(&tm_dqkv, sV, 2 * C + h * HS, jt * TILE, b);
        tma_store_3d(&tm_dqkv, sK, C + h * HS, jt * TILE, b);
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

__global__ void __launch_bounds__(256, 2)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_kv64,
                   const __grid_constant__ CUtensorMap tm_dqkv, bf16* __restrict__ dqkv, const float* __restrict__ lse,
                   const float* __restrict__ dsum, int T, int C, int NH, int causal, int accumulate) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sQ = base, sdO = sQ + TILE_BYTES, sK = sdO + TILE_BYTES, sV = sK + 2 * SUB_BYTES;
    const uint32_t bar0 = sV + 2 * SUB_BYTES;
    const uint32_t bar_qdo = bar0, bar_kv = bar0 + 8, bar_s = bar0 + 24, bar_acc = bar0 + 32;  // bar_kv[2]
    volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(gen + (bar0 - base) + 40);
    constexpr uint32_t TMEM_COLS = 256, cS = 0, cDP = 64, cDQ = 128;

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform: warp 0 issues the tcgen05 instructions converged
    const int NT = (T + TILE - 1) / TILE;
    const int qt = blockIdx.x % NT, bh = blockIdx.x / NT, b = bh / NH, h = bh - b * NH;
    const int kmax = causal ? min(T, (qt + 1) * TILE) : T;  // keys this query tile can see
    const int NS = (kmax + SUB - 1) / SUB;
    if (tid == 0) {
        tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_do); tma_prefetch_desc(&tm_kv64); tma_prefetch_desc(&tm_dqkv);
        mbar_init(bar_qdo, 1); mbar_init(bar_kv, 1); mbar_init(bar_kv + 8, 1); mbar_init(bar_s, 1); mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(smem_u32((const void*)slot), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *slot, 0);
    const bool leader = tid == 0;
    auto load_sub = [&](int s, int buf) {
        mbar_expect_tx(bar_kv + 8 * buf, 2 * SUB_BYTES);
        tma_load_3d(sK + buf * SUB_BYTES, &tm_kv64, bar_kv + 8 * buf, C + h * HS, s * SUB, b);
        tma_load_3d(sV + buf * SUB_BYTES, &tm_kv64, bar_kv + 8 * buf, 2 * C + h * HS, s * SUB, b);
    };
    if (leader) {
        mbar_expect_tx(bar_qdo, 2 * TILE_BYTES);
        tma_load_3d(sQ, &tm_q, bar_qdo, h * HS, qt * TILE, b);
        tma_load_3d(sdO, &tm_do, bar_qdo, h * HS, qt * TILE, b);
        load_sub(0, 0);
        if (1 < NS) load_sub(1, 1);
    }
    const int r = tid & 127, half = warp >> 2;
    const int q = qt * TILE + r;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const float scale = 1.0f / sqrtf((float)HS);
    const float sl2 = kLog2e * scale;
    const float lse2 = q < T ? lse[(long)bh * T + q] * kLog2e : 0.f;
    const float dq_ = q < T ? dsum[(long)bh * T + q] : 0.f;
    const int warp_q_min = qt * TILE + (warp & 3) * 32;  // first query of this warp's 32 TMEM lanes
    const bool warp_q_live = warp_q_min + 31 < T;

    for (int s = 0; s < NS; ++s) {
        const int buf = s & 1;
        const uint32_t ph = (uint32_t)((s >> 1) & 1);
        const int nk = min(SUB, T - s * SUB), nk16 = (nk + 15) & ~15;
        if (warp == 0) {
            if (s == 0) mbar_wait(bar_qdo, 0);
            mbar_wait(bar_kv + 8 * buf, ph);
            tc_fence_after();
            const uint32_t idesc = make_idesc(TILE, nk16, 0, 0);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < HS / 16; ++k)  // S = Q_i K_s^T
                    umma_bf16(tmem_base + cS, make_desc(sQ + k * 32, 0, 1024), make_desc(sK + buf * SUB_BYTES + k * 32, 0, 1024), idesc, k > 0);
#pragma unroll
                for (int k = 0; k < HS / 16; ++k)  // dP = dO_i V_s^T
                    umma_bf16(tmem_base + cDP, make_desc(sdO + k * 32, 0, 1024), make_desc(sV + buf * SUB_BYTES + k * 32, 0, 1024), idesc, k > 0);
                umma_commit(bar_s);  // also covers the dQ MMAs of the previous sub-tile
            }
            __syncwarp();
        }
        mbar_wait(bar_s, (uint32_t)(s & 1));
        tc_fence_after();
        if (leader && s >= 1 && s + 1 < NS) load_sub(s + 1, buf ^ 1);
        const bool mine = half * 32 < nk16;
        uint32_t sv[32], dpv[32];
        if (mine) {
            tmem_ld32(tmem_base + lane_off + cS + half * 32, sv);
            tmem_ld32(tmem_base + lane_off + cDP + half * 32, dpv);
        }
        // (no barrier: each thread packs dS over score columns it has read itself, the second half of a row from column 32 on)
        if (mine) {
            uint32_t dd[16];
            // warp-uniform: every query of the warp and every key of the chunk is live and (causal) on or below the diagonal
            const bool fast = warp_q_live && half * 32 + 32 <= nk && (!causal || s * SUB + half * 32 + 31 <= warp_q_min);
            if (fast) {
                const float2 sl22 = splat2(sl2), nl2 = splat2(-lse2), sc2 = splat2(scale), nd2 = splat2(-dq_ * scale);
#pragma unroll
                for (int c = 0; c < 16; ++c) {  // packed f32x2 arithmetic: P = exp2(S sl2 - lse2), dS = P (dP scale - D scale)
                    const float2 a = fma2(make_float2(__uint_as_float(sv[2 * c]), __uint_as_float(sv[2 * c + 1])), sl22, nl2);
                    const float2 pr = make_float2(ex2(a.x), ex2(a.y));
                    const float2 t = fma2(make_float2(__uint_as_float(dpv[2 * c]), __uint_as_float(dpv[2 * c + 1])), sc2, nd2);
                    const float2 d2 = mul2(pr, t);
                    dd[c] = pack_bf16(d2.x, d2.y);
                }
            } else {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    float dv[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int kc = half * 32 + 2 * c + e;
                        const int key = s * SUB + kc;
                        const bool live = q < T && kc < nk && (!causal || key <= q);
                        const float p = live ? ex2(__uint_as_float(sv[2 * c + e]) * sl2 - lse2) : 0.f;
                        dv[e] = p * (__uint_as_float(dpv[2 * c + e]) - dq_) * scale;
                    }
                    dd[c] = pack_bf16(dv[0], dv[1]);
                }
            }
            tmem_st16(tmem_base + lane_off + cS + half * 32, dd);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            const uint32_t idesc = make_idesc(TILE, HS, 0, 1);
            if (elect_one()) {
                for (int k16 = 0; k16 < nk16 / 16; ++k16)  // dQ_i += dS K_s (keys [32, 64) were packed from column 32 on)
                    umma_bf16_ts(tmem_base + cDQ, tmem_base + cS + (uint32_t)((k16 >> 1) * 32 + (k16 & 1) * 8),
                                 make_desc(sK + buf * SUB_BYTES + k16 * 2048, SUB_BYTES, 1024), idesc, (s > 0 || k16 > 0) ? 1u : 0u);
                if (s + 1 == NS) umma_commit(bar_acc);
            }
            __syncwarp();
        }
    }
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    {
        uint32_t v[32];
        const bf16* old = (accumulate && q < T) ? dqkv + ((long)b * T + q) * 3 * C + h * HS + half * 32 : nullptr;
        tmem_ld32(tmem_base + lane_off + cDQ + half * 32, v);
        stage_half_row(sQ, r, half, v, old);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (leader) {
        tma_store_3d(&tm_dqkv, sQ, h * HS, qt * TILE, b);
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

int encode_map3d(vitrs_ctx* ctx, CUtensorMap* map, const void* basep, uint64_t cols, uint64_t rows, uint64_t batch, uint32_t box_rows = TILE) {
    const uint64_t dims[3] = {cols, rows, batch};
    const uint64_t strides[2] = {cols * 2, rows * cols * 2};
    const uint32_t box[3] = {HS, box_rows, 1};
    return vitrs_tensor_map(ctx, map, 3, basep, dims, strides, box);
}

bool tc_shape_ok(const void* a, const void* b_, int t, int c, int nh) {
    return nh > 0 && c % nh == 0 && c / nh == HS && t >= 1 && t <= 4096 && ((uintptr_t)a & 15) == 0 && ((uintptr_t)b_ & 15) == 0;
}

}  // namespace

int op_attention_forward_tc(vitrs_ctx* ctx, bf16* out, float* lse, const bf16* qkv, int b, int t, int c, int nh, int causal) {
    if (!tc_shape_ok(out, qkv, t, c, nh)) return VITRS_ERR_UNSUPPORTED;
    if (b <= 0) return VITRS_OK;
    CUtensorMap tm, tm_out;
    VITRS_TRY(encode_map3d(ctx, &tm, qkv, 3 * (uint64_t)c, t, b));
    VITRS_TRY(encode_map3d(ctx, &tm_out, out, (uint64_t)c, t, b));
    const int NT = (t + TILE - 1) / TILE;
    const int NK = (t + 15) & ~15;
    if (t > 2 * TILE || ctx->env_attn_fwd_stream) {  // VITRS_ATTN_FWD_STREAM (test aid): the streaming kernel at any T
        const size_t smem = (size_t)5 * TILE_BYTES + 64 + 1024;
        VITRS_TRY(vitrs_func_smem(ctx, (const void*)attn_fwd_stream_kernel, smem));
        attn_fwd_stream_kernel<<<b * nh * NT, 128, smem, ctx->stream>>>(tm, tm_out, lse, t, c, nh, causal);
        VITRS_LAUNCHED(ctx);
        return VITRS_OK;
    }
    if (NT == 2 && !causal && !ctx->env_attn_fwd_legacy) {  // VITRS_ATTN_FWD_LEGACY (A/B aid): one CTA per query tile
        const size_t smem = (size_t)12 * TILE_BYTES + 256 + 2 * 2 * TILE * 3 * sizeof(float) + 1024;
        const int heads = b * nh, grid = heads < ctx->sm_count ? heads : ctx->sm_count;
        if (ctx->env_attn_fwd_nostagger) {
            VITRS_TRY(vitrs_func_smem(ctx, (const void*)attn_fwd_persist_kernel<false, false>, smem));
            attn_fwd_persist_kernel<false, false><<<grid, kFwdThreads, smem, ctx->stream>>>(tm, tm_out, lse, t, c, nh, NK, heads);
        } else if (ctx->env_attn_fwd_nosplit) {  // VITRS_ATTN_FWD_NOSPLIT (A/B aid): one thread per query row
            VITRS_TRY(vitrs_func_smem(ctx, (const void*)attn_fwd_persist_kernel<true, false>, smem));
            attn_fwd_persist_kernel<true, false><<<grid, kFwdThreads, smem, ctx->stream>>>(tm, tm_out, lse, t, c, nh, NK, heads);
        } else {
            VITRS_TRY(vitrs_func_smem(ctx, (const void*)attn_fwd_persist_kernel<true, true>, smem));
            attn_fwd_persist_kernel<true, true><<<grid, kFwdSplitThreads, smem, ctx->stream>>>(tm, tm_out, lse, t, c, nh, NK, heads);
        }
        VITRS_LAUNCHED(ctx);
        return VITRS_OK;
    }
    const size_t smem = (size_t)(1 + 2 * NT) * TILE_BYTES + 64 + 1024;
    const uint32_t tmem_cols = NK <= 128 ? 128 : 256;
    VITRS_TRY(vitrs_func_smem(ctx, (const void*)attn_fwd_tc2_kernel, smem));
    attn_fwd_tc2_kernel<<<b * nh * NT, 128, smem, ctx->stream>>>(tm, tm_out, lse, t, c, nh, causal, NT, NK, tmem_cols);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}

// D[b,h,q] = sum_i dO * O over the head slice (the row term of the softmax backward)
int op_attention_bwd_prep(vitrs_ctx* ctx, float* dsum, const bf16* dout, const bf16* out, int b, int t, int c, int nh) {
    if (b <= 0) return VITRS_OK;
    attn_bwd_prep_kernel<<<ceil_div((long)b * t, 8), 256, 0, ctx->stream>>>(dsum, dout, out, b, t, c, nh);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}

// dsum_ready: D already computed by the producer of dout (the attproj dX GEMM's EPI_ROWDOT epilogue); null -> computed here
int op_attention_backward_tc(vitrs_ctx* ctx, bf16* dqkv, const bf16* dout, const bf16* out, const bf16* qkv, const float* lse, int b,
                             int t, int c, int nh, int causal, int accumulate, const float* dsum_ready) {
    if (!tc_shape_ok(dqkv, qkv, t, c, nh) || ((uintptr_t)dout & 15) || ((uintptr_t)out & 15)) return VITRS_ERR_UNSUPPORTED;
    if (b <= 0) return VITRS_OK;
    const float* dsum = dsum_ready;
    if (!dsum) {
        VITRS_TRY(vitrs_ensure_scratch(ctx, (size_t)b * nh * t));
        VITRS_TRY(op_attention_bwd_prep(ctx, ctx->scratch, dout, out, b, t, c, nh));
        dsum = ctx->scratch;
    }
    const int NT = (t + TILE - 1) / TILE;
    // T <= 256: the one-CTA-per-head kernel (5 matmuls, dQ kept in TMEM) is faster (measured 139.6 vs 143.8 ms per
    // ViT-B/16 step); longer sequences take the two streaming kernels.  VITRS_ATTN_BWD_STREAM forces them (test / A-B aid).
    if (t > 2 * TILE || ctx->env_attn_bwd_stream) {
        CUtensorMap tm_q128, tm_q64, tm_do128, tm_do64, tm_dqkv;
        VITRS_TRY(encode_map3d(ctx, &tm_q128, qkv, 3 * (uint64_t)c, t, b, TILE));
        VITRS_TRY(encode_map3d(ctx, &tm_q64, qkv, 3 * (uint64_t)c, t, b, SUB));
        VITRS_TRY(encode_map3d(ctx, &tm_do128, dout, (uint64_t)c, t, b, TILE));
        VITRS_TRY(encode_map3d(ctx, &tm_do64, dout, (uint64_t)c, t, b, SUB));
        VITRS_TRY(encode_map3d(ctx, &tm_dqkv, dqkv, 3 * (uint64_t)c, t, b, TILE));
        const size_t smem_kv = (size_t)2 * TILE_BYTES + 4 * SUB_BYTES + 64 + 2 * (size_t)((t + SUB - 1) & ~(SUB - 1)) * 4 + 1024;  // + lse, D of the head
        const size_t smem_q = (size_t)2 * TILE_BYTES + 4 * SUB_BYTES + 64 + 1024;
        VITRS_TRY(vitrs_func_smem(ctx, (const void*)attn_bwd_dkv_kernel, smem_kv));
        VITRS_TRY(vitrs_func_smem(ctx, (const void*)attn_bwd_dq_kernel, smem_q));
        attn_bwd_dkv_kernel<<<b * nh * NT, kThreads, smem_kv, ctx->stream>>>(tm_q128, tm_q64, tm_do64, tm_dqkv, dqkv, lse, dsum, t, c, nh, causal,
                                                                           accumulate);
        VITRS_LAUNCHED(ctx);
        attn_bwd_dq_kernel<<<b * nh * NT, kThreads, smem_q, ctx->stream>>>(tm_q128, tm_do128, tm_q64, tm_dqkv, dqkv, lse, dsum, t, c, nh, causal,
                                                                         accumulate);
        VITRS_LAUNCHED(ctx);
        return VITRS_OK;
    }
    CUtensorMap tm_qkv, tm_do;
    VITRS_TRY(encode_map3d(ctx, &tm_qkv, qkv, 3 * (uint64_t)c, t, b));
    VITRS_TRY(encode_map3d(ctx, &tm_do, dout, (uint64_t)c, t, b));
    if (!causal) {
        const size_t smem_s = (size_t)NT * 4 * TILE_BYTES + 5 * TILE_BYTES + 4 * 256 * 4 + 256 + 1024;
        VITRS_TRY(vitrs_func_smem(ctx, (const void*)attn_bwd_persist_kernel, smem_s));
        CUtensorMap tm_dq;
        VITRS_TRY(encode_map3d(ctx, &tm_dq, dqkv, 3 * (uint64_t)c, t, b));
        const int heads = b * nh, grid = heads < ctx->sm_count ? heads : ctx->sm_count;
        attn_bwd_persist_kernel<<<grid, kPipeThreads, smem_s, ctx->stream>>>(tm_qkv, tm_do, tm_dq, dqkv, lse, dsum, t, c, nh, NT, accumulate, heads);
        VITRS_LAUNCHED(ctx);
        return VITRS_OK;
    }
    const size_t smem = (size_t)NT * 4 * TILE_BYTES + 4 * TILE_BYTES + 2 * 256 * 4 + 64 + 1024;
    VITRS_TRY(vitrs_func_smem(ctx, (const void*)attn_bwd_tc_kernel, smem));
    attn_bwd_tc_kernel<<<b * nh, kThreads, smem, ctx->stream>>>(tm_qkv, tm_do, dqkv, lse, dsum, t, c, nh, causal, NT, accumulate);
    VITRS_LAUNCHED(ctx);
    return VITRS_OK;
}
