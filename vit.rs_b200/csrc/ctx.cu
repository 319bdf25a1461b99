// ctx.cu — context lifetime, error text, device memory helpers, NCCL binding (dlopen).
#include <dlfcn.h>
#include <stdarg.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"

int vitrs_set_error(vitrs_ctx* ctx, int code, const char* fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

static char g_create_err[512] = "no error";

// ---- tensor-map cache ------------------------------------------------------------------------------
struct MapKey {
    const void* base;
    uint64_t dims[3], strides[2];
    uint32_t box[3], rank;
};
struct vitrs_map_entry {
    MapKey key;
    CUtensorMap map;
    int used;
};
static constexpr int kMapCacheSlots = 8192;  // power of two; cleared when half full

int vitrs_tensor_map(vitrs_ctx* ctx, CUtensorMap* out, int rank, const void* base, const uint64_t* dims, const uint64_t* strides,
                     const uint32_t* box) {
    MapKey key;
    memset(&key, 0, sizeof(key));
    key.base = base;
    key.rank = (uint32_t)rank;
    for (int i = 0; i < rank; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) key.strides[i] = strides[i];
    uint64_t h = 1469598103934665603ull;
    const unsigned char* kb = reinterpret_cast<const unsigned char*>(&key);
    for (size_t i = 0; i < sizeof(key); ++i) h = (h ^ kb[i]) * 1099511628211ull;
    int slot = (int)(h & (kMapCacheSlots - 1));
    if (!ctx->env_no_map_cache) {
        for (int probe = 0; probe < kMapCacheSlots; ++probe, slot = (slot + 1) & (kMapCacheSlots - 1)) {
            vitrs_map_entry& e = ctx->map_cache[slot];
            if (!e.used) break;
            if (memcmp(&e.key, &key, sizeof(key)) == 0) {
                *out = e.map;
                return VITRS_OK;
            }
        }
    }
    cuuint64_t d[3], st[2];
    cuuint32_t bx[3], estr[3] = {1, 1, 1};
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) st[i] = strides[i];
    CUresult r = ctx->encode_tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, st, bx, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return vitrs_set_error(ctx, VITRS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): rank=%d dims=%llu,%llu,%llu stride0=%llu box=%u,%u,%u",
                               (int)r, rank, (unsigned long long)key.dims[0], (unsigned long long)key.dims[1],
                               (unsigned long long)key.dims[2], (unsigned long long)key.strides[0], key.box[0], key.box[1], key.box[2]);
    if (!ctx->env_no_map_cache) {
        if (ctx->map_cache_used >= kMapCacheSlots / 2) {
            memset(ctx->map_cache, 0, sizeof(vitrs_map_entry) * kMapCacheSlots);
            ctx->map_cache_used = 0;
            slot = (int)(h & (kMapCacheSlots - 1));
        }
        while (ctx->map_cache[slot].used) slot = (slot + 1) & (kMapCacheSlots - 1);
        ctx->map_cache[slot].key = key;
        ctx->map_cache[slot].map = *out;
        ctx->map_cache[slot].used = 1;
        ctx->map_cache_used++;
    }
    return VITRS_OK;
}

// The dynamic shared memory opt-in is an attribute of (function, device), shared by every context of the process on that device:
// it is recorded per device here (not per context — a second context on the same GPU that asked for less would lower it under
// the first one's feet) and only ever raised.
namespace {
struct FuncSmem { const void* fn; size_t bytes; };
constexpr int kMaxDevices = 16, kMaxFuncs = 128;
FuncSmem g_func_smem[kMaxDevices][kMaxFuncs];
int g_func_smem_count[kMaxDevices];
std::mutex g_func_smem_mu;
}  // namespace

int vitrs_func_smem(vitrs_ctx* ctx, const void* fn, size_t bytes) {
    const int dev = ctx->device >= 0 && ctx->device < kMaxDevices ? ctx->device : 0;
    std::lock_guard<std::mutex> lk(g_func_smem_mu);
    FuncSmem* tab = g_func_smem[dev];
    int& n = g_func_smem_count[dev];
    for (int i = 0; i < n; ++i) {
        if (tab[i].fn == fn) {
            if (tab[i].bytes >= bytes) return VITRS_OK;
            VITRS_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            tab[i].bytes = bytes;
            return VITRS_OK;
        }
    }
    if (n >= kMaxFuncs) {  // (table full: ask the runtime what the function has and never lower it)
        cudaFuncAttributes at;
        VITRS_CUDA(ctx, cudaFuncGetAttributes(&at, fn));
        if ((size_t)at.maxDynamicSharedSizeBytes >= bytes) return VITRS_OK;
        VITRS_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        return VITRS_OK;
    }
    VITRS_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    tab[n].fn = fn;
    tab[n].bytes = bytes;
    ++n;
    return VITRS_OK;
}

extern "C" const char* vitrs_version(void) { return "vitrs-b200 0.1 (sm_100a)"; }

extern "C" const char* vitrs_last_error(vitrs_ctx* ctx) { return ctx ? ctx->err : g_create_err; }

extern "C" int vitrs_ctx_create(vitrs_ctx** out, int device) {
    if (!out) return VITRS_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        // no CPU fallback: fail loudly
        snprintf(g_create_err, sizeof(g_create_err), "no CUDA device: %s", cudaGetErrorString(e));
        return VITRS_ERR_CUDA;
    }
    if (device < 0 || device >= count) {
        snprintf(g_create_err, sizeof(g_create_err), "device %d out of range (%d devices)", device, count);
        return VITRS_ERR_ARG;
    }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        snprintf(g_create_err, sizeof(g_create_err), "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
        return VITRS_ERR_CUDA;
    }
    if (prop.major != 10) {
        snprintf(g_create_err, sizeof(g_create_err), "device %d is sm_%d%d; this library is sm_100a only",
                 device, prop.major, prop.minor);
        return VITRS_ERR_UNSUPPORTED;
    }
    vitrs_ctx* ctx = (vitrs_ctx*)calloc(1, sizeof(vitrs_ctx));
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    strcpy(ctx->err, "no error");
    cudaSetDevice(device);
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking) != cudaSuccess) {
        snprintf(g_create_err, sizeof(g_create_err), "stream creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        free(ctx);
        return VITRS_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    // cuTensorMapEncodeTiled through the runtime, so the library needs no link to libcuda
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        snprintf(g_create_err, sizeof(g_create_err), "cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
        free(ctx);
        return VITRS_ERR_CUDA;
    }
    ctx->encode_tiled = (PFN_encodeTiled)fn;
    ctx->world = 1;
    ctx->map_cache = (vitrs_map_entry*)calloc(kMapCacheSlots, sizeof(vitrs_map_entry));
    auto flag = [](const char* name) { return getenv(name) != nullptr ? 1 : 0; };
    ctx->env_gemm_cg1 = getenv("VITRS_GEMM_CG") && atoi(getenv("VITRS_GEMM_CG")) == 1;
    ctx->env_gemm_splits = getenv("VITRS_GEMM_SPLITS") ? atoi(getenv("VITRS_GEMM_SPLITS")) : 0;
    ctx->env_dp_defer = flag("VITRS_DP_DEFER");
    ctx->env_attn_fwd_stream = flag("VITRS_ATTN_FWD_STREAM");
    ctx->env_attn_bwd_stream = flag("VITRS_ATTN_BWD_STREAM");
    ctx->env_attn_bwd_overwrite = flag("VITRS_ATTN_BWD_OVERWRITE");
    ctx->env_no_map_cache = flag("VITRS_NO_MAP_CACHE");
    ctx->env_attn_fwd_legacy = flag("VITRS_ATTN_FWD_LEGACY");
    ctx->env_attn_fwd_nostagger = flag("VITRS_ATTN_FWD_NOSTAGGER");
    ctx->env_gemm_static = flag("VITRS_GEMM_STATIC");
    ctx->env_no_step_graph = flag("VITRS_NO_STEP_GRAPH");
    ctx->env_gemm_no_small = flag("VITRS_GEMM_NO_SMALL");
    ctx->env_attn_fwd_nosplit = flag("VITRS_ATTN_FWD_NOSPLIT");
    ctx->env_gemm_patch_tc = flag("VITRS_GEMM_PATCH_TC");
    if (cudaMalloc(&ctx->dev_flags, 128) != cudaSuccess || cudaMemset(ctx->dev_flags, 0, 128) != cudaSuccess) {
        snprintf(g_create_err, sizeof(g_create_err), "cudaMalloc(flags) failed: %s", cudaGetErrorString(cudaGetLastError()));
        free(ctx->map_cache);
        free(ctx);
        return VITRS_ERR_CUDA;
    }
    ctx->d_hyper = reinterpret_cast<AdamHyper*>(ctx->dev_flags + 16);
    ctx->gemm_sched = reinterpret_cast<unsigned int*>(ctx->dev_flags + 8);
    *out = ctx;
    return VITRS_OK;
}

extern "C" int vitrs_ctx_destroy(vitrs_ctx* ctx) {
    if (!ctx) return VITRS_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    vitrs_comm_destroy(ctx);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->dev_flags) cudaFree(ctx->dev_flags);
    free(ctx->map_cache);
    if (ctx->prof_ev) {
        for (int i = 0; i < 2 * ctx->prof_cap; ++i) cudaEventDestroy(ctx->prof_ev[i]);
        free(ctx->prof_ev);
        free(ctx->prof_flops);
    }
    cudaStreamDestroy(ctx->own_stream);
    cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->comm_stream);
    free(ctx);
    return VITRS_OK;
}

extern "C" int vitrs_ctx_set_stream(vitrs_ctx* ctx, void* s) {
    if (!ctx) return VITRS_ERR_ARG;
    ctx->stream = (cudaStream_t)s;
    return VITRS_OK;
}

extern "C" int vitrs_ctx_reset_stream(vitrs_ctx* ctx) {
    if (!ctx) return VITRS_ERR_ARG;
    ctx->stream = ctx->own_stream;
    return VITRS_OK;
}

extern "C" void* vitrs_ctx_stream(vitrs_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" int vitrs_ctx_synchronize(vitrs_ctx* ctx) {
    if (!ctx) return VITRS_ERR_ARG;
    VITRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VITRS_OK;
}

extern "C" uint64_t vitrs_launch_count(vitrs_ctx* ctx) { return ctx ? ctx->launches : 0; }

void vitrs_prof_before(vitrs_ctx* ctx, double flops) {
    if (!ctx->prof_on || ctx->prof_count >= ctx->prof_cap) return;
    ctx->prof_flops[ctx->prof_count] = flops;
    cudaEventRecord(ctx->prof_ev[2 * ctx->prof_count], ctx->stream);
}
void vitrs_prof_after(vitrs_ctx* ctx) {
    if (!ctx->prof_on || ctx->prof_count >= ctx->prof_cap) return;
    cudaEventRecord(ctx->prof_ev[2 * ctx->prof_count + 1], ctx->stream);
    ctx->prof_count++;
}

extern "C" int vitrs_profile_begin(vitrs_ctx* ctx) {
    if (!ctx) return VITRS_ERR_ARG;
    if (!ctx->prof_ev) {
        ctx->prof_cap = 4096;
        ctx->prof_ev = (cudaEvent_t*)calloc(2 * ctx->prof_cap, sizeof(cudaEvent_t));
        ctx->prof_flops = (double*)calloc(ctx->prof_cap, sizeof(double));
        for (int i = 0; i < 2 * ctx->prof_cap; ++i) VITRS_CUDA(ctx, cudaEventCreate(&ctx->prof_ev[i]));
    }
    ctx->prof_count = 0;
    ctx->prof_on = 1;
    return VITRS_OK;
}

extern "C" int vitrs_profile_end(vitrs_ctx* ctx, double* gemm_ms, double* gemm_flops, int* gemm_launches) {
    if (!ctx) return VITRS_ERR_ARG;
    ctx->prof_on = 0;
    VITRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double ms = 0.0, fl = 0.0;
    for (int i = 0; i < ctx->prof_count; ++i) {
        float t = 0.f;
        VITRS_CUDA(ctx, cudaEventElapsedTime(&t, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]));
        ms += t;
        fl += ctx->prof_flops[i];
    }
    if (gemm_ms) *gemm_ms = ms;
    if (gemm_flops) *gemm_flops = fl;
    if (gemm_launches) *gemm_launches = ctx->prof_count;
    return VITRS_OK;
}

int vitrs_ensure_scratch(vitrs_ctx* ctx, size_t floats) {
    if (ctx->scratch_floats >= floats) return VITRS_OK;
    if (ctx->scratch) {
        VITRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        VITRS_CUDA(ctx, cudaFree(ctx->scratch));
        ctx->scratch = nullptr;
        ctx->scratch_floats = 0;
    }
    VITRS_CUDA(ctx, cudaMalloc(&ctx->scratch, floats * sizeof(float)));
    ctx->scratch_floats = floats;
    ctx->scratch_gen++;
    return VITRS_OK;
}

extern "C" int vitrs_malloc(vitrs_ctx* ctx, void** p, size_t bytes) {
    VITRS_ARG(ctx, ctx && p);
    VITRS_CUDA(ctx, cudaSetDevice(ctx->device));
    VITRS_CUDA(ctx, cudaMalloc(p, bytes));
    return VITRS_OK;
}
extern "C" int vitrs_free(vitrs_ctx* ctx, void* p) {
    VITRS_ARG(ctx, ctx != nullptr);
    VITRS_CUDA(ctx, cudaFree(p));
    return VITRS_OK;
}
extern "C" int vitrs_malloc_host(vitrs_ctx* ctx, void** p, size_t bytes) {
    VITRS_ARG(ctx, ctx && p);
    VITRS_CUDA(ctx, cudaMallocHost(p, bytes));
    return VITRS_OK;
}
extern "C" int vitrs_free_host(vitrs_ctx* ctx, void* p) {
    VITRS_ARG(ctx, ctx != nullptr);
    VITRS_CUDA(ctx, cudaFreeHost(p));
    return VITRS_OK;
}
extern "C" int vitrs_memcpy_h2d(vitrs_ctx* ctx, void* dst, const void* src, size_t bytes) {
    VITRS_ARG(ctx, ctx != nullptr);
    VITRS_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return VITRS_OK;
}
extern "C" int vitrs_memcpy_d2h(vitrs_ctx* ctx, void* dst, const void* src, size_t bytes) {
    VITRS_ARG(ctx, ctx != nullptr);
    VITRS_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    VITRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VITRS_OK;
}
extern "C" int vitrs_memset(vitrs_ctx* ctx, void* dst, int value, size_t bytes) {
    VITRS_ARG(ctx, ctx != nullptr);
    VITRS_CUDA(ctx, cudaMemsetAsync(dst, value, bytes, ctx->stream));
    return VITRS_OK;
}

// ---- NCCL through dlopen ---------------------------------------------------------------------
// Only the entry points the data-parallel / ZeRO-1 step needs; types restated from nccl.h
// (ncclUniqueId is 128 opaque bytes; ncclFloat32 = 7, ncclBfloat16 = 9, ncclSum = 0).
typedef struct { char internal[128]; } nccl_uid;
// ncclConfig_t as of NCCL 2.19+ (size / magic / version header, then plain ints; later versions append fields and accept
// shorter structs by `size`): used only to cap the CTAs NCCL may occupy while the persistent GEMMs hold the SMs
typedef struct {
    size_t size;
    unsigned int magic, version;
    int blocking, cgaClusterSize, minCTAs, maxCTAs;
    const char* netName;
    int splitShare, trafficClass;  // (trafficClass fills what would be padding; NCCL ignores it for version < 2.23)
} nccl_config_v219;
typedef int (*PFN_ncclGetUniqueId)(nccl_uid*);
typedef int (*PFN_ncclCommInitRank)(void**, int, nccl_uid, int);
typedef int (*PFN_ncclCommInitRankConfig)(void**, int, nccl_uid, int, void*);
typedef int (*PFN_ncclCommDestroy)(void*);
typedef int (*PFN_ncclAllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*PFN_ncclReduceScatter)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*PFN_ncclAllGather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*PFN_ncclGroup)(void);
typedef int (*PFN_ncclCommGetAsyncError)(void*, int*);
typedef int (*PFN_ncclGetVersion)(int*);
typedef const char* (*PFN_ncclGetErrorString)(int);

struct NcclApi {
    PFN_ncclGetUniqueId get_uid;
    PFN_ncclCommInitRank init_rank;
    PFN_ncclCommInitRankConfig init_rank_config;
    PFN_ncclCommDestroy destroy;
    PFN_ncclAllReduce all_reduce;
    PFN_ncclReduceScatter reduce_scatter;
    PFN_ncclAllGather all_gather;
    PFN_ncclGroup group_start, group_end;
    PFN_ncclCommGetAsyncError async_error;
    PFN_ncclGetVersion get_version;
    PFN_ncclGetErrorString err_str;
};
static NcclApi g_nccl;

static int load_nccl(vitrs_ctx* ctx) {
    if (ctx->nccl_lib) return VITRS_OK;
    // RTLD_NOLOAD first: inside a torch process this is torch's own libnccl (one NCCL per process)
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return vitrs_set_error(ctx, VITRS_ERR_NCCL, "dlopen(libnccl.so.2) failed: %s", dlerror());
    g_nccl.get_uid = (PFN_ncclGetUniqueId)dlsym(h, "ncclGetUniqueId");
    g_nccl.init_rank = (PFN_ncclCommInitRank)dlsym(h, "ncclCommInitRank");
    g_nccl.init_rank_config = (PFN_ncclCommInitRankConfig)dlsym(h, "ncclCommInitRankConfig");
    g_nccl.destroy = (PFN_ncclCommDestroy)dlsym(h, "ncclCommDestroy");
    g_nccl.all_reduce = (PFN_ncclAllReduce)dlsym(h, "ncclAllReduce");
    g_nccl.reduce_scatter = (PFN_ncclReduceScatter)dlsym(h, "ncclReduceScatter");
    g_nccl.all_gather = (PFN_ncclAllGather)dlsym(h, "ncclAllGather");
    g_nccl.group_start = (PFN_ncclGroup)dlsym(h, "ncclGroupStart");
    g_nccl.group_end = (PFN_ncclGroup)dlsym(h, "ncclGroupEnd");
    g_nccl.async_error = (PFN_ncclCommGetAsyncError)dlsym(h, "ncclCommGetAsyncError");
    g_nccl.get_version = (PFN_ncclGetVersion)dlsym(h, "ncclGetVersion");
    g_nccl.err_str = (PFN_ncclGetErrorString)dlsym(h, "ncclGetErrorString");
    if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.destroy || !g_nccl.all_reduce || !g_nccl.reduce_scatter ||
        !g_nccl.all_gather || !g_nccl.group_start || !g_nccl.group_end)
        return vitrs_set_error(ctx, VITRS_ERR_NCCL, "libnccl.so.2 lacks a required symbol");
    ctx->nccl_lib = h;
    return VITRS_OK;
}

#define VITRS_NCCL(ctx, expr)                                                                        \
    do {                                                                                             \
        int r__ = (expr);                                                                            \
        if (r__ != 0)                                                                                \
            return vitrs_set_error(ctx, VITRS_ERR_NCCL, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, \
                                   g_nccl.err_str ? g_nccl.err_str(r__) : "nccl error");             \
    } while (0)

extern "C" int vitrs_comm_unique_id(vitrs_ctx* ctx, void* id128) {
    VITRS_ARG(ctx, ctx && id128);
    VITRS_TRY(load_nccl(ctx));
    VITRS_NCCL(ctx, g_nccl.get_uid((nccl_uid*)id128));
    return VITRS_OK;
}

// max_ctas > 0 caps the thread blocks NCCL may use per collective (ncclConfig_t.maxCTAs).  The persistent GEMMs and attention
// kernels fill every SM with one large-shared-memory CTA, so each CTA NCCL occupies displaces a CTA pair of the next GEMM;
// with NVSwitch a handful of CTAs already saturates the gradient exchange.  0 = NCCL's default.
extern "C" int vitrs_comm_init_config(vitrs_ctx* ctx, const void* id128, int rank, int world, int max_ctas) {
    VITRS_ARG(ctx, ctx && id128 && world >= 1 && rank >= 0 && rank < world && max_ctas >= 0);
    VITRS_TRY(load_nccl(ctx));
    VITRS_CUDA(ctx, cudaSetDevice(ctx->device));
    nccl_uid uid;
    memcpy(&uid, id128, sizeof(uid));
    if (max_ctas > 0 && g_nccl.init_rank_config) {
        nccl_config_v219 cfg;
        cfg.size = sizeof(cfg);
        cfg.magic = 0xcafebeefu;
        cfg.version = 2 * 10000 + 19 * 100 + 0;  // NCCL_VERSION(2, 19, 0)
        const int undef = (int)0x80000000;       // NCCL_CONFIG_UNDEF_INT
        cfg.blocking = undef; cfg.cgaClusterSize = undef; cfg.minCTAs = undef; cfg.maxCTAs = max_ctas;
        cfg.netName = nullptr; cfg.splitShare = undef; cfg.trafficClass = undef;
        VITRS_NCCL(ctx, g_nccl.init_rank_config(&ctx->nccl_comm, world, uid, rank, &cfg));
    } else {
        VITRS_NCCL(ctx, g_nccl.init_rank(&ctx->nccl_comm, world, uid, rank));
    }
    ctx->rank = rank;
    ctx->world = world;
    ctx->nccl_max_ctas = max_ctas;
    return VITRS_OK;
}

extern "C" int vitrs_comm_init(vitrs_ctx* ctx, const void* id128, int rank, int world) {
    // VITRS_NCCL_MAX_CTAS overrides the default cap (0 = leave NCCL alone)
    // 8 GPUs: 66.9 k images/s at 2 CTAs against 65.9 k at 8 (one sample each); at 2 a 14 MB bucket takes 0.65 ms (N = 2 timeline), which
    // is what the last bucket exposes behind backward, so the default sits between
    int max_ctas = 4;
    if (const char* ov = getenv("VITRS_NCCL_MAX_CTAS")) max_ctas = atoi(ov);
    return vitrs_comm_init_config(ctx, id128, rank, world, max_ctas < 0 ? 0 : max_ctas);
}

extern "C" int vitrs_comm_destroy(vitrs_ctx* ctx) {
    if (ctx && ctx->nccl_comm) {
        g_nccl.destroy(ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
        ctx->world = 1;
        ctx->rank = 0;
    }
    return VITRS_OK;
}

extern "C" int vitrs_comm_world(vitrs_ctx* ctx, int* rank, int* world) {
    VITRS_ARG(ctx, ctx != nullptr);
    if (rank) *rank = ctx->rank;
    if (world) *world = ctx->world;
    return VITRS_OK;
}

// ncclCommGetAsyncError: 0 = no communicator or healthy; a non-zero NCCL result means a peer died or the network failed and
// the communicator must be torn down (the collectives already queued would otherwise hang)
extern "C" int vitrs_comm_async_error(vitrs_ctx* ctx, int* nccl_result) {
    VITRS_ARG(ctx, ctx && nccl_result);
    *nccl_result = 0;
    if (!ctx->nccl_comm || !g_nccl.async_error) return VITRS_OK;
    VITRS_NCCL(ctx, g_nccl.async_error(ctx->nccl_comm, nccl_result));
    if (*nccl_result != 0)
        return vitrs_set_error(ctx, VITRS_ERR_NCCL, "NCCL asynchronous error: %s", g_nccl.err_str ? g_nccl.err_str(*nccl_result) : "?");
    return VITRS_OK;
}

static inline int nccl_dtype(int dtype) { return dtype == 1 ? /*ncclBfloat16*/ 9 : /*ncclFloat32*/ 7; }

// grouped sum all-reduce of `count` fp32 slices on the comm stream; the caller orders it against
// the compute stream with events
int vitrs_nccl_allreduce_group(vitrs_ctx* ctx, float* const* bufs, const size_t* counts, int count) {
    if (!ctx->nccl_comm) return VITRS_OK;
    VITRS_NCCL(ctx, g_nccl.group_start());
    for (int i = 0; i < count; ++i)
        VITRS_NCCL(ctx, g_nccl.all_reduce(bufs[i], bufs[i], counts[i], /*ncclFloat32*/ 7, /*ncclSum*/ 0, ctx->nccl_comm,
                                          ctx->comm_stream));
    VITRS_NCCL(ctx, g_nccl.group_end());
    return VITRS_OK;
}

// the three collectives of the step, on the comm stream (dtype: 0 = fp32, 1 = bf16)
int vitrs_nccl_allreduce(vitrs_ctx* ctx, const void* send, void* recv, size_t count, int dtype) {
    if (!ctx->nccl_comm) return VITRS_OK;
    VITRS_NCCL(ctx, g_nccl.all_reduce(send, recv, count, nccl_dtype(dtype), 0, ctx->nccl_comm, ctx->comm_stream));
    return VITRS_OK;
}
int vitrs_nccl_reduce_scatter(vitrs_ctx* ctx, const void* send, void* recv, size_t recv_count, int dtype) {
    if (!ctx->nccl_comm) return VITRS_OK;
    VITRS_NCCL(ctx, g_nccl.reduce_scatter(send, recv, recv_count, nccl_dtype(dtype), 0, ctx->nccl_comm, ctx->comm_stream));
    return VITRS_OK;
}
int vitrs_nccl_all_gather(vitrs_ctx* ctx, const void* send, void* recv, size_t send_count, int dtype) {
    if (!ctx->nccl_comm) return VITRS_OK;
    VITRS_NCCL(ctx, g_nccl.all_gather(send, recv, send_count, nccl_dtype(dtype), ctx->nccl_comm, ctx->comm_stream));
    return VITRS_OK;
}
int vitrs_nccl_group(vitrs_ctx* ctx, int begin) {
    if (!ctx->nccl_comm) return VITRS_OK;
    VITRS_NCCL(ctx, begin ? g_nccl.group_start() : g_nccl.group_end());
    return VITRS_OK;
}

extern "C" int vitrs_allreduce_f32(vitrs_ctx* ctx, float* buf, size_t n) {
    VITRS_ARG(ctx, ctx != nullptr);
    if (!ctx->nccl_comm) return VITRS_OK;
    cudaEvent_t ev;
    VITRS_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    VITRS_CUDA(ctx, cudaEventRecord(ev, ctx->stream));
    VITRS_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ev, 0));
    float* bufs[1] = {buf};
    size_t counts[1] = {n};
    VITRS_TRY(vitrs_nccl_allreduce_group(ctx, bufs, counts, 1));
    VITRS_CUDA(ctx, cudaEventRecord(ev, ctx->comm_stream));
    VITRS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev, 0));
    VITRS_CUDA(ctx, cudaEventDestroy(ev));
    return VITRS_OK;
}

// device-side error flags raised by kernels since the last call (bit 0: class label outside [0, classes)); synchronises
extern "C" int vitrs_ctx_error_flags(vitrs_ctx* ctx, int* flags) {
    VITRS_ARG(ctx, ctx && flags);
    VITRS_CUDA(ctx, cudaSetDevice(ctx->device));
    VITRS_CUDA(ctx, cudaMemcpyAsync(flags, ctx->dev_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    VITRS_CUDA(ctx, cudaMemsetAsync(ctx->dev_flags, 0, sizeof(int), ctx->stream));
    VITRS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VITRS_OK;
}
