// ctx.cu — context lifetime, error text, device memory helpers, NCCL binding (dlopen).
#include <dlfcn.h>
#include <stdarg.h>
#include <stdlib.h>

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
    *out = ctx;
    return VITRS_OK;
}

extern "C" int vitrs_ctx_destroy(vitrs_ctx* ctx) {
    if (!ctx) return VITRS_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    vitrs_comm_destroy(ctx);
    if (ctx->scratch) cudaFree(ctx->scratch);
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
// Only the five entry points the data-parallel step needs; types restated from nccl.h
// (ncclUniqueId is 128 opaque bytes; ncclFloat32 = 7, ncclSum = 0).
typedef struct { char internal[128]; } nccl_uid;
typedef int (*PFN_ncclGetUniqueId)(nccl_uid*);
typedef int (*PFN_ncclCommInitRank)(void**, int, nccl_uid, int);
typedef int (*PFN_ncclCommDestroy)(void*);
typedef int (*PFN_ncclAllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*PFN_ncclGroup)(void);
typedef const char* (*PFN_ncclGetErrorString)(int);

struct NcclApi {
    PFN_ncclGetUniqueId get_uid;
    PFN_ncclCommInitRank init_rank;
    PFN_ncclCommDestroy destroy;
    PFN_ncclAllReduce all_reduce;
    PFN_ncclGroup group_start, group_end;
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
    g_nccl.destroy = (PFN_ncclCommDestroy)dlsym(h, "ncclCommDestroy");
    g_nccl.all_reduce = (PFN_ncclAllReduce)dlsym(h, "ncclAllReduce");
    g_nccl.group_start = (PFN_ncclGroup)dlsym(h, "ncclGroupStart");
    g_nccl.group_end = (PFN_ncclGroup)dlsym(h, "ncclGroupEnd");
    g_nccl.err_str = (PFN_ncclGetErrorString)dlsym(h, "ncclGetErrorString");
    if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.destroy || !g_nccl.all_reduce || !g_nccl.group_start ||
        !g_nccl.group_end)
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

extern "C" int vitrs_comm_init(vitrs_ctx* ctx, const void* id128, int rank, int world) {
    VITRS_ARG(ctx, ctx && id128 && world >= 1 && rank >= 0 && rank < world);
    VITRS_TRY(load_nccl(ctx));
    VITRS_CUDA(ctx, cudaSetDevice(ctx->device));
    nccl_uid uid;
    memcpy(&uid, id128, sizeof(uid));
    VITRS_NCCL(ctx, g_nccl.init_rank(&ctx->nccl_comm, world, uid, rank));
    ctx->rank = rank;
    ctx->world = world;
    return VITRS_OK;
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

// grouped sum all-reduce of `count` slices on the comm stream; the caller orders it against
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
