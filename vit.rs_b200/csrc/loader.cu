// loader.cu — the data path in front of the step.  The reference's train loop takes in-memory buffers
// (`inputs: &[usize]`, rusty_vit.rs:269) and has no reader; this is the piece a training job needs in its place: a reader of
// fixed-size image records (the CIFAR-10 / CIFAR-100 binary layout: `label_bytes` label bytes, then 3 x H x W uint8 samples,
// channel-major) and a loader thread that assembles shuffled batches into a ring of pinned host slots, so that
// vitrs_model_train_step_loader finds the next batch ready, stages the one after it on the copy stream and never waits on
// the file system.  Pixels stay uint8 until the im2col pass of the patch embedding normalises them on the device.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "model.cuh"

namespace {
constexpr int kSlots = 4;

inline uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
}  // namespace

struct vitrs_loader {
    vitrs_ctx* ctx;  // NULL: pageable slots (tools and CPU tests); otherwise pinned
    std::vector<uint8_t> data;  // every record of every file, back to back
    size_t record_bytes, image_bytes, num_records;
    int label_bytes, batch, shuffle, drop_last, num_classes_seen;
    int rank, world;  // data parallel: every rank walks the same (seed, epoch) order and takes batches rank, rank + world, ...
    uint64_t seed;
    // transform between the stored record and the delivered image (vitrs_loader_open_transform); all zero = plain copy
    int in_size, out_size;      // side of the stored / delivered images
    size_t out_image_bytes;     // 3 * out_size^2
    int random_flip, crop_pad;  // augmentation: a function of (seed, epoch, record) only, whatever the thread count
    bool transform;
    std::vector<int> rx0, rx1;  // bilinear taps of every output column (= row: the images are square) and the weight of the second
    std::vector<float> rw;
    // helper threads that share the images of the batch being assembled with the loader thread
    std::vector<std::thread> helpers;
    std::mutex job_mu;
    std::condition_variable job_cv;
    uint64_t job_seq;            // bumped per batch
    int job_slot, job_n, job_done_helpers;
    size_t job_at;
    uint64_t job_epoch;
    const uint32_t* job_order;
    std::atomic<int> job_next;   // next image of the batch to take
    // ring of host slots: the loader thread fills `filled`, the consumer hands slots back through `released`
    uint8_t* images[kSlots];
    int* labels[kSlots];
    int count[kSlots];
    uint64_t epoch_of[kSlots];
    uint64_t produced, consumed, released;  // slot sequence numbers (slot = n % kSlots)
    int prefetched;                          // 1: slot `consumed` has already been staged on the device by the previous step
    std::atomic<bool> stop;
    std::mutex mu;
    std::condition_variable cv;
    std::thread worker;
};

namespace {

// ---- record -> delivered image ------------------------------------------------------------------------------------------
// random crop of the zero-padded record (the standard CIFAR augmentation), horizontal flip, then a bilinear resize with
// half-pixel centres (what torch.nn.functional.interpolate(mode="bilinear", align_corners=False) computes), separable: rows
// first into a float strip, then columns.  Randomness comes from (seed, epoch, record index), so a batch is the same bytes
// whichever thread assembles it.
void bilinear_taps(vitrs_loader* L) {
    const int S = L->in_size, O = L->out_size;
    L->rx0.resize(O); L->rx1.resize(O); L->rw.resize(O);
    const float scale = (float)S / (float)O;
    for (int o = 0; o < O; ++o) {
        float src = ((float)o + 0.5f) * scale - 0.5f;
        if (src < 0.f) src = 0.f;
        int i0 = (int)src;
        if (i0 > S - 1) i0 = S - 1;
        L->rx0[o] = i0;
        L->rx1[o] = i0 + 1 < S ? i0 + 1 : S - 1;
        L->rw[o] = src - (float)i0;
    }
}

void transform_image(const vitrs_loader* L, const uint8_t* rec_pixels, uint8_t* dst, uint64_t epoch, uint32_t record) {
    const int S = L->in_size, O = L->out_size;
    uint64_t s = (L->seed ^ 0xA5A5A5A5A5A5A5A5ull) + epoch * 0x9E3779B97F4A7C15ull + (uint64_t)record * 0xD1B54A32D192ED03ull;
    const uint64_t r = splitmix64(s);
    const bool flip = L->random_flip && (r & 1);
    const int span = 2 * L->crop_pad + 1;
    const int dx = L->crop_pad ? (int)((r >> 8) % (uint64_t)span) - L->crop_pad : 0;   // crop window offset in the record's frame
    const int dy = L->crop_pad ? (int)((r >> 32) % (uint64_t)span) - L->crop_pad : 0;
    static thread_local std::vector<uint8_t> staged_buf;  // cropped + flipped record
    static thread_local std::vector<float> strip;         // one channel, resized along x
    const uint8_t* src = rec_pixels;
    if (flip || dx || dy) {
        staged_buf.resize((size_t)3 * S * S);
        uint8_t* staged = staged_buf.data();
        for (int c = 0; c < 3; ++c)
            for (int y = 0; y < S; ++y) {
                const int sy = y + dy;
                uint8_t* row = staged + ((size_t)c * S + y) * S;
                if (sy < 0 || sy >= S) { memset(row, 0, S); continue; }
                const uint8_t* srow = rec_pixels + ((size_t)c * S + sy) * S;
                for (int x = 0; x < S; ++x) {
                    const int sx = (flip ? S - 1 - x : x) + dx;
                    row[x] = (sx < 0 || sx >= S) ? 0 : srow[sx];
                }
            }
        src = staged;
    }
    if (O == S) { memcpy(dst, src, (size_t)3 * S * S); return; }
    strip.resize((size_t)S * O);
    const int* __restrict__ x0 = L->rx0.data(); const int* __restrict__ x1 = L->rx1.data(); const float* __restrict__ w = L->rw.data();
    for (int c = 0; c < 3; ++c) {
        for (int y = 0; y < S; ++y) {
            const uint8_t* __restrict__ srow = src + ((size_t)c * S + y) * S;
            float* __restrict__ out = strip.data() + (size_t)y * O;
            for (int o = 0; o < O; ++o) out[o] = (float)srow[x0[o]] + w[o] * ((float)srow[x1[o]] - (float)srow[x0[o]]);
        }
        for (int oy = 0; oy < O; ++oy) {
            // (restrict: a byte store may alias anything, which would keep this loop — 3 x O x O elements per image — scalar)
            const float* __restrict__ a = strip.data() + (size_t)x0[oy] * O;
            const float* __restrict__ b = strip.data() + (size_t)x1[oy] * O;
            const float wy = w[oy];
            uint8_t* __restrict__ out = dst + ((size_t)c * O + oy) * O;
            for (int o = 0; o < O; ++o) out[o] = (uint8_t)(int)(a[o] + wy * (b[o] - a[o]) + 0.5f);
        }
    }
}

// images of the current job, taken one at a time by whoever is free (the loader thread and its helpers)
void work_on_job(vitrs_loader* L, int slot, size_t at, int n, uint64_t epoch, const uint32_t* order) {
    for (;;) {
        const int k = L->job_next.fetch_add(1);
        if (k >= n) return;
        const uint32_t id = order[at + k];
        const uint8_t* rec = L->data.data() + (size_t)id * L->record_bytes;
        L->labels[slot][k] = rec[L->label_bytes - 1];
        transform_image(L, rec + L->label_bytes, L->images[slot] + (size_t)k * L->out_image_bytes, epoch, id);
    }
}

void helper_loop(vitrs_loader* L) {
    uint64_t seen = 0;
    for (;;) {
        int slot, n; size_t at; uint64_t epoch; const uint32_t* order;
        {
            std::unique_lock<std::mutex> lk(L->job_mu);
            L->job_cv.wait(lk, [&] { return L->stop || L->job_seq != seen; });
            if (L->stop) return;
            seen = L->job_seq;
            slot = L->job_slot; n = L->job_n; at = L->job_at; epoch = L->job_epoch; order = L->job_order;
        }
        work_on_job(L, slot, at, n, epoch, order);
        {
            std::lock_guard<std::mutex> lk(L->job_mu);
            L->job_done_helpers++;
        }
        L->job_cv.notify_all();
    }
}

// assemble one batch through the transform: post the job, work on it, wait for the helpers to drain it
void run_job(vitrs_loader* L, int slot, size_t at, int n, uint64_t epoch, const uint32_t* order) {
    {
        std::lock_guard<std::mutex> lk(L->job_mu);
        L->job_slot = slot; L->job_n = n; L->job_at = at; L->job_epoch = epoch; L->job_order = order;
        L->job_next.store(0);
        L->job_done_helpers = 0;
        L->job_seq++;
    }
    L->job_cv.notify_all();
    work_on_job(L, slot, at, n, epoch, order);
    std::unique_lock<std::mutex> lk(L->job_mu);
    L->job_cv.wait(lk, [&] { return L->stop || L->job_done_helpers == (int)L->helpers.size(); });
}

void fill_loop(vitrs_loader* L) {
    std::vector<uint32_t> order(L->num_records);
    uint64_t epoch = 0;
    for (;;) {
        for (size_t i = 0; i < L->num_records; ++i) order[i] = (uint32_t)i;
        if (L->shuffle) {  // Fisher-Yates from a per-epoch stream: the order is a function of (seed, epoch) only
            uint64_t s = L->seed * 0x100000001B3ull + epoch;
            for (size_t i = L->num_records - 1; i > 0; --i) {
                const size_t j = (size_t)(splitmix64(s) % (i + 1));
                const uint32_t t = order[i]; order[i] = order[j]; order[j] = t;
            }
        }
        // sharded: only whole rounds of `world` batches, so every rank sees the same number of batches per epoch
        const size_t full_batches = L->num_records / (size_t)L->batch;
        const size_t round_limit = L->world > 1 ? full_batches / L->world * L->world * (size_t)L->batch : L->num_records;
        size_t bi = 0;
        for (size_t at = 0; at < round_limit; at += (size_t)L->batch, ++bi) {
            const size_t n = L->num_records - at < (size_t)L->batch ? L->num_records - at : (size_t)L->batch;
            if (n < (size_t)L->batch && L->drop_last) break;
            if ((int)(bi % (size_t)L->world) != L->rank) continue;
            int slot;
            {
                std::unique_lock<std::mutex> lk(L->mu);
                L->cv.wait(lk, [&] { return L->stop || L->produced - L->released < (uint64_t)kSlots; });
                if (L->stop) return;
                slot = (int)(L->produced % kSlots);
            }
            if (L->transform) {
                run_job(L, slot, at, (int)n, epoch, order.data());
                if (L->stop) return;
            } else {
                for (size_t k = 0; k < n; ++k) {
                    const uint8_t* rec = L->data.data() + (size_t)order[at + k] * L->record_bytes;
                    L->labels[slot][k] = rec[L->label_bytes - 1];  // CIFAR-100: (coarse, fine) -> the fine label
                    memcpy(L->images[slot] + k * L->image_bytes, rec + L->label_bytes, L->image_bytes);
                }
            }
            {
                std::lock_guard<std::mutex> lk(L->mu);
                L->count[slot] = (int)n;
                L->epoch_of[slot] = epoch;
                L->produced++;
            }
            L->cv.notify_all();
        }
        ++epoch;
    }
}

}  // namespace

extern "C" {

static int open_impl(vitrs_ctx* ctx, const char* const* paths, int num_paths, const vitrs_loader_options& o, vitrs_loader** out) {
    const int image_size = o.image_size, label_bytes = o.label_bytes, batch = o.batch, world = o.world, rank = o.rank;
    const int out_size = o.out_size > 0 ? o.out_size : o.image_size;
    if (!out || !paths || num_paths < 1 || image_size < 1 || label_bytes < 1 || label_bytes > 4 || batch < 1 || world < 1 || rank < 0 ||
        rank >= world || out_size < 1 || out_size > 4096 || o.workers < 0 || o.workers > 256 || o.crop_pad < 0 || o.crop_pad > image_size) {
        if (ctx) vitrs_set_error(ctx, VITRS_ERR_ARG, "vitrs_loader_open: bad argument");
        return VITRS_ERR_ARG;
    }
    *out = nullptr;
    vitrs_loader* L = new vitrs_loader();
    L->ctx = ctx;
    L->stop = false;
    L->image_bytes = (size_t)3 * image_size * image_size;
    L->record_bytes = L->image_bytes + label_bytes;
    L->label_bytes = label_bytes; L->batch = batch; L->shuffle = o.shuffle; L->drop_last = o.drop_last; L->seed = o.seed;
    L->rank = rank; L->world = world;
    L->in_size = image_size; L->out_size = out_size;
    L->out_image_bytes = (size_t)3 * out_size * out_size;
    L->random_flip = o.random_flip != 0; L->crop_pad = o.crop_pad;
    L->transform = out_size != image_size || L->random_flip || L->crop_pad > 0 || o.workers > 1;
    if (L->transform) bilinear_taps(L);
    auto fail = [&](int code, const char* what, const char* path) {
        if (ctx) vitrs_set_error(ctx, code, "vitrs_loader_open: %s (%s)", what, path);
        for (int s = 0; s < kSlots; ++s) {  // (slots allocated before the failure; the rest are still null)
            if (ctx) { cudaFreeHost(L->images[s]); cudaFreeHost(L->labels[s]); }
            else { free(L->images[s]); free(L->labels[s]); }
        }
        delete L;
        return code;
    };
    for (int i = 0; i < num_paths; ++i) {
        FILE* f = fopen(paths[i], "rb");
        if (!f) return fail(VITRS_ERR_ARG, "cannot open", paths[i]);
        fseek(f, 0, SEEK_END);
        const long bytes = ftell(f);
        fseek(f, 0, SEEK_SET);
        if (bytes <= 0 || (size_t)bytes % L->record_bytes != 0) {
            fclose(f);
            return fail(VITRS_ERR_ARG, "file size is not a whole number of records", paths[i]);
        }
        const size_t at = L->data.size();
        L->data.resize(at + (size_t)bytes);
        const size_t got = fread(L->data.data() + at, 1, (size_t)bytes, f);
        fclose(f);
        if (got != (size_t)bytes) return fail(VITRS_ERR_ARG, "short read", paths[i]);
    }
    L->num_records = L->data.size() / L->record_bytes;
    if (o.drop_last && L->num_records < (size_t)batch) return fail(VITRS_ERR_ARG, "fewer records than one batch", paths[0]);
    if (world > 1 && L->num_records / (size_t)batch < (size_t)world) return fail(VITRS_ERR_ARG, "fewer batches than ranks", paths[0]);
    int max_label = 0;
    for (size_t r = 0; r < L->num_records; ++r) {
        const int lb = L->data[r * L->record_bytes + label_bytes - 1];
        if (lb > max_label) max_label = lb;
    }
    L->num_classes_seen = max_label + 1;
    for (int s = 0; s < kSlots; ++s) {
        const size_t ib = L->out_image_bytes * batch, lb = sizeof(int) * (size_t)batch;
        if (ctx) {
            cudaSetDevice(ctx->device);
            if (cudaMallocHost(&L->images[s], ib) != cudaSuccess || cudaMallocHost(&L->labels[s], lb) != cudaSuccess)
                return fail(VITRS_ERR_CUDA, "cudaMallocHost of a batch slot failed", paths[0]);
        } else {
            L->images[s] = (uint8_t*)malloc(ib);
            L->labels[s] = (int*)malloc(lb);
        }
    }
    for (int h = 1; h < o.workers; ++h) L->helpers.emplace_back(helper_loop, L);  // the loader thread is worker 0
    L->worker = std::thread(fill_loop, L);
    *out = L;
    return VITRS_OK;
}

int vitrs_loader_open(vitrs_ctx* ctx, const char* const* paths, int num_paths, int image_size, int label_bytes, int batch, int shuffle,
                      uint64_t seed, int drop_last, vitrs_loader** out) {
    return vitrs_loader_open_sharded(ctx, paths, num_paths, image_size, label_bytes, batch, shuffle, seed, drop_last, 0, 1, out);
}

int vitrs_loader_open_sharded(vitrs_ctx* ctx, const char* const* paths, int num_paths, int image_size, int label_bytes, int batch,
                              int shuffle, uint64_t seed, int drop_last, int rank, int world, vitrs_loader** out) {
    vitrs_loader_options o = {};
    o.image_size = image_size; o.label_bytes = label_bytes; o.batch = batch; o.shuffle = shuffle; o.seed = seed;
    o.drop_last = drop_last; o.rank = rank; o.world = world;
    return open_impl(ctx, paths, num_paths, o, out);
}

int vitrs_loader_open_transform(vitrs_ctx* ctx, const char* const* paths, int num_paths, const vitrs_loader_options* options,
                                vitrs_loader** out) {
    if (!options) return VITRS_ERR_ARG;
    return open_impl(ctx, paths, num_paths, *options, out);
}

int vitrs_loader_close(vitrs_loader* L) {
    if (!L) return VITRS_OK;
    {
        std::lock_guard<std::mutex> lk(L->mu);
        L->stop = true;
    }
    L->cv.notify_all();
    { std::lock_guard<std::mutex> lk(L->job_mu); }  // a helper is either before its predicate (sees stop) or already waiting
    L->job_cv.notify_all();
    if (L->worker.joinable()) L->worker.join();
    for (std::thread& h : L->helpers)
        if (h.joinable()) h.join();
    for (int s = 0; s < kSlots; ++s) {
        if (L->ctx) { cudaFreeHost(L->images[s]); cudaFreeHost(L->labels[s]); }
        else { free(L->images[s]); free(L->labels[s]); }
    }
    delete L;
    return VITRS_OK;
}

int vitrs_loader_image_size(vitrs_loader* L) { return L ? L->out_size : 0; }

int vitrs_loader_info(vitrs_loader* L, size_t* num_records, int* batches_per_epoch, int* num_classes_seen) {
    if (!L) return VITRS_ERR_ARG;
    if (num_records) *num_records = L->num_records;
    if (batches_per_epoch)
        *batches_per_epoch = L->world > 1 ? (int)(L->num_records / L->batch / L->world)
                                          : (int)(L->drop_last ? L->num_records / L->batch : (L->num_records + L->batch - 1) / L->batch);
    if (num_classes_seen) *num_classes_seen = L->num_classes_seen;
    return VITRS_OK;
}

// Blocks until the next batch is assembled; the pointers stay valid until the call after the next one (the slot of batch n is
// handed back to the loader thread when batch n + 1 is taken).
int vitrs_loader_next(vitrs_loader* L, const uint8_t** h_images, const int** h_labels, int* b, uint64_t* epoch) {
    if (!L || !h_images || !h_labels || !b) return VITRS_ERR_ARG;
    std::unique_lock<std::mutex> lk(L->mu);
    if (L->consumed > 0 && L->released < L->consumed - 1) {  // batch n - 2 is dead now
        L->released = L->consumed - 1;
        L->cv.notify_all();
    }
    L->cv.wait(lk, [&] { return L->produced > L->consumed; });
    const int slot = (int)(L->consumed % kSlots);
    *h_images = L->images[slot];
    *h_labels = L->labels[slot];
    *b = L->count[slot];
    if (epoch) *epoch = L->epoch_of[slot];
    L->consumed++;
    return VITRS_OK;
}

// One training step fed by the loader: take the next batch, stage the one after it (if already assembled) on the copy stream so
// that its H2D transfer overlaps this step, run the uint8 host step (H2D / step / D2H of the loss).
int vitrs_model_train_step_loader(vitrs_model* m, vitrs_loader* L, float lr, float beta1, float beta2, float eps, float weight_decay,
                                  float* loss_out, int* batch_out) {
    if (!m || !L) return VITRS_ERR_ARG;
    if (L->out_size != m->cfg.image_size || L->batch > m->max_batch)  // (the step copies batch x 3 x image^2 bytes out of the slot)
        return vitrs_set_error(m->ctx, VITRS_ERR_ARG, "loader delivers %d x %d images in batches of %d; the model takes %d x %d, at most %d",
                               L->out_size, L->out_size, L->batch, m->cfg.image_size, m->cfg.image_size, m->max_batch);
    const uint8_t* img;
    const int* lab;
    int b;
    VITRS_TRY(vitrs_loader_next(L, &img, &lab, &b, nullptr));
    if (batch_out) *batch_out = b;
    const uint8_t* nimg = nullptr;
    const int* nlab = nullptr;
    int nb = 0;
    {
        std::lock_guard<std::mutex> lk(L->mu);
        if (L->produced > L->consumed) {  // peek: do not consume
            const int slot = (int)(L->consumed % kSlots);
            nimg = L->images[slot]; nlab = L->labels[slot]; nb = L->count[slot];
        }
    }
    // the batch being consumed now was staged by the previous call when it was already there (prefetch_host_u8 matches by pointer)
    if (!L->prefetched) VITRS_TRY(vitrs_model_prefetch_host_u8(m, img, lab, b));
    L->prefetched = 0;
    if (nimg) {
        VITRS_TRY(vitrs_model_prefetch_host_u8(m, nimg, nlab, nb));
        L->prefetched = 1;
    }
    return vitrs_model_train_step_host_u8(m, img, 0 /* NCHW records */, lab, b, lr, beta1, beta2, eps, weight_decay, loss_out);
}

}  // extern "C"
