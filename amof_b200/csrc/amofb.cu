// amofb.cu -- the C ABI of libamofb.so (include/amofb.h): context, memory helpers and the three analyses.
// Build: see amof_b200/build.py (nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -Xcompiler -ffp-contract=off).
#include "common.cuh"
#include "prep.cuh"
#include "pair.cuh"
#include "pair_tiled.cuh"
#include "neigh.cuh"
#include "bad.cuh"
#include "msd.cuh"

#include <algorithm>
#include <thread>
#include <new>

// ================================================================================================
// lifetime and helpers
// ================================================================================================

static const char *k_no_ctx = "null context";

extern "C" const char *amofb_version(void) { return AMOFB_VERSION_STRING; }

extern "C" const char *amofb_last_error(const amofb_ctx *ctx) { return ctx ? ctx->err.c_str() : k_no_ctx; }

extern "C" int amofb_create(int device, amofb_ctx **out) {
    if (!out) return AMOFB_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) return AMOFB_ERR_CUDA;   // no CPU fallback
    if (device < 0 || device >= count) return AMOFB_ERR_ARG;
    amofb_ctx *ctx = new (std::nothrow) amofb_ctx();
    if (!ctx) return AMOFB_ERR_MEMORY;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return AMOFB_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return AMOFB_ERR_CUDA; }
    ctx->num_sms = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&ctx->s_compute, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->s_copy, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return AMOFB_ERR_CUDA;
    }
    const char *g = getenv("AMOFB_GUARD");
    ctx->guard = g && *g && *g != '0';
    *out = ctx;
    return AMOFB_OK;
}

static void pair_release(amofb_ctx *ctx);
static void bad_release(amofb_ctx *ctx);
static void msd_release(amofb_ctx *ctx);
static void neigh_release(amofb_ctx *ctx);
static void pool_destroy(amofb_ctx *ctx);

extern "C" int amofb_destroy(amofb_ctx *ctx) {
    if (!ctx) return AMOFB_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    pair_release(ctx);
    bad_release(ctx);
    msd_release(ctx);
    neigh_release(ctx);
    for (auto &p : ctx->pending_pair_events) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    for (auto &t : ctx->timer) if (t) cudaEventDestroy(t);
    pool_destroy(ctx);
    if (ctx->s_compute) cudaStreamDestroy(ctx->s_compute);
    if (ctx->s_copy) cudaStreamDestroy(ctx->s_copy);
    delete ctx;
    return AMOFB_OK;
}

static int drain_pair_events(amofb_ctx *ctx) {
    for (auto &p : ctx->pending_pair_events) {
        float ms = 0.f;
        CUDA_TRY(ctx, cudaEventSynchronize(p.second));
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, p.first, p.second));
        ctx->pair_ms += ms;
        ctx->pair_launches += 1;
        cudaEventDestroy(p.first);
        cudaEventDestroy(p.second);
    }
    ctx->pending_pair_events.clear();
    return AMOFB_OK;
}

extern "C" int amofb_sync(amofb_ctx *ctx) {
    if (!ctx) return AMOFB_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->s_copy));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->s_compute));
    return AMOFB_OK;
}

extern "C" int amofb_sync_copies(amofb_ctx *ctx) {
    if (!ctx) return AMOFB_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->s_copy));
    return AMOFB_OK;
}

extern "C" int64_t amofb_launch_count(const amofb_ctx *ctx) { return ctx ? ctx->launches : -1; }
extern "C" int64_t amofb_guard_violations(const amofb_ctx *ctx) { return ctx ? (ctx->guard ? ctx->guard_violations : -1) : -1; }

extern "C" int amofb_set_option(amofb_ctx *ctx, int option, int value) {
    if (!ctx) return AMOFB_ERR_ARG;
    switch (option) {
        case AMOFB_OPT_RDF_BIN_RULE:
            if (value != 0 && value != 1) return amofb_fail(ctx, AMOFB_ERR_ARG, "AMOFB_OPT_RDF_BIN_RULE takes 0 (divide) or 1 (multiply)");
            if (ctx->pair) return amofb_fail(ctx, AMOFB_ERR_STATE, "options cannot change while a pair analysis is open");
            ctx->rdf_bin_rule = value;
            return AMOFB_OK;
        default:
            return amofb_fail(ctx, AMOFB_ERR_ARG, "unknown option %d", option);
    }
}

extern "C" int amofb_set_profiling(amofb_ctx *ctx, int enabled) {
    if (!ctx) return AMOFB_ERR_ARG;
    ctx->profiling = enabled != 0;
    return AMOFB_OK;
}

extern "C" int amofb_pair_kernel_time(amofb_ctx *ctx, double *total_ms, int64_t *launches, int reset) {
    if (!ctx) return AMOFB_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    AMOFB_TRY(drain_pair_events(ctx));
    if (total_ms) *total_ms = ctx->pair_ms;
    if (launches) *launches = ctx->pair_launches;
    if (reset) { ctx->pair_ms = 0.0; ctx->pair_launches = 0; }
    return AMOFB_OK;
}

extern "C" int amofb_timer_mark(amofb_ctx *ctx, int slot) {
    if (!ctx || slot < 0 || slot >= 8) return AMOFB_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!ctx->timer[slot]) CUDA_TRY(ctx, cudaEventCreate(&ctx->timer[slot]));
    CUDA_TRY(ctx, cudaEventRecord(ctx->timer[slot], ctx->s_compute));
    return AMOFB_OK;
}
extern "C" int amofb_timer_elapsed(amofb_ctx *ctx, int from_slot, int to_slot, double *ms) {
    if (!ctx || !ms || from_slot < 0 || from_slot >= 8 || to_slot < 0 || to_slot >= 8) return AMOFB_ERR_ARG;
    if (!ctx->timer[from_slot] || !ctx->timer[to_slot]) return amofb_fail(ctx, AMOFB_ERR_STATE, "timer slot never marked");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->timer[from_slot]));
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->timer[to_slot]));
    float f = 0.f;
    CUDA_TRY(ctx, cudaEventElapsedTime(&f, ctx->timer[from_slot], ctx->timer[to_slot]));
    *ms = f;
    return AMOFB_OK;
}

extern "C" int amofb_host_alloc(amofb_ctx *ctx, uint64_t bytes, void **out) {
    if (!ctx || !out) return AMOFB_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return AMOFB_OK;
}
extern "C" int amofb_host_free(amofb_ctx *ctx, void *ptr) {
    if (!ctx) return AMOFB_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaFreeHost(ptr));
    return AMOFB_OK;
}
extern "C" int amofb_device_alloc(amofb_ctx *ctx, uint64_t bytes, void **out) {
    if (!ctx || !out) return AMOFB_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return amofb_fail(ctx, AMOFB_ERR_MEMORY, "cudaMalloc of %llu bytes failed", (unsigned long long)bytes); }
    CUDA_TRY(ctx, e);
    return AMOFB_OK;
}
extern "C" int amofb_device_free(amofb_ctx *ctx, void *ptr) {
    if (!ctx) return AMOFB_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaFree(ptr));
    return AMOFB_OK;
}
extern "C" int amofb_memcpy_h2d(amofb_ctx *ctx, void *dst_device, const void *src_host, uint64_t bytes) {
    if (!ctx) return AMOFB_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaMemcpyAsync(dst_device, src_host, bytes, cudaMemcpyHostToDevice, ctx->s_copy));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->s_copy));
    return AMOFB_OK;
}
extern "C" int amofb_memcpy_d2h(amofb_ctx *ctx, void *dst_host, const void *src_device, uint64_t bytes) {
    if (!ctx) return AMOFB_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->s_compute));
    CUDA_TRY(ctx, cudaMemcpyAsync(dst_host, src_device, bytes, cudaMemcpyDeviceToHost, ctx->s_copy));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->s_copy));
    return AMOFB_OK;
}

// AMOFB_GUARD (the pool's compute-sanitizer is closed, profiles/r02_sanitizer_refused.log): POOL_GUARD canary bytes before and
// after every pooled device block, compared when the block comes back; a block is only reused at its exact size
static int pool_guard_alloc(amofb_ctx *ctx, PoolBlock &blk) {
    char *base = nullptr;
    cudaError_t e = cudaMalloc(&base, blk.bytes + 2 * POOL_GUARD);
    if (e != cudaSuccess) return e == cudaErrorMemoryAllocation ? (cudaGetLastError(), AMOFB_ERR_MEMORY) : AMOFB_ERR_CUDA;
    cudaMemset(base, 0xA5, POOL_GUARD);
    cudaMemset(base + POOL_GUARD + blk.bytes, 0xA5, POOL_GUARD);
    blk.p = base + POOL_GUARD;
    blk.guarded = true;
    return AMOFB_OK;
}
static void pool_guard_check(amofb_ctx *ctx, const PoolBlock &blk) {
    if (!blk.guarded) return;
    static unsigned char h[2 * POOL_GUARD];
    cudaDeviceSynchronize();
    cudaMemcpy(h, (char *)blk.p - POOL_GUARD, POOL_GUARD, cudaMemcpyDeviceToHost);
    cudaMemcpy(h + POOL_GUARD, (char *)blk.p + blk.bytes, POOL_GUARD, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 2 * POOL_GUARD; ++i)
        if (h[i] != 0xA5) {
            ++ctx->guard_violations;
            fprintf(stderr, "amofb guard: block of %zu bytes written %s its bounds (byte %d of the canary)\n", blk.bytes,
                    i < POOL_GUARD ? "before" : "after", i < POOL_GUARD ? POOL_GUARD - i : i - POOL_GUARD);
            cudaMemset((char *)blk.p - POOL_GUARD, 0xA5, POOL_GUARD);
            cudaMemset((char *)blk.p + blk.bytes, 0xA5, POOL_GUARD);
            break;
        }
}
static void pool_block_free(const PoolBlock &b) {
    if (b.pinned) cudaFreeHost(b.p);
    else cudaFree(b.guarded ? (char *)b.p - POOL_GUARD : (char *)b.p);
}

static int pool_get(amofb_ctx *ctx, void **out, size_t bytes, bool pinned) {
    *out = nullptr;
    if (bytes < 256) bytes = 256;
    const bool guarded = ctx->guard && !pinned;
    int best = -1;
    for (int i = 0; i < (int)ctx->pool_idle.size(); ++i) {
        const PoolBlock &b = ctx->pool_idle[i];
        if (b.pinned == pinned && b.bytes >= bytes && (guarded ? b.bytes == bytes : b.bytes <= 2 * bytes + 4096) &&
            (best < 0 || b.bytes < ctx->pool_idle[best].bytes))
            best = i;
    }
    PoolBlock blk;
    if (best >= 0) {
        blk = ctx->pool_idle[best];
        ctx->pool_idle.erase(ctx->pool_idle.begin() + best);
    } else if (guarded) {
        blk.bytes = bytes;
        blk.pinned = false;
        int rc = pool_guard_alloc(ctx, blk);
        if (rc == AMOFB_ERR_MEMORY) {
            for (auto &b : ctx->pool_idle) pool_block_free(b);
            ctx->pool_idle.clear();
            rc = pool_guard_alloc(ctx, blk);
        }
        if (rc) return amofb_fail(ctx, rc, "guarded device allocation of %zu bytes failed", bytes);
    } else {
        blk.bytes = bytes;
        blk.pinned = pinned;
        cudaError_t e = pinned ? cudaHostAlloc(&blk.p, bytes, cudaHostAllocDefault) : cudaMalloc(&blk.p, bytes);
        if (e == cudaErrorMemoryAllocation) {
            // give idle blocks back to the driver and retry once
            cudaGetLastError();
            for (auto &b : ctx->pool_idle) pool_block_free(b);
            ctx->pool_idle.clear();
            e = pinned ? cudaHostAlloc(&blk.p, bytes, cudaHostAllocDefault) : cudaMalloc(&blk.p, bytes);
        }
        if (e == cudaErrorMemoryAllocation) {
            cudaGetLastError();
            return amofb_fail(ctx, AMOFB_ERR_MEMORY, "%s allocation of %zu bytes failed", pinned ? "pinned host" : "device", bytes);
        }
        CUDA_TRY(ctx, e);
    }
    ctx->pool_live[blk.p] = blk;
    *out = blk.p;
    return AMOFB_OK;
}

// callers make sure no enqueued work still uses p (the release functions synchronise first)
static void pool_put(amofb_ctx *ctx, void *p) {
    if (!p) return;
    auto it = ctx->pool_live.find(p);
    if (it == ctx->pool_live.end()) return;
    pool_guard_check(ctx, it->second);
    ctx->pool_idle.push_back(it->second);
    ctx->pool_live.erase(it);
}

static void pool_destroy(amofb_ctx *ctx) {
    for (auto &kv : ctx->pool_live) { pool_guard_check(ctx, kv.second); ctx->pool_idle.push_back(kv.second); }
    ctx->pool_live.clear();
    for (auto &b : ctx->pool_idle) pool_block_free(b);
    ctx->pool_idle.clear();
}

template <typename T>
static int dev_alloc(amofb_ctx *ctx, T **p, size_t count) {
    return pool_get(ctx, (void **)p, count * sizeof(T), false);
}

template <typename T>
static int pinned_alloc(amofb_ctx *ctx, T **p, size_t count) {
    return pool_get(ctx, (void **)p, count * sizeof(T), true);
}

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    if (!v || !*v) return dflt;
    return atoi(v);
}

// ================================================================================================
// frame batches: the streaming unit shared by the pair and bond-angle analyses
//   raw positions (H2D on s_copy) -> cell list (3 kernels on s_compute) -> analysis kernel
// two slots alternate so the copy of batch k+1 overlaps the kernels of batch k
// ================================================================================================

struct BatchSlot {
    double *d_raw = nullptr;
    FrameGeom *d_geom = nullptr, *h_geom = nullptr;
    uint32_t *d_cell_count = nullptr, *d_cell_start = nullptr, *d_cid = nullptr, *d_rank = nullptr;
    SAtom *d_sorted = nullptr;
    uint32_t *d_orig = nullptr;            // only when the batcher was asked for it (want_orig)
    int *d_wraps = nullptr;                // idem: [frames][n_atoms][3]
    unsigned long long *d_out = nullptr, *h_out = nullptr;   // per-frame outputs of the batch
    cudaEvent_t ev_h2d = nullptr, ev_done = nullptr;
    int frames = 0;        // frames of the batch in flight (0 = idle)
    int64_t first = 0;     // global index of its first frame
};

struct Batcher {
    int n_atoms = 0, cap_frames = 0, per_frame_out = 0;
    size_t cells_per_frame = 0;
    double rcut = 0.0;
    int cell_div = 1;
    double cell_widen = 1.0;             // cells this much wider than rcut / cell_div (bond angles on species-filtered frames)
    uint8_t *d_species = nullptr;
    uint8_t *d_species_keep = nullptr;   // optional species filter of the cell list (bond angles)
    int *d_keep_idx = nullptr;           // with the filter: original indices of the atoms that pass it
    int n_keep = 0;                      // atoms per frame that pass it (= n_atoms without a filter)
    int *d_centre_rank = nullptr;        // optional (bond angles): rank of every kept atom among the possible centres of a frame, or -1
    unsigned *d_centre_list = nullptr;   //   and the compact list the scatter kernel fills: [cap_frames * n_centres]
    int n_centres = 0;
    std::vector<int> keep_host;          // host copy of d_keep_idx
    bool gather = false;                 // host frames: only the kept atoms cross PCIe (gathered into h_compact by host threads)
    double *h_compact[2] = {nullptr, nullptr};      // per slot: [cap_frames][n_keep][3], page-locked
    int n_lists = 1;                     // cell lists per frame (bond angles: one per kept species)
    uint8_t list_of[AMOFB_MAX_SPECIES] = {0};
    bool want_orig = false;              // keep the original index of every sorted atom
    BatchSlot slot[2];
    int next = 0;
    int64_t frames_seen = 0;
    double volume_sum = 0.0;
    std::vector<unsigned long long> out_all;   // harvested per-frame outputs, frame-major
};

static void batcher_release(amofb_ctx *ctx, Batcher &b) {
    for (auto &s : b.slot) {
        pool_put(ctx, s.d_raw); pool_put(ctx, s.d_geom); pool_put(ctx, s.h_geom);
        pool_put(ctx, s.d_cell_count); pool_put(ctx, s.d_cell_start); pool_put(ctx, s.d_cid); pool_put(ctx, s.d_rank);
        pool_put(ctx, s.d_sorted); pool_put(ctx, s.d_orig); pool_put(ctx, s.d_wraps); pool_put(ctx, s.d_out); pool_put(ctx, s.h_out);
        if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
        if (s.ev_done) cudaEventDestroy(s.ev_done);
        s = BatchSlot();
    }
    pool_put(ctx, b.d_species); pool_put(ctx, b.d_species_keep); pool_put(ctx, b.d_keep_idx);
    pool_put(ctx, b.d_centre_rank); pool_put(ctx, b.d_centre_list);
    pool_put(ctx, b.h_compact[0]); pool_put(ctx, b.h_compact[1]);
    b.h_compact[0] = b.h_compact[1] = nullptr;
    b.d_species = nullptr; b.d_species_keep = nullptr; b.d_keep_idx = nullptr; b.d_centre_rank = nullptr; b.d_centre_list = nullptr;
}

// n_work: atoms per frame that enter the cell list (< n_atoms under a species filter): the batch is sized by the work, the raw
// frames of a batch are bounded by 1 GiB
static int batcher_init(amofb_ctx *ctx, Batcher &b, int n_atoms, const uint8_t *species, double rcut, int cell_div,
                        int per_frame_out, int max_frames = 0, int n_work = -1) {
    b.n_atoms = n_atoms;
    b.n_keep = n_atoms;
    b.rcut = rcut;
    b.cell_div = cell_div;
    b.per_frame_out = per_frame_out;
    long long target = env_int("AMOFB_BATCH_ATOMS", 1 << 22);     // 4 Mi atoms per batch: measured +22 % (BAD), +3 % (RDF) over 1 Mi
    long long cap = target / std::max(n_work > 0 ? n_work : n_atoms, 1);
    cap = std::min<long long>(cap, (1ll << 30) / (24ll * std::max(n_atoms, 1)));
    b.cap_frames = (int)std::min<long long>(std::max<long long>(cap, 1), 8192);
    if (max_frames > 0 && b.cap_frames > max_frames) b.cap_frames = max_frames;
    b.cells_per_frame = (size_t)(4.0 * n_atoms + 64.0 * AMOFB_MAX_SPECIES) + 4;      // per list at most 4 cells per atom + 64 (host_fill_geom)
    AMOFB_TRY(dev_alloc(ctx, &b.d_species, (size_t)n_atoms));
    CUDA_TRY(ctx, cudaMemcpy(b.d_species, species, (size_t)n_atoms, cudaMemcpyHostToDevice));
    size_t na = (size_t)b.cap_frames * n_atoms;
    for (auto &s : b.slot) {
        AMOFB_TRY(dev_alloc(ctx, &s.d_raw, na * 3));
        AMOFB_TRY(dev_alloc(ctx, &s.d_geom, (size_t)b.cap_frames));
        AMOFB_TRY(pinned_alloc(ctx, &s.h_geom, (size_t)b.cap_frames));
        AMOFB_TRY(dev_alloc(ctx, &s.d_cell_count, b.cells_per_frame * b.cap_frames));
        AMOFB_TRY(dev_alloc(ctx, &s.d_cell_start, b.cells_per_frame * b.cap_frames));
        AMOFB_TRY(dev_alloc(ctx, &s.d_cid, na));
        AMOFB_TRY(dev_alloc(ctx, &s.d_rank, na));
        AMOFB_TRY(dev_alloc(ctx, &s.d_sorted, na));
        if (b.want_orig) AMOFB_TRY(dev_alloc(ctx, &s.d_orig, na));
        if (b.want_orig) AMOFB_TRY(dev_alloc(ctx, &s.d_wraps, na * 3));
        if (per_frame_out > 0) {
            AMOFB_TRY(dev_alloc(ctx, &s.d_out, (size_t)b.cap_frames * per_frame_out));
            AMOFB_TRY(pinned_alloc(ctx, &s.h_out, (size_t)b.cap_frames * per_frame_out));
        }
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&s.ev_h2d, cudaEventDisableTiming));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
    }
    return AMOFB_OK;
}

// species filter of the cell list: atoms whose species has keep[] == 0 never enter it
// centre_mask (optional, [AMOFB_MAX_SPECIES]): species that can be the centre of an angle; the scatter kernel then also lists them
static int batcher_set_filter(amofb_ctx *ctx, Batcher &b, const uint8_t *species, const uint8_t *keep, const unsigned long long *centre_mask = nullptr) {
    std::vector<int> idx;
    idx.reserve((size_t)b.n_atoms);
    for (int i = 0; i < b.n_atoms; ++i)
        if (keep[species[i]]) idx.push_back(i);
    b.n_keep = (int)idx.size();
    if (centre_mask && !idx.empty()) {
        std::vector<int> crank(idx.size());
        int nc = 0;
        for (size_t k = 0; k < idx.size(); ++k) crank[k] = centre_mask[species[idx[k]]] ? nc++ : -1;
        b.n_centres = nc;
        // every kept atom a possible centre (Bad on 'Zn-N' asks for N-Zn-N and Zn-N-Zn): a list would only replace the cell order of
        // the threads by the file order (measured on C4: 277 k instead of 311 k frames/s)
        if (nc == (int)idx.size()) b.n_centres = 0;
        else {
            AMOFB_TRY(dev_alloc(ctx, &b.d_centre_rank, crank.size()));
            CUDA_TRY(ctx, cudaMemcpy(b.d_centre_rank, crank.data(), sizeof(int) * crank.size(), cudaMemcpyHostToDevice));
            AMOFB_TRY(dev_alloc(ctx, &b.d_centre_list, (size_t)std::max(nc, 1) * (size_t)b.cap_frames));
        }
    }
    // frames in host memory: when under 60 % of the atoms are kept, host threads gather them and only they are copied (C4 'Zn-N':
    // 29 % of 1.175 MB per frame; the kept atoms come in runs of ten between others, too short for a strided DMA)
    b.keep_host = idx;
    b.gather = !idx.empty() && 10 * (long long)idx.size() < 6 * (long long)b.n_atoms && !env_int("AMOFB_NO_HOST_GATHER", 0);
    AMOFB_TRY(dev_alloc(ctx, &b.d_species_keep, (size_t)AMOFB_MAX_SPECIES));
    CUDA_TRY(ctx, cudaMemcpy(b.d_species_keep, keep, AMOFB_MAX_SPECIES, cudaMemcpyHostToDevice));
    AMOFB_TRY(dev_alloc(ctx, &b.d_keep_idx, idx.size() + 1));
    if (!idx.empty()) CUDA_TRY(ctx, cudaMemcpy(b.d_keep_idx, idx.data(), sizeof(int) * idx.size(), cudaMemcpyHostToDevice));
    return AMOFB_OK;
}

// wait for the slot's batch and move its per-frame outputs to out_all
static int batcher_harvest(amofb_ctx *ctx, Batcher &b, BatchSlot &s) {
    if (s.frames == 0) return AMOFB_OK;
    CUDA_TRY(ctx, cudaEventSynchronize(s.ev_done));
    if (b.per_frame_out > 0) {
        size_t need = (size_t)(s.first + s.frames) * b.per_frame_out;
        if (b.out_all.size() < need) b.out_all.resize(need, 0ull);
        memcpy(b.out_all.data() + (size_t)s.first * b.per_frame_out, s.h_out,
               sizeof(unsigned long long) * (size_t)s.frames * b.per_frame_out);
    }
    s.frames = 0;
    return AMOFB_OK;
}

#ifndef HOST_GATHER_AHEAD
#define HOST_GATHER_AHEAD 64        // measured on C4 (16 threads): none 59-61 k, 8 ahead 66-69 k, 64 ahead 70-72 k frames/s end to end
#endif
// entries a frame takes in the batch-wide cell_count / cell_start arrays: its n_lists * ncell + 1 counters rounded up to 16 bytes, so
// that every frame's array starts aligned and k_cell_scan takes its 16-byte path (with the odd natural size three frames in four did not)
static inline int cs_stride(int n_lists, int ncell) { return (n_lists * ncell + 1 + 3) & ~3; }

// out[f][k] = pos[f][keep[k]] for nf frames, split over host threads by frame
static void host_gather_atoms(const double *pos, int nf, int n_atoms, const std::vector<int> &keep, double *out) {
    const size_t nk = keep.size();
    auto work = [&](int f0, int f1) {
        for (int f = f0; f < f1; ++f) {
            const double *src = pos + 3 * (size_t)f * n_atoms;
            double *dst = out + 3 * (size_t)f * nk;
            for (size_t k = 0; k < nk; ++k) {
                // the kept atoms are scattered: without a hint a core only has its few demand misses in flight
                if (k + HOST_GATHER_AHEAD < nk) __builtin_prefetch(src + 3 * (size_t)keep[k + HOST_GATHER_AHEAD], 0, 0);
                const double *p = src + 3 * (size_t)keep[k];
                dst[3 * k] = p[0]; dst[3 * k + 1] = p[1]; dst[3 * k + 2] = p[2];
            }
        }
    };
    unsigned hc = std::thread::hardware_concurrency();
    int threads = (int)std::min<unsigned>(hc ? hc : 1u, 16u);
    if (int f = env_int("AMOFB_HOST_THREADS", 0)) threads = f;
    threads = std::max(1, std::min(threads, nf));
    if (threads == 1) { work(0, nf); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(work, (int)((long long)nf * t / threads), (int)((long long)nf * (t + 1) / threads));
    for (auto &t : pool) t.join();
}

// Stage one batch (<= cap_frames frames): geometry, H2D (or adopt a device pointer) and the cell list.
// On return the slot's sorted atoms / cell_start are valid on s_compute; the caller enqueues its analysis kernel
// and then calls batcher_commit.
static int batcher_stage(amofb_ctx *ctx, Batcher &b, int nf, const double *pos, bool pos_on_device, const double *cell,
                         BatchSlot **out_slot, const double **out_raw) {
    BatchSlot &s = b.slot[b.next];
    AMOFB_TRY(batcher_harvest(ctx, b, s));
    int cs_off = 0;
    for (int f = 0; f < nf; ++f) {
        FrameGeom &g = s.h_geom[f];
        if (!host_fill_geom(g, cell + 9 * (size_t)f, b.rcut, b.cell_div, std::max(1, b.n_keep / b.n_lists), b.cell_widen))
            return amofb_fail(ctx, AMOFB_ERR_GEOMETRY, "frame %lld: singular cell or cell far smaller than the cutoff %g",
                              (long long)(b.frames_seen + f), b.rcut);
        g.cs_off = cs_off;
        g.frame_id = (int)(b.frames_seen + f);
        cs_off += cs_stride(b.n_lists, g.ncell);
        b.volume_sum += host_cell_volume(cell + 9 * (size_t)f);
    }
    const double *raw = pos;
    size_t bytes = sizeof(double) * 3 * (size_t)nf * b.n_atoms;
    bool compact = false;
    if (!pos_on_device && b.gather) {
        double *&h = b.h_compact[b.next];
        if (!h) AMOFB_TRY(pinned_alloc(ctx, &h, (size_t)b.cap_frames * b.n_keep * 3));
        host_gather_atoms(pos, nf, b.n_atoms, b.keep_host, h);
        pos = h;
        bytes = sizeof(double) * 3 * (size_t)nf * b.n_keep;
        compact = true;
    }
    if (!pos_on_device) {
        CUDA_TRY(ctx, cudaMemcpyAsync(s.d_raw, pos, bytes, cudaMemcpyHostToDevice, ctx->s_copy));
        CUDA_TRY(ctx, cudaEventRecord(s.ev_h2d, ctx->s_copy));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->s_compute, s.ev_h2d, 0));
        raw = s.d_raw;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(s.d_geom, s.h_geom, sizeof(FrameGeom) * nf, cudaMemcpyHostToDevice, ctx->s_compute));
    CUDA_TRY(ctx, cudaMemsetAsync(s.d_cell_count, 0, sizeof(uint32_t) * (size_t)cs_off, ctx->s_compute));
    if (b.per_frame_out > 0)
        CUDA_TRY(ctx, cudaMemsetAsync(s.d_out, 0, sizeof(unsigned long long) * (size_t)nf * b.per_frame_out, ctx->s_compute));
    PrepArgs pa;
    pa.raw = raw; pa.geom = s.d_geom; pa.species = b.d_species;
    pa.cell_count = s.d_cell_count; pa.cell_start = s.d_cell_start; pa.cid = s.d_cid; pa.rank = s.d_rank;
    pa.sorted = s.d_sorted; pa.n_atoms = b.n_atoms; pa.n_frames = nf;
    pa.species_keep = b.d_species_keep;
    pa.keep_idx = b.d_keep_idx;
    pa.n_keep = b.n_keep;
    pa.orig = s.d_orig;
    pa.wraps = s.d_wraps;
    pa.slot = nullptr;
    pa.centre_rank = b.d_centre_rank; pa.centre_list = b.d_centre_list; pa.n_centres = b.n_centres;
    pa.n_lists = b.n_lists;
    memcpy(pa.list_of, b.list_of, sizeof pa.list_of);
    pa.raw_compact = compact ? 1 : 0;
    long long total = (long long)nf * (b.d_keep_idx ? b.n_keep : b.n_atoms);
    int blocks = (int)std::min<long long>((total + 255) / 256, (long long)ctx->num_sms * 16);
    if (blocks < 1) blocks = 1;
    if (total > 0) {
        k_cell_assign<<<blocks, 256, 0, ctx->s_compute>>>(pa);
        k_cell_scan<<<nf, 1024, 0, ctx->s_compute>>>(pa);
        k_cell_scatter<<<blocks, 256, 0, ctx->s_compute>>>(pa);
        ctx->launches += 3;
        CUDA_TRY(ctx, cudaGetLastError());
    } else {
        k_cell_scan<<<nf, 1024, 0, ctx->s_compute>>>(pa);
        ctx->launches += 1;
        CUDA_TRY(ctx, cudaGetLastError());
    }
    *out_slot = &s;
    *out_raw = raw;
    return AMOFB_OK;
}

static int batcher_commit(amofb_ctx *ctx, Batcher &b, BatchSlot &s, int nf) {
    if (b.per_frame_out > 0)
        CUDA_TRY(ctx, cudaMemcpyAsync(s.h_out, s.d_out, sizeof(unsigned long long) * (size_t)nf * b.per_frame_out,
                                      cudaMemcpyDeviceToHost, ctx->s_compute));
    CUDA_TRY(ctx, cudaEventRecord(s.ev_done, ctx->s_compute));
    s.frames = nf;
    s.first = b.frames_seen;
    b.frames_seen += nf;
    b.next ^= 1;
    return AMOFB_OK;
}

static int batcher_drain(amofb_ctx *ctx, Batcher &b) {
    // harvest in submission order
    BatchSlot *a = &b.slot[b.next], *c = &b.slot[b.next ^ 1];
    AMOFB_TRY(batcher_harvest(ctx, b, *a));
    AMOFB_TRY(batcher_harvest(ctx, b, *c));
    return AMOFB_OK;
}

static inline int fold_key(int a, int b, int S) {
    int lo = a < b ? a : b, hi = a < b ? b : a;
    return lo * S - lo * (lo - 1) / 2 + (hi - lo);
}

// ================================================================================================
// pair analysis
// ================================================================================================

struct PairState {
    Batcher bt;
    int n_species = 0, nkeys = 0, nbins = 0;
    bool has_rdf = false, has_cn = false, smem_hist = false;
    double rmax = 0.0;
    double *d_edge2 = nullptr, *d_cnthr2 = nullptr;
    uint16_t *d_keyidx = nullptr;
    unsigned long long *d_slabs = nullptr, *d_hist = nullptr;
    int grid = 0;
    size_t smem = 0;
    double r2search = 0.0, r2max = 0.0;
    float inv_dr_f = 0.f, bin_margin = 0.f;
    double cn_r2max = 0.0;
    // tiled path (pair_tiled.cuh)
    bool tiled = false, cn_wide = false;
    PairTile *d_tiles = nullptr;
    int *d_ntiles = nullptr, *d_flags = nullptr;
    uint8_t *d_hard = nullptr;
    int tile_cap = 0, tile_grid = 0, max_tiles = 0;
    size_t tile_smem = 0, hard_bytes = 0;
};

static void pair_release(amofb_ctx *ctx) {
    PairState *p = ctx->pair;
    if (!p) return;
    cudaStreamSynchronize(ctx->s_copy);
    cudaStreamSynchronize(ctx->s_compute);
    batcher_release(ctx, p->bt);
    pool_put(ctx, p->d_edge2); pool_put(ctx, p->d_cnthr2); pool_put(ctx, p->d_keyidx); pool_put(ctx, p->d_slabs); pool_put(ctx, p->d_hist);
    pool_put(ctx, p->d_tiles); pool_put(ctx, p->d_ntiles); pool_put(ctx, p->d_flags); pool_put(ctx, p->d_hard);
    delete p;
    ctx->pair = nullptr;
}

static const void *tiled_kernel(bool has_cn, bool cn_wide) {
    if (!has_cn) return (const void *)k_pair_tiled<false, false>;
    return cn_wide ? (const void *)k_pair_tiled<true, true> : (const void *)k_pair_tiled<true, false>;
}

template <bool R, bool C, bool M>
static int pair_configure(amofb_ctx *ctx, PairState *p) {
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_pair<R, C, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
    int per_sm = 0;
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pair<R, C, M>, PAIR_TILE, p->smem));
    if (per_sm < 1) return amofb_fail(ctx, AMOFB_ERR_CUDA, "pair kernel does not fit on an SM (smem %zu)", p->smem);
    int force = env_int("AMOFB_PAIR_BLOCKS_PER_SM", 0);
    if (force > 0 && force < per_sm) per_sm = force;
    p->grid = ctx->num_sms * per_sm;
    return AMOFB_OK;
}

template <bool R, bool C, bool M>
static void pair_launch(amofb_ctx *ctx, PairState *p, const PairArgs &a, int grid) {
    k_pair<R, C, M><<<grid, PAIR_TILE, p->smem, ctx->s_compute>>>(a);
}

extern "C" int amofb_pair_begin(amofb_ctx *ctx, int n_atoms, int n_species, const uint8_t *species, double rmax,
                                int nbins, const double *cn_cutoff) {
    if (!ctx) return AMOFB_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (ctx->pair) return amofb_fail(ctx, AMOFB_ERR_STATE, "pair analysis already open; call amofb_pair_finish first");
    if (n_atoms < 0 || n_species < 1 || n_species > AMOFB_MAX_SPECIES || (n_atoms > 0 && !species))
        return amofb_fail(ctx, AMOFB_ERR_ARG, "bad n_atoms/n_species (n_species must be 1..%d)", AMOFB_MAX_SPECIES);
    if (nbins < 0 || (nbins > 0 && !(rmax > 0.0 && isfinite(rmax))))
        return amofb_fail(ctx, AMOFB_ERR_ARG, "nbins > 0 needs a finite rmax > 0");
    if (nbins == 0 && !cn_cutoff) return amofb_fail(ctx, AMOFB_ERR_ARG, "nothing to compute: nbins == 0 and no cutoffs");
    for (int i = 0; i < n_atoms; ++i)
        if (species[i] >= n_species) return amofb_fail(ctx, AMOFB_ERR_ARG, "species[%d] = %d out of range", i, species[i]);
    const int S = n_species;
    double cut_max = 0.0;
    if (cn_cutoff) {
        for (int a = 0; a < S; ++a)
            for (int b = 0; b < S; ++b) {
                double c = cn_cutoff[a * S + b];
                if (!(c >= 0.0) || !isfinite(c)) return amofb_fail(ctx, AMOFB_ERR_ARG, "cutoff[%d][%d] must be finite and >= 0", a, b);
                if (c != cn_cutoff[b * S + a]) return amofb_fail(ctx, AMOFB_ERR_ARG, "cutoff matrix must be symmetric");
                cut_max = std::max(cut_max, c);
            }
    }
    PairState *p = new (std::nothrow) PairState();
    if (!p) return AMOFB_ERR_MEMORY;
    ctx->pair = p;
    p->n_species = S;
    p->nkeys = S * (S + 1) / 2;
    p->nbins = nbins;
    p->has_rdf = nbins > 0;
    p->has_cn = cn_cutoff != nullptr;
    p->rmax = rmax;

    // exact thresholds in d2 (P4, P5)
    std::vector<double> edge2((size_t)nbins + 1, 0.0), cnthr((size_t)p->nkeys, 0.0);
    if (p->has_rdf) {
        const double dr = rmax / (double)nbins, inv_dr = (double)nbins / rmax;
        const int rule = ctx->rdf_bin_rule;
        if (ctx->edge_nbins == nbins && ctx->edge_rmax == rmax && ctx->edge_rule == rule && (int)ctx->edge_cache.size() == nbins + 1) {
            edge2 = ctx->edge_cache;                  // same axis as the previous analysis: reuse the bisected table
        } else {
            for (int b = 1; b <= nbins; ++b) {
                const double fb = (double)b;
                double guess = (fb * dr) * (fb * dr);
                // pin U1 (what asap3 evaluates is not on disk): the quotient d / dr, or the product d * (nBins / rMax)
                if (rule == 0) edge2[b] = host_threshold(guess, [&](double t) { return sqrt(t) / dr >= fb; });
                else edge2[b] = host_threshold(guess, [&](double t) { return sqrt(t) * inv_dr >= fb; });
            }
            ctx->edge_cache = edge2; ctx->edge_rmax = rmax; ctx->edge_nbins = nbins; ctx->edge_rule = rule;
        }
        p->r2max = edge2[nbins];
        p->inv_dr_f = (float)((double)nbins / rmax);
        // fp32 estimate of d/dr: relative error < 2^-21 (conversion, MUFU.SQRT, multiply, fma) -> absolute < nbins*2^-21
        double err = (double)nbins * 4.8e-7;
        p->bin_margin = err < 0.2 ? (float)(2.0 * err + 1e-3) : 0.f;    // 0 = use the searching fallback
    }
    if (p->has_cn)
        for (int a = 0; a < S; ++a)
            for (int b = a; b < S; ++b) {
                double c = cn_cutoff[a * S + b];
                cnthr[fold_key(a, b, S)] = c > 0.0 ? host_threshold(c * c, [&](double t) { return sqrt(t) >= c; }) : 0.0;
            }
    p->r2search = p->r2max;
    for (double t : cnthr) { p->r2search = std::max(p->r2search, t); p->cn_r2max = std::max(p->cn_r2max, t); }
    std::vector<uint16_t> keyidx((size_t)S * S);
    for (int a = 0; a < S; ++a)
        for (int b = 0; b < S; ++b) keyidx[a * S + b] = (uint16_t)fold_key(a, b, S);

    int rc = AMOFB_OK;
    auto fail = [&](int code) { pair_release(ctx); return code; };
    double rcut = std::max(p->has_rdf ? rmax : 0.0, cut_max);
    if (!(rcut > 0.0)) rcut = 1e-3;   // all cutoffs zero: nothing will be counted, any grid works
    int cell_div = env_int("AMOFB_CELL_DIV", p->has_rdf ? 2 : 1);
    if (cell_div < 1) cell_div = 1;
    if ((rc = batcher_init(ctx, p->bt, n_atoms, species, rcut, cell_div, p->has_cn ? p->nkeys : 0))) return fail(rc);
    if (!p->has_rdf && !env_int("AMOFB_CN_NO_FILTER", 0)) {
        // counts only (amof.cn): a species without a positive cutoff towards any species can neither count nor be counted,
        // so it never enters the cell list (as for the bond angles)
        uint8_t keep[AMOFB_MAX_SPECIES];
        memset(keep, 0, sizeof keep);
        for (int x = 0; x < S; ++x)
            for (int y = 0; y < S; ++y)
                if (cn_cutoff[x * S + y] > 0.0) keep[x] = 1;
        if ((rc = batcher_set_filter(ctx, p->bt, species, keep))) return fail(rc);
    }
    if ((rc = dev_alloc(ctx, &p->d_edge2, edge2.size()))) return fail(rc);
    if ((rc = dev_alloc(ctx, &p->d_cnthr2, cnthr.size()))) return fail(rc);
    if ((rc = dev_alloc(ctx, &p->d_keyidx, keyidx.size()))) return fail(rc);
    cudaMemcpy(p->d_edge2, edge2.data(), sizeof(double) * edge2.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(p->d_cnthr2, cnthr.data(), sizeof(double) * cnthr.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(p->d_keyidx, keyidx.data(), sizeof(uint16_t) * keyidx.size(), cudaMemcpyHostToDevice);

    const size_t hist_n = (size_t)p->nkeys * nbins;
    size_t smem_full = sizeof(double) * ((size_t)nbins + 1) + sizeof(double) * p->nkeys + sizeof(uint32_t) * hist_n +
                       sizeof(uint32_t) * p->nkeys + sizeof(uint16_t) * S * S + 16;
    size_t smem_lite = sizeof(double) * p->nkeys + sizeof(uint32_t) * p->nkeys + sizeof(uint16_t) * S * S + 16;
    size_t budget = (size_t)ctx->max_smem_optin > 2048 ? (size_t)ctx->max_smem_optin - 2048 : 0;
    p->smem_hist = p->has_rdf && smem_full <= budget && !env_int("AMOFB_FORCE_GLOBAL_HIST", 0);
    p->smem = (p->smem_hist ? smem_full : smem_lite);
    if (p->has_rdf && p->has_cn) rc = p->smem_hist ? pair_configure<true, true, true>(ctx, p) : pair_configure<true, true, false>(ctx, p);
    else if (p->has_rdf) rc = p->smem_hist ? pair_configure<true, false, true>(ctx, p) : pair_configure<true, false, false>(ctx, p);
    else rc = pair_configure<false, true, false>(ctx, p);
    if (rc) return fail(rc);
    if ((rc = dev_alloc(ctx, &p->d_hist, hist_n))) return fail(rc);
    if (hist_n) cudaMemset(p->d_hist, 0, sizeof(unsigned long long) * hist_n);
    // tiled kernel: shared memory = fixed tables + as many staged atoms as still let two blocks share an SM
    if (p->smem_hist && n_atoms > 0 && p->bin_margin > 0.f && !env_int("AMOFB_PAIR_GENERIC", 0)) {
        size_t fixed = smem_full + sizeof(int) * (TILE_OFF_WORDS + TILE_PRE_WORDS) + 64;
        const size_t per_atom = sizeof(SAtom) + 1;                            // + the image code byte
        int per_sm_target = env_int("AMOFB_TILE_BLOCKS_PER_SM", TILE_MIN_BLOCKS);
        size_t sm_total = (size_t)ctx->max_smem_optin + 1024;                 // 227 KB opt-in + 1 KB reserved per block
        size_t per_block = sm_total / std::max(per_sm_target, 1) - 1024 - TILE_STATIC_SMEM;     // driver reserve + the kernel's static shared memory
        if (per_block + TILE_STATIC_SMEM > (size_t)ctx->max_smem_optin) per_block = (size_t)ctx->max_smem_optin - TILE_STATIC_SMEM;
        long long cap = per_block > fixed ? (long long)((per_block - fixed) / per_atom) : 0;
        int cap_env = env_int("AMOFB_TILE_CAP", 0);
        if (cap_env > 0 && cap_env < cap) cap = cap_env;
        if (cap > 2000) cap = 2000;     // staged indices and flat candidate indices are packed into 16 bits each
        if (cap >= 256) {
            p->tile_cap = (int)cap;
            p->tile_smem = fixed + per_atom * (size_t)cap;
            int per_sm = 0;
            p->cn_wide = p->has_cn && p->cn_r2max > p->r2max;
            const void *kfn = tiled_kernel(p->has_cn, p->cn_wide);
            cudaError_t e1 = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->tile_smem);
            if (e1 == cudaSuccess) e1 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, TILE_THREADS, p->tile_smem);
            if (e1 == cudaSuccess && per_sm >= 1) {
                p->tiled = true;
                p->tile_grid = ctx->num_sms * per_sm;
                // a tile has at least one home atom and usually ~100 (a column chunk); N/2 + 1024 per frame is ample, and an
                // overflow is detected and reported (d_flags) rather than silently dropped
                p->max_tiles = (int)std::min<size_t>(std::min<size_t>(p->bt.cells_per_frame, (size_t)n_atoms / 2 + 1024) * p->bt.cap_frames,
                                                     (size_t)1 << 26);
                p->hard_bytes = p->bt.cells_per_frame * p->bt.cap_frames;
                if ((rc = dev_alloc(ctx, &p->d_tiles, (size_t)p->max_tiles))) return fail(rc);
                if ((rc = dev_alloc(ctx, &p->d_ntiles, 4))) return fail(rc);
                if ((rc = dev_alloc(ctx, &p->d_flags, 1))) return fail(rc);
                if ((rc = dev_alloc(ctx, &p->d_hard, p->hard_bytes))) return fail(rc);
                cudaMemset(p->d_flags, 0, sizeof(int));
            } else cudaGetLastError();
        }
    }
    if (p->smem_hist) {
        size_t nslab = (size_t)std::max(p->grid, p->tile_grid);
        if ((rc = dev_alloc(ctx, &p->d_slabs, hist_n * nslab))) return fail(rc);
        cudaMemset(p->d_slabs, 0, sizeof(unsigned long long) * hist_n * nslab);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { amofb_fail(ctx, AMOFB_ERR_CUDA, "pair_begin: %s", cudaGetErrorString(e)); return fail(AMOFB_ERR_CUDA); }
    return AMOFB_OK;
}

static int pair_push_impl(amofb_ctx *ctx, int n_frames, const double *pos, bool on_device, const double *cell) {
    if (!ctx) return AMOFB_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    PairState *p = ctx->pair;
    if (!p) return amofb_fail(ctx, AMOFB_ERR_STATE, "amofb_pair_push before amofb_pair_begin");
    if (n_frames < 0 || (n_frames > 0 && (!cell || (!pos && p->bt.n_atoms > 0))))
        return amofb_fail(ctx, AMOFB_ERR_ARG, "bad push arguments");
    Batcher &b = p->bt;
    for (int done = 0; done < n_frames;) {
        int nf = std::min(b.cap_frames, n_frames - done);
        BatchSlot *s = nullptr;
        const double *raw = nullptr;
        AMOFB_TRY(batcher_stage(ctx, b, nf, pos + 3 * (size_t)done * b.n_atoms, on_device, cell + 9 * (size_t)done, &s, &raw));
        PairArgs a;
        a.sorted = s->d_sorted; a.geom = s->d_geom; a.cell_start = s->d_cell_start;
        a.edge2 = p->d_edge2; a.cn_thr2 = p->d_cnthr2; a.keyidx = p->d_keyidx;
        a.slabs = p->d_slabs; a.ghist = p->d_hist; a.cn_out = s->d_out;
        a.r2search = p->r2search; a.r2max = p->r2max; a.inv_dr_f = p->inv_dr_f; a.bin_margin = p->bin_margin; a.cn_r2max = p->cn_r2max;
        a.n_atoms = b.n_atoms; a.n_frames = nf; a.n_species = p->n_species; a.nkeys = p->nkeys; a.nbins = p->nbins;
        a.n_sorted = b.n_keep;
        a.tiles_per_frame = (b.n_keep + PAIR_TILE - 1) / PAIR_TILE;
        a.hard_mask = nullptr; a.n_hard = nullptr;
        long long tiles = (long long)nf * a.tiles_per_frame;
        if (tiles > 0) {
            // smem-histogram mode always launches the full grid: every block owns a slab
            int grid = p->smem_hist ? p->grid : (int)std::min<long long>(tiles, p->grid);
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            if (ctx->profiling) {
                CUDA_TRY(ctx, cudaEventCreate(&e0));
                CUDA_TRY(ctx, cudaEventCreate(&e1));
                CUDA_TRY(ctx, cudaEventRecord(e0, ctx->s_compute));
            }
            // the tiled kernel needs every frame's half stencil to fit its offset table
            bool tiled = p->tiled;
            size_t ncell_total = 0;
            long long columns = 0;
            int uniform_cols = nf > 0 ? s->h_geom[0].nc[0] * s->h_geom[0].nc[1] : 0;
            for (int f = 0; f < nf; ++f)
                if (s->h_geom[f].nc[0] * s->h_geom[f].nc[1] != uniform_cols) uniform_cols = 0;
            for (int f = 0; f < nf && tiled; ++f) {
                const FrameGeom &g = s->h_geom[f];
                int R = (g.m[1] + 1) + g.m[0] * (2 * g.m[1] + 1);
                if (R > TILE_MAX_ROWS || TILE_MAX_ENTRIES / R < 2 * g.m[2] + 1) tiled = false;
                if (g.m[0] > g.nc[0] || g.m[1] > g.nc[1] || g.m[2] > g.nc[2]) tiled = false;     // image shifts beyond +-1 cell vector: generic kernel
                ncell_total += (size_t)cs_stride(1, g.ncell);
                columns += (long long)g.nc[0] * g.nc[1];
            }
            if (tiled) {
                CUDA_TRY(ctx, cudaMemsetAsync(p->d_ntiles, 0, sizeof(int) * 4, ctx->s_compute));
                CUDA_TRY(ctx, cudaMemsetAsync(p->d_hard, 0, ncell_total, ctx->s_compute));
                PlanArgs pl;
                pl.geom = s->d_geom; pl.cell_start = s->d_cell_start; pl.tiles = p->d_tiles; pl.n_tiles = p->d_ntiles;
                pl.flags = p->d_flags; pl.hard = p->d_hard; pl.n_frames = nf; pl.cap = p->tile_cap; pl.max_tiles = p->max_tiles;
                pl.uniform_cols = uniform_cols;
                k_pair_plan<<<(unsigned)((columns + 3) / 4), 128, 0, ctx->s_compute>>>(pl);      // one warp per column
                TiledArgs ta;
                ta.p = a; ta.tiles = p->d_tiles; ta.n_tiles = p->d_ntiles; ta.cap = p->tile_cap; ta.max_tiles = p->max_tiles;
                ta.p.hard_mask = nullptr; ta.p.n_hard = nullptr;
                {
                    void *kargs[] = {(void *)&ta};
                    CUDA_TRY(ctx, cudaLaunchKernel(tiled_kernel(p->has_cn, p->cn_wide), dim3(p->tile_grid), dim3(TILE_THREADS), kargs,
                                                   p->tile_smem, ctx->s_compute));
                }
                ctx->launches += 2;
                CUDA_TRY(ctx, cudaGetLastError());
                a.hard_mask = p->d_hard;          // clean-up launch: only the home cells the plan could not tile
                a.n_hard = p->d_ntiles + 1;
            }
            if (p->has_rdf && p->has_cn) { if (p->smem_hist) pair_launch<true, true, true>(ctx, p, a, grid); else pair_launch<true, true, false>(ctx, p, a, grid); }
            else if (p->has_rdf) { if (p->smem_hist) pair_launch<true, false, true>(ctx, p, a, grid); else pair_launch<true, false, false>(ctx, p, a, grid); }
            else pair_launch<false, true, false>(ctx, p, a, grid);
            ctx->launches += 1;
            CUDA_TRY(ctx, cudaGetLastError());
            if (ctx->profiling) {
                CUDA_TRY(ctx, cudaEventRecord(e1, ctx->s_compute));
                ctx->pending_pair_events.emplace_back(e0, e1);
            }
        }
        AMOFB_TRY(batcher_commit(ctx, b, *s, nf));
        done += nf;
    }
    return AMOFB_OK;
}

extern "C" int amofb_pair_push(amofb_ctx *ctx, int n_frames, const double *pos, const double *cell) {
    nvtx_range rng("amofb_pair_push");
    return pair_push_impl(ctx, n_frames, pos, false, cell);
}
extern "C" int amofb_pair_push_device(amofb_ctx *ctx, int n_frames, const double *pos_device, const double *cell) {
    nvtx_range rng("amofb_pair_push_device");
    return pair_push_impl(ctx, n_frames, pos_device, true, cell);
}

// finish and take share this: drain the batches and hand out what has accumulated; `reset` empties the accumulators and
// keeps the analysis open (amof.rdf.CoordinationNumber wants one histogram per frame, rdf.py:181-186)
static int pair_collect(amofb_ctx *ctx, uint64_t *hist, uint64_t *cn_counts, int64_t cn_frames, int64_t *n_frames_out,
                        double *volume_sum_out, bool reset) {
    if (!ctx) return AMOFB_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    PairState *p = ctx->pair;
    if (!p) return amofb_fail(ctx, AMOFB_ERR_STATE, "amofb_pair_finish / _take before amofb_pair_begin");
    int rc = AMOFB_OK;
    auto body = [&]() -> int {
        Batcher &b = p->bt;
        AMOFB_TRY(batcher_drain(ctx, b));
        const int S = p->n_species;
        if (p->d_flags) {
            // tile list overflow: some pairs were never evaluated, so NEITHER output may be handed out
            int flags = 0;
            CUDA_TRY(ctx, cudaMemcpyAsync(&flags, p->d_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->s_compute));
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->s_compute));
            if (flags) return amofb_fail(ctx, AMOFB_ERR_GEOMETRY, "tile list overflow (extremely inhomogeneous frame); rerun with AMOFB_PAIR_GENERIC=1");
        }
        if (cn_counts) {
            if (!p->has_cn) return amofb_fail(ctx, AMOFB_ERR_ARG, "cn_counts requested but no cutoffs were given at begin");
            if (cn_frames != b.frames_seen)
                return amofb_fail(ctx, AMOFB_ERR_ARG, "cn_frames = %lld but %lld frames were pushed", (long long)cn_frames, (long long)b.frames_seen);
            for (int64_t f = 0; f < b.frames_seen; ++f)
                for (int x = 0; x < S; ++x)
                    for (int y = 0; y < S; ++y) {
                        unsigned long long u = b.out_all[(size_t)f * p->nkeys + fold_key(x, y, S)];
                        cn_counts[((size_t)f * S + x) * S + y] = x == y ? 2 * u : u;
                    }
        }
        if (hist) {
            if (!p->has_rdf) return amofb_fail(ctx, AMOFB_ERR_ARG, "hist requested but nbins was 0 at begin");
            const size_t hist_n = (size_t)p->nkeys * p->nbins;
            if (p->smem_hist) {
                k_slab_reduce<<<ctx->num_sms * 2, 256, 0, ctx->s_compute>>>(p->d_slabs, std::max(p->grid, p->tile_grid), (int)hist_n, p->d_hist);
                ctx->launches += 1;
                CUDA_TRY(ctx, cudaGetLastError());
            }
            std::vector<unsigned long long> u(hist_n);
            CUDA_TRY(ctx, cudaMemcpyAsync(u.data(), p->d_hist, sizeof(unsigned long long) * hist_n, cudaMemcpyDeviceToHost, ctx->s_compute));
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->s_compute));
            for (int x = 0; x < S; ++x)
                for (int y = 0; y < S; ++y) {
                    const unsigned long long *src = u.data() + (size_t)fold_key(x, y, S) * p->nbins;
                    uint64_t *dst = hist + ((size_t)x * S + y) * p->nbins;
                    for (int k = 0; k < p->nbins; ++k) dst[k] = x == y ? 2 * src[k] : src[k];
                }
        }
        if (n_frames_out) *n_frames_out = b.frames_seen;
        if (volume_sum_out) *volume_sum_out = b.volume_sum;
        if (reset) {
            const size_t hist_n = (size_t)p->nkeys * p->nbins;
            if (hist_n) CUDA_TRY(ctx, cudaMemsetAsync(p->d_hist, 0, sizeof(unsigned long long) * hist_n, ctx->s_compute));
            if (p->d_slabs)
                CUDA_TRY(ctx, cudaMemsetAsync(p->d_slabs, 0, sizeof(unsigned long long) * hist_n * std::max(p->grid, p->tile_grid), ctx->s_compute));
            b.frames_seen = 0;
            b.volume_sum = 0.0;
            b.out_all.clear();
        }
        return AMOFB_OK;
    };
    rc = body();
    if (ctx->profiling) drain_pair_events(ctx);
    if (!reset || rc) pair_release(ctx);
    return rc;
}

extern "C" int amofb_pair_finish(amofb_ctx *ctx, uint64_t *hist, uint64_t *cn_counts, int64_t cn_frames,
                                 int64_t *n_frames_out, double *volume_sum_out) {
    nvtx_range rng("amofb_pair_finish");
    return pair_collect(ctx, hist, cn_counts, cn_frames, n_frames_out, volume_sum_out, false);
}
extern "C" int amofb_pair_take(amofb_ctx *ctx, uint64_t *hist, uint64_t *cn_counts, int64_t cn_frames,
                               int64_t *n_frames_out, double *volume_sum_out) {
    nvtx_range rng("amofb_pair_take");
    return pair_collect(ctx, hist, cn_counts, cn_frames, n_frames_out, volume_sum_out, true);
}

extern "C" int amofb_rdf_begin(amofb_ctx *ctx, int n_atoms, int n_species, const uint8_t *species, double rmax, int nbins) {
    if (ctx && nbins < 1) return amofb_fail(ctx, AMOFB_ERR_ARG, "amofb_rdf_begin needs nbins >= 1");
    return amofb_pair_begin(ctx, n_atoms, n_species, species, rmax, nbins, nullptr);
}
extern "C" int amofb_rdf_push(amofb_ctx *ctx, int n_frames, const double *pos, const double *cell) {
    return amofb_pair_push(ctx, n_frames, pos, cell);
}
extern "C" int amofb_rdf_finish(amofb_ctx *ctx, uint64_t *hist, int64_t *n_frames_out, double *volume_sum_out) {
    return amofb_pair_finish(ctx, hist, nullptr, 0, n_frames_out, volume_sum_out);
}
extern "C" int amofb_cn_begin(amofb_ctx *ctx, int n_atoms, int n_species, const uint8_t *species, const double *cn_cutoff) {
    if (ctx && !cn_cutoff) return amofb_fail(ctx, AMOFB_ERR_ARG, "amofb_cn_begin needs a cutoff matrix");
    return amofb_pair_begin(ctx, n_atoms, n_species, species, 0.0, 0, cn_cutoff);
}
extern "C" int amofb_cn_push(amofb_ctx *ctx, int n_frames, const double *pos, const double *cell) {
    return amofb_pair_push(ctx, n_frames, pos, cell);
}
extern "C" int amofb_cn_finish(amofb_ctx *ctx, uint64_t *cn_counts, int64_t cn_frames) {
    return amofb_pair_finish(ctx, nullptr, cn_counts, cn_frames, nullptr, nullptr);
}

#include "bad_host.inl"
#include "msd_host.inl"
#include "neigh_host.inl"
