// common.cuh -- context, error plumbing and the pinned host-side geometry of libamofb.
//
// Host-side fp64 arithmetic here is bin-deciding (inverse cell, image shifts, bin thresholds), so
// this translation unit is compiled with -Xcompiler -ffp-contract=off and the device code with
// -fmad=false: every expression below is evaluated operation by operation in IEEE-754 binary64,
// in the order written, and mirrors pins P1-P8 of oracle/amof_oracle.c.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/amofb.h"

// NVTX ranges around the host-side phases (header-only NVTX3: no library to link; a no-op unless a tool is attached)
#if __has_include(<nvtx3/nvToolsExt.h>)
#include <nvtx3/nvToolsExt.h>
struct nvtx_range {
    explicit nvtx_range(const char *name) { nvtxRangePushA(name); }
    ~nvtx_range() { nvtxRangePop(); }
};
#else
struct nvtx_range { explicit nvtx_range(const char *) {} };
#endif

#define AMOFB_VERSION_STRING "amofb 0.1 (sm_100a)"

struct PairState;
struct BadState;
struct MsdState;
struct NeighState;

// Buffers released by an analysis are kept by the context and handed to the next one: begin/finish pairs run once
// per trajectory pass, and cudaMalloc / cudaHostAlloc / cudaFree each cost more than a whole batch of kernels.
struct PoolBlock {
    void *p;
    size_t bytes;
    bool pinned;
    bool guarded = false;   // AMOFB_GUARD: p sits POOL_GUARD bytes inside the allocation, canaries on both sides
};
#define POOL_GUARD 4096

struct amofb_ctx {
    int device = 0;
    // last RDF threshold table (amof.rdf.CoordinationNumber opens one analysis per frame with the same rmax / bins)
    double edge_rmax = 0.0;
    int edge_nbins = 0, edge_rule = 0;
    int rdf_bin_rule = 0;         // AMOFB_OPT_RDF_BIN_RULE: 0 bin = (int)(d / (rmax/nbins)), 1 bin = (int)(d * (nbins/rmax))  (pin U1)
    std::vector<double> edge_cache;
    // last bond-angle threshold table (3 600 bisections through libm acos: a few ms, identical for every analysis with
    // the same dtheta / bin count)
    double tthr_dtheta = 0.0;
    int tthr_nbins = 0;
    std::vector<double> tthr_cache;
    std::vector<PoolBlock> pool_idle;
    std::unordered_map<void *, PoolBlock> pool_live;
    bool guard = false;            // AMOFB_GUARD=1: canary bytes around every pooled device block, checked when it is returned
    int64_t guard_violations = 0;
    cudaStream_t s_compute = nullptr;
    cudaStream_t s_copy = nullptr;
    std::string err;
    int64_t launches = 0;
    int num_sms = 0;
    int max_smem_optin = 0;
    bool profiling = false;
    double pair_ms = 0.0;
    int64_t pair_launches = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending_pair_events;
    cudaEvent_t timer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    PairState *pair = nullptr;
    BadState *bad = nullptr;
    MsdState *msd = nullptr;
    NeighState *neigh = nullptr;
};

static int amofb_fail(amofb_ctx *ctx, int code, const char *fmt, ...) __attribute__((format(printf, 3, 4)));
static int amofb_fail(amofb_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

#define CUDA_TRY(ctx, expr)                                                                              \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return amofb_fail((ctx), AMOFB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                              __FILE__, __LINE__);                                                       \
    } while (0)

#define AMOFB_TRY(expr)            \
    do {                           \
        int rc__ = (expr);         \
        if (rc__ != AMOFB_OK) return rc__; \
    } while (0)

// ------------------------------------------------------------------------------------------------
// Per-frame geometry, computed on the host and uploaded with every batch.
struct FrameGeom {
    double cell[9];   // rows = lattice vectors
    double inv[9];    // P1
    int nc[3];        // linked cells per axis
    int m[3];         // stencil half-width per axis (cells)
    int ncell;        // nc0*nc1*nc2
    int cs_off;       // offset of this frame's cell_start array (ncell+1 entries) in the batch-wide array
    int frame_id;     // global frame index (for per-frame outputs)
    int pad_;
};

// P1: inverse by cofactors over the determinant, fractional coordinate f = p . inv
static inline bool host_cell_inverse(const double *c, double *inv) {
    double m00 = c[4] * c[8] - c[5] * c[7];
    double m01 = c[3] * c[8] - c[5] * c[6];
    double m02 = c[3] * c[7] - c[4] * c[6];
    double det = (c[0] * m00 - c[1] * m01) + c[2] * m02;
    if (!(det != 0.0) || !isfinite(det)) return false;
    inv[0] = m00 / det;
    inv[1] = (c[2] * c[7] - c[1] * c[8]) / det;
    inv[2] = (c[1] * c[5] - c[2] * c[4]) / det;
    inv[3] = (c[5] * c[6] - c[3] * c[8]) / det;
    inv[4] = (c[0] * c[8] - c[2] * c[6]) / det;
    inv[5] = (c[2] * c[3] - c[0] * c[5]) / det;
    inv[6] = m02 / det;
    inv[7] = (c[1] * c[6] - c[0] * c[7]) / det;
    inv[8] = (c[0] * c[4] - c[1] * c[3]) / det;
    for (int k = 0; k < 9; ++k)
        if (!isfinite(inv[k])) return false;
    return true;
}

static inline double host_cell_volume(const double *c) {
    double m00 = c[4] * c[8] - c[5] * c[7];
    double m01 = c[3] * c[8] - c[5] * c[6];
    double m02 = c[3] * c[7] - c[4] * c[6];
    return fabs((c[0] * m00 - c[1] * m01) + c[2] * m02);
}

// perpendicular heights: h_k = 1/|column k of inv|
static inline void host_cell_heights(const double *inv, double *h) {
    for (int k = 0; k < 3; ++k) {
        double s = (inv[0 + k] * inv[0 + k] + inv[3 + k] * inv[3 + k]) + inv[6 + k] * inv[6 + k];
        h[k] = 1.0 / sqrt(s);
    }
}

// Fill the linked-cell grid of one frame for a search radius rcut.
//   cells of width >= rcut*(1+1e-9)/cell_div along every perpendicular direction, so a stencil of
//   half-width m_k = ceil(rcut*(1+1e-9)/w_k) cells provably covers every pair within rcut even though
//   atoms sit in their cell only up to fp rounding.
//   widen > 1 makes the cells that much wider than rcut/cell_div (sparse, species-filtered frames: fewer, fuller cells)
static inline bool host_fill_geom(FrameGeom &g, const double *cell, double rcut, int cell_div, int n_atoms, double widen = 1.0) {
    memcpy(g.cell, cell, sizeof(double) * 9);
    if (!host_cell_inverse(cell, g.inv)) return false;
    double h[3];
    host_cell_heights(g.inv, h);
    double rpad = rcut * (1.0 + 1e-9) + 1e-300;
    double total = 1.0;
    for (int k = 0; k < 3; ++k) {
        double n = floor(h[k] * cell_div / (rpad * (widen > 1.0 ? widen : 1.0)));
        if (!(n >= 1.0)) n = 1.0;
        if (n > 1024.0) n = 1024.0;
        g.nc[k] = (int)n;
        total *= n;
    }
    // keep the grid from exploding on sparse boxes: at most ~4 cells per atom (+64)
    double cap = 4.0 * (double)n_atoms + 64.0;
    if (total > cap) {
        double s = cbrt(cap / total);
        for (int k = 0; k < 3; ++k) {
            int n = (int)floor(g.nc[k] * s);
            g.nc[k] = n < 1 ? 1 : n;
        }
    }
    for (int k = 0; k < 3; ++k) {
        double w = h[k] / g.nc[k];
        int m = (int)ceil(rpad / w);
        if (m < 1) m = 1;
        if (m > 60) return false;   // box far smaller than the cutoff: refuse instead of looping forever
        g.m[k] = m;
    }
    g.ncell = g.nc[0] * g.nc[1] * g.nc[2];
    g.pad_ = 0;
    return true;
}

// Smallest non-negative double t with pred(t) true, pred monotone (false ... false true ... true) in t.
// Bisection on the bit pattern, which is order-preserving for non-negative doubles.
template <typename Pred>
static inline double host_threshold(double hi_guess, Pred pred) {
    uint64_t lo = 0, hi;
    double h = hi_guess;
    if (!(h > 0.0)) h = 1e-300;
    while (!pred(h)) h *= 2.0;
    memcpy(&hi, &h, 8);
    if (pred(0.0)) return 0.0;
    // invariant: pred(lo) false, pred(hi) true
    while (hi - lo > 1) {
        uint64_t mid = lo + (hi - lo) / 2;
        double t;
        memcpy(&t, &mid, 8);
        if (pred(t)) hi = mid; else lo = mid;
    }
    double t;
    memcpy(&t, &hi, 8);
    return t;
}

// ------------------------------------------------------------------------------------------------
// device helpers shared by the kernels

__device__ __forceinline__ int floordiv_i(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

// floor(t / n) and t mod n for stencil coordinates: |t| is within a few n of [0, n) (usually inside it), so two
// compare loops beat an integer division (~25 instructions on the SM)
__device__ __forceinline__ void wrap_cell(int t, int n, int &s, int &q) {
    s = 0; q = t;
    while (q < 0) { q += n; --s; }
    while (q >= n) { q -= n; ++s; }
}

__device__ __forceinline__ unsigned long long atomicAdd64(unsigned long long *p, unsigned long long v) {
    return atomicAdd(p, v);
}

// ---- TMA 1-D bulk copies (cp.async.bulk) completing on an mbarrier ----------------------------------------
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned mbar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst_smem, const void *src, unsigned bytes, unsigned mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
// the same with a suspend-time hint: the warp sleeps in the barrier unit (up to ~ns nanoseconds per attempt) instead of spinning
// through the issue slots of the warps that have work
__device__ __forceinline__ void mbar_wait_hint(unsigned mbar, unsigned parity, unsigned ns) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAITH_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra MBAR_DONEH_%=;\n"
        "bra MBAR_WAITH_%=;\n"
        "MBAR_DONEH_%=:\n"
        "}\n" :: "r"(mbar), "r"(parity), "r"(ns) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE_%=;\n"
        "bra MBAR_WAIT_%=;\n"
        "MBAR_DONE_%=:\n"
        "}\n" :: "r"(mbar), "r"(parity) : "memory");
}
