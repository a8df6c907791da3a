// pair_tiled.cuh -- the production RDF(+CN) pair kernel: linked-cell tiles staged in shared memory.
//
// The generic kernel of pair.cuh reads every candidate atom through L1/L2 once per home atom; at 10 A the candidate
// set of a 256-atom tile (~200 KB) does not fit in L1, so it runs at L2 speed.  Here a block owns a HOME TILE = a run
// of `zlen` consecutive cells of one column along the fastest cell axis, and first copies the tile's whole half
// stencil -- R rows (neighbour columns) x (zlen + 2*m2) virtual cells -- into shared memory with TMA 1-D bulk copies
// (one cp.async.bulk per contiguous run of a row, completing on an mbarrier), so that each row is one contiguous run
// ordered by virtual cell.  Every candidate is then read from shared memory.
//
//   k_pair_plan   one warp per column: greedy split of the column into tiles whose staged atoms fit `cap`;
//                 cells too dense even alone are split by rows, or marked "hard" and left to the generic kernel.
//   k_pair_tiled  persistent blocks over the tile list.  Work item = (home cell, row); a warp takes one item:
//                 lanes = (home atom i, sub-lane s) with G = 32 / n_home sub-lanes per home atom striding the
//                 candidate run, so control flow is warp-uniform and shared-memory reads are G-way contiguous.
//
// Arithmetic, thresholds, folding of species pairs and histogram privatisation are exactly those of pair.cuh.
// Compile-time switches kept for the measurements in DESIGN.md section 8: TILE_QUEUE (warp-aggregated hit queue),
// TILE_FLAT (one flat scan per item), TILE_PAIRED (two candidates per trip), TILE_TMA (bulk-copy staging).
#pragma once
#include "pair.cuh"

#ifndef TILE_THREADS
#define TILE_THREADS 512     // 2 blocks x 16 warps per SM at <= 64 registers: measured best on B200 (tools/sweep_pair.sh)
#endif
#ifndef TILE_MIN_BLOCKS
#define TILE_MIN_BLOCKS 2
#endif
#ifndef TILE_UNROLL
#define TILE_UNROLL 1
#endif
#ifndef TILE_QUEUE
#define TILE_QUEUE 0
#endif
#ifndef TILE_PAIRED
#define TILE_PAIRED 1
#endif
#ifndef TILE_QUAD
#define TILE_QUAD 0
#endif
#define TILE_PRAGMA_(x) _Pragma(#x)
#define TILE_PRAGMA_UNROLL(n) TILE_PRAGMA_(unroll n)
#define TILE_MAX_ENTRIES 1024     // rows x virtual cells per tile
#define TILE_MAX_ZLEN 62
#define TILE_MAX_ROWS 256
#define TILE_OFF_WORDS (TILE_MAX_ENTRIES + TILE_MAX_ZLEN + 2)   // offsets: entries + home cells + 1

struct PairTile {
    int frame, c0, c1, z0, zlen, rb, re, pad;
};

struct PlanArgs {
    const FrameGeom *geom;
    const uint32_t *cell_start;
    PairTile *tiles;
    int *n_tiles;            // [0] tiles, [1] hard cells
    int *flags;              // sticky: bit 0 = tile list overflow
    uint8_t *hard;           // batch-wide per-cell mask
    int n_frames, cap, max_tiles;
    int zlen_max;            // <= TILE_MAX_ZLEN; smaller when the fp32 path bounds the tile extent
    const unsigned char *sel;   // optional per-frame selector: only frames with sel[f] == want are planned
    int want;
    int uniform_cols;           // > 0: every frame has this many columns (nc0 * nc1), so frame = warp / uniform_cols
};

__device__ __forceinline__ int tile_rows(const FrameGeom &g) { return (g.m[1] + 1) + g.m[0] * (2 * g.m[1] + 1); }

__device__ __forceinline__ void tile_row_offset(const FrameGeom &g, int r, int &d0, int &d1) {
    if (r <= g.m[1]) { d0 = 0; d1 = r; }
    else {
        int rr = r - (g.m[1] + 1);
        const int w = 2 * g.m[1] + 1;
        d0 = 1;
        while (rr >= w) { rr -= w; ++d0; }      // d0 <= m0: a couple of steps, cheaper than a division
        d1 = rr - g.m[1];
    }
}

// atoms in the virtual cells [va, vb] of the column starting at cs[colbase]
__device__ __forceinline__ int column_count(const uint32_t *cs, int colbase, int nc2, int va, int vb) {
    const int total = (int)(cs[colbase + nc2] - cs[colbase]);
    const int fa = floordiv_i(va, nc2), fb = floordiv_i(vb + 1, nc2);
    const int qa = va - fa * nc2, qb = vb + 1 - fb * nc2;
    return (fb - fa) * total + (int)cs[colbase + qb] - (int)cs[colbase + qa];
}

__global__ void __launch_bounds__(128) k_pair_plan(PlanArgs a) {
    // warp -> (frame, column); the lanes share the stencil rows when a tile's population is summed
    const int lane = threadIdx.x & 31;
    long long t = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    int f = 0;
    if (a.uniform_cols > 0) {
        // the usual case (one cell for the whole batch): no walk over the frames
        f = (int)(t / a.uniform_cols);
        if (f >= a.n_frames) return;
        t -= (long long)f * a.uniform_cols;
        if (a.sel && a.sel[f] != (unsigned char)a.want) return;
    } else {
        // frames may have different grids: walk the frames
        for (; f < a.n_frames; ++f) {
            long long cols = (a.sel && a.sel[f] != (unsigned char)a.want) ? 0 : (long long)a.geom[f].nc[0] * a.geom[f].nc[1];
            if (t < cols) break;
            t -= cols;
        }
        if (f >= a.n_frames) return;
    }
    const FrameGeom &g = a.geom[f];
    const uint32_t *cs = a.cell_start + g.cs_off;
    const int nc1 = g.nc[1], nc2 = g.nc[2], m2 = g.m[2];
    const int c0 = (int)(t / nc1), c1 = (int)(t - (long long)c0 * nc1);
    const int R = tile_rows(g);
    const int vmax = TILE_MAX_ENTRIES / R;          // virtual cells per row that the offset table can hold
    const int homebase = (c0 * nc1 + c1) * nc2;
    int z = 0;
    while (z < nc2) {
        int zlen = min(min(nc2 - z, vmax - 2 * m2), a.zlen_max);
        if (zlen < 1) zlen = 1;                     // host guarantees vmax >= 2*m2 + 1
        int total = 0;
        for (;;) {
            total = 0;
            for (int r = lane; r < R; r += 32) {
                int d0, d1;
                tile_row_offset(g, r, d0, d1);
                const int t0 = c0 + d0, t1 = c1 + d1;
                const int q0 = t0 - floordiv_i(t0, g.nc[0]) * g.nc[0], q1 = t1 - floordiv_i(t1, nc1) * nc1;
                total += column_count(cs, (q0 * nc1 + q1) * nc2, nc2, z - m2, z + zlen - 1 + m2);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
            total += (int)(cs[homebase + z + zlen] - cs[homebase + z]);   // the home cells are staged once more
            if (total <= a.cap || zlen == 1) break;
            --zlen;
        }
        const int home = (int)(cs[homebase + z + zlen] - cs[homebase + z]);
        if (home > 0 && lane == 0) {                // every lane holds the same total; lane 0 records the decision
            if (total <= a.cap) {
                int k = atomicAdd(&a.n_tiles[0], 1);
                if (k < a.max_tiles) a.tiles[k] = PairTile{f, c0, c1, z, zlen, 0, R, 0};
                else atomicOr(a.flags, 1);
            } else {
                // a single home cell whose stencil does not fit: split by rows; a row that does not fit alone -> hard cell
                bool hard = false;
                for (int r = 0; r < R && !hard; ++r) {
                    int d0, d1;
                    tile_row_offset(g, r, d0, d1);
                    const int t0 = c0 + d0, t1 = c1 + d1;
                    const int q0 = t0 - floordiv_i(t0, g.nc[0]) * g.nc[0], q1 = t1 - floordiv_i(t1, nc1) * nc1;
                    if (column_count(cs, (q0 * nc1 + q1) * nc2, nc2, z - m2, z + m2) + home > a.cap) hard = true;
                }
                if (hard) {
                    a.hard[g.cs_off + homebase + z] = 1;
                    atomicAdd(&a.n_tiles[1], 1);
                } else {
                    int rb = 0;
                    while (rb < R) {
                        int acc = home, re = rb;
                        while (re < R) {
                            int d0, d1;
                            tile_row_offset(g, re, d0, d1);
                            const int t0 = c0 + d0, t1 = c1 + d1;
                            const int q0 = t0 - floordiv_i(t0, g.nc[0]) * g.nc[0], q1 = t1 - floordiv_i(t1, nc1) * nc1;
                            const int c = column_count(cs, (q0 * nc1 + q1) * nc2, nc2, z - m2, z + m2);
                            if (acc + c > a.cap) break;
                            acc += c;
                            ++re;
                        }
                        int k = atomicAdd(&a.n_tiles[0], 1);
                        if (k < a.max_tiles) a.tiles[k] = PairTile{f, c0, c1, z, 1, rb, re, 0};
                        else atomicOr(a.flags, 1);
                        rb = re;
                    }
                }
            }
        }
        z += zlen;
    }
}

// fp32 fast path (see scan_run_f32): thresholds in d2 with their certainty bands, all computed on the host in fp64
struct F32Params {
    const float2 *cn_band;   // [nkeys]: d2 < x -> certainly under the cutoff, d2 >= y -> certainly not; (0, 0) = not listed
    float r2hi;              // d2 >= r2hi: certainly outside every range of interest
    float r2max_lo, r2max_hi;// certainly inside / certainly outside the RDF range
    float cn_hi;             // d2 >= cn_hi: certainly above every cutoff
    float margin;            // |fp32 estimate of d/dr - exact quotient| < margin (bins)
    int enabled;
};

struct TiledArgs {
    PairArgs p;
    F32Params f;
    const PairTile *tiles;
    const int *n_tiles;
    int cap;                 // staged atoms per tile
    int max_tiles;
};

// G = 32 / ng sub-lanes per home atom and the magic multiplier of x / G (exact for x < 2048), by table: two integer
// divisions per work item cost ~50 instructions on the SM
__constant__ unsigned char c_sub_lanes[33] = {32, 32, 16, 10, 8, 6, 5, 4, 4, 3, 3, 2, 2, 2, 2, 2, 2,
                                               1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
__constant__ unsigned short c_div_magic[33] = {0, 65535, 32768, 21846, 16384, 13108, 10923, 9363, 8192, 7282, 6554, 5958, 5462,
                                                5042, 4682, 4370, 4096, 3856, 3641, 3450, 3277, 3121, 2979, 2850, 2731, 2622,
                                                2521, 2428, 2341, 2260, 2185, 2115, 2048};

// ---- TMA 1-D bulk copies (cp.async.bulk) completing on an mbarrier ----------------------------------------
#ifndef TILE_TMA
#define TILE_TMA 1
#endif
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

// shared-memory loads through an explicit 32-bit shared address kept in a register: the compiler otherwise
// re-derives the shared window base (S2R SR_CgaCtaId + LEA) inside the candidate loop
__device__ __forceinline__ void lds_xyz(unsigned addr, double &x, double &y, double &z) {
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
    asm("ld.shared.f64 %0, [%1+16];" : "=d"(z) : "r"(addr));
}
__device__ __forceinline__ int lds_species(unsigned addr) {
    unsigned v;
    asm("ld.shared.u8 %0, [%1+24];" : "=r"(v) : "r"(addr));
    return (int)v;
}
__device__ __forceinline__ int lds_u16(unsigned addr) {
    unsigned short v;
    asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return (int)v;
}
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void reds_inc(unsigned addr) {
    asm volatile("red.shared.add.u32 [%0], 1;" :: "r"(addr) : "memory");
}
// 32-bit shared addresses of the kernel's tables, derived once per kernel from an opaque base (the asm keeps the
// compiler from re-materialising S2UR SR_CgaCtaId + ULEA chains next to every use)
struct SmemAddr {
    unsigned atoms, edge, cnthr, hist, cn, key;
};
__device__ __forceinline__ unsigned opaque_u32(unsigned v) {
    unsigned r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

// rdf_bin (pair.cuh) on a shared-memory threshold table given by its 32-bit shared address; margin > 0 path only
__device__ __forceinline__ int rdf_bin_s(double d2, unsigned edge_addr, float inv_dr_f, float margin) {
    const int b = (int)fmaf(sqrt_approx((float)d2), inv_dr_f, -margin);
    return b + (d2 >= lds_f64(edge_addr + (unsigned)(b + 1) * 8u) ? 1 : 0);
}

// ---- hit queue -------------------------------------------------------------------------------------------
// The candidate loop only decides "d2 < r2search" and appends the hits (d2 + who) to a per-warp ring of 64
// 16-byte entries in shared memory, at warp-aggregated positions (one ballot per iteration).  Whenever 32 hits are
// waiting, all 32 lanes bin one each: the expensive part (bin search, species-pair lookup, shared atomics) always
// runs with full warps instead of the ~1/4-full warps a branch inside the loop would leave.
struct HitQueue {
    ulonglong2 *q;       // [64]
    unsigned head, tail; // warp-uniform
};

template <bool HAS_CN>
__device__ __forceinline__ void hits_flush(const PairArgs &a, const SAtom *__restrict__ s_atoms, const double *__restrict__ s_edge2,
                                           const double *__restrict__ s_cnthr, uint32_t *__restrict__ s_hist, uint32_t *__restrict__ s_cn,
                                           const uint16_t *__restrict__ s_key, HitQueue &hq, int lane, int count) {
    __syncwarp();
    if (lane < count) {
        const ulonglong2 e = hq.q[(hq.head + lane) & 63u];
        const double dd = __longlong_as_double((long long)e.x);
        const int j = (int)(e.y & 0xffffu), hi = (int)(e.y >> 16);
        const int sj = reinterpret_cast<const unsigned char *>(s_atoms + j)[24];
        const int si = reinterpret_cast<const unsigned char *>(s_atoms + hi)[24];
        const int key = s_key[si * a.n_species + sj];
        if (dd < a.r2max) {
            const int b = rdf_bin(dd, s_edge2, a.inv_dr_f, a.bin_margin, a.nbins);
            atomicAdd(&s_hist[key * a.nbins + b], 1u);
        }
        if (HAS_CN && dd < a.cn_r2max && dd < s_cnthr[key]) atomicAdd(&s_cn[key], 1u);
    }
    hq.head += count;
    __syncwarp();
}

// One run of staged candidates [jb, je) against the home atoms of this warp: lane (il, sub) takes jb+sub, +G, ...
//   SHIFT: the run sits in a periodic image (T != 0); without it (pj - pi) + 0 == pj - pi bit for bit, so the adds go.
//   AFTER: the run starts with the home cell itself: only partners staged after me (index > ism) count.
// CN_WIDE: some cutoff exceeds rmax, so a candidate inside r2search can still be outside the RDF range
template <bool HAS_CN, bool CN_WIDE, bool SHIFT, bool AFTER>
__device__ __forceinline__ void scan_run(const PairArgs &a, const SAtom *__restrict__ s_atoms, const double *__restrict__ s_edge2,
                                         const double *__restrict__ s_cnthr, uint32_t *__restrict__ s_hist, uint32_t *__restrict__ s_cn,
                                         const uint16_t *__restrict__ s_key, HitQueue &hq, const SAtom &me, double Tx, double Ty, double Tz,
                                         int jb, int je, int G, int n_iter, int sub, bool active, int ism, int hidx, int lane,
                                         unsigned lt_mask, const SmemAddr &sa) {
    const double r2search = a.r2search;
#if TILE_QUEUE
    const unsigned long long who_hi = (unsigned long long)hidx << 16;
    int j = jb + sub;
    TILE_PRAGMA_UNROLL(TILE_UNROLL)
    for (int k = 0; k < n_iter; ++k, j += G) {
        const bool valid = active && j < je;
        const double2 *q = reinterpret_cast<const double2 *>(s_atoms + (valid ? j : jb));
        const double2 o0 = q[0];
        const double oz = reinterpret_cast<const double *>(q)[2];
        double dx = o0.x - me.x, dy = o0.y - me.y, dz = oz - me.z;
        if (SHIFT) { dx += Tx; dy += Ty; dz += Tz; }
        const double dd = (dx * dx + dy * dy) + dz * dz;
        const bool hit = valid && dd < r2search && !(AFTER && j <= ism);
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            const unsigned pos = (hq.tail + __popc(m & lt_mask)) & 63u;
            hq.q[pos] = make_ulonglong2((unsigned long long)__double_as_longlong(dd), who_hi | (unsigned long long)j);
        }
        hq.tail += __popc(m);
        if (hq.tail - hq.head >= 32u) hits_flush<HAS_CN>(a, s_atoms, s_edge2, s_cnthr, s_hist, s_cn, s_key, hq, lane, 32);
    }
#else
    if (!active) return;
    const double r2max = a.r2max, cn_r2max = a.cn_r2max;
    const float inv_dr_f = a.inv_dr_f, margin = a.bin_margin;
    const int nbins = a.nbins;
    const unsigned abase = sa.atoms, edge_addr = sa.edge, hist_addr = sa.hist, cnthr_addr = sa.cnthr, cn_addr = sa.cn;
    const unsigned krow_addr = sa.key + 2u * (unsigned)((int)(me.s & 0xff) * a.n_species);
    const unsigned astep = (unsigned)G * 32u;
    const unsigned aend = abase + (unsigned)je * 32u;
    // AFTER: partners staged at or before me do not count (ism < 0: I am staged before this whole run -> nothing to skip)
    const unsigned askip = ism >= 0 ? abase + (unsigned)ism * 32u : 0u;
    // the hit work (exact bin + shared-memory increment), shared by the paired and the tail iteration
    auto hit = [&](unsigned addr, double dd) {
        const int key = lds_u16(krow_addr + 2u * (unsigned)lds_species(addr));
        if (!CN_WIDE || dd < r2max) {             // r2search == r2max unless a cutoff reaches beyond rmax
            const int b = rdf_bin_s(dd, edge_addr, inv_dr_f, margin);          // the host only selects this kernel when margin > 0
            reds_inc(hist_addr + 4u * (unsigned)(key * nbins + b));
        }
        if (HAS_CN && dd < cn_r2max && dd < lds_f64(cnthr_addr + 8u * (unsigned)key)) reds_inc(cn_addr + 4u * (unsigned)key);
    };
    unsigned addr = abase + (unsigned)(jb + sub) * 32u;
#if TILE_QUAD
    // four candidates per trip (measured against two: see DESIGN.md section 8)
    for (; addr + 3u * astep < aend; addr += 4u * astep) {
        const unsigned addr1 = addr + astep, addr2 = addr1 + astep, addr3 = addr2 + astep;
        double ox0, oy0, oz0, ox1, oy1, oz1, ox2, oy2, oz2, ox3, oy3, oz3;
        lds_xyz(addr, ox0, oy0, oz0);
        lds_xyz(addr1, ox1, oy1, oz1);
        lds_xyz(addr2, ox2, oy2, oz2);
        lds_xyz(addr3, ox3, oy3, oz3);
        double dx0 = ox0 - me.x, dy0 = oy0 - me.y, dz0 = oz0 - me.z;
        double dx1 = ox1 - me.x, dy1 = oy1 - me.y, dz1 = oz1 - me.z;
        double dx2 = ox2 - me.x, dy2 = oy2 - me.y, dz2 = oz2 - me.z;
        double dx3 = ox3 - me.x, dy3 = oy3 - me.y, dz3 = oz3 - me.z;
        if (SHIFT) {
            dx0 += Tx; dy0 += Ty; dz0 += Tz; dx1 += Tx; dy1 += Ty; dz1 += Tz;
            dx2 += Tx; dy2 += Ty; dz2 += Tz; dx3 += Tx; dy3 += Ty; dz3 += Tz;
        }
        const double dd0 = (dx0 * dx0 + dy0 * dy0) + dz0 * dz0;
        const double dd1 = (dx1 * dx1 + dy1 * dy1) + dz1 * dz1;
        const double dd2 = (dx2 * dx2 + dy2 * dy2) + dz2 * dz2;
        const double dd3 = (dx3 * dx3 + dy3 * dy3) + dz3 * dz3;
        if (dd0 < r2search && !(AFTER && addr <= askip)) hit(addr, dd0);
        if (dd1 < r2search && !(AFTER && addr1 <= askip)) hit(addr1, dd1);
        if (dd2 < r2search && !(AFTER && addr2 <= askip)) hit(addr2, dd2);
        if (dd3 < r2search && !(AFTER && addr3 <= askip)) hit(addr3, dd3);
    }
#endif
#if TILE_PAIRED
    // two candidates per trip: both distance chains are in flight together (the fp64 chain is latency-bound at
    // 8 warps per scheduler), and the loop overhead is paid once per pair
    for (; addr + astep < aend; addr += 2u * astep) {
        const unsigned addr1 = addr + astep;
        double ox0, oy0, oz0, ox1, oy1, oz1;
        lds_xyz(addr, ox0, oy0, oz0);
        lds_xyz(addr1, ox1, oy1, oz1);
        double dx0 = ox0 - me.x, dy0 = oy0 - me.y, dz0 = oz0 - me.z;
        double dx1 = ox1 - me.x, dy1 = oy1 - me.y, dz1 = oz1 - me.z;
        if (SHIFT) { dx0 += Tx; dy0 += Ty; dz0 += Tz; dx1 += Tx; dy1 += Ty; dz1 += Tz; }
        const double dd0 = (dx0 * dx0 + dy0 * dy0) + dz0 * dz0;
        const double dd1 = (dx1 * dx1 + dy1 * dy1) + dz1 * dz1;
        if (dd0 < r2search && !(AFTER && addr <= askip)) hit(addr, dd0);
        if (dd1 < r2search && !(AFTER && addr1 <= askip)) hit(addr1, dd1);
    }
#endif
    TILE_PRAGMA_UNROLL(TILE_UNROLL)
    for (; addr < aend; addr += astep) {
        double ox, oy, oz;
        lds_xyz(addr, ox, oy, oz);
        double dx = ox - me.x, dy = oy - me.y, dz = oz - me.z;
        if (SHIFT) { dx += Tx; dy += Ty; dz += Tz; }
        const double dd = (dx * dx + dy * dy) + dz * dz;
        if (dd < r2search && !(AFTER && addr <= askip)) hit(addr, dd);
    }
#endif
}

// ---- flat scan -------------------------------------------------------------------------------------------
// A work item (home cell x a group of rows) is a handful of candidate runs.  Setting each run up separately costs more
// instructions than scanning it (a run is ~10 iterations), so the runs of an item are described once in a small
// per-warp table {image shift, first flat index, staged offset} and the lanes walk ONE flat index over all of them,
// stepping to the next table entry when they cross a run boundary.  The image shift is always added (0.0 for the home
// image: (pj - pi) + 0.0 == pj - pi bit for bit), and the "partners after me" rule of the home cell is a per-entry
// index bound instead of a separate loop variant.
#ifndef TILE_FLAT
#define TILE_FLAT 0      // measured: 4.27 ms vs 4.07 ms per 214 C2 frames for the per-run variant
#endif
#define FLAT_MAXE 8
struct __align__(16) FlatRun {
    double Tx, Ty, Tz;
    int kbeg;        // flat index of the run's first candidate (entry n_runs holds the total)
    int jofs;        // staged index = flat index + jofs; bit 30 of kbeg set: the run starts with the home cell itself
};

template <bool HAS_CN, bool CN_WIDE>
__device__ __forceinline__ void scan_flat(const PairArgs &a, const SmemAddr &sa, const FlatRun *__restrict__ runs, int total,
                                          const SAtom &me, int sub, int G, int ism) {
    const double r2search = a.r2search, r2max = a.r2max, cn_r2max = a.cn_r2max;
    const float inv_dr_f = a.inv_dr_f, margin = a.bin_margin;
    const int nbins = a.nbins;
    const unsigned krow_addr = sa.key + 2u * (unsigned)((int)(me.s & 0xff) * a.n_species);
    const unsigned abase = sa.atoms;
    int e = 0;
    FlatRun cur = runs[0];
    int kend = runs[1].kbeg & 0x3fffffff;
    int jskip = (cur.kbeg >> 30) ? ism : -1;
    for (int k = sub; k < total; k += G) {
        while (k >= kend) {                       // crossed into the next run (rare: runs are ~40 candidates long)
            ++e;
            cur = runs[e];
            kend = runs[e + 1].kbeg & 0x3fffffff;
            jskip = (cur.kbeg >> 30) ? ism : -1;
        }
        const int j = k + cur.jofs;
        const unsigned addr = abase + (unsigned)j * 32u;
        double ox, oy, oz;
        lds_xyz(addr, ox, oy, oz);
        const double dx = (ox - me.x) + cur.Tx;
        const double dy = (oy - me.y) + cur.Ty;
        const double dz = (oz - me.z) + cur.Tz;
        const double dd = (dx * dx + dy * dy) + dz * dz;
        if (dd < r2search && j > jskip) {
            const int key = lds_u16(krow_addr + 2u * (unsigned)lds_species(addr));
            if (!CN_WIDE || dd < r2max) {
                const int b = rdf_bin_s(dd, sa.edge, inv_dr_f, margin);
                reds_inc(sa.hist + 4u * (unsigned)(key * nbins + b));
            }
            if (HAS_CN && dd < cn_r2max && dd < lds_f64(sa.cnthr + 8u * (unsigned)key)) reds_inc(sa.cn + 4u * (unsigned)key);
        }
    }
}

// ---- fp32 fast path -----------------------------------------------------------------------------------------
// After staging, every staged atom also gets a 16-byte fp32 record {x, y, z, species} holding (p + T) - O, with T the
// image shift of its stencil entry and O the tile origin (first home atom), formed in fp64 and rounded once.  Local
// coordinates stay below 128 A (the planner bounds the tile), so a coordinate is off by <= 2^-24 * 128 and a squared
// distance by a bound E(d) the host evaluates; the host turns E into certainty bands around every threshold.  The
// candidate loop then runs entirely in fp32 on half the shared-memory bytes: a pair whose fp32 d2 is outside all bands
// and whose estimated d/dr has a fractional part farther than `margin` from 0 and 1 is binned from fp32 -- provably
// the bin the fp64 arithmetic of pin P3/P4 gives -- and the ~1-2 % of pairs inside a band are re-evaluated in fp64
// exactly as scan_run does.  Counts therefore stay bit-exact while the FP64 pipe and the threshold table leave the
// common path.
#ifndef TILE_F32
#define TILE_F32 0      // measured: 3.80 ms vs 2.60 ms per 214 C2 frames for the fp64 path (same instruction count per
#endif                  // candidate, one more pass and barrier per tile, smaller tiles): correct, kept for reference, off

__device__ __forceinline__ float4 lds_f4(unsigned addr) {
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds_f2(unsigned addr) {
    float2 v;
    asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}

template <bool HAS_CN, bool CN_WIDE, bool AFTER>
__device__ __forceinline__ void scan_run_f32(const PairArgs &a, const F32Params &f, const FrameGeom &geom, const SAtom *__restrict__ s_atoms,
                                             unsigned f32_addr, unsigned band_addr, unsigned edge_addr, unsigned hist_addr, unsigned cn_addr,
                                             unsigned cnthr_addr, unsigned krow_addr, float mx, float my, float mz, int hidx,
                                             int s0, int s1, int s2, int jb, int je, int G, int sub, int ism) {
    const float r2hi = f.r2hi, r2max_lo = f.r2max_lo, r2max_hi = f.r2max_hi, cn_hi = f.cn_hi, margin = f.margin;
    const float inv_dr_f = a.inv_dr_f;
    const int nbins = a.nbins;
    // the rare exact re-evaluation: the fp64 arithmetic of scan_run on the fp64 records
    auto exact = [&](int j) {
        const SAtom me = s_atoms[hidx], o = s_atoms[j];
        double dx = o.x - me.x, dy = o.y - me.y, dz = o.z - me.z;
        if ((s0 | s1 | s2) != 0) {
            const double fs0 = (double)s0, fs1 = (double)s1, fs2 = (double)s2;
            dx += (fs0 * geom.cell[0] + fs1 * geom.cell[3]) + fs2 * geom.cell[6];
            dy += (fs0 * geom.cell[1] + fs1 * geom.cell[4]) + fs2 * geom.cell[7];
            dz += (fs0 * geom.cell[2] + fs1 * geom.cell[5]) + fs2 * geom.cell[8];
        }
        const double dd = (dx * dx + dy * dy) + dz * dz;
        if (dd < a.r2search) {
            const int key = lds_u16(krow_addr + 2u * (unsigned)(o.s & 0xff));
            if (!CN_WIDE || dd < a.r2max) {
                const int b = rdf_bin_s(dd, edge_addr, a.inv_dr_f, a.bin_margin);
                reds_inc(hist_addr + 4u * (unsigned)(key * nbins + b));
            }
            if (HAS_CN && dd < a.cn_r2max && dd < lds_f64(cnthr_addr + 8u * (unsigned)key)) reds_inc(cn_addr + 4u * (unsigned)key);
        }
    };
    for (int j = jb + sub; j < je; j += G) {
        const float4 c = lds_f4(f32_addr + (unsigned)j * 16u);
        const float dx = c.x - mx, dy = c.y - my, dz = c.z - mz;
        const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (d2 < r2hi && !(AFTER && j <= ism)) {
            const bool in_rdf = d2 < r2max_lo;
            bool amb = !in_rdf && (!CN_WIDE || d2 < r2max_hi);
            const float t = sqrt_approx(d2) * inv_dr_f;
            const int b = (int)t;
            amb = amb || (in_rdf && fabsf((t - (float)b) - 0.5f) > 0.5f - margin);
            const int key = lds_u16(krow_addr + 2u * (unsigned)(__float_as_int(c.w) & 0xff));
            bool cn_yes = false;
            if (HAS_CN && d2 < cn_hi) {
                const float2 band = lds_f2(band_addr + 8u * (unsigned)key);
                cn_yes = d2 < band.x;
                amb = amb || (!cn_yes && d2 < band.y);
            }
            if (amb) exact(j);
            else {
                if (in_rdf) reds_inc(hist_addr + 4u * (unsigned)(key * nbins + b));
                if (HAS_CN && cn_yes) reds_inc(cn_addr + 4u * (unsigned)key);
            }
        }
    }
}

template <bool HAS_CN, bool CN_WIDE, bool F32>
__global__ void __launch_bounds__(TILE_THREADS, TILE_MIN_BLOCKS) k_pair_tiled(TiledArgs ta) {
    const PairArgs &a = ta.p;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: atoms[cap] (32 B) | edge2[nbins+1] | cn_thr2[nkeys] | hist[nkeys*nbins] u32 | cn_cnt[nkeys] u32 | off[] int | hit queues | keyidx[S*S] u16
    // carved by byte offsets from the shared base, so every pointer keeps the shared address space
    size_t off = 0;
    SAtom *s_atoms = reinterpret_cast<SAtom *>(smem_raw + off);            off += sizeof(SAtom) * (size_t)ta.cap;
    float4 *s_f32 = reinterpret_cast<float4 *>(smem_raw + off);            off += sizeof(float4) * (size_t)(F32 ? ta.cap : 0);
    float2 *s_band = reinterpret_cast<float2 *>(smem_raw + off);           off += sizeof(float2) * (size_t)(F32 && HAS_CN ? a.nkeys : 0);
    double *s_edge2 = reinterpret_cast<double *>(smem_raw + off);          off += sizeof(double) * (size_t)(a.nbins + 1);
    double *s_cnthr = reinterpret_cast<double *>(smem_raw + off);          off += sizeof(double) * (size_t)(HAS_CN ? a.nkeys : 0);
    off = (off + 15) & ~(size_t)15;
    ulonglong2 *s_queue = reinterpret_cast<ulonglong2 *>(smem_raw + off);  off += sizeof(ulonglong2) * (TILE_QUEUE ? 64 * (TILE_THREADS / 32) : 0);
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(smem_raw + off);       off += sizeof(uint32_t) * (size_t)a.nkeys * a.nbins;
    uint32_t *s_cn = reinterpret_cast<uint32_t *>(smem_raw + off);         off += sizeof(uint32_t) * (size_t)(HAS_CN ? a.nkeys : 0);
    int *s_off = reinterpret_cast<int *>(smem_raw + off);                  off += sizeof(int) * TILE_OFF_WORDS;
    off = (off + 15) & ~(size_t)15;
    FlatRun *s_runs = reinterpret_cast<FlatRun *>(smem_raw + off);         off += sizeof(FlatRun) * (FLAT_MAXE + 1) * (TILE_THREADS / 32);
    uint16_t *s_key = reinterpret_cast<uint16_t *>(smem_raw + off);
    (void)s_runs;                                  // only the flat-scan variant uses it
    SmemAddr sa;
    {
        const unsigned sb = opaque_u32((unsigned)__cvta_generic_to_shared(smem_raw));
        sa.atoms = sb + (unsigned)(reinterpret_cast<unsigned char *>(s_atoms) - smem_raw);
        sa.edge = sb + (unsigned)(reinterpret_cast<unsigned char *>(s_edge2) - smem_raw);
        sa.cnthr = sb + (unsigned)(reinterpret_cast<unsigned char *>(s_cnthr) - smem_raw);
        sa.hist = sb + (unsigned)(reinterpret_cast<unsigned char *>(s_hist) - smem_raw);
        sa.cn = sb + (unsigned)(reinterpret_cast<unsigned char *>(s_cn) - smem_raw);
        sa.key = sb + (unsigned)(reinterpret_cast<unsigned char *>(s_key) - smem_raw);
    }
    __shared__ FrameGeom s_geom;
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ int s_rowimg[TILE_MAX_ROWS];        // per staged row: image (s0 | s1 << 16) of its column

    const int S = a.n_species;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = TILE_THREADS / 32;
    const unsigned lt_mask = (1u << lane) - 1u;
    HitQueue hq;
    hq.q = s_queue + 64 * warp;
    hq.head = hq.tail = 0u;
    for (int k = threadIdx.x; k <= a.nbins; k += blockDim.x) s_edge2[k] = a.edge2[k];
    for (int k = threadIdx.x; k < a.nkeys * a.nbins; k += blockDim.x) s_hist[k] = 0u;
    if (HAS_CN)
        for (int k = threadIdx.x; k < a.nkeys; k += blockDim.x) { s_cnthr[k] = a.cn_thr2[k]; s_cn[k] = 0u; }
    for (int k = threadIdx.x; k < S * S; k += blockDim.x) s_key[k] = a.keyidx[k];
    if (F32 && HAS_CN)
        for (int k = threadIdx.x; k < a.nkeys; k += blockDim.x) s_band[k] = ta.f.cn_band[k];

    const unsigned f32_a = (unsigned)__cvta_generic_to_shared(s_f32), band_a = (unsigned)__cvta_generic_to_shared(s_band);
    const unsigned edge_a = (unsigned)__cvta_generic_to_shared(s_edge2), hist_a = (unsigned)__cvta_generic_to_shared(s_hist);
    const unsigned cn_a = (unsigned)__cvta_generic_to_shared(s_cn), cnthr_a = (unsigned)__cvta_generic_to_shared(s_cnthr);
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(&s_mbar);
    if (TILE_TMA && threadIdx.x == 0) mbar_init(mbar, 1);
    unsigned tma_phase = 0;
    const int n_tiles = min(*ta.n_tiles, ta.max_tiles);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();     // previous tile fully consumed (atoms, offsets, geometry, cn counters)
        // every thread reads the 32-byte tile record itself (one broadcast line from L2): no shared copy, no barrier
        const int4 t_lo = __ldg(reinterpret_cast<const int4 *>(&ta.tiles[tile]));
        const int4 t_hi = __ldg(reinterpret_cast<const int4 *>(&ta.tiles[tile]) + 1);
        PairTile s_tile;
        s_tile.frame = t_lo.x; s_tile.c0 = t_lo.y; s_tile.c1 = t_lo.z; s_tile.z0 = t_lo.w;
        s_tile.zlen = t_hi.x; s_tile.rb = t_hi.y; s_tile.re = t_hi.z; s_tile.pad = 0;
        const int f = s_tile.frame;
        if (threadIdx.x < (int)(sizeof(FrameGeom) / sizeof(int)))
            reinterpret_cast<int *>(&s_geom)[threadIdx.x] = reinterpret_cast<const int *>(&a.geom[f])[threadIdx.x];
        __syncthreads();
        const SAtom *fr = a.sorted + (long long)f * a.n_atoms;
        const uint32_t *cs = a.cell_start + s_geom.cs_off;
        const int nc0 = s_geom.nc[0], nc1 = s_geom.nc[1], nc2 = s_geom.nc[2];
        const int m2 = s_geom.m[2];
        const int c0 = s_tile.c0, c1 = s_tile.c1, z0 = s_tile.z0, zlen = s_tile.zlen, rb = s_tile.rb;
        const int RR = s_tile.re - rb;              // rows staged
        const int V = zlen + 2 * m2;                // virtual cells per row
        const int E = RR * V;

        // ---- stage ----------------------------------------------------------------------------------------
        // entries: RR rows x V virtual cells, then the zlen home cells once more (so the compute phase never
        // touches global memory).  1) every thread fetches populations, 2) warp 0 scans them in shared memory,
        // 3) one warp per row copies the row: consecutive virtual cells are consecutive in the sorted frame until
        // the column wraps, so a row is a few long coalesced runs.
        const int EH = E + zlen;                    // + home entries
        const int homebase = (c0 * nc1 + c1) * nc2;
        for (int t = threadIdx.x; t < RR; t += blockDim.x) {
            int d0, d1, s0_, s1_, q0_, q1_;
            tile_row_offset(s_geom, rb + t, d0, d1);
            wrap_cell(c0 + d0, nc0, s0_, q0_);
            wrap_cell(c1 + d1, nc1, s1_, q1_);
            s_rowimg[t] = (s0_ & 0xffff) | (s1_ << 16);
        }
        for (int e = threadIdx.x; e < EH; e += blockDim.x) {
            int cell;
            if (e < E) {
                const int r = rb + e / V, v = e - (e / V) * V;
                int d0, d1;
                tile_row_offset(s_geom, r, d0, d1);
                const int t0 = c0 + d0, t1 = c1 + d1, t2 = z0 - m2 + v;
                const int q0 = t0 - floordiv_i(t0, nc0) * nc0, q1 = t1 - floordiv_i(t1, nc1) * nc1, q2 = t2 - floordiv_i(t2, nc2) * nc2;
                cell = (q0 * nc1 + q1) * nc2 + q2;
            } else cell = homebase + z0 + (e - E);
            s_off[e + 1] = (int)(cs[cell + 1] - cs[cell]);
        }
        __syncthreads();
        if (warp == 0) {
            int carry = 0;
            for (int e0 = 0; e0 < EH; e0 += 32) {
                const int e = e0 + lane;
                const int cnt = e < EH ? s_off[e + 1] : 0;
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                __syncwarp();
                if (e < EH) s_off[e + 1] = carry + incl;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) s_off[0] = 0;
        }
        __syncthreads();
#if TILE_TMA
        // One warp issues the copies: lane = task (row), each contiguous run of a row is ONE cp.async.bulk whose bytes
        // complete on the block's mbarrier; everybody then waits on the barrier phase instead of moving the atoms
        // through registers (3 200 load/store pairs per tile otherwise).
        if (warp == 0) {
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // earlier generic reads of the buffer are done
                mbar_arrive_expect_tx(mbar, (unsigned)s_off[EH] * (unsigned)sizeof(SAtom));
            }
            __syncwarp();
            const unsigned abase = (unsigned)__cvta_generic_to_shared(s_atoms);
            for (int task = lane; task < RR + 1; task += 32) {
                int colbase, va, vb, ebase;
                if (task < RR) {
                    int d0, d1, s0_, s1_, q0, q1;
                    tile_row_offset(s_geom, rb + task, d0, d1);
                    wrap_cell(c0 + d0, nc0, s0_, q0);
                    wrap_cell(c1 + d1, nc1, s1_, q1);
                    colbase = (q0 * nc1 + q1) * nc2; va = z0 - m2; vb = z0 + zlen + m2; ebase = task * V;
                } else { colbase = homebase; va = z0; vb = z0 + zlen; ebase = E; }
                int v = va;
                while (v < vb) {                          // one contiguous run per wrap of the column
                    int sdum, q;
                    wrap_cell(v, nc2, sdum, q);
                    const int run = min(vb - v, nc2 - q);
                    const int src = (int)cs[colbase + q], n = (int)cs[colbase + q + run] - src;
                    if (n > 0) bulk_g2s(abase + (unsigned)s_off[ebase + (v - va)] * (unsigned)sizeof(SAtom), fr + src, (unsigned)n * (unsigned)sizeof(SAtom), mbar);
                    v += run;
                }
            }
        }
        if (threadIdx.x == 0) mbar_wait(mbar, tma_phase);     // one poller; 511 spinning threads would eat issue slots
        tma_phase ^= 1u;
        __syncthreads();
#else
        for (int task = warp; task < RR + 1; task += nwarp) {
            // task < RR: row rb+task, virtual cells [z0-m2, z0+zlen+m2); task == RR: the home cells [z0, z0+zlen)
            int colbase, va, vb, ebase;
            if (task < RR) {
                int d0, d1;
                tile_row_offset(s_geom, rb + task, d0, d1);
                const int t0 = c0 + d0, t1 = c1 + d1;
                const int q0 = t0 - floordiv_i(t0, nc0) * nc0, q1 = t1 - floordiv_i(t1, nc1) * nc1;
                colbase = (q0 * nc1 + q1) * nc2; va = z0 - m2; vb = z0 + zlen + m2; ebase = task * V;
            } else { colbase = homebase; va = z0; vb = z0 + zlen; ebase = E; }
            int v = va;
            while (v < vb) {                          // one contiguous run per wrap of the column
                const int q = v - floordiv_i(v, nc2) * nc2;
                const int run = min(vb - v, nc2 - q);
                const int src = (int)cs[colbase + q], n = (int)cs[colbase + q + run] - src;
                const double2 *gp = reinterpret_cast<const double2 *>(fr + src);
                double2 *sp = reinterpret_cast<double2 *>(s_atoms + s_off[ebase + (v - va)]);
#pragma unroll 4
                for (int k = lane; k < 2 * n; k += 32) sp[k] = __ldg(gp + k);
                v += run;
            }
        }
        __syncthreads();
#endif

        // ---- fp32 records: (p + T) - O per staged atom, one warp per row, T per stencil entry ----
        if (F32) {
            const SAtom o0 = s_atoms[s_off[E]];                       // tile origin: the first staged home atom
            for (int task = warp; task < RR + 1; task += nwarp) {
                int s0_ = 0, s1_ = 0, vlo = z0, vcount = zlen, ebase = E;
                if (task < RR) {
                    const int img = s_rowimg[task];
                    s0_ = (int)(short)(img & 0xffff); s1_ = img >> 16;
                    vlo = z0 - m2; vcount = V; ebase = task * V;
                }
                for (int v = 0; v < vcount; ++v) {
                    int s2_ = 0, q2_ = 0;
                    if (task < RR) wrap_cell(vlo + v, nc2, s2_, q2_);
                    double Tx = 0.0, Ty = 0.0, Tz = 0.0;
                    if ((s0_ | s1_ | s2_) != 0) {
                        const double fs0 = (double)s0_, fs1 = (double)s1_, fs2 = (double)s2_;
                        Tx = (fs0 * s_geom.cell[0] + fs1 * s_geom.cell[3]) + fs2 * s_geom.cell[6];
                        Ty = (fs0 * s_geom.cell[1] + fs1 * s_geom.cell[4]) + fs2 * s_geom.cell[7];
                        Tz = (fs0 * s_geom.cell[2] + fs1 * s_geom.cell[5]) + fs2 * s_geom.cell[8];
                    }
                    const int kb = s_off[ebase + v], ke = s_off[ebase + v + 1];
                    for (int k = kb + lane; k < ke; k += 32) {
                        const SAtom pk = s_atoms[k];
                        float4 c;
                        c.x = (float)((pk.x + Tx) - o0.x);
                        c.y = (float)((pk.y + Ty) - o0.y);
                        c.z = (float)((pk.z + Tz) - o0.z);
                        c.w = __int_as_float((int)(pk.s & 0xff));
                        s_f32[k] = c;
                    }
                }
            }
            __syncthreads();
        }

#if TILE_FLAT && !TILE_QUEUE
        // ---- compute: work item = (home cell, group of RG staged rows), scanned as one flat candidate list ----
        {
            int RG = (zlen * RR) / (4 * nwarp);                 // rows per item: aim at >= 4 items per warp
            RG = RG < 1 ? 1 : (RG > 4 ? 4 : RG);
            const int groups = (RR + RG - 1) / RG;
            const int items = zlen * groups;
            const unsigned g_magic0 = (65536u + (unsigned)groups - 1u) / (unsigned)groups;
            FlatRun *runs = s_runs + warp * (FLAT_MAXE + 1);
            for (int item = warp; item < items; item += nwarp) {
                const int hz = (int)(((unsigned)item * g_magic0) >> 16), grp = item - hz * groups;
                const int z = z0 + hz;
                const int hb = s_off[E + hz], nh = s_off[E + hz + 1] - hb;      // staged home cell
                if (nh == 0) continue;
                const int rr_lo = grp * RG, rr_hi = min(RR, rr_lo + RG);
                int rr = rr_lo, d2 = 0;
                bool row_open = false;
                int s0 = 0, s1 = 0;
                int own_off = 0;
                while (rr < rr_hi) {
                    // ---- describe up to FLAT_MAXE runs (all lanes compute the same scalars; lane 0 writes them)
                    int ne = 0, total = 0;
                    __syncwarp();
                    while (rr < rr_hi && ne < FLAT_MAXE) {
                        const int r = rb + rr;
                        if (!row_open) {
                            int d0, d1, q0_, q1_;
                            tile_row_offset(s_geom, r, d0, d1);
                            wrap_cell(c0 + d0, nc0, s0, q0_);
                            wrap_cell(c1 + d1, nc1, s1, q1_);
                            d2 = (r == 0) ? 0 : -m2;
                            row_open = true;
                        }
                        int s2, q2;
                        wrap_cell(z + d2, nc2, s2, q2);
                        const int len = min(m2 - d2, nc2 - 1 - q2) + 1;
                        const int v = hz + m2 + d2;
                        const int jb = s_off[rr * V + v], je = s_off[rr * V + v + len];
                        const bool after_me = (r == 0 && d2 == 0);
                        if (after_me) own_off = jb;
                        if (je > jb) {
                            if (lane == 0) {
                                FlatRun fr_;
                                if ((s0 | s1 | s2) != 0) {
                                    const double fs0 = (double)s0, fs1 = (double)s1, fs2 = (double)s2;
                                    fr_.Tx = (fs0 * s_geom.cell[0] + fs1 * s_geom.cell[3]) + fs2 * s_geom.cell[6];
                                    fr_.Ty = (fs0 * s_geom.cell[1] + fs1 * s_geom.cell[4]) + fs2 * s_geom.cell[7];
                                    fr_.Tz = (fs0 * s_geom.cell[2] + fs1 * s_geom.cell[5]) + fs2 * s_geom.cell[8];
                                } else fr_.Tx = fr_.Ty = fr_.Tz = 0.0;
                                fr_.kbeg = total | (after_me ? (1 << 30) : 0);
                                fr_.jofs = jb - total;
                                runs[ne] = fr_;
                            }
                            total += je - jb;
                            ++ne;
                        }
                        d2 += len;
                        if (d2 > m2) { ++rr; row_open = false; }
                    }
                    if (lane == 0) { FlatRun end_; end_.Tx = end_.Ty = end_.Tz = 0.0; end_.kbeg = total; end_.jofs = 0; runs[ne] = end_; }
                    __syncwarp();
                    if (total == 0) continue;
                    // ---- scan them, 32 home atoms at a time
                    for (int h0 = 0; h0 < nh; h0 += 32) {
                        const int ng = min(32, nh - h0);
                        const int G = 32 / ng;
                        const unsigned g_magic = (65536u + (unsigned)G - 1u) / (unsigned)G;
                        const int il = (int)(((unsigned)lane * g_magic) >> 16), sub = lane - il * G;
                        if (il < ng) {
                            const SAtom me = s_atoms[hb + h0 + il];
                            scan_flat<HAS_CN, CN_WIDE>(a, sa, runs, total, me, sub, G, own_off + h0 + il);
                        }
                    }
                }
            }
        }
#else
        // ---- compute: work item = (home cell, staged row); everything is read from shared memory ----
        const int items = zlen * RR;
        const unsigned rr_magic = (65536u + (unsigned)RR - 1u) / (unsigned)RR;    // exact x / RR for x < 2048
        for (int item = warp; item < items; item += nwarp) {     // (dynamic hand-out via an smem counter measured no better)
            const int hz = (int)(((unsigned)item * rr_magic) >> 16), rr = item - hz * RR, r = rb + rr;
            const int z = z0 + hz;
            const int hb = s_off[E + hz], nh = s_off[E + hz + 1] - hb;      // staged home cell
            if (nh == 0) continue;
            const int img01 = s_rowimg[rr];                               // image of the row's column, prepared per tile
            const int s0 = (int)(short)(img01 & 0xffff), s1 = img01 >> 16;
            const bool home_row = (r == 0);
            const int own_off = home_row ? s_off[rr * V + hz + m2] : 0;   // position of the home cell inside row 0
            for (int h0 = 0; h0 < nh; h0 += 32) {
                const int ng = min(32, nh - h0);                 // home atoms in this group
                const int G = c_sub_lanes[ng];
                const unsigned g_magic = G == 1 ? 65536u : c_div_magic[G];
                const int il = (int)(((unsigned)lane * g_magic) >> 16), sub = lane - il * G;
                const bool active = il < ng;
                const int hidx = hb + h0 + (active ? il : 0);
                const SAtom me = s_atoms[hidx];
                const int ism = own_off + h0 + il;
                int d2 = home_row ? 0 : -m2;
                while (d2 <= m2) {
                    int s2, q2;
                    wrap_cell(z + d2, nc2, s2, q2);
                    const int len = min(m2 - d2, nc2 - 1 - q2) + 1;
                    const int v = hz + m2 + d2;
                    const int jb = s_off[rr * V + v], je = s_off[rr * V + v + len];
                    const int n_iter = (int)(((unsigned)(je - jb + G - 1) * g_magic) >> 16);
                    const bool after_me = home_row && d2 == 0;   // own cell leads this run: partners after me only
                    if (F32) {
                        if (active) {
                            const float4 mef = s_f32[hidx];
                            const unsigned krow_a = (unsigned)__cvta_generic_to_shared(s_key) + 2u * (unsigned)((__float_as_int(mef.w) & 0xff) * S);
                            if (after_me) scan_run_f32<HAS_CN, CN_WIDE, true>(a, ta.f, s_geom, s_atoms, f32_a, band_a, edge_a, hist_a, cn_a, cnthr_a, krow_a, mef.x, mef.y, mef.z, hidx, s0, s1, s2, jb, je, G, sub, ism);
                            else scan_run_f32<HAS_CN, CN_WIDE, false>(a, ta.f, s_geom, s_atoms, f32_a, band_a, edge_a, hist_a, cn_a, cnthr_a, krow_a, mef.x, mef.y, mef.z, hidx, s0, s1, s2, jb, je, G, sub, ism);
                        }
                        d2 += len;
                        continue;
                    }
                    if ((s0 | s1 | s2) != 0) {
                        const double fs0 = (double)s0, fs1 = (double)s1, fs2 = (double)s2;
                        const double Tx = (fs0 * s_geom.cell[0] + fs1 * s_geom.cell[3]) + fs2 * s_geom.cell[6];
                        const double Ty = (fs0 * s_geom.cell[1] + fs1 * s_geom.cell[4]) + fs2 * s_geom.cell[7];
                        const double Tz = (fs0 * s_geom.cell[2] + fs1 * s_geom.cell[5]) + fs2 * s_geom.cell[8];
                        if (after_me) scan_run<HAS_CN, CN_WIDE, true, true>(a, s_atoms, s_edge2, s_cnthr, s_hist, s_cn, s_key, hq, me, Tx, Ty, Tz, jb, je, G, n_iter, sub, active, ism, hidx, lane, lt_mask, sa);
                        else scan_run<HAS_CN, CN_WIDE, true, false>(a, s_atoms, s_edge2, s_cnthr, s_hist, s_cn, s_key, hq, me, Tx, Ty, Tz, jb, je, G, n_iter, sub, active, ism, hidx, lane, lt_mask, sa);
                    } else {
                        if (after_me) scan_run<HAS_CN, CN_WIDE, false, true>(a, s_atoms, s_edge2, s_cnthr, s_hist, s_cn, s_key, hq, me, 0.0, 0.0, 0.0, jb, je, G, n_iter, sub, active, ism, hidx, lane, lt_mask, sa);
                        else scan_run<HAS_CN, CN_WIDE, false, false>(a, s_atoms, s_edge2, s_cnthr, s_hist, s_cn, s_key, hq, me, 0.0, 0.0, 0.0, jb, je, G, n_iter, sub, active, ism, hidx, lane, lt_mask, sa);
                    }
                    d2 += len;
                }
            }
        }
#endif
        // the staged atoms are about to be replaced: bin what is still queued
        if (hq.tail != hq.head) hits_flush<HAS_CN>(a, s_atoms, s_edge2, s_cnthr, s_hist, s_cn, s_key, hq, lane, (int)(hq.tail - hq.head));
        if (HAS_CN) {
            __syncthreads();
            for (int k = threadIdx.x; k < a.nkeys; k += blockDim.x) {
                const uint32_t v = s_cn[k];
                if (v) {
                    atomicAdd(&a.cn_out[(size_t)f * a.nkeys + k], (unsigned long long)v);
                    s_cn[k] = 0u;
                }
            }
        }
    }
    __syncthreads();
    unsigned long long *slab = a.slabs + (size_t)blockIdx.x * a.nkeys * a.nbins;
    for (int k = threadIdx.x; k < a.nkeys * a.nbins; k += blockDim.x) {
        const uint32_t v = s_hist[k];
        if (v) slab[k] += v;
    }
}
