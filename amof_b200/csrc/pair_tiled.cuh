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
//   k_pair_tiled  persistent blocks over the tile list.
//
// Compute phase ("candidates in registers, home atoms broadcast"): the candidates of one home cell -- the runs of all
// staged rows -- form ONE flat index space; a work item is (home cell, chunk of 32*TILE_K flat candidates).  A warp
// loads its chunk once (TILE_K candidates per lane, kept in registers) and then walks the home atoms of the cell: the
// home atom is read with a broadcast shared-memory load (one wavefront), every lane forms its TILE_K distances.  All 32
// lanes hold candidates whatever the cell populations, and shared-memory traffic is one broadcast per 32*TILE_K
// distances instead of one 32-byte record per lane and distance (the round-1 kernel: lanes = home atom x sub-lane, one
// item per (home cell, row); experiments/csrc/pair_tiled_r1.cuh).  A two-pass variant that first drops the candidates
// farther than the search radius from the bounding sphere of the home atoms (41 % of them) and compacts the rest is in
// experiments/csrc/pair_tiled_twopass.cuh: bit-exact, but slower.
//
// Arithmetic, thresholds, folding of species pairs and histogram privatisation are exactly those of pair.cuh.
#pragma once
#include "pair.cuh"

#ifndef TILE_THREADS
#define TILE_THREADS 512     // 2 blocks x 16 warps per SM at <= 64 registers
#endif
#ifndef TILE_MIN_BLOCKS
#define TILE_MIN_BLOCKS 2
#endif
#ifndef TILE_K
#define TILE_K 2             // candidates per lane
#endif
#ifndef TILE_HUNROLL
#define TILE_HUNROLL 1       // home atoms per trip of the inner loop
#endif
#define TILE_STATIC_SMEM 2560     // upper bound of the kernel's static shared memory (host-side budget)
#define TILE_CHUNK (32 * TILE_K)
#define TILE_PRAGMA_(x) _Pragma(#x)
#define TILE_PRAGMA_UNROLL(n) TILE_PRAGMA_(unroll n)
#define TILE_MAX_ENTRIES 1024     // rows x virtual cells per tile
#define TILE_MAX_ZLEN 32           // home cells per tile: one lane per home cell in the work split
#define TILE_MAX_ROWS 256
#define TILE_OFF_WORDS (TILE_MAX_ENTRIES + TILE_MAX_ZLEN + 2)       // offsets: entries + home cells + 1
#define TILE_PRE_WORDS (TILE_MAX_ENTRIES + 2 * TILE_MAX_ZLEN + 2)   // per home cell: (rows + 1) packed row descriptors

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
    int uniform_cols;        // > 0: every frame has this many columns (nc0 * nc1), so frame = warp / uniform_cols
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
    } else {
        // frames may have different grids: walk the frames
        for (; f < a.n_frames; ++f) {
            long long cols = (long long)a.geom[f].nc[0] * a.geom[f].nc[1];
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
        int zlen = min(min(nc2 - z, vmax - 2 * m2), TILE_MAX_ZLEN);
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

struct TiledArgs {
    PairArgs p;
    const PairTile *tiles;
    const int *n_tiles;
    int cap;                 // staged atoms per tile
    int max_tiles;
};

// shared-memory accesses through an explicit 32-bit shared address kept in a register: the compiler otherwise
// re-derives the shared window base (S2R SR_CgaCtaId + LEA) inside the loops
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
__device__ __forceinline__ int lds_s32(unsigned addr) {
    int v;
    asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
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
    unsigned atoms, edge, cnthr, hist, cn, key, off, pre, code, tvec;
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

__device__ __forceinline__ int warp_incl_scan_i(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// The K candidates of every lane (in registers) against the nh home atoms of one cell, read by broadcast loads.
//   SHIFT: some candidate of the chunk sits in a periodic image; without it (pj - pi) + 0 == pj - pi bit for bit (up to
//          the sign of a zero, which the squares drop), so the adds go.
//   AFTER: the chunk holds candidates of the home cell itself: the pair (home h, own-cell candidate f) counts iff f > h,
//          passed as hlim[k] = f (nh for every other candidate).
//   CN_WIDE: some cutoff exceeds rmax, so a candidate inside r2search can still be outside the RDF range.
template <bool HAS_CN, bool CN_WIDE, bool SHIFT, bool AFTER, int K>
__device__ __forceinline__ void bcast_scan(const PairArgs &a, const SmemAddr &sa, unsigned haddr, int nh,
                                           const double (&cx)[K], const double (&cy)[K], const double (&cz)[K],
                                           const unsigned (&krow)[K], const double (&Tx)[K], const double (&Ty)[K],
                                           const double (&Tz)[K], const int (&hlim)[K]) {
    const double r2search = a.r2search, r2max = a.r2max, cn_r2max = a.cn_r2max;
    const float inv_dr_f = a.inv_dr_f, margin = a.bin_margin;
    const int nbins = a.nbins;
    const unsigned edge_addr = sa.edge, hist_addr = sa.hist, cnthr_addr = sa.cnthr, cn_addr = sa.cn;
    // the hit work: exact bin + shared-memory increment (the host only selects this kernel when margin > 0)
    auto hit = [&](double dd, unsigned kaddr) {
        const int key = lds_u16(kaddr);
        if (!CN_WIDE || dd < r2max) {             // r2search == r2max unless a cutoff reaches beyond rmax
            const int b = rdf_bin_s(dd, edge_addr, inv_dr_f, margin);
            reds_inc(hist_addr + 4u * (unsigned)(key * nbins + b));
        }
        if (HAS_CN && dd < cn_r2max && dd < lds_f64(cnthr_addr + 8u * (unsigned)key)) reds_inc(cn_addr + 4u * (unsigned)key);
    };
    TILE_PRAGMA_UNROLL(TILE_HUNROLL)
    for (int h = 0; h < nh; ++h, haddr += 32u) {
        double hx, hy, hz;
        lds_xyz(haddr, hx, hy, hz);                                   // same address in every lane: one wavefront
        const unsigned si2 = 2u * (unsigned)lds_species(haddr);
        double dd[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {                                    // all distance chains first: they overlap in the FP64 pipe
            double dx = cx[k] - hx, dy = cy[k] - hy, dz = cz[k] - hz;    // P3: (pj - pi) + T, j the candidate
            if (SHIFT) { dx += Tx[k]; dy += Ty[k]; dz += Tz[k]; }
            dd[k] = (dx * dx + dy * dy) + dz * dz;
        }
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (dd[k] < r2search && (!AFTER || h < hlim[k])) hit(dd[k], krow[k] + si2);
    }
}

__device__ __forceinline__ int lds_u8(unsigned addr) {
    unsigned v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return (int)v;
}

// One chunk: the flat candidates [base, base + 32*K) of the home cell described by pre_a, against its nh home atoms.
// rr0: the row that holds flat index `base`.
template <bool HAS_CN, bool CN_WIDE, int K>
__device__ __forceinline__ void chunk(const PairArgs &a, const SmemAddr &sa, unsigned pre_a, int rr0, int base, int F, int lane,
                                      unsigned haddr, int nh, bool after, bool tile_img) {
    double cx[K], cy[K], cz[K], Tx[K], Ty[K], Tz[K];
    unsigned krow[K];
    int hlim[K];
    bool shifted = false;
    // walk state: staged index = flat index + joff while flat index < nstart
    int rr = rr0 - 1, next = lds_s32(pre_a + 4u * (unsigned)rr0), joff = 0, nstart = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int fl = base + k * 32 + lane;
        Tx[k] = Ty[k] = Tz[k] = 0.0;
        hlim[k] = nh;
        if (fl < F) {
            while (fl >= nstart) {
                ++rr;
                joff = (next >> 16) - (next & 0xffff);
                next = lds_s32(pre_a + 4u * (unsigned)(rr + 1));
                nstart = next & 0xffff;
            }
            const int j = fl + joff;
            const unsigned addr = sa.atoms + (unsigned)j * 32u;
            lds_xyz(addr, cx[k], cy[k], cz[k]);
            krow[k] = sa.key + 2u * (unsigned)(lds_species(addr) * a.n_species);
            if (after && fl < nh) hlim[k] = fl;                          // a candidate of the home cell itself
            if (tile_img) {
                const unsigned code = (unsigned)lds_u8(sa.code + (unsigned)j);
                if (code != 13u) {
                    const unsigned ta = sa.tvec + 24u * code;
                    Tx[k] = lds_f64(ta); Ty[k] = lds_f64(ta + 8u); Tz[k] = lds_f64(ta + 16u);
                    shifted = true;
                }
            }
        } else {
            cx[k] = cy[k] = cz[k] = 1e300;                               // d2 = inf: below no threshold
            krow[k] = sa.key;
        }
    }
    if (tile_img && __any_sync(0xffffffffu, shifted)) {
        if (after) bcast_scan<HAS_CN, CN_WIDE, true, true, K>(a, sa, haddr, nh, cx, cy, cz, krow, Tx, Ty, Tz, hlim);
        else bcast_scan<HAS_CN, CN_WIDE, true, false, K>(a, sa, haddr, nh, cx, cy, cz, krow, Tx, Ty, Tz, hlim);
    } else {
        if (after) bcast_scan<HAS_CN, CN_WIDE, false, true, K>(a, sa, haddr, nh, cx, cy, cz, krow, Tx, Ty, Tz, hlim);
        else bcast_scan<HAS_CN, CN_WIDE, false, false, K>(a, sa, haddr, nh, cx, cy, cz, krow, Tx, Ty, Tz, hlim);
    }
}

template <bool HAS_CN, bool CN_WIDE>
__global__ void __launch_bounds__(TILE_THREADS, TILE_MIN_BLOCKS) k_pair_tiled(TiledArgs ta) {
    const PairArgs &a = ta.p;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: atoms[cap] (32 B) | edge2[nbins+1] | cn_thr2[nkeys] | hist[nkeys*nbins] u32 | cn_cnt[nkeys] u32 | off[] int | pre[] int |
    //         keyidx[S*S] u16 | image code[cap] u8; carved by byte offsets from the shared base, so every pointer keeps the shared address space
    size_t off = 0;
    SAtom *s_atoms = reinterpret_cast<SAtom *>(smem_raw + off);            off += sizeof(SAtom) * (size_t)ta.cap;
    double *s_edge2 = reinterpret_cast<double *>(smem_raw + off);          off += sizeof(double) * (size_t)(a.nbins + 1);
    double *s_cnthr = reinterpret_cast<double *>(smem_raw + off);          off += sizeof(double) * (size_t)(HAS_CN ? a.nkeys : 0);
    off = (off + 15) & ~(size_t)15;
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(smem_raw + off);       off += sizeof(uint32_t) * (size_t)a.nkeys * a.nbins;
    uint32_t *s_cn = reinterpret_cast<uint32_t *>(smem_raw + off);         off += sizeof(uint32_t) * (size_t)(HAS_CN ? a.nkeys : 0);
    int *s_off = reinterpret_cast<int *>(smem_raw + off);                  off += sizeof(int) * TILE_OFF_WORDS;
    int *s_pre = reinterpret_cast<int *>(smem_raw + off);                  off += sizeof(int) * TILE_PRE_WORDS;
    uint16_t *s_key = reinterpret_cast<uint16_t *>(smem_raw + off);        off += sizeof(uint16_t) * (size_t)(a.n_species * a.n_species);
    unsigned char *s_code = smem_raw + off;
    __shared__ FrameGeom s_geom;
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ int s_rowimg[TILE_MAX_ROWS];        // per staged row: image (s0 | s1 << 16) of its column
    __shared__ double s_tvec[27 * 3];              // image shift T(s0, s1, s2), s in {-1, 0, 1}^3, code = (s0+1) + 3 (s1+1) + 9 (s2+1)
    SmemAddr sa;
    {
        const unsigned sb = opaque_u32((unsigned)__cvta_generic_to_shared(smem_raw));
        sa.atoms = sb + (unsigned)(reinterpret_cast<unsigned char *>(s_atoms) - smem_raw);
        sa.edge = sb + (unsigned)(reinterpret_cast<unsigned char *>(s_edge2) - smem_raw);
        sa.cnthr = sb + (unsigned)(reinterpret_cast<unsigned char *>(s_cnthr) - smem_raw);
        sa.hist = sb + (unsigned)(reinterpret_cast<unsigned char *>(s_hist) - smem_raw);
        sa.cn = sb + (unsigned)(reinterpret_cast<unsigned char *>(s_cn) - smem_raw);
        sa.key = sb + (unsigned)(reinterpret_cast<unsigned char *>(s_key) - smem_raw);
        sa.off = sb + (unsigned)(reinterpret_cast<unsigned char *>(s_off) - smem_raw);
        sa.pre = sb + (unsigned)(reinterpret_cast<unsigned char *>(s_pre) - smem_raw);
        sa.code = sb + (unsigned)(s_code - smem_raw);
        sa.tvec = opaque_u32((unsigned)__cvta_generic_to_shared(s_tvec));
    }

    const int S = a.n_species;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = TILE_THREADS / 32;
    for (int k = threadIdx.x; k <= a.nbins; k += blockDim.x) s_edge2[k] = a.edge2[k];
    for (int k = threadIdx.x; k < a.nkeys * a.nbins; k += blockDim.x) s_hist[k] = 0u;
    if (HAS_CN)
        for (int k = threadIdx.x; k < a.nkeys; k += blockDim.x) { s_cnthr[k] = a.cn_thr2[k]; s_cn[k] = 0u; }
    for (int k = threadIdx.x; k < S * S; k += blockDim.x) s_key[k] = a.keyidx[k];

    const unsigned mbar = (unsigned)__cvta_generic_to_shared(&s_mbar);
    if (threadIdx.x == 0) mbar_init(mbar, 1);
    unsigned tma_phase = 0;
    const int n_tiles = min(*ta.n_tiles, ta.max_tiles);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();     // previous tile fully consumed (atoms, offsets, geometry, cn counters)
        // every thread reads the 32-byte tile record itself (one broadcast line from L2): no shared copy, no barrier
        const int4 t_lo = __ldg(reinterpret_cast<const int4 *>(&ta.tiles[tile]));
        const int4 t_hi = __ldg(reinterpret_cast<const int4 *>(&ta.tiles[tile]) + 1);
        const int f = t_lo.x, c0 = t_lo.y, c1 = t_lo.z, z0 = t_lo.w, zlen = t_hi.x, rb = t_hi.y;
        const int RR = t_hi.z - rb;                 // rows staged
        if (threadIdx.x < (int)(sizeof(FrameGeom) / sizeof(int)))
            reinterpret_cast<int *>(&s_geom)[threadIdx.x] = reinterpret_cast<const int *>(&a.geom[f])[threadIdx.x];
        __syncthreads();
        const SAtom *fr = a.sorted + (long long)f * a.n_atoms;
        const uint32_t *cs = a.cell_start + s_geom.cs_off;
        const int nc0 = s_geom.nc[0], nc1 = s_geom.nc[1], nc2 = s_geom.nc[2];
        const int m2 = s_geom.m[2];
        const int V = zlen + 2 * m2;                // virtual cells per row
        const int E = RR * V;
        // does any staged cell sit in a periodic image?  (interior tiles skip every image lookup)
        const bool tile_img = (c0 + s_geom.m[0] >= nc0) || (c1 - s_geom.m[1] < 0) || (c1 + s_geom.m[1] >= nc1) ||
                              (z0 - m2 < 0) || (z0 + zlen - 1 + m2 >= nc2);

        // ---- stage ----------------------------------------------------------------------------------------
        // entries: RR rows x V virtual cells, then the zlen home cells once more (so the compute phase never
        // touches global memory).  1) every thread fetches populations, 2) warp 0 scans them in shared memory,
        // 3) warp 0 issues one bulk copy per contiguous run of a row: consecutive virtual cells are consecutive in the
        // sorted frame until the column wraps, so a row is a few long runs.
        const int EH = E + zlen;                    // + home entries
        const int homebase = (c0 * nc1 + c1) * nc2;
        for (int t = threadIdx.x; t < RR; t += blockDim.x) {
            int d0, d1, s0_, s1_, q0_, q1_;
            tile_row_offset(s_geom, rb + t, d0, d1);
            wrap_cell(c0 + d0, nc0, s0_, q0_);
            wrap_cell(c1 + d1, nc1, s1_, q1_);
            s_rowimg[t] = (s0_ + 1) + 3 * (s1_ + 1);     // (s0, s1) part of the image code
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
                const int incl = warp_incl_scan_i(cnt, lane);
                __syncwarp();
                if (e < EH) s_off[e + 1] = carry + incl;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) s_off[0] = 0;
        }
        __syncthreads();
        // One warp issues the copies: lane = task (row), each contiguous run of a row is ONE cp.async.bulk whose bytes
        // complete on the block's mbarrier; everybody then waits on the barrier phase instead of moving the atoms
        // through registers.
        if (warp == 0) {
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // earlier generic reads of the buffer are done
                mbar_arrive_expect_tx(mbar, (unsigned)s_off[EH] * (unsigned)sizeof(SAtom));
            }
            __syncwarp();
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
                    if (n > 0) bulk_g2s(sa.atoms + (unsigned)s_off[ebase + (v - va)] * (unsigned)sizeof(SAtom), fr + src, (unsigned)n * (unsigned)sizeof(SAtom), mbar);
                    v += run;
                }
            }
        } else {
            // meanwhile the other warps lay out the flat candidate space of every home cell: one warp per home cell, lanes =
            // staged rows; row descriptor = (flat index of the row's first candidate) | (its staged index << 16), then
            // the total.  Row 0 of the stencil (the home column itself) starts at the home cell: partners before it were
            // counted from their side.
            for (int hz = warp - 1; hz < zlen; hz += nwarp - 1) {
                int carry = 0;
                int *pre = s_pre + hz * (RR + 1);
                for (int r0 = 0; r0 < RR; r0 += 32) {
                    const int rr = r0 + lane;
                    int jb = 0, len = 0;
                    if (rr < RR) {
                        jb = s_off[rr * V + hz + (rb + rr == 0 ? m2 : 0)];
                        len = s_off[rr * V + hz + 2 * m2 + 1] - jb;
                    }
                    const int incl = warp_incl_scan_i(len, lane);
                    if (rr < RR) pre[rr] = (carry + incl - len) | (jb << 16);
                    carry += __shfl_sync(0xffffffffu, incl, 31);
                }
                if (lane == 0) pre[RR] = carry;
            }
            if (tile_img) {
                // image shift of every image code (P3), and the image code of every staged atom: one thread per staged cell
                const int t = threadIdx.x - 32;
                if (t < 27) {
                    const double fs0 = (double)(t % 3 - 1), fs1 = (double)((t / 3) % 3 - 1), fs2 = (double)(t / 9 - 1);
                    s_tvec[3 * t] = (fs0 * s_geom.cell[0] + fs1 * s_geom.cell[3]) + fs2 * s_geom.cell[6];
                    s_tvec[3 * t + 1] = (fs0 * s_geom.cell[1] + fs1 * s_geom.cell[4]) + fs2 * s_geom.cell[7];
                    s_tvec[3 * t + 2] = (fs0 * s_geom.cell[2] + fs1 * s_geom.cell[5]) + fs2 * s_geom.cell[8];
                }
                for (int e = t; e < E; e += TILE_THREADS - 32) {
                    const int r = e / V, t2 = z0 - m2 + (e - r * V);
                    const int code = s_rowimg[r] + 9 * (t2 < 0 ? 0 : (t2 >= nc2 ? 2 : 1));
                    for (int k = s_off[e]; k < s_off[e + 1]; ++k) s_code[k] = (unsigned char)code;
                }
            }
        }
        if (threadIdx.x == 0) mbar_wait(mbar, tma_phase);     // one poller; 511 spinning threads would eat issue slots
        tma_phase ^= 1u;
        __syncthreads();

        // ---- compute --------------------------------------------------------------------------------------
        // items = (home cell, chunk of its flat candidate space) in home-cell order; lane = home cell: every warp keeps the
        // inclusive prefix of the chunk counts in a register and finds the cell of an item with one ballot
        int nc_l = 0;
        if (lane < zlen && s_off[E + lane + 1] > s_off[E + lane]) nc_l = (s_pre[lane * (RR + 1) + RR] + TILE_CHUNK - 1) / TILE_CHUNK;
        const int ci_l = warp_incl_scan_i(nc_l, lane);
        const int n_items = __shfl_sync(0xffffffffu, ci_l, 31);
        for (int item = warp; item < n_items; item += nwarp) {
            const int hz = __popc(__ballot_sync(0xffffffffu, ci_l <= item));      // cells wholly before this item
            const int base = (item - __shfl_sync(0xffffffffu, ci_l - nc_l, hz)) * TILE_CHUNK;
            const int hb = s_off[E + hz], nh = s_off[E + hz + 1] - hb;
            const unsigned pre_a = sa.pre + 4u * (unsigned)(hz * (RR + 1));
            const int F = lds_s32(pre_a + 4u * (unsigned)RR);
            // first row of the chunk: the last row whose flat start is <= base (row starts are non-decreasing)
            int rr0 = -1;
            for (int r0 = 0; r0 < RR; r0 += 32) {
                const int p = r0 + lane < RR ? (lds_s32(pre_a + 4u * (unsigned)(r0 + lane)) & 0xffff) : 0x7fffffff;
                rr0 += __popc(__ballot_sync(0xffffffffu, p <= base));
            }
            const unsigned haddr = sa.atoms + (unsigned)hb * 32u;
            const bool after = rb == 0 && base < nh;          // row 0 opens with the home cell itself
            if (TILE_K > 1 && F - base <= 32) chunk<HAS_CN, CN_WIDE, 1>(a, sa, pre_a, rr0, base, F, lane, haddr, nh, after, tile_img);
            else chunk<HAS_CN, CN_WIDE, TILE_K>(a, sa, pre_a, rr0, base, F, lane, haddr, nh, after, tile_img);
        }
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
