// pair_pipe.cuh -- the tiled RDF(+CN) pair kernel as a producer/consumer pipeline.
//
// k_pair_tiled (pair_tiled.cuh) alternates, per tile, a staging phase (populations, scan, TMA copies) and a compute
// phase, separated by block barriers; ncu shows ~20 % of the warp time parked at those barriers and a third of the
// executed instructions in staging and set-up.  Here the two phases run CONCURRENTLY:
//
//   * one block of PIPE_THREADS per SM, PIPE_STAGES tile buffers in shared memory;
//   * warp 0 is the PRODUCER: for tile t+1 it derives the offset table straight from the frame's cell_start[] (a row of
//     the stencil is contiguous in the sorted frame, so offsets inside a row are differences of cell_start[]; only the
//     per-row totals need a scan), writes the per-row image shifts, and issues one TMA bulk copy per contiguous run,
//     all completing on the buffer's `full` mbarrier;
//   * the other warps are CONSUMERS: they wait on `full`, take work items (home cell x stencil row) from a shared
//     counter and scan them exactly as k_pair_tiled does (same arithmetic, same thresholds, same privatised
//     histogram); the last warp to run out of items flushes the tile's coordination counters and arrives on the
//     buffer's `empty` mbarrier, which lets the producer overwrite it.
//
// No block-wide barrier is executed between the prologue and the final histogram merge: a warp that finishes its
// share of tile t moves on to tile t+1 as soon as that buffer is full.
#pragma once
#include "pair_tiled.cuh"

#ifndef PIPE_THREADS
#define PIPE_THREADS 1024
#endif
#ifndef PIPE_STAGES
#define PIPE_STAGES 2
#endif
#define PIPE_CONSUMERS (PIPE_THREADS / 32 - 1)

struct __align__(16) PipeMeta {
    double cell[9];          // lattice vectors of the tile's frame (image shifts along the fastest axis)
    int frame, z0, zlen, rb, RR, V, E, nc2, m2, items;
    unsigned rr_magic;       // ceil(2^32 / RR): item / RR = umulhi(item, rr_magic) for RR > 1
    int pad;
};

// byte offsets of the dynamic shared memory, one definition for the host (size) and the kernel (carve-up)
struct PipeLayout {
    unsigned atoms, edge, cnthr, hist, cn, off, rowT, rowimg, key, scratch, total;
};
__host__ __device__ inline PipeLayout pipe_layout(int cap, int nbins, int nkeys, int S, bool has_cn) {
    PipeLayout L;
    unsigned o = 0;
    L.atoms = o;   o += (unsigned)sizeof(SAtom) * (unsigned)cap * PIPE_STAGES;
    L.edge = o;    o += 8u * (unsigned)(nbins + 1);
    L.cnthr = o;   o += 8u * (unsigned)(has_cn ? nkeys : 0);
    L.rowT = o;    o += 24u * TILE_MAX_ROWS * PIPE_STAGES;
    L.hist = o;    o += 4u * (unsigned)nkeys * (unsigned)nbins;
    L.cn = o;      o += 4u * (unsigned)(has_cn ? nkeys : 0) * PIPE_STAGES;
    L.off = o;     o += 4u * TILE_OFF_WORDS * PIPE_STAGES;
    L.rowimg = o;  o += 4u * TILE_MAX_ROWS * PIPE_STAGES;
    L.scratch = o; o += 4u * 2u * (TILE_MAX_ROWS + 1);          // producer only: row bases and column bases
    L.key = o;     o += 2u * (unsigned)(S * S);
    L.total = (o + 15u) & ~15u;
    return L;
}

__device__ __forceinline__ void mbar_arrive(unsigned mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(mbar) : "memory");
}

// One run of staged candidates [jb, je) against this warp's home atoms; see scan_run (pair_tiled.cuh) for the rules.
template <bool HAS_CN, bool CN_WIDE, bool SHIFT, bool AFTER>
__device__ __forceinline__ void pipe_scan_run(const PairArgs &a, const SmemAddr &sa, double mx, double my, double mz, unsigned krow_addr,
                                              double Tx, double Ty, double Tz, int jb, int je, int G, int sub, int ism) {
    const double r2search = a.r2search, r2max = a.r2max, cn_r2max = a.cn_r2max;
    const float inv_dr_f = a.inv_dr_f, margin = a.bin_margin;
    const int nbins = a.nbins;
    const unsigned abase = sa.atoms;
    const unsigned astep = (unsigned)G * 32u;
    const unsigned aend = abase + (unsigned)je * 32u;
    const unsigned askip = ism >= 0 ? abase + (unsigned)ism * 32u : 0u;
    auto hit = [&](unsigned addr, double dd) {
        const int key = lds_u16(krow_addr + 2u * (unsigned)lds_species(addr));
        if (!CN_WIDE || dd < r2max) {
            const int b = rdf_bin_s(dd, sa.edge, inv_dr_f, margin);
            reds_inc(sa.hist + 4u * (unsigned)(key * nbins + b));
        }
        if (HAS_CN && dd < cn_r2max && dd < lds_f64(sa.cnthr + 8u * (unsigned)key)) reds_inc(sa.cn + 4u * (unsigned)key);
    };
    unsigned addr = abase + (unsigned)(jb + sub) * 32u;
    for (; addr + astep < aend; addr += 2u * astep) {
        const unsigned addr1 = addr + astep;
        double ox0, oy0, oz0, ox1, oy1, oz1;
        lds_xyz(addr, ox0, oy0, oz0);
        lds_xyz(addr1, ox1, oy1, oz1);
        double dx0 = ox0 - mx, dy0 = oy0 - my, dz0 = oz0 - mz;
        double dx1 = ox1 - mx, dy1 = oy1 - my, dz1 = oz1 - mz;
        if (SHIFT) { dx0 += Tx; dy0 += Ty; dz0 += Tz; dx1 += Tx; dy1 += Ty; dz1 += Tz; }
        const double dd0 = (dx0 * dx0 + dy0 * dy0) + dz0 * dz0;
        const double dd1 = (dx1 * dx1 + dy1 * dy1) + dz1 * dz1;
        if (dd0 < r2search && !(AFTER && addr <= askip)) hit(addr, dd0);
        if (dd1 < r2search && !(AFTER && addr1 <= askip)) hit(addr1, dd1);
    }
    if (addr < aend) {
        double ox, oy, oz;
        lds_xyz(addr, ox, oy, oz);
        double dx = ox - mx, dy = oy - my, dz = oz - mz;
        if (SHIFT) { dx += Tx; dy += Ty; dz += Tz; }
        const double dd = (dx * dx + dy * dy) + dz * dz;
        if (dd < r2search && !(AFTER && addr <= askip)) hit(addr, dd);
    }
}

template <bool HAS_CN, bool CN_WIDE>
__global__ void __launch_bounds__(PIPE_THREADS, 1) k_pair_pipe(TiledArgs ta) {
    const PairArgs &a = ta.p;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ PipeMeta s_meta[PIPE_STAGES];
    __shared__ __align__(8) unsigned long long s_full[PIPE_STAGES], s_empty[PIPE_STAGES];
    __shared__ int s_next[PIPE_STAGES], s_done[PIPE_STAGES];

    const int S = a.n_species, cap = ta.cap;
    const PipeLayout L = pipe_layout(cap, a.nbins, a.nkeys, S, HAS_CN);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned sb = opaque_u32((unsigned)__cvta_generic_to_shared(smem_raw));

    // ---- prologue: tables and a zeroed private histogram ---------------------------------------------------
    {
        double *s_edge2 = reinterpret_cast<double *>(smem_raw + L.edge);
        double *s_cnthr = reinterpret_cast<double *>(smem_raw + L.cnthr);
        uint32_t *s_hist = reinterpret_cast<uint32_t *>(smem_raw + L.hist);
        uint32_t *s_cn = reinterpret_cast<uint32_t *>(smem_raw + L.cn);
        uint16_t *s_key = reinterpret_cast<uint16_t *>(smem_raw + L.key);
        for (int k = threadIdx.x; k <= a.nbins; k += blockDim.x) s_edge2[k] = a.edge2[k];
        for (int k = threadIdx.x; k < a.nkeys * a.nbins; k += blockDim.x) s_hist[k] = 0u;
        if (HAS_CN) {
            for (int k = threadIdx.x; k < a.nkeys; k += blockDim.x) s_cnthr[k] = a.cn_thr2[k];
            for (int k = threadIdx.x; k < a.nkeys * PIPE_STAGES; k += blockDim.x) s_cn[k] = 0u;
        }
        for (int k = threadIdx.x; k < S * S; k += blockDim.x) s_key[k] = a.keyidx[k];
        if (threadIdx.x == 0) {
            for (int b = 0; b < PIPE_STAGES; ++b) {
                mbar_init((unsigned)__cvta_generic_to_shared(&s_full[b]), 1);
                mbar_init((unsigned)__cvta_generic_to_shared(&s_empty[b]), 1);
                s_next[b] = 0; s_done[b] = 0;
            }
        }
    }
    __syncthreads();
    const int n_tiles = min(*ta.n_tiles, ta.max_tiles);

    if (warp == 0) {
        // ================================ producer ==========================================================
        int *s_rowbase = reinterpret_cast<int *>(smem_raw + L.scratch);
        int *s_colbase = s_rowbase + (TILE_MAX_ROWS + 1);
        int k = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
            const int buf = k % PIPE_STAGES;
            if (k >= PIPE_STAGES) {                  // the buffer's previous tile must be fully consumed
                if (lane == 0) mbar_wait((unsigned)__cvta_generic_to_shared(&s_empty[buf]), (unsigned)((k / PIPE_STAGES - 1) & 1));
                __syncwarp();
            }
            const int4 t_lo = __ldg(reinterpret_cast<const int4 *>(&ta.tiles[tile]));
            const int4 t_hi = __ldg(reinterpret_cast<const int4 *>(&ta.tiles[tile]) + 1);
            const int f = t_lo.x, c0 = t_lo.y, c1 = t_lo.z, z0 = t_lo.w, zlen = t_hi.x, rb = t_hi.y, RR = t_hi.z - t_hi.y;
            const FrameGeom &g = a.geom[f];
            const SAtom *fr = a.sorted + (long long)f * a.n_atoms;
            const uint32_t *cs = a.cell_start + g.cs_off;
            const int nc0 = g.nc[0], nc1 = g.nc[1], nc2 = g.nc[2], m2 = g.m[2];
            const int V = zlen + 2 * m2, E = RR * V, EH = E + zlen;
            const int homebase = (c0 * nc1 + c1) * nc2;
            int *s_off = reinterpret_cast<int *>(smem_raw + L.off) + buf * TILE_OFF_WORDS;
            int *s_rowimg = reinterpret_cast<int *>(smem_raw + L.rowimg) + buf * TILE_MAX_ROWS;
            double *s_rowT = reinterpret_cast<double *>(smem_raw + L.rowT) + buf * (3 * TILE_MAX_ROWS);
            const double cx0 = g.cell[0], cy0 = g.cell[1], cz0 = g.cell[2], cx1 = g.cell[3], cy1 = g.cell[4], cz1 = g.cell[5];

            // rows: column, image, population -> exclusive scan of the row totals
            int carry = 0;
            for (int r0 = 0; r0 < RR; r0 += 32) {
                const int r = r0 + lane;
                int cnt = 0;
                if (r < RR) {
                    int d0, d1, s0_, s1_, q0, q1;
                    tile_row_offset(g, rb + r, d0, d1);
                    wrap_cell(c0 + d0, nc0, s0_, q0);
                    wrap_cell(c1 + d1, nc1, s1_, q1);
                    const int colbase = (q0 * nc1 + q1) * nc2;
                    cnt = column_count(cs, colbase, nc2, z0 - m2, z0 + zlen + m2 - 1);
                    s_colbase[r] = colbase;
                    s_rowimg[r] = (s0_ & 0xffff) | (s1_ << 16);
                    const double fs0 = (double)s0_, fs1 = (double)s1_;
                    s_rowT[3 * r + 0] = fs0 * cx0 + fs1 * cx1;       // P3: (s0*a + s1*b) first, + s2*c by the consumer
                    s_rowT[3 * r + 1] = fs0 * cy0 + fs1 * cy1;
                    s_rowT[3 * r + 2] = fs0 * cz0 + fs1 * cz1;
                }
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                if (r < RR) s_rowbase[r] = carry + incl - cnt;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            const int homestart = carry;
            const int home0 = (int)cs[homebase + z0];
            const int total = homestart + (int)cs[homebase + z0 + zlen] - home0;
            __syncwarp();
            // entries: offset = row base + population of the row's virtual cells before this one
            for (int e = lane; e <= EH; e += 32) {
                int o;
                if (e < E) {
                    const int r = e / V, v = e - r * V;
                    o = s_rowbase[r] + (v > 0 ? column_count(cs, s_colbase[r], nc2, z0 - m2, z0 - m2 + v - 1) : 0);
                } else o = homestart + (int)cs[homebase + z0 + (e - E)] - home0;
                s_off[e] = o;
            }
            if (lane == 0) {
                PipeMeta &mt = s_meta[buf];
#pragma unroll
                for (int q = 0; q < 9; ++q) mt.cell[q] = g.cell[q];
                mt.frame = f; mt.z0 = z0; mt.zlen = zlen; mt.rb = rb; mt.RR = RR; mt.V = V; mt.E = E; mt.nc2 = nc2; mt.m2 = m2;
                mt.items = zlen * RR;
                mt.rr_magic = RR > 1 ? (unsigned)((0x100000000ull + (unsigned long long)RR - 1ull) / (unsigned long long)RR) : 0u;
                s_next[buf] = 0;
            }
            __syncwarp();
            const unsigned full = (unsigned)__cvta_generic_to_shared(&s_full[buf]);
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the consumers' generic reads of the buffer are done
                mbar_arrive_expect_tx(full, (unsigned)total * (unsigned)sizeof(SAtom));
            }
            __syncwarp();
            const unsigned abase = sb + L.atoms + (unsigned)buf * (unsigned)cap * (unsigned)sizeof(SAtom);
            for (int task = lane; task < RR + 1; task += 32) {
                int colbase, va, vb, ebase;
                if (task < RR) { colbase = s_colbase[task]; va = z0 - m2; vb = z0 + zlen + m2; ebase = task * V; }
                else { colbase = homebase; va = z0; vb = z0 + zlen; ebase = E; }
                int v = va;
                while (v < vb) {                          // one contiguous run per wrap of the column
                    int sdum, q;
                    wrap_cell(v, nc2, sdum, q);
                    const int run = min(vb - v, nc2 - q);
                    const int src = (int)cs[colbase + q], n = (int)cs[colbase + q + run] - src;
                    if (n > 0) bulk_g2s(abase + (unsigned)s_off[ebase + (v - va)] * (unsigned)sizeof(SAtom), fr + src, (unsigned)n * (unsigned)sizeof(SAtom), full);
                    v += run;
                }
            }
            __syncwarp();
        }
    } else {
        // ================================ consumers =========================================================
        SmemAddr sa;
        sa.edge = sb + L.edge; sa.cnthr = sb + L.cnthr; sa.hist = sb + L.hist; sa.key = sb + L.key;
        int k = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
            const int buf = k % PIPE_STAGES;
            if (lane == 0) mbar_wait((unsigned)__cvta_generic_to_shared(&s_full[buf]), (unsigned)((k / PIPE_STAGES) & 1));
            __syncwarp();
            const PipeMeta &mt = s_meta[buf];
            const int z0 = mt.z0, rb = mt.rb, RR = mt.RR, V = mt.V, E = mt.E, nc2 = mt.nc2, m2 = mt.m2, items = mt.items;
            const unsigned rr_magic = mt.rr_magic;
            const int *s_off = reinterpret_cast<const int *>(smem_raw + L.off) + buf * TILE_OFF_WORDS;
            const int *s_rowimg = reinterpret_cast<const int *>(smem_raw + L.rowimg) + buf * TILE_MAX_ROWS;
            const double *s_rowT = reinterpret_cast<const double *>(smem_raw + L.rowT) + buf * (3 * TILE_MAX_ROWS);
            const SAtom *s_atoms = reinterpret_cast<const SAtom *>(smem_raw + L.atoms) + (size_t)buf * cap;
            sa.atoms = sb + L.atoms + (unsigned)buf * (unsigned)cap * (unsigned)sizeof(SAtom);
            sa.cn = sb + L.cn + (unsigned)buf * 4u * (unsigned)a.nkeys;

            int item = 0;
            if (lane == 0) item = atomicAdd(&s_next[buf], 1);
            item = __shfl_sync(0xffffffffu, item, 0);
            while (item < items) {
                int nxt = 0;
                if (lane == 0) nxt = atomicAdd(&s_next[buf], 1);      // in flight while this item is scanned
                const int hz = RR > 1 ? (int)__umulhi((unsigned)item, rr_magic) : item;
                const int rr = item - hz * RR, r = rb + rr;
                const int z = z0 + hz;
                const int hb = s_off[E + hz], nh = s_off[E + hz + 1] - hb;      // staged home cell
                if (nh > 0) {
                    const int img01 = s_rowimg[rr];
                    const double Rx = s_rowT[3 * rr], Ry = s_rowT[3 * rr + 1], Rz = s_rowT[3 * rr + 2];
                    const bool home_row = (r == 0);
                    const int own_off = home_row ? s_off[rr * V + hz + m2] : 0;   // position of the home cell inside row 0
                    for (int h0 = 0; h0 < nh; h0 += 32) {
                        const int ng = min(32, nh - h0);                 // home atoms in this group
                        const int G = c_sub_lanes[ng];
                        const unsigned g_magic = G == 1 ? 65536u : c_div_magic[G];
                        const int il = (int)(((unsigned)lane * g_magic) >> 16), sub = lane - il * G;
                        if (il < ng) {
                            const int hidx = hb + h0 + il;
                            const SAtom me = s_atoms[hidx];
                            const unsigned krow_addr = sa.key + 2u * (unsigned)((int)(me.s & 0xff) * S);
                            const int ism = own_off + h0 + il;
                            int d2 = home_row ? 0 : -m2;
                            while (d2 <= m2) {
                                int s2, q2;
                                wrap_cell(z + d2, nc2, s2, q2);
                                const int len = min(m2 - d2, nc2 - 1 - q2) + 1;
                                const int v = hz + m2 + d2;
                                const int jb = s_off[rr * V + v], je = s_off[rr * V + v + len];
                                const bool after_me = home_row && d2 == 0;   // own cell leads this run: partners after me only
                                if ((img01 | s2) != 0) {
                                    double Tx = Rx, Ty = Ry, Tz = Rz;
                                    if (s2 != 0) {
                                        const double fs2 = (double)s2;
                                        Tx = Rx + fs2 * mt.cell[6]; Ty = Ry + fs2 * mt.cell[7]; Tz = Rz + fs2 * mt.cell[8];
                                    }
                                    if (after_me) pipe_scan_run<HAS_CN, CN_WIDE, true, true>(a, sa, me.x, me.y, me.z, krow_addr, Tx, Ty, Tz, jb, je, G, sub, ism);
                                    else pipe_scan_run<HAS_CN, CN_WIDE, true, false>(a, sa, me.x, me.y, me.z, krow_addr, Tx, Ty, Tz, jb, je, G, sub, ism);
                                } else {
                                    if (after_me) pipe_scan_run<HAS_CN, CN_WIDE, false, true>(a, sa, me.x, me.y, me.z, krow_addr, 0.0, 0.0, 0.0, jb, je, G, sub, ism);
                                    else pipe_scan_run<HAS_CN, CN_WIDE, false, false>(a, sa, me.x, me.y, me.z, krow_addr, 0.0, 0.0, 0.0, jb, je, G, sub, ism);
                                }
                                d2 += len;
                            }
                        }
                    }
                }
                __syncwarp();
                item = __shfl_sync(0xffffffffu, nxt, 0);
            }
            // ---- release: the last warp out flushes the tile's coordination counters and frees the buffer ----
            __syncwarp();
            int last = 0;
            if (lane == 0) {
                __threadfence_block();
                last = (atomicAdd(&s_done[buf], 1) == PIPE_CONSUMERS - 1) ? 1 : 0;
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
                __threadfence_block();
                if (HAS_CN) {
                    uint32_t *s_cn = reinterpret_cast<uint32_t *>(smem_raw + L.cn) + buf * a.nkeys;
                    const int f = mt.frame;
                    for (int q = lane; q < a.nkeys; q += 32) {
                        const uint32_t v = s_cn[q];
                        if (v) {
                            atomicAdd(&a.cn_out[(size_t)f * a.nkeys + q], (unsigned long long)v);
                            s_cn[q] = 0u;
                        }
                    }
                }
                if (lane == 0) s_done[buf] = 0;
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    mbar_arrive((unsigned)__cvta_generic_to_shared(&s_empty[buf]));
                }
            }
        }
    }
    __syncthreads();
    const uint32_t *s_hist = reinterpret_cast<const uint32_t *>(smem_raw + L.hist);
    unsigned long long *slab = a.slabs + (size_t)blockIdx.x * a.nkeys * a.nbins;
    for (int q = threadIdx.x; q < a.nkeys * a.nbins; q += blockDim.x) {
        const uint32_t v = s_hist[q];
        if (v) slab[q] += v;
    }
}
