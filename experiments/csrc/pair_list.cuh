// pair_list.cuh -- RDF(+CN) from a pair list with a skin, reused over a segment of consecutive frames
// (AMOFB_PAIR_LIST=1).
//
// Three quarters of the candidates k_pair_tiled evaluates are rejects (the half stencil of rc/2 cells has 3.7x the
// volume of the cutoff sphere).  Consecutive MD frames are correlated, so the unordered pairs (i, j, S) within
// rc + skin found at a REFERENCE frame R stay a complete candidate set for frame t as long as no atom moved by more
// than skin/2 since R.  A batch is cut into segments of seg_len frames, the first of each being its reference:
//
//   k_regroup      every frame t of a segment is rewritten in the cell order OF ITS REFERENCE (refsorted[t][slot_R(i)]):
//                  same tile layout for the whole segment, so the TMA staging and tile-local indices of R stay valid.
//                  The record carries m_i = round(frac(pw_i(t) - pw_i(R))), the lattice translation between the atom's
//                  wrapped position now and the image of it that continues its position at R, and the frame's largest
//                  minimum-image displacement is recorded (validity).
//   k_list_build   per tile of R: the tiled candidate scan with radius rc + skin; a hit appends one 28-bit entry
//                  {partner index, home index, shift id} to the tile's list.
//   k_list_scan    per (tile of R, frame t of its segment): stage the tile from refsorted[t], then full warps walk the
//                  list: dv = (pw_j - pw_i) + T(S_t), S_t = S_R - m_j + m_i.  That is the oracle's own expression (P3) on
//                  the oracle's own operands -- P2-wrapped positions of frame t and an integer image -- so d2 and
//                  therefore every count is bit-identical; only the enumeration differs.
//   completeness   a pair closer than rc at t was closer than rc + 2 max|disp| <= rc + skin at R, so it is listed.
//   fallback       frames that moved too far (or segments whose list overflowed / whose tiles do not fit) are masked in
//                  and go through k_pair_plan + k_pair_tiled; everything is decided on the device, no host sync.
// Only used when every frame of the batch has the same cell.
#pragma once
#include "pair_tiled.cuh"

#define LIST_MAX_SHIFT 64      // shift ids: row * 3 + (s2 + 1)

struct ListArgs {
    TiledArgs t;                  // t.tiles / t.n_tiles: the tiles of the reference frames
    const SAtom *refsorted;       // [F][N] frames in their reference's cell order; .s = species | (m0+8)<<8 | (m1+8)<<12 | (m2+8)<<16
    const int *ref_of;            // [F]
    const unsigned char *valid;   // [F] 1 = served by the list
    unsigned *entries;            // [max_tiles][list_cap]
    int *counts;                  // [max_tiles]
    int *flags;                   // [0] bit 0: reference tile list overflow, bit 1: a pair list overflowed, bit 2: unsupported tile; [1] largest list seen
    int list_cap, seg_len;
    double r2list;                // (rc + skin)^2, padded
};

struct RegroupArgs {
    const SAtom *sorted;          // [F][N] every frame in its own cell order
    const uint32_t *slot;         // [F][N] atom -> position in its frame's order
    const FrameGeom *geom;
    const int *ref_of;
    SAtom *refsorted;
    unsigned *maxdisp2;           // [F] float bits of the largest squared displacement (monotone for non-negative floats)
    int n_atoms, n_frames;
};

__global__ void __launch_bounds__(256) k_regroup(RegroupArgs a) {
    const long long total = (long long)a.n_frames * a.n_atoms;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(idx / a.n_atoms), i = (int)(idx - (long long)t * a.n_atoms);
        const int R = a.ref_of[t];
        const uint32_t posT = a.slot[idx], posR = a.slot[(long long)R * a.n_atoms + i];
        const SAtom rt = load_satom(a.sorted + (long long)t * a.n_atoms + posT);
        SAtom out = rt;
        int m0 = 0, m1 = 0, m2 = 0;
        float d2f = 0.f;
        if (t != R) {
            const SAtom rr = load_satom(a.sorted + (long long)R * a.n_atoms + posR);
            const FrameGeom &g = a.geom[t];
            const double dx = rt.x - rr.x, dy = rt.y - rr.y, dz = rt.z - rr.z;
            const double f0 = (dx * g.inv[0] + dy * g.inv[3]) + dz * g.inv[6];
            const double f1 = (dx * g.inv[1] + dy * g.inv[4]) + dz * g.inv[7];
            const double f2 = (dx * g.inv[2] + dy * g.inv[5]) + dz * g.inv[8];
            const double r0 = rint(f0), r1 = rint(f1), r2 = rint(f2);
            m0 = (int)r0; m1 = (int)r1; m2 = (int)r2;
            const double ex = dx - ((r0 * g.cell[0] + r1 * g.cell[3]) + r2 * g.cell[6]);
            const double ey = dy - ((r0 * g.cell[1] + r1 * g.cell[4]) + r2 * g.cell[7]);
            const double ez = dz - ((r0 * g.cell[2] + r1 * g.cell[5]) + r2 * g.cell[8]);
            d2f = __double2float_ru((ex * ex + ey * ey) + ez * ez);
            if (m0 < -7 || m0 > 7 || m1 < -7 || m1 > 7 || m2 < -7 || m2 > 7 || !(d2f >= 0.f)) d2f = 3.0e38f;   // not representable: frame falls back
        }
        out.s = (rt.s & 0xff) | ((long long)(m0 + 8) << 8) | ((long long)(m1 + 8) << 12) | ((long long)(m2 + 8) << 16);
        a.refsorted[(long long)t * a.n_atoms + posR] = out;
        // one atomic per warp and frame: the lanes of a warp almost always share the frame
        const unsigned act = __activemask();
        bool warp_path = false;
        if (act == 0xffffffffu) {
            const int t0 = __shfl_sync(0xffffffffu, t, 0);
            warp_path = __all_sync(0xffffffffu, t == t0);
        }
        if (warp_path) {
            unsigned v = __float_as_uint(d2f);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
            if ((threadIdx.x & 31) == 0 && v > 0u) atomicMax(&a.maxdisp2[t], v);
        } else if (d2f > 0.f) atomicMax(&a.maxdisp2[t], __float_as_uint(d2f));
    }
}

// valid[t] = 1 iff frame t may go through its reference's list
__global__ void k_list_mask(const unsigned *maxdisp2, const int *flags, const int *ntiles_ref, float lim2, int n_frames, unsigned char *valid) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_frames) return;
    const bool bad = flags[0] != 0 || ntiles_ref[1] != 0;           // overflow / unsupported tile / cells the planner could not tile
    valid[t] = (!bad && __uint_as_float(maxdisp2[t]) <= lim2) ? 1 : 0;
}

// ---- staging shared by the two list kernels: the tile of reference frame geometry `s_geom`/`cs`, atoms read from `fr` ----
struct ListTile {
    int c0, c1, z0, zlen, rb, RR, V, E, EH, nc2, m2;
};

__device__ __forceinline__ void list_stage(const ListTile &lt, const FrameGeom &s_geom, const uint32_t *__restrict__ cs, const SAtom *__restrict__ fr,
                                           SAtom *s_atoms, int *s_off, int *s_rowimg, unsigned mbar, unsigned &tma_phase) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nc0 = s_geom.nc[0], nc1 = s_geom.nc[1], nc2 = lt.nc2, m2 = lt.m2;
    const int homebase = (lt.c0 * nc1 + lt.c1) * nc2;
    for (int t = threadIdx.x; t < lt.RR; t += blockDim.x) {
        int d0, d1, s0_, s1_, q0_, q1_;
        tile_row_offset(s_geom, lt.rb + t, d0, d1);
        wrap_cell(lt.c0 + d0, nc0, s0_, q0_);
        wrap_cell(lt.c1 + d1, nc1, s1_, q1_);
        s_rowimg[t] = (s0_ & 0xffff) | (s1_ << 16);
    }
    for (int e = threadIdx.x; e < lt.EH; e += blockDim.x) {
        int cell;
        if (e < lt.E) {
            const int r = lt.rb + e / lt.V, v = e - (e / lt.V) * lt.V;
            int d0, d1;
            tile_row_offset(s_geom, r, d0, d1);
            const int t0 = lt.c0 + d0, t1 = lt.c1 + d1, t2 = lt.z0 - m2 + v;
            const int q0 = t0 - floordiv_i(t0, nc0) * nc0, q1 = t1 - floordiv_i(t1, nc1) * nc1, q2 = t2 - floordiv_i(t2, nc2) * nc2;
            cell = (q0 * nc1 + q1) * nc2 + q2;
        } else cell = homebase + lt.z0 + (e - lt.E);
        s_off[e + 1] = (int)(cs[cell + 1] - cs[cell]);
    }
    __syncthreads();
    if (warp == 0) {
        int carry = 0;
        for (int e0 = 0; e0 < lt.EH; e0 += 32) {
            const int e = e0 + lane;
            const int cnt = e < lt.EH ? s_off[e + 1] : 0;
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            __syncwarp();
            if (e < lt.EH) s_off[e + 1] = carry + incl;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) s_off[0] = 0;
    }
    __syncthreads();
    if (warp == 0) {
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive_expect_tx(mbar, (unsigned)s_off[lt.EH] * (unsigned)sizeof(SAtom));
        }
        __syncwarp();
        const unsigned abase = (unsigned)__cvta_generic_to_shared(s_atoms);
        for (int task = lane; task < lt.RR + 1; task += 32) {
            int colbase, va, vb, ebase;
            if (task < lt.RR) {
                int d0, d1, s0_, s1_, q0, q1;
                tile_row_offset(s_geom, lt.rb + task, d0, d1);
                wrap_cell(lt.c0 + d0, nc0, s0_, q0);
                wrap_cell(lt.c1 + d1, nc1, s1_, q1);
                colbase = (q0 * nc1 + q1) * nc2; va = lt.z0 - m2; vb = lt.z0 + lt.zlen + m2; ebase = task * lt.V;
            } else { colbase = homebase; va = lt.z0; vb = lt.z0 + lt.zlen; ebase = lt.E; }
            int v = va;
            while (v < vb) {
                int sdum, q;
                wrap_cell(v, nc2, sdum, q);
                const int run = min(vb - v, nc2 - q);
                const int src = (int)cs[colbase + q], n = (int)cs[colbase + q + run] - src;
                if (n > 0) bulk_g2s(abase + (unsigned)s_off[ebase + (v - va)] * (unsigned)sizeof(SAtom), fr + src, (unsigned)n * (unsigned)sizeof(SAtom), mbar);
                v += run;
            }
        }
    }
    if (threadIdx.x == 0) mbar_wait(mbar, tma_phase);
    tma_phase ^= 1u;
    __syncthreads();
}

// smem carve-up of the list kernels: the same buffer sizes as k_pair_tiled (host passes ta.cap and the same byte count)
struct ListSmem {
    SAtom *atoms; double *edge2, *cnthr; uint32_t *hist, *cn; int *off; double *ttab; uint16_t *key;
};
__device__ __forceinline__ ListSmem list_carve(unsigned char *smem_raw, const TiledArgs &ta, bool has_cn) {
    const PairArgs &a = ta.p;
    ListSmem m;
    size_t off = 0;
    m.atoms = reinterpret_cast<SAtom *>(smem_raw + off);      off += sizeof(SAtom) * (size_t)ta.cap;
    m.edge2 = reinterpret_cast<double *>(smem_raw + off);     off += sizeof(double) * (size_t)(a.nbins + 1);
    m.cnthr = reinterpret_cast<double *>(smem_raw + off);     off += sizeof(double) * (size_t)(has_cn ? a.nkeys : 0);
    off = (off + 15) & ~(size_t)15;
    m.hist = reinterpret_cast<uint32_t *>(smem_raw + off);    off += sizeof(uint32_t) * (size_t)a.nkeys * a.nbins;
    m.cn = reinterpret_cast<uint32_t *>(smem_raw + off);      off += sizeof(uint32_t) * (size_t)(has_cn ? a.nkeys : 0);
    m.off = reinterpret_cast<int *>(smem_raw + off);          off += sizeof(int) * TILE_OFF_WORDS;
    off = (off + 15) & ~(size_t)15;
    m.ttab = reinterpret_cast<double *>(smem_raw + off);      off += sizeof(FlatRun) * (FLAT_MAXE + 1) * (TILE_THREADS / 32);   // 4.5 KB >= 64 * 32 B
    m.key = reinterpret_cast<uint16_t *>(smem_raw + off);
    return m;
}

__device__ __forceinline__ bool list_tile_supported(const PairTile &t, const FrameGeom &g) {
    const int R = tile_rows(g);
    return t.rb == 0 && t.re == R && R * 3 <= LIST_MAX_SHIFT && g.nc[2] >= 2 * g.m[2] + 1;
}

// ---- list build: one pass over the tiles of the reference frames ----------------------------------------------------
__global__ void __launch_bounds__(TILE_THREADS, TILE_MIN_BLOCKS) k_list_build(ListArgs la) {
    const TiledArgs &ta = la.t;
    const PairArgs &a = ta.p;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const ListSmem sm = list_carve(smem_raw, ta, true);
    __shared__ FrameGeom s_geom;
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ int s_rowimg[TILE_MAX_ROWS];
    __shared__ unsigned s_count;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = TILE_THREADS / 32;
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(&s_mbar);
    if (threadIdx.x == 0) mbar_init(mbar, 1);
    unsigned tma_phase = 0;
    const int n_tiles = min(*ta.n_tiles, ta.max_tiles);
    const unsigned abase = (unsigned)__cvta_generic_to_shared(sm.atoms);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();
        const PairTile tl = ta.tiles[tile];
        const int f = tl.frame;
        if (threadIdx.x < (int)(sizeof(FrameGeom) / sizeof(int)))
            reinterpret_cast<int *>(&s_geom)[threadIdx.x] = reinterpret_cast<const int *>(&a.geom[f])[threadIdx.x];
        if (threadIdx.x == 0) s_count = 0u;
        __syncthreads();
        if (!list_tile_supported(tl, s_geom)) {                 // uniform: row-split tile, too many rows, box narrower than the stencil
            if (threadIdx.x == 0) { atomicOr(la.flags, 4); la.counts[tile] = 0; }
            continue;
        }
        ListTile lt;
        lt.c0 = tl.c0; lt.c1 = tl.c1; lt.z0 = tl.z0; lt.zlen = tl.zlen; lt.rb = 0; lt.RR = tl.re;
        lt.nc2 = s_geom.nc[2]; lt.m2 = s_geom.m[2];
        lt.V = lt.zlen + 2 * lt.m2; lt.E = lt.RR * lt.V; lt.EH = lt.E + lt.zlen;
        list_stage(lt, s_geom, a.cell_start + s_geom.cs_off, a.sorted + (long long)f * a.n_atoms, sm.atoms, sm.off, s_rowimg, mbar, tma_phase);
        const int *s_off = sm.off;
        unsigned *out = la.entries + (size_t)tile * la.list_cap;
        const int items = lt.zlen * lt.RR;
        for (int item = warp; item < items; item += nwarp) {
            const int hz = item / lt.RR, rr = item - hz * lt.RR;
            const int hb = s_off[lt.E + hz], nh = s_off[lt.E + hz + 1] - hb;
            if (nh == 0) continue;
            const int img01 = s_rowimg[rr];
            const int s0 = (int)(short)(img01 & 0xffff), s1 = img01 >> 16;
            const bool home_row = (rr == 0);
            const int own_off = home_row ? s_off[rr * lt.V + hz + lt.m2] : 0;
            for (int h0 = 0; h0 < nh; h0 += 32) {
                const int ng = min(32, nh - h0);
                const int G = c_sub_lanes[ng];
                const unsigned g_magic = G == 1 ? 65536u : c_div_magic[G];
                const int il = (int)(((unsigned)lane * g_magic) >> 16), sub = lane - il * G;
                if (il >= ng) continue;
                const int hidx = hb + h0 + il;
                const SAtom me = sm.atoms[hidx];
                int d2 = home_row ? 0 : -lt.m2;
                while (d2 <= lt.m2) {
                    int s2, q2;
                    wrap_cell(lt.z0 + hz + d2, lt.nc2, s2, q2);
                    const int len = min(lt.m2 - d2, lt.nc2 - 1 - q2) + 1;
                    const int v = hz + lt.m2 + d2;
                    const int jb = s_off[rr * lt.V + v], je = s_off[rr * lt.V + v + len];
                    const bool after_me = home_row && d2 == 0;
                    const int jskip = after_me ? own_off + h0 + il : -1;
                    const double fs0 = (double)s0, fs1 = (double)s1, fs2 = (double)s2;
                    const double Tx = (fs0 * s_geom.cell[0] + fs1 * s_geom.cell[3]) + fs2 * s_geom.cell[6];
                    const double Ty = (fs0 * s_geom.cell[1] + fs1 * s_geom.cell[4]) + fs2 * s_geom.cell[7];
                    const double Tz = (fs0 * s_geom.cell[2] + fs1 * s_geom.cell[5]) + fs2 * s_geom.cell[8];
                    const unsigned sid = (unsigned)(rr * 3 + (s2 + 1));
                    for (int j = jb + sub; j < je; j += G) {
                        double ox, oy, oz;
                        lds_xyz(abase + (unsigned)j * 32u, ox, oy, oz);
                        const double dx = (ox - me.x) + Tx, dy = (oy - me.y) + Ty, dz = (oz - me.z) + Tz;
                        const double dd = (dx * dx + dy * dy) + dz * dz;
                        if (dd < la.r2list && j > jskip) {
                            // one shared-memory atomic per group of hitting lanes, not per hit
                            const unsigned act = __activemask();
                            const int leader = __ffs(act) - 1;
                            unsigned base = 0;
                            if (lane == leader) base = atomicAdd(&s_count, (unsigned)__popc(act));
                            base = __shfl_sync(act, base, leader);
                            const unsigned pos = base + (unsigned)__popc(act & ((1u << lane) - 1u));
                            if (pos < (unsigned)la.list_cap) out[pos] = (unsigned)j | ((unsigned)hidx << 11) | (sid << 22);
                        }
                    }
                    d2 += len;
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            la.counts[tile] = (int)min(s_count, (unsigned)la.list_cap);
            if (s_count > (unsigned)la.list_cap) { atomicOr(la.flags, 2); atomicMax(la.flags + 1, (int)s_count); }
        }
    }
}

// ---- list scan ----------------------------------------------------------------------------------------------------
template <bool HAS_CN, bool CN_WIDE>
__global__ void __launch_bounds__(TILE_THREADS, TILE_MIN_BLOCKS) k_list_scan(ListArgs la) {
    const TiledArgs &ta = la.t;
    const PairArgs &a = ta.p;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const ListSmem sm = list_carve(smem_raw, ta, HAS_CN);
    __shared__ FrameGeom s_geom;
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ int s_rowimg[TILE_MAX_ROWS];
    const int S = a.n_species;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = TILE_THREADS / 32;
    for (int k = threadIdx.x; k <= a.nbins; k += blockDim.x) sm.edge2[k] = a.edge2[k];
    for (int k = threadIdx.x; k < a.nkeys * a.nbins; k += blockDim.x) sm.hist[k] = 0u;
    if (HAS_CN)
        for (int k = threadIdx.x; k < a.nkeys; k += blockDim.x) { sm.cnthr[k] = a.cn_thr2[k]; sm.cn[k] = 0u; }
    for (int k = threadIdx.x; k < S * S; k += blockDim.x) sm.key[k] = a.keyidx[k];
    SmemAddr sa;
    {
        const unsigned sb = opaque_u32((unsigned)__cvta_generic_to_shared(smem_raw));
        sa.atoms = sb + (unsigned)(reinterpret_cast<unsigned char *>(sm.atoms) - smem_raw);
        sa.edge = sb + (unsigned)(reinterpret_cast<unsigned char *>(sm.edge2) - smem_raw);
        sa.cnthr = sb + (unsigned)(reinterpret_cast<unsigned char *>(sm.cnthr) - smem_raw);
        sa.hist = sb + (unsigned)(reinterpret_cast<unsigned char *>(sm.hist) - smem_raw);
        sa.cn = sb + (unsigned)(reinterpret_cast<unsigned char *>(sm.cn) - smem_raw);
        sa.key = sb + (unsigned)(reinterpret_cast<unsigned char *>(sm.key) - smem_raw);
    }
    const unsigned ttab_addr = sa.atoms - (unsigned)(reinterpret_cast<unsigned char *>(sm.atoms) - smem_raw)
                               + (unsigned)(reinterpret_cast<unsigned char *>(sm.ttab) - smem_raw);
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(&s_mbar);
    if (threadIdx.x == 0) mbar_init(mbar, 1);
    unsigned tma_phase = 0;
    const double r2search = a.r2search, r2max = a.r2max, cn_r2max = a.cn_r2max;
    const float inv_dr_f = a.inv_dr_f, margin = a.bin_margin;
    const int nbins = a.nbins;
    const int n_tiles = min(*ta.n_tiles, ta.max_tiles);
    const long long total = (long long)n_tiles * la.seg_len;
    for (long long w = blockIdx.x; w < total; w += gridDim.x) {
        const int q = (int)(w / la.seg_len), u = (int)(w - (long long)q * la.seg_len);
        const PairTile tl = ta.tiles[q];
        const int R = tl.frame, t = R + u;
        if (t >= a.n_frames || la.ref_of[t] != R || !la.valid[t]) continue;        // uniform over the block
        const int count = la.counts[q];
        __syncthreads();     // previous work item fully consumed
        if (threadIdx.x < (int)(sizeof(FrameGeom) / sizeof(int)))
            reinterpret_cast<int *>(&s_geom)[threadIdx.x] = reinterpret_cast<const int *>(&a.geom[R])[threadIdx.x];
        __syncthreads();
        ListTile lt;
        lt.c0 = tl.c0; lt.c1 = tl.c1; lt.z0 = tl.z0; lt.zlen = tl.zlen; lt.rb = 0; lt.RR = tl.re;
        lt.nc2 = s_geom.nc[2]; lt.m2 = s_geom.m[2];
        lt.V = lt.zlen + 2 * lt.m2; lt.E = lt.RR * lt.V; lt.EH = lt.E + lt.zlen;
        // image shifts by id (row, s2): the staging's barriers publish them together with the atoms
        for (int k = threadIdx.x; k < lt.RR * 3; k += blockDim.x) {
            const int rr = k / 3, s2 = k - rr * 3 - 1;
            int d0, d1, s0_, s1_, q0_, q1_;
            tile_row_offset(s_geom, rr, d0, d1);
            wrap_cell(lt.c0 + d0, s_geom.nc[0], s0_, q0_);
            wrap_cell(lt.c1 + d1, s_geom.nc[1], s1_, q1_);
            const double fs0 = (double)s0_, fs1 = (double)s1_, fs2 = (double)s2;
            sm.ttab[4 * k] = (fs0 * s_geom.cell[0] + fs1 * s_geom.cell[3]) + fs2 * s_geom.cell[6];
            sm.ttab[4 * k + 1] = (fs0 * s_geom.cell[1] + fs1 * s_geom.cell[4]) + fs2 * s_geom.cell[7];
            sm.ttab[4 * k + 2] = (fs0 * s_geom.cell[2] + fs1 * s_geom.cell[5]) + fs2 * s_geom.cell[8];
            sm.ttab[4 * k + 3] = 0.0;
        }
        list_stage(lt, s_geom, a.cell_start + s_geom.cs_off, la.refsorted + (long long)t * a.n_atoms, sm.atoms, sm.off, s_rowimg, mbar, tma_phase);
        const unsigned *ent = la.entries + (size_t)q * la.list_cap;
        // the entry of the next trip is fetched (from L2) before this trip's pair is evaluated
        unsigned en_next = (warp * 32 + lane) < count ? __ldg(ent + warp * 32 + lane) : 0u;
        for (int e0 = warp * 32; e0 < count; e0 += nwarp * 32) {
            const int e = e0 + lane;
            const unsigned en = en_next;
            const int e2 = e + nwarp * 32;
            en_next = e2 < count ? __ldg(ent + e2) : 0u;
            if (e >= count) continue;
            const unsigned aj = sa.atoms + (en & 2047u) * 32u, ai = sa.atoms + ((en >> 11) & 2047u) * 32u;
            const unsigned sid = en >> 22;
            double xj, yj, zj, xi, yi, zi, sjd, sid_d;
            asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(xj), "=d"(yj) : "r"(aj));
            asm("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(zj), "=d"(sjd) : "r"(aj));
            asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(xi), "=d"(yi) : "r"(ai));
            asm("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(zi), "=d"(sid_d) : "r"(ai));
            const unsigned wj = (unsigned)__double_as_longlong(sjd), wi = (unsigned)__double_as_longlong(sid_d);
            double Tx, Ty, Tz;
            if ((wj >> 8) == (wi >> 8)) {                      // both atoms moved by the same lattice translation since R (almost always 0)
                double pad;
                asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(Tx), "=d"(Ty) : "r"(ttab_addr + sid * 32u));
                asm("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(Tz), "=d"(pad) : "r"(ttab_addr + sid * 32u));
            } else {
                // S_t = S_R - m_j + m_i, then T as P3 forms it
                const int rr = (int)(sid / 3u);
                const int img01 = s_rowimg[rr];
                const int s0 = (int)(short)(img01 & 0xffff) - (int)((wj >> 8) & 15u) + (int)((wi >> 8) & 15u);
                const int s1 = (img01 >> 16) - (int)((wj >> 12) & 15u) + (int)((wi >> 12) & 15u);
                const int s2 = (int)(sid - 3u * (unsigned)rr) - 1 - (int)((wj >> 16) & 15u) + (int)((wi >> 16) & 15u);
                const double fs0 = (double)s0, fs1 = (double)s1, fs2 = (double)s2;
                Tx = (fs0 * s_geom.cell[0] + fs1 * s_geom.cell[3]) + fs2 * s_geom.cell[6];
                Ty = (fs0 * s_geom.cell[1] + fs1 * s_geom.cell[4]) + fs2 * s_geom.cell[7];
                Tz = (fs0 * s_geom.cell[2] + fs1 * s_geom.cell[5]) + fs2 * s_geom.cell[8];
            }
            const double dx = (xj - xi) + Tx, dy = (yj - yi) + Ty, dz = (zj - zi) + Tz;
            const double dd = (dx * dx + dy * dy) + dz * dz;
            if (dd < r2search) {
                const int key = lds_u16(sa.key + 2u * ((wi & 0xffu) * (unsigned)S + (wj & 0xffu)));
                if (!CN_WIDE || dd < r2max) {
                    const int b = rdf_bin_s(dd, sa.edge, inv_dr_f, margin);
                    reds_inc(sa.hist + 4u * (unsigned)(key * nbins + b));
                }
                if (HAS_CN && dd < cn_r2max && dd < lds_f64(sa.cnthr + 8u * (unsigned)key)) reds_inc(sa.cn + 4u * (unsigned)key);
            }
        }
        if (HAS_CN) {
            __syncthreads();
            for (int k = threadIdx.x; k < a.nkeys; k += blockDim.x) {
                const uint32_t v = sm.cn[k];
                if (v) {
                    atomicAdd(&a.cn_out[(size_t)t * a.nkeys + k], (unsigned long long)v);
                    sm.cn[k] = 0u;
                }
            }
        }
    }
    __syncthreads();
    unsigned long long *slab = a.slabs + (size_t)blockIdx.x * a.nkeys * a.nbins;
    for (int k = threadIdx.x; k < a.nkeys * a.nbins; k += blockDim.x) {
        const uint32_t v = sm.hist[k];
        if (v) slab[k] += v;
    }
}
