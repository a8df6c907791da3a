// pair_warp.cuh -- barrier-free variant of the RDF(+CN) pair kernel: every WARP streams its own candidates.
//
// k_pair_tiled (pair_tiled.cuh) stages a whole tile per block and pays for it with block-wide barriers (ncu: ~25 % of
// the warp time waits at __syncthreads) and a planner launch.  Here a warp owns one home cell at a time and walks its
// half stencil run by run: a run (the cells of one neighbour column that fall inside the stencil, contiguous in the
// cell-sorted frame) is fetched in chunks of WCHUNK atoms into a private double-buffered slice of shared memory --
// the loads of chunk c+1 are issued before chunk c is computed, so L2 latency hides behind the fp64 work -- and
// scanned with the same lane mapping (home atom x sub-lane), arithmetic and exact binning as the tiled kernel.
// No block barrier is executed between the histogram set-up and the final merge.
#pragma once
#include "pair_tiled.cuh"

#ifndef WARP_THREADS
#define WARP_THREADS 512
#endif
#define WCHUNK 48                      // atoms per chunk: 96 double2 = 3 loads per lane

struct WarpArgs {
    PairArgs p;
    long long total_cells;             // sum over the frames of the batch of ncell
};

// iterator over the chunks of one home cell's half stencil (all state warp-uniform)
struct ChunkIter {
    int r, d2;          // current row, next stencil offset along the fastest axis
    int pos, end;       // remaining atoms of the current run (indices into the sorted frame)
    int s0, s1, s2;     // image of the current run
    int rowbase;
    bool after;         // current run starts with the home cell itself
};

template <bool HAS_CN>
__global__ void __launch_bounds__(WARP_THREADS, 2) k_pair_warp(WarpArgs wa) {
    const PairArgs &a = wa.p;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: chunk buffers [nwarp][2][WCHUNK] (32 B) | edge2[nbins+1] | cn_thr2[nkeys] | hist u32 | cn_cnt[nwarp][nkeys] u32 | keyidx u16
    const int nwarp = WARP_THREADS / 32;
    SAtom *s_buf = reinterpret_cast<SAtom *>(smem_raw);
    double *s_edge2 = reinterpret_cast<double *>(s_buf + nwarp * 2 * WCHUNK);
    double *s_cnthr = s_edge2 + a.nbins + 1;
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(s_cnthr + (HAS_CN ? a.nkeys : 0));
    uint32_t *s_cn_all = s_hist + a.nkeys * a.nbins;
    uint16_t *s_key = reinterpret_cast<uint16_t *>(s_cn_all + (HAS_CN ? nwarp * a.nkeys : 0));

    const int S = a.n_species;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k <= a.nbins; k += blockDim.x) s_edge2[k] = a.edge2[k];
    for (int k = threadIdx.x; k < a.nkeys * a.nbins; k += blockDim.x) s_hist[k] = 0u;
    if (HAS_CN) {
        for (int k = threadIdx.x; k < a.nkeys; k += blockDim.x) s_cnthr[k] = a.cn_thr2[k];
        for (int k = threadIdx.x; k < nwarp * a.nkeys; k += blockDim.x) s_cn_all[k] = 0u;
    }
    for (int k = threadIdx.x; k < S * S; k += blockDim.x) s_key[k] = a.keyidx[k];
    __syncthreads();

    SmemAddr sa;
    {
        const unsigned sb = opaque_u32((unsigned)__cvta_generic_to_shared(smem_raw));
        sa.edge = sb + (unsigned)(reinterpret_cast<unsigned char *>(s_edge2) - smem_raw);
        sa.cnthr = sb + (unsigned)(reinterpret_cast<unsigned char *>(s_cnthr) - smem_raw);
        sa.hist = sb + (unsigned)(reinterpret_cast<unsigned char *>(s_hist) - smem_raw);
        sa.key = sb + (unsigned)(reinterpret_cast<unsigned char *>(s_key) - smem_raw);
        sa.atoms = 0u; sa.cn = 0u;      // per warp / per chunk, set below
    }
    SAtom *buf = s_buf + warp * 2 * WCHUNK;
    uint32_t *s_cn = s_cn_all + warp * a.nkeys;
    HitQueue hq;                                   // unused in the direct mode; keeps scan_run's signature
    hq.q = nullptr; hq.head = hq.tail = 0u;

    int f = 0;
    long long f_lo = 0, f_hi = a.n_frames > 0 ? (long long)a.geom[0].ncell : 0;     // items [f_lo, f_hi) belong to frame f
    const long long stride = (long long)gridDim.x * nwarp;
    for (long long item = (long long)blockIdx.x * nwarp + warp; item < wa.total_cells; item += stride) {
        while (item >= f_hi) { ++f; f_lo = f_hi; f_hi += (long long)a.geom[f].ncell; }
        const FrameGeom &G = a.geom[f];
        const uint32_t *cs = a.cell_start + G.cs_off;
        const int hcell = (int)(item - f_lo);
        const int hb = (int)cs[hcell], nh = (int)cs[hcell + 1] - hb;
        if (nh == 0) continue;
        const SAtom *fr = a.sorted + (long long)f * a.n_atoms;
        const int nc0 = G.nc[0], nc1 = G.nc[1], nc2 = G.nc[2];
        const int m0 = G.m[0], m1 = G.m[1], m2 = G.m[2];
        const int c2 = hcell % nc2, c01 = hcell / nc2, c1 = c01 % nc1, c0 = c01 / nc1;
        const int R = (m1 + 1) + m0 * (2 * m1 + 1);

        for (int h0 = 0; h0 < nh; h0 += 32) {
            const int ng = min(32, nh - h0);
            const int G_ = 32 / ng;
            const unsigned g_magic = (65536u + (unsigned)G_ - 1u) / (unsigned)G_;
            const int il = (int)(((unsigned)lane * g_magic) >> 16), sub = lane - il * G_;
            const bool active = il < ng;
            const int iabs = hb + h0 + (active ? il : 0);
            const SAtom me = load_satom(fr + iabs);

            // ---- chunk pipeline: descriptor + loads of the next chunk are issued before the current one is scanned
            ChunkIter it;
            it.r = -1; it.d2 = m2 + 1; it.pos = it.end = 0; it.s0 = it.s1 = it.s2 = 0; it.rowbase = 0; it.after = false;
            int cur = 0;
            // chunk descriptors (current / next)
            int c_jb = 0, c_n = 0, c_s0 = 0, c_s1 = 0, c_s2 = 0; bool c_after = false;
            int n_jb = 0, n_n = 0, n_s0 = 0, n_s1 = 0, n_s2 = 0; bool n_after = false;
            double2 pre[3];
            bool have_next;
            // advance the iterator to the next non-empty chunk; returns false when the stencil is exhausted
            auto next_chunk = [&]() -> bool {
                for (;;) {
                    if (it.pos < it.end) {
                        n_jb = it.pos; n_n = min(WCHUNK, it.end - it.pos); it.pos += n_n;
                        n_s0 = it.s0; n_s1 = it.s1; n_s2 = it.s2; n_after = it.after;
                        return true;
                    }
                    if (it.d2 > m2) {                       // next row
                        ++it.r;
                        if (it.r >= R) return false;
                        int d0, d1;
                        if (it.r <= m1) { d0 = 0; d1 = it.r; }
                        else { const int rr = it.r - (m1 + 1), w = 2 * m1 + 1; d0 = 1 + rr / w; d1 = rr - (d0 - 1) * w - m1; }
                        int q0, q1;
                        wrap_cell(c0 + d0, nc0, it.s0, q0);
                        wrap_cell(c1 + d1, nc1, it.s1, q1);
                        it.rowbase = (q0 * nc1 + q1) * nc2;
                        it.d2 = (it.r == 0) ? 0 : -m2;
                    }
                    // next segment of the row: cells up to the column wrap
                    int q2;
                    wrap_cell(c2 + it.d2, nc2, it.s2, q2);
                    const int len = min(m2 - it.d2, nc2 - 1 - q2) + 1;
                    it.pos = (int)cs[it.rowbase + q2];
                    it.end = (int)cs[it.rowbase + q2 + len];
                    it.after = (it.r == 0 && it.d2 == 0);
                    it.d2 += len;
                }
            };
            auto issue_loads = [&]() {
                const double2 *gp = reinterpret_cast<const double2 *>(fr + n_jb);
#pragma unroll
                for (int u = 0; u < 3; ++u)
                    if (lane + 32 * u < 2 * n_n) pre[u] = __ldg(gp + lane + 32 * u);
            };
            have_next = next_chunk();
            if (have_next) issue_loads();
            while (have_next) {
                // publish the prefetched chunk
                double2 *sp = reinterpret_cast<double2 *>(buf + cur * WCHUNK);
#pragma unroll
                for (int u = 0; u < 3; ++u)
                    if (lane + 32 * u < 2 * n_n) sp[lane + 32 * u] = pre[u];
                c_jb = n_jb; c_n = n_n; c_s0 = n_s0; c_s1 = n_s1; c_s2 = n_s2; c_after = n_after;
                __syncwarp();
                have_next = next_chunk();
                if (have_next) issue_loads();
                // scan the current chunk
                const SAtom *cb = buf + cur * WCHUNK;
                sa.atoms = (unsigned)__cvta_generic_to_shared(cb);
                sa.cn = (unsigned)__cvta_generic_to_shared(s_cn);
                const int G2 = G_;
                const int n_iter = (int)(((unsigned)(c_n + G2 - 1) * g_magic) >> 16);
                const int ism = iabs - c_jb;                   // chunk-relative index of "me" (meaningful when c_after)
                if ((c_s0 | c_s1 | c_s2) != 0) {
                    const double fs0 = (double)c_s0, fs1 = (double)c_s1, fs2 = (double)c_s2;
                    const double Tx = (fs0 * G.cell[0] + fs1 * G.cell[3]) + fs2 * G.cell[6];
                    const double Ty = (fs0 * G.cell[1] + fs1 * G.cell[4]) + fs2 * G.cell[7];
                    const double Tz = (fs0 * G.cell[2] + fs1 * G.cell[5]) + fs2 * G.cell[8];
                    if (c_after) scan_run<HAS_CN, true, true, true>(a, cb, s_edge2, s_cnthr, s_hist, s_cn, s_key, hq, me, Tx, Ty, Tz, 0, c_n, G2, n_iter, sub, active, ism, 0, lane, 0u, sa);
                    else scan_run<HAS_CN, true, true, false>(a, cb, s_edge2, s_cnthr, s_hist, s_cn, s_key, hq, me, Tx, Ty, Tz, 0, c_n, G2, n_iter, sub, active, ism, 0, lane, 0u, sa);
                } else {
                    if (c_after) scan_run<HAS_CN, true, false, true>(a, cb, s_edge2, s_cnthr, s_hist, s_cn, s_key, hq, me, 0.0, 0.0, 0.0, 0, c_n, G2, n_iter, sub, active, ism, 0, lane, 0u, sa);
                    else scan_run<HAS_CN, true, false, false>(a, cb, s_edge2, s_cnthr, s_hist, s_cn, s_key, hq, me, 0.0, 0.0, 0.0, 0, c_n, G2, n_iter, sub, active, ism, 0, lane, 0u, sa);
                }
                __syncwarp();
                cur ^= 1;
            }
        }
        if (HAS_CN) {
            __syncwarp();
            for (int k = lane; k < a.nkeys; k += 32) {
                const uint32_t v = s_cn[k];
                if (v) {
                    atomicAdd(&a.cn_out[(size_t)f * a.nkeys + k], (unsigned long long)v);
                    s_cn[k] = 0u;
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    unsigned long long *slab = a.slabs + (size_t)blockIdx.x * a.nkeys * a.nbins;
    for (int k = threadIdx.x; k < a.nkeys * a.nbins; k += blockDim.x) {
        const uint32_t v = s_hist[k];
        if (v) slab[k] += v;
    }
}
