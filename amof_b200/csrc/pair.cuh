// pair.cuh -- the pair kernel (K2 + K3 of SURVEY.md 2.1): partial-RDF histograms and cutoff neighbour counts.
//
// One thread owns one home atom of the cell-sorted frame and walks the HALF stencil of linked cells around it
// (every unordered pair {(i,j,S),(j,i,-S)} is visited exactly once; P3 of the oracle makes the two orientations
// bit-identical, so directed counts are recovered exactly at finish time as U[a][b] (a != b) or 2*U[a][a]).
// Along the fastest cell axis the stencil cells of one row are contiguous in the sorted order, so a row is one
// or two index ranges, not 2m+1 cell visits.
//
// Bin assignment is exact without fp64 sqrt or divide: the host bisects, once per analysis, the smallest d2 for
// which (int)(sqrt(d2)/dr) reaches each bin (edge2[], P4) and the smallest d2 with sqrt(d2) >= cutoff (cn_thr2[],
// P5); both are monotone in d2 because IEEE sqrt and divide are.  The kernel guesses the bin in fp32 and corrects
// it against edge2[] in shared memory.  d2 itself is computed in fp64 exactly as P3 orders it (-fmad=false).
//
// Histograms are privatised per block in shared memory (u32, species pairs folded to a <= b), merged once per
// launch into that block's own u64 slab in global memory without atomics; slabs are summed at finish.
#pragma once
#include "prep.cuh"

#define PAIR_TILE 256

struct PairArgs {
    const SAtom *sorted;          // [F*N]
    const FrameGeom *geom;        // [F]
    const uint32_t *cell_start;   // batch-wide
    const double *edge2;          // [nbins+1]
    const double *cn_thr2;        // [nkeys]   (0 = pair not listed)
    const uint16_t *keyidx;       // [S*S] -> folded key
    unsigned long long *slabs;    // [gridDim.x][nkeys*nbins]   (smem-histogram mode)
    unsigned long long *ghist;    // [nkeys*nbins]              (global-atomic mode)
    unsigned long long *cn_out;   // [F][nkeys]
    const uint8_t *hard_mask;     // optional: batch-wide per-cell mask, only atoms of marked home cells are processed
    const int *n_hard;            // optional: number of marked cells (0 -> the launch returns at once)
    double r2search;              // a pair matters iff d2 < r2search
    double r2max;                 // = edge2[nbins]
    double cn_r2max;              // max of cn_thr2: cheap pre-test before the per-pair threshold
    float inv_dr_f;
    float bin_margin;             // see rdf_bin
    int n_atoms, n_frames, n_species, nkeys, nbins;
    int tiles_per_frame;
    int n_sorted;                 // atoms per frame in the cell list (< n_atoms when the species filter of a CN-only analysis is on)
};

__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));   // one MUFU.SQRT, relative error <= 2^-23
    return r;
}

// exact P4 bin of d2 (precondition: d2 < edge2[nbins]).
// The fp32 estimate r*inv_dr is within nbins*2^-21 of the true quotient; subtracting `margin` (> that bound, < 1/2,
// checked on the host) makes floor() land on the true bin or the one below it, so ONE comparison against the exact
// threshold of the next bin settles it.  Falls back to a search when the bin count is too large for that argument.
__device__ __forceinline__ int rdf_bin(double d2, const double *__restrict__ edge2, float inv_dr_f, float margin, int nbins) {
    if (margin > 0.f) {
        // truncation of a value > -1 -> max(floor, 0); d2 < edge2[nbins] and margin > the estimate's error keep it <= nbins-1
        const int b = (int)fmaf(sqrt_approx((float)d2), inv_dr_f, -margin);
        return b + (d2 >= edge2[b + 1] ? 1 : 0);                        // edge2[nbins] > d2: never overshoots
    }
    int b = (int)(sqrt_approx((float)d2) * inv_dr_f);
    b = b > nbins - 1 ? nbins - 1 : b;
    while (d2 < edge2[b]) --b;          // edge2[0] == 0 stops it
    while (d2 >= edge2[b + 1]) ++b;     // edge2[nbins] > d2 stops it
    return b;
}

template <bool HAS_RDF, bool HAS_CN, bool SMEM_HIST>
__global__ void __launch_bounds__(PAIR_TILE, 2) k_pair(PairArgs a) {
    if (a.n_hard && *a.n_hard == 0) return;   // clean-up launch behind the tiled kernel with nothing to clean up
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: edge2[nbins+1] (f64) | cn_thr2[nkeys] (f64) | hist[nkeys*nbins] (u32) | cn_cnt[nkeys] (u32) | keyidx[S*S] (u16)
    double *s_edge2 = reinterpret_cast<double *>(smem_raw);
    double *s_cnthr = s_edge2 + (SMEM_HIST && HAS_RDF ? a.nbins + 1 : 0);
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(s_cnthr + (HAS_CN ? a.nkeys : 0));
    uint32_t *s_cn = s_hist + (SMEM_HIST && HAS_RDF ? a.nkeys * a.nbins : 0);
    uint16_t *s_key = reinterpret_cast<uint16_t *>(s_cn + (HAS_CN ? a.nkeys : 0));
    __shared__ FrameGeom s_geom;

    const int S = a.n_species;
    if (SMEM_HIST && HAS_RDF) {
        for (int k = threadIdx.x; k <= a.nbins; k += blockDim.x) s_edge2[k] = a.edge2[k];
        for (int k = threadIdx.x; k < a.nkeys * a.nbins; k += blockDim.x) s_hist[k] = 0u;
    }
    if (HAS_CN)
        for (int k = threadIdx.x; k < a.nkeys; k += blockDim.x) { s_cnthr[k] = a.cn_thr2[k]; s_cn[k] = 0u; }
    for (int k = threadIdx.x; k < S * S; k += blockDim.x) s_key[k] = a.keyidx[k];
    const double *edge2 = (SMEM_HIST && HAS_RDF) ? s_edge2 : a.edge2;

    const long long total_tiles = (long long)a.n_frames * a.tiles_per_frame;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int f = (int)(tile / a.tiles_per_frame);
        const int tin = (int)(tile - (long long)f * a.tiles_per_frame);
        __syncthreads();   // previous tile's cn flush and geometry reads are done
        if (threadIdx.x < (int)(sizeof(FrameGeom) / sizeof(int)))
            reinterpret_cast<int *>(&s_geom)[threadIdx.x] = reinterpret_cast<const int *>(&a.geom[f])[threadIdx.x];
        __syncthreads();
        const int i = tin * PAIR_TILE + threadIdx.x;   // sorted index inside the frame
        if (i < a.n_sorted) {
            const SAtom *fr = a.sorted + (long long)f * a.n_atoms;
            const uint32_t *cs = a.cell_start + s_geom.cs_off;
            const SAtom me = load_satom(fr + i);
            const int si = (int)(me.s & 0xff);
            const int c0 = (int)((me.s >> 8) & 0xfff), c1 = (int)((me.s >> 20) & 0xfff), c2 = (int)((me.s >> 32) & 0xfff);
            const int nc0 = s_geom.nc[0], nc1 = s_geom.nc[1], nc2 = s_geom.nc[2];
            const int m0 = (a.hard_mask && !a.hard_mask[s_geom.cs_off + (c0 * nc1 + c1) * nc2 + c2]) ? -1 : s_geom.m[0];
            const int m1 = s_geom.m[1], m2 = s_geom.m[2];
            const uint16_t *krow = s_key + si * S;
            for (int d0 = 0; d0 <= m0; ++d0) {
                const int t0 = c0 + d0, s0 = floordiv_i(t0, nc0), q0 = t0 - s0 * nc0;
                for (int d1 = (d0 == 0 ? 0 : -m1); d1 <= m1; ++d1) {
                    const int t1 = c1 + d1, s1 = floordiv_i(t1, nc1), q1 = t1 - s1 * nc1;
                    const bool home_row = (d0 == 0 && d1 == 0);
                    int d2 = home_row ? 0 : -m2;
                    const int rowbase = (q0 * nc1 + q1) * nc2;
                    while (d2 <= m2) {
                        const int t2 = c2 + d2, s2 = floordiv_i(t2, nc2), q2 = t2 - s2 * nc2;
                        const int len = min(m2 - d2, nc2 - 1 - q2) + 1;
                        int jb = (int)cs[rowbase + q2];
                        const int je = (int)cs[rowbase + q2 + len];
                        if (home_row && d2 == 0) jb = i + 1;   // own cell: partners after me; following cells whole
                        // P3 image shift
                        const double fs0 = (double)s0, fs1 = (double)s1, fs2 = (double)s2;
                        const double Tx = (fs0 * s_geom.cell[0] + fs1 * s_geom.cell[3]) + fs2 * s_geom.cell[6];
                        const double Ty = (fs0 * s_geom.cell[1] + fs1 * s_geom.cell[4]) + fs2 * s_geom.cell[7];
                        const double Tz = (fs0 * s_geom.cell[2] + fs1 * s_geom.cell[5]) + fs2 * s_geom.cell[8];
#pragma unroll 2
                        for (int j = jb; j < je; ++j) {
                            const SAtom o = load_satom(fr + j);
                            const double dx = (o.x - me.x) + Tx;
                            const double dy = (o.y - me.y) + Ty;
                            const double dz = (o.z - me.z) + Tz;
                            const double dd = (dx * dx + dy * dy) + dz * dz;
                            if (dd < a.r2search) {
                                const int key = krow[(int)(o.s & 0xff)];
                                if (HAS_RDF && dd < a.r2max) {
                                    const int b = rdf_bin(dd, edge2, a.inv_dr_f, a.bin_margin, a.nbins);
                                    if (SMEM_HIST) atomicAdd(&s_hist[key * a.nbins + b], 1u);
                                    else atomicAdd(&a.ghist[(size_t)key * a.nbins + b], 1ull);
                                }
                                if (HAS_CN && dd < a.cn_r2max && dd < s_cnthr[key]) atomicAdd(&s_cn[key], 1u);
                            }
                        }
                        d2 += len;
                    }
                }
            }
        }
        if (HAS_CN) {
            __syncthreads();
            for (int k = threadIdx.x; k < a.nkeys; k += blockDim.x) {
                uint32_t v = s_cn[k];
                if (v) {
                    atomicAdd(&a.cn_out[(size_t)f * a.nkeys + k], (unsigned long long)v);
                    s_cn[k] = 0u;
                }
            }
        }
    }
    if (SMEM_HIST && HAS_RDF) {
        __syncthreads();
        unsigned long long *slab = a.slabs + (size_t)blockIdx.x * a.nkeys * a.nbins;
        for (int k = threadIdx.x; k < a.nkeys * a.nbins; k += blockDim.x) {
            uint32_t v = s_hist[k];
            if (v) slab[k] += v;
        }
    }
}

// sum the per-block slabs into out[nkeys*nbins] (+= so the global-atomic mode can share the buffer)
__global__ void __launch_bounds__(256) k_slab_reduce(const unsigned long long *slabs, int n_slabs, int n,
                                                     unsigned long long *out) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        unsigned long long s = 0;
        for (int b = 0; b < n_slabs; ++b) s += slabs[(size_t)b * n + k];
        out[k] += s;
    }
}
