// msd.cuh -- mean-squared-displacement kernels (K5-K7 of SURVEY.md 2.1).
//
// The trajectory lives on the device ATOM-MAJOR: P[a][k][c], a = local atom, k = frame, c = x/y/z, so the whole
// time series of one atom is one contiguous run of T*24 bytes.  That is what lets the window kernel read every
// position from HBM exactly once (the series is staged in shared memory and all (k, k-m) pairs are formed there)
// instead of streaming the trajectory once per window length.
//
//   k_msd_transpose   frame-major staging [count][n][3] -> atom-major slab (coalesced both ways through smem)
//   k_msd_scan        one warp per atom: displacement wrap (P8) + running sum along time, in place.
//                     mode UNWRAP : p_k <- p_0 + sum_{j<=k} wrap(p_j - p_{j-1})            (msd.py:222-230)
//                     mode PREPARE: q = p - com;  R_k <- sum_{j<=k} delta_j, delta_0 = q_0  (msd.py:235-237, trajectory.py:285-303)
//   k_msd_frame_sums  per-frame weighted sums over atoms (centre of mass; DirectMsd totals), deterministic 2-stage
//   k_msd_window      persistent blocks; per atom: series -> smem (SoA), all windows, per-species partial sums
//   k_msd_direct      DirectMsd recurrence (orthogonal cells, msd.py:81-105), one thread per atom
// and, further down, the STREAMING path of WindowMsd(unwrap=False) -- what the headline C5 number runs on:
//   k_msd_slab_sums / k_msd_slab_commit   ingest: per-frame mass sums; shift, wrap, running sum and transposition into an
//                     atom-major store of three arrays x[Tp] | y[Tp] | z[Tp] per atom, relative to the first frame
//   k_msd_window_wide autocorrelation form on wide register tiles, two series buffers filled by bulk copies, no block barrier
//   k_msd_window_soa  its narrow predecessor: difference form (fallback) and series too long for two buffers
//
// MSD is compared at 1e-12 relative (north_star), not bit-exactly: running sums are warp scans and the squared
// norms use explicit FMAs.  The wrap itself follows P8 operation by operation.
#pragma once
#include "common.cuh"

struct MsdGeom {   // per frame
    double cell[9];
    double inv[9];
};

// ---- transpose ---------------------------------------------------------------------------------
// src [count][n][3]; dst atom-major with T frames per atom, writing frames [first, first+count)
__global__ void __launch_bounds__(256) k_msd_transpose(const double *__restrict__ src, double *__restrict__ dst, int n, int T,
                                                       int first, int count) {
    __shared__ double tile[32][97];   // [frame][atom*3 + c], padded
    const int a0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int na = min(32, n - a0), nk = min(32, count - k0);
    for (int idx = threadIdx.x; idx < nk * 96; idx += blockDim.x) {
        int kk = idx / 96, e = idx - kk * 96;
        if (e < na * 3) tile[kk][e] = src[((size_t)(k0 + kk) * n + a0) * 3 + e];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < na * nk * 3; idx += blockDim.x) {
        int aa = idx / (nk * 3), r = idx - aa * (nk * 3);
        int kk = r / 3, c = r - kk * 3;
        dst[((size_t)(a0 + aa) * T + first + k0 + kk) * 3 + c] = tile[kk][aa * 3 + c];
    }
}

// atom-major -> frame-major (for amofb_msd_get_positions), frames [first, first+count)
__global__ void __launch_bounds__(256) k_msd_untranspose(const double *__restrict__ src, double *__restrict__ dst, int n, int T,
                                                         int first, int count) {
    __shared__ double tile[32][97];
    const int a0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int na = min(32, n - a0), nk = min(32, count - k0);
    for (int idx = threadIdx.x; idx < na * nk * 3; idx += blockDim.x) {
        int aa = idx / (nk * 3), r = idx - aa * (nk * 3);
        int kk = r / 3, c = r - kk * 3;
        tile[kk][aa * 3 + c] = src[((size_t)(a0 + aa) * T + first + k0 + kk) * 3 + c];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < nk * 96; idx += blockDim.x) {
        int kk = idx / 96, e = idx - kk * 96;
        if (e < na * 3) dst[((size_t)(k0 + kk) * n + a0) * 3 + e] = tile[kk][e];
    }
}

// ---- P8 ----------------------------------------------------------------------------------------
__device__ __forceinline__ double np_mod1(double g) {
    double r = g - trunc(g);            // == fmod(g, 1.0), exact
    if (r != 0.0) { if (r < 0.0) r += 1.0; }
    else r = 0.0;
    return r;
}

__device__ __forceinline__ void wrap_disp(const MsdGeom &G, double dx, double dy, double dz, double &ox, double &oy, double &oz) {
    const double shift = (0.0 - 0.5) - 1e-7;
    double g0 = (dx * G.inv[0] + dy * G.inv[3]) + dz * G.inv[6];
    double g1 = (dx * G.inv[1] + dy * G.inv[4]) + dz * G.inv[7];
    double g2 = (dx * G.inv[2] + dy * G.inv[5]) + dz * G.inv[8];
    g0 = np_mod1(g0 - shift) + shift;
    g1 = np_mod1(g1 - shift) + shift;
    g2 = np_mod1(g2 - shift) + shift;
    ox = (g0 * G.cell[0] + g1 * G.cell[3]) + g2 * G.cell[6];
    oy = (g0 * G.cell[1] + g1 * G.cell[4]) + g2 * G.cell[7];
    oz = (g0 * G.cell[2] + g1 * G.cell[5]) + g2 * G.cell[8];
}

// The same for a cell whose six off-diagonal entries (and therefore those of its inverse, P1) are zero: every dropped
// product is +-0 and x + (+-0) == x, so the results are those of wrap_disp bit for bit (up to the sign of a zero).
__device__ __forceinline__ void wrap_disp_diag(double i0, double i4, double i8, double c0, double c4, double c8,
                                               double dx, double dy, double dz, double &ox, double &oy, double &oz) {
    const double shift = (0.0 - 0.5) - 1e-7;
    ox = (np_mod1(dx * i0 - shift) + shift) * c0;
    oy = (np_mod1(dy * i4 - shift) + shift) * c4;
    oz = (np_mod1(dz * i8 - shift) + shift) * c8;
}

__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// one warp per atom; PREPARE subtracts com[k] first (translate(-cg)), UNWRAP does not.
// Per round the warp takes 32*SCAN_F frames: (1) coalesced read of the 768*... contiguous doubles into a padded
// shared-memory slab, (2) every lane walks ITS SCAN_F consecutive frames serially (wrap P8 + running sum, in place in
// the slab), (3) one warp scan of the 32 lane totals, (4) coalesced write-back adding each owner lane's offset.
// One shuffle scan per 256 frames instead of one per 32, and every global access is a full 128-byte line.
#define SCAN_F 8
#define SCAN_WARPS 6
#define SCAN_LANE_STRIDE (3 * SCAN_F + 1)                  // doubles per lane in the slab (+1 pad: conflict-free)
// FIXED_CELL: every frame has the same cell (NVT runs): the 144-byte geometry is read once into registers instead
// of once per frame and lane (it would otherwise be 6x the traffic of the positions themselves).
template <bool PREPARE, bool FIXED_CELL>
__global__ void __launch_bounds__(32 * SCAN_WARPS) k_msd_scan(double *__restrict__ P, const MsdGeom *__restrict__ geom,
                                                              const double *__restrict__ com, int n, int T) {
    __shared__ double s_slab[SCAN_WARPS][32 * SCAN_LANE_STRIDE];
    __shared__ double s_off[SCAN_WARPS][32][3];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double *slab = s_slab[wib];
    MsdGeom g0;
    if (FIXED_CELL) g0 = geom[0];
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long a = warp; a < n; a += nwarps) {
        double *p = P + (size_t)a * T * 3;
        double cx = 0.0, cy = 0.0, cz = 0.0;      // running sum carried between rounds
        double lx = 0.0, ly = 0.0, lz = 0.0;      // (shifted) original position of the last frame of the previous round
        for (int k0 = 0; k0 < T; k0 += 32 * SCAN_F) {
            const int nf = min(32 * SCAN_F, T - k0);               // frames in this round
            const int ne = 3 * nf;
            const double *src = p + 3 * (size_t)k0;
            __syncwarp();
#pragma unroll 4
            for (int e = lane; e < ne; e += 32) {
                double v = src[e];
                if (PREPARE) v -= com[3 * (size_t)k0 + e];
                slab[e + e / (3 * SCAN_F)] = v;
            }
            __syncwarp();
            const int kb = lane * SCAN_F;                          // first frame (within the round) of this lane
            const int nv = max(0, min(SCAN_F, nf - kb));
            double *mine = slab + lane * SCAN_LANE_STRIDE;
            // the original position just before my first frame: last frame of the lane before me / of the previous round
            double px = lx, py = ly, pz = lz;
            if (lane > 0 && nv > 0) { const double *q = mine - SCAN_LANE_STRIDE + 3 * (SCAN_F - 1); px = q[0]; py = q[1]; pz = q[2]; }
            // remember my own last original position before it is overwritten (the next lane read it above; order matters)
            double ex = 0.0, ey = 0.0, ez = 0.0;
            if (nv > 0) { ex = mine[3 * (nv - 1)]; ey = mine[3 * (nv - 1) + 1]; ez = mine[3 * (nv - 1) + 2]; }
            __syncwarp();
            double sx = 0.0, sy = 0.0, sz = 0.0;
            for (int i = 0; i < nv; ++i) {
                const double x = mine[3 * i], y = mine[3 * i + 1], z = mine[3 * i + 2];
                double dx, dy, dz;
                const int k = k0 + kb + i;
                if (k == 0) { dx = x; dy = y; dz = z; }                                       // delta_0 = first positions
                else wrap_disp(FIXED_CELL ? g0 : geom[k - 1], x - px, y - py, z - pz, dx, dy, dz);   // cell of frame k-1 wraps k-1 -> k
                px = x; py = y; pz = z;
                sx += dx; sy += dy; sz += dz;
                mine[3 * i] = sx; mine[3 * i + 1] = sy; mine[3 * i + 2] = sz;
            }
            const double ix = warp_incl_scan(sx, lane), iy = warp_incl_scan(sy, lane), iz = warp_incl_scan(sz, lane);
            s_off[wib][lane][0] = cx + (ix - sx);                  // everything before this lane
            s_off[wib][lane][1] = cy + (iy - sy);
            s_off[wib][lane][2] = cz + (iz - sz);
            __syncwarp();
            double *dst = p + 3 * (size_t)k0;
#pragma unroll 4
            for (int e = lane; e < ne; e += 32) {
                const int owner = e / (3 * SCAN_F), c = e % 3;
                dst[e] = slab[e + owner] + s_off[wib][owner][c];
            }
            cx += __shfl_sync(0xffffffffu, ix, 31); cy += __shfl_sync(0xffffffffu, iy, 31); cz += __shfl_sync(0xffffffffu, iz, 31);
            const int last_lane = (nf - 1) / SCAN_F;               // lane that owns the round's last frame
            lx = __shfl_sync(0xffffffffu, ex, last_lane); ly = __shfl_sync(0xffffffffu, ey, last_lane); lz = __shfl_sync(0xffffffffu, ez, last_lane);
        }
    }
}

// ---- per-frame sums over atoms -------------------------------------------------------------------
// partial[blockIdx.y][k][c] = sum over this block's atoms of w[a] * P[a][k][c], c < NC (NC = 3: all components, NC = 1: x only)
// grid.x = ceil(T/32) frame chunks, grid.y = atom groups; warps of a block take atoms round-robin and are
// combined in warp order, groups are combined in order by k_msd_sum_groups: the result does not depend on timing.
template <int NC>
__global__ void __launch_bounds__(256) k_msd_frame_sums(const double *__restrict__ P, const double *__restrict__ w, int n, int T,
                                                        double *__restrict__ partial) {
    __shared__ double red[8][32][NC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = blockIdx.x * 32 + lane;
    const int per = (n + gridDim.y - 1) / gridDim.y;
    const int a_lo = blockIdx.y * per, a_hi = min(n, a_lo + per);
    double acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.0;
    if (k < T)
        for (int a = a_lo + warp; a < a_hi; a += 8) {
            const double wa = w[a];
            if (wa == 0.0) continue;
            const double *p = P + ((size_t)a * T + k) * 3;
#pragma unroll
            for (int c = 0; c < NC; ++c) acc[c] += wa * p[c];
        }
#pragma unroll
    for (int c = 0; c < NC; ++c) red[warp][lane][c] = acc[c];
    __syncthreads();
    if (warp == 0 && k < T) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            double s = 0.0;
            for (int q = 0; q < 8; ++q) s += red[q][lane][c];
            partial[((size_t)blockIdx.y * T + k) * NC + c] = s;
        }
    }
}

__global__ void __launch_bounds__(256) k_msd_sum_groups(const double *__restrict__ partial, int groups, int len, double *__restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int g = 0; g < groups; ++g) s += partial[(size_t)g * len + i];
        out[i] = s;
    }
}

// ---- window MSD ------------------------------------------------------------------------------------
// partial[block][S][nw] = sum over the block's atoms of species s, over k = m+1..T-1, of |R_k - R_{k-m}|^2.
// One atom at a time: its series is staged in shared memory (SoA), a thread takes frames k = tid+1, +blockDim, ...
// keeps R_k in registers and walks the window lengths, so each pair costs one shared-memory read of R_{k-m}.
// The per-window sums stay in registers (MSD_NW per pass) across all atoms of one species; they are reduced over the
// block only when the species changes, in a fixed order, so the result does not depend on scheduling.
#ifndef MSD_NW
#define MSD_NW 32
#endif
#ifndef MSD_THREADS
#define MSD_THREADS 512
#endif

template <bool SMEM>
__global__ void __launch_bounds__(MSD_THREADS) k_msd_window(const double *__restrict__ P, const uint8_t *__restrict__ species, int n, int T,
                                                            const int *__restrict__ window, int nw, int S, double *__restrict__ partial) {
    extern __shared__ double sm[];
    double *s_acc = sm + (SMEM ? 3 * (size_t)T + 1 : 0);    // [S][nw] block accumulators (after the staged series)
    double *s_red = s_acc + (size_t)S * nw;                 // [nwarp][MSD_NW]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int i = threadIdx.x; i < S * nw; i += blockDim.x) s_acc[i] = 0.0;
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int a_lo = blockIdx.x * per, a_hi = min(n, a_lo + per);
    for (int w0 = 0; w0 < nw; w0 += MSD_NW) {
        int mw[MSD_NW];
        double acc[MSD_NW];
#pragma unroll
        for (int w = 0; w < MSD_NW; ++w) {
            const int m = (w0 + w < nw) ? window[w0 + w] : -1;
            mw[w] = (m >= 0 && m < T) ? m : T;              // T: no frame pair qualifies
            acc[w] = 0.0;
        }
        int cur_sp = -1;
        for (int a = a_lo; a <= a_hi; ++a) {
            const int sp = a < a_hi ? (int)species[a] : -2;
            if (sp != cur_sp) {
                if (cur_sp >= 0) {                          // species changed (or done): reduce the register sums
#pragma unroll
                    for (int w = 0; w < MSD_NW; ++w) {
                        double v = acc[w];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                        if (lane == 0) s_red[warp * MSD_NW + w] = v;
                        acc[w] = 0.0;
                    }
                    __syncthreads();
                    if (threadIdx.x < MSD_NW && w0 + threadIdx.x < nw) {
                        double t = 0.0;
                        for (int q = 0; q < nwarp; ++q) t += s_red[q * MSD_NW + threadIdx.x];
                        s_acc[cur_sp * nw + w0 + threadIdx.x] += t;
                    }
                    __syncthreads();
                }
                cur_sp = sp;
            }
            if (a >= a_hi) break;
            const double *p = P + (size_t)a * T * 3;
            if (SMEM) {
                __syncthreads();                            // everyone finished the previous atom
                // stage the series as it is (AoS, frame k at sm[3k..3k+2]): wide loads, 8 in flight per thread, so the
                // whole 24*T bytes are requested within one memory latency
                if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
                    const double2 *p2 = reinterpret_cast<const double2 *>(p);
                    double2 *s2 = reinterpret_cast<double2 *>(sm);
                    const int n2 = (3 * T) >> 1;
                    for (int i0 = threadIdx.x; i0 < n2; i0 += 8 * MSD_THREADS) {
                        double2 v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) if (i0 + u * MSD_THREADS < n2) v[u] = __ldg(p2 + i0 + u * MSD_THREADS);
#pragma unroll
                        for (int u = 0; u < 8; ++u) if (i0 + u * MSD_THREADS < n2) s2[i0 + u * MSD_THREADS] = v[u];
                    }
                    if ((3 * T) & 1) { if (threadIdx.x == 0) sm[3 * T - 1] = p[3 * T - 1]; }
                } else {
                    for (int i = threadIdx.x; i < 3 * T; i += blockDim.x) sm[i] = p[i];
                }
                __syncthreads();
            }
            for (int k = 1 + threadIdx.x; k < T; k += blockDim.x) {
                double xk, yk, zk;
                if (SMEM) { xk = sm[3 * k]; yk = sm[3 * k + 1]; zk = sm[3 * k + 2]; }
                else { xk = p[3 * (size_t)k]; yk = p[3 * (size_t)k + 1]; zk = p[3 * (size_t)k + 2]; }
                // branch-free over the windows (32 independent chains in flight): a pair that does not qualify
                // (k - m < 1) reads frame k itself and contributes exactly 0
#pragma unroll
                for (int w = 0; w < MSD_NW; ++w) {
                    const int j0 = k - mw[w];
                    const int j = j0 >= 1 ? j0 : k;
                    double dx, dy, dz;
                    if (SMEM) { dx = xk - sm[3 * j]; dy = yk - sm[3 * j + 1]; dz = zk - sm[3 * j + 2]; }
                    else { dx = xk - p[3 * (size_t)j]; dy = yk - p[3 * (size_t)j + 1]; dz = zk - p[3 * (size_t)j + 2]; }
                    acc[w] += __fma_rn(dz, dz, __fma_rn(dy, dy, dx * dx));
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S * nw; i += blockDim.x) partial[(size_t)blockIdx.x * S * nw + i] = s_acc[i];
}

// ---- window MSD, window lengths in arithmetic progression -------------------------------------------------
// WindowMsd always asks for m = 0, D, 2D, ... (msd.py:176-178: window = arange(0, max_time, delta_time)).  k_msd_window
// reads one R_{k-m} from shared memory per frame pair and is bound by the shared-memory pipe (ncu: 85 % busy, the
// 24-byte stride costing a 2-way bank conflict).  With lags that are multiples of D a thread can keep a TILE of KB
// frames k_i = b + i*D in registers and walk the partners j_u = b - (wlo + u)*D: the pair (k_i, j_u) has lag
// (wlo + u + i)*D, i.e. it belongs to window wlo + u + i, so one shared-memory read of R_j feeds KB pairs.
//   * the windows of a pass are split into NG groups of <= MSD_AP_NWT; a warp belongs to ONE group and holds only that
//     group's sums in registers (8 doubles instead of 32 -> three times the warps per SM); every group walks all tasks;
//   * full super-rows: frames [1 + s*KB*D, 1 + (s+1)*KB*D) all exist -> task (s, r) owns k_i = 1 + s*KB*D + r + i*D;
//     the remaining frames (fewer than KB*D) are handled one frame per task (KB = 1);
//   * a partner before frame 1 (the reference starts at k = m + 1, msd.py:197) is clamped and weighted by 0.0 through
//     the accumulating FMA (fma(1.0, d2, acc) == acc + d2 bit for bit).
// Everything is fully unrolled, so the window sums stay in registers under static indices.
#ifndef MSD_AP_KB
#define MSD_AP_KB 4
#endif
#ifndef MSD_AP_FMA_ACC
#define MSD_AP_FMA_ACC 1      // measured: 14.9 vs 15.5 ms per C5/100k step; sums stay within 1e-12 of the oracle
#endif
#define MSD_AP_NWT_MAX 13       // window sums per thread: instantiated for 5, 7, 9, 11, 13 (128 registers at 13)
#ifndef MSD_AP_THREADS
#define MSD_AP_THREADS 512
#endif

// sm: the staged series as three arrays x[tp], y[tp], z[tp] (consecutive threads -> consecutive words: no bank conflict)
// EDGE: some partner of this tile lies before frame 1 (the reference starts at k = m + 1, msd.py:197); partners only move
// backwards with u, so the lane simply stops at the first one that does not exist.
template <int KB, int NWT, bool EDGE>
__device__ __forceinline__ void msd_ap_tile(const double *__restrict__ sm, int tp, int b, int delta, int lag0, double (&acc)[NWT]) {
    double kx[KB], ky[KB], kz[KB];
#pragma unroll
    for (int i = 0; i < KB; ++i) {
        const double *q = sm + (b + i * delta);
        kx[i] = q[0]; ky[i] = q[tp]; kz[i] = q[2 * tp];
    }
#pragma unroll
    for (int u = -(KB - 1); u < NWT; ++u) {
        const int j0 = b - lag0 - u * delta;
        if (EDGE && j0 < 1) break;
        const double *q = sm + j0;
        const double jx = q[0], jy = q[tp], jz = q[2 * tp];
#pragma unroll
        for (int i = 0; i < KB; ++i) {
            const int w = u + i;
            if (w >= 0 && w < NWT) {
                const double dx = kx[i] - jx, dy = ky[i] - jy, dz = kz[i] - jz;
#if MSD_AP_FMA_ACC
                acc[w] = __fma_rn(dz, dz, __fma_rn(dy, dy, __fma_rn(dx, dx, acc[w])));      // 6 instead of 7 FP64 instructions per pair
#else
                acc[w] += __fma_rn(dz, dz, __fma_rn(dy, dy, dx * dx));
#endif
            }
        }
    }
}

// ng groups of NWT windows per pass (ng * NWT windows; those beyond nw are formed and dropped); blockDim.x = 32 * ng * (warps per group)
// prepare != 0: P still holds the positions; the centre-of-mass shift, the displacement wrap (P8) and the running sum
// along time that k_msd_scan<PREPARE> would write back to HBM are done here, on the staged series, every time an atom
// is loaded: one 24*N*T read instead of a read, a write and another read.
template <int KB, int NWT, int CELL>      // CELL: 0 = one cell per frame, 1 = the same cell in every frame, 2 = the same orthorhombic cell
__global__ void __launch_bounds__(MSD_AP_THREADS) k_msd_window_ap(const double *__restrict__ P, const uint8_t *__restrict__ species, int n, int T,
                                                                  int delta, int nw, int ng, int S, double *__restrict__ partial,
                                                                  const MsdGeom *__restrict__ geom, const double *__restrict__ com, int prepare) {
    constexpr int nwt = NWT;
    extern __shared__ double sm[];
    const int tp = (T + 1) & ~1;                            // padded series length (keeps y[] and z[] 16-byte aligned)
    double *s_acc = sm + 3 * (size_t)tp;                    // [S][nw] block accumulators (after the staged series)
    double *s_red = s_acc + (size_t)S * nw;                 // [nwarp][NWT]
    double *s_scan = s_red + (size_t)MSD_NW * (MSD_AP_THREADS / 32);   // [32][3] warp totals of the in-block running sum
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int grp = warp % ng, wpg = nwarp / ng;            // my window group; warps per group
    const int tg = (warp / ng) * 32 + lane, gs = wpg * 32;  // my index among the group's threads
    for (int i = threadIdx.x; i < S * nw; i += blockDim.x) s_acc[i] = 0.0;
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int a_lo = blockIdx.x * per, a_hi = min(n, a_lo + per);
    const int span = KB * delta;
    const int nsr = (T - 1) / span;                         // full super-rows over k = 1 .. T-1
    const int ntask = nsr * delta;
    const int k_rem = 1 + nsr * span;                       // first frame of the remainder
    for (int w0 = 0; w0 < nw; w0 += ng * nwt) {
        const int wlo = w0 + grp * nwt;                     // first window of my group in this pass
        const bool mine = wlo < nw;                         // my group has at least one requested window
        const int lag0 = wlo * delta;
        const int lag_far = lag0 + (NWT - 1) * delta;       // the longest lag of my group
        double acc[NWT];
#pragma unroll
        for (int w = 0; w < NWT; ++w) acc[w] = 0.0;
        int cur_sp = -1;
        for (int a = a_lo; a <= a_hi; ++a) {
            const int sp = a < a_hi ? (int)species[a] : -2;
            if (sp != cur_sp) {
                if (cur_sp >= 0) {                          // species changed (or done): reduce the register sums
#pragma unroll
                    for (int w = 0; w < NWT; ++w) {
                        double v = acc[w];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                        if (lane == 0) s_red[warp * NWT + w] = v;
                        acc[w] = 0.0;
                    }
                    __syncthreads();
                    if (threadIdx.x < ng * nwt) {           // thread -> (group, window of the group); warps of a group in order
                        const int g = threadIdx.x / nwt, w = threadIdx.x - g * nwt, wi = w0 + g * nwt + w;
                        if (wi < nw) {
                            double t = 0.0;
                            for (int q = 0; q < wpg; ++q) t += s_red[(q * ng + g) * NWT + w];
                            s_acc[cur_sp * nw + wi] += t;
                        }
                    }
                    __syncthreads();
                }
                cur_sp = sp;
            }
            if (a >= a_hi) break;
            const double *p = P + (size_t)a * T * 3;
            __syncthreads();                                // everyone finished the previous atom
            // stage the series AoS -> SoA: wide loads, 4 in flight per thread; element e of the stream is component
            // e % 3 of frame e / 3
            if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
                const double2 *p2 = reinterpret_cast<const double2 *>(p);
                const int n2 = (3 * T) >> 1;
                // prepare: the centre of mass [T][3] has the layout of the series itself, so the shift p - com is taken here,
                // element by element, with the same coalesced access (com stays in L2)
                const double2 *c2 = reinterpret_cast<const double2 *>(com);
                for (int i0 = threadIdx.x; i0 < n2; i0 += 8 * blockDim.x) {
                    double2 v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) if (i0 + u * blockDim.x < n2) v[u] = __ldg(p2 + i0 + u * blockDim.x);
                    if (prepare) {
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (i0 + u * blockDim.x < n2) { const double2 c = __ldg(c2 + i0 + u * blockDim.x); v[u].x -= c.x; v[u].y -= c.y; }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = i0 + u * blockDim.x;
                        if (i < n2) {
                            const unsigned e = 2u * (unsigned)i, f0 = e / 3u, c0 = e - 3u * f0;
                            const unsigned f1 = c0 == 2u ? f0 + 1u : f0, c1 = c0 == 2u ? 0u : c0 + 1u;
                            sm[c0 * tp + f0] = v[u].x;
                            sm[c1 * tp + f1] = v[u].y;
                        }
                    }
                }
                if ((3 * T) & 1) { if (threadIdx.x == 0) sm[2 * tp + (T - 1)] = p[3 * T - 1] - (prepare ? com[3 * T - 1] : 0.0); }
            } else {
                for (int i = threadIdx.x; i < 3 * T; i += blockDim.x) { const int f = i / 3; sm[(i - 3 * f) * tp + f] = p[i] - (prepare ? com[i] : 0.0); }
            }
            if (a + 1 < a_hi) {                             // pull the next series into L2 while this one is worked on
                const char *nx = reinterpret_cast<const char *>(p + (size_t)T * 3);
                const int nlines = (3 * T * 8 + 127) >> 7;
                for (int i = threadIdx.x; i < nlines; i += blockDim.x)
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(nx + ((size_t)i << 7)));
            }
            __syncthreads();
            if (prepare) {
                // thread t owns the frames [t*C, t*C + C), C odd (conflict-free 8-byte strides): wrap + local running sum in
                // place, then one block-wide exclusive scan of the per-thread totals.  (Measured: forming the displacements
                // in parallel over frames first, with coalesced centre-of-mass reads, is 8 % slower on C5.)
                const int C = ((T + (int)blockDim.x - 1) / (int)blockDim.x) | 1;
                const int k0 = threadIdx.x * C, k1 = min(T, k0 + C);
                double px = 0.0, py = 0.0, pz = 0.0;        // (shifted) position of the frame before my first one
                if (k0 >= 1 && k0 < T) { px = sm[k0 - 1]; py = sm[tp + k0 - 1]; pz = sm[2 * tp + k0 - 1]; }
                __syncthreads();                            // every predecessor is read before anybody overwrites it
                double sx = 0.0, sy = 0.0, sz = 0.0;
                {
                    MsdGeom g0;
                    if (CELL == 1) g0 = geom[0];
                    double i0 = 0.0, i4 = 0.0, i8 = 0.0, c0 = 0.0, c4 = 0.0, c8 = 0.0;
                    if (CELL == 2) { i0 = geom[0].inv[0]; i4 = geom[0].inv[4]; i8 = geom[0].inv[8]; c0 = geom[0].cell[0]; c4 = geom[0].cell[4]; c8 = geom[0].cell[8]; }
                    // (Measured: four frames per trip with their wraps overlapped is 20 % slower -- register spills.)
                    for (int k = k0; k < k1; ++k) {
                        const double x = sm[k], y = sm[tp + k], z = sm[2 * tp + k];       // already shifted by the centre of mass
                        double dx, dy, dz;
                        if (k == 0) { dx = x; dy = y; dz = z; }                                     // delta_0 = first positions
                        else if (CELL == 2) wrap_disp_diag(i0, i4, i8, c0, c4, c8, x - px, y - py, z - pz, dx, dy, dz);
                        else wrap_disp(CELL == 1 ? g0 : geom[k - 1], x - px, y - py, z - pz, dx, dy, dz);   // cell of frame k-1 wraps k-1 -> k
                        px = x; py = y; pz = z;
                        sx += dx; sy += dy; sz += dz;
                        sm[k] = sx; sm[tp + k] = sy; sm[2 * tp + k] = sz;
                    }
                }
                const double ix = warp_incl_scan(sx, lane), iy = warp_incl_scan(sy, lane), iz = warp_incl_scan(sz, lane);
                if (lane == 31) { s_scan[3 * warp] = ix; s_scan[3 * warp + 1] = iy; s_scan[3 * warp + 2] = iz; }
                __syncthreads();
                double ox, oy, oz;
                {
                    const double tx = lane < nwarp ? s_scan[3 * lane] : 0.0, ty = lane < nwarp ? s_scan[3 * lane + 1] : 0.0,
                                 tz = lane < nwarp ? s_scan[3 * lane + 2] : 0.0;
                    const double wx = warp_incl_scan(tx, lane), wy = warp_incl_scan(ty, lane), wz = warp_incl_scan(tz, lane);
                    ox = __shfl_sync(0xffffffffu, wx - tx, warp) + (ix - sx);       // warps before mine + lanes before me
                    oy = __shfl_sync(0xffffffffu, wy - ty, warp) + (iy - sy);
                    oz = __shfl_sync(0xffffffffu, wz - tz, warp) + (iz - sz);
                }
                for (int k = k0; k < k1; ++k) { sm[k] += ox; sm[tp + k] += oy; sm[2 * tp + k] += oz; }
                __syncthreads();
            }
            if (mine) {
                for (int q = tg; q < ntask; q += gs) {
                    const int s = q / delta, r = q - s * delta;
                    const int b = 1 + s * span + r;
                    if (b - lag_far >= 1) msd_ap_tile<KB, NWT, false>(sm, tp, b, delta, lag0, acc);
                    else msd_ap_tile<KB, NWT, true>(sm, tp, b, delta, lag0, acc);
                }
                for (int k = k_rem + tg; k < T; k += gs) {
                    if (k - lag_far >= 1) msd_ap_tile<1, NWT, false>(sm, tp, k, delta, lag0, acc);
                    else msd_ap_tile<1, NWT, true>(sm, tp, k, delta, lag0, acc);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S * nw; i += blockDim.x) partial[(size_t)blockIdx.x * S * nw + i] = s_acc[i];
}

// ---- DirectMsd ----------------------------------------------------------------------------------------
// one thread per atom, sequential in t; overwrites P[a][t].x with |r_t - r_0|^2 (P[a][0].x = 0)
__global__ void __launch_bounds__(128) k_msd_direct(double *__restrict__ P, const MsdGeom *__restrict__ geom, int n, int T) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    double *p = P + (size_t)a * T * 3;
    double r0[3] = {p[0], p[1], p[2]}, r[3] = {p[0], p[1], p[2]};
    p[0] = 0.0;
    for (int t = 1; t < T; ++t) {
        double s2 = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double L = geom[t].cell[4 * j];
            const double prev = r[j];
            double pm = fmod(prev, L);                                   // numpy %: result takes the divisor's sign
            if (pm != 0.0 && ((L < 0.0) != (pm < 0.0))) pm += L;
            double d = p[3 * (size_t)t + j] - pm;
            if (d > L / 2) d -= L; else if (d < -L / 2) d += L;
            const double rt = d + prev;
            r[j] = rt;
            const double dd = rt - r0[j];
            s2 += dd * dd;
        }
        p[3 * (size_t)t] = s2;
    }
}

// =====================================================================================================================
// Streaming path of WindowMsd(unwrap=False): the trajectory arrives slab by slab (frame-major), and everything that is
// a pass over the positions happens ON THE WAY IN:
//   k_msd_slab_sums    per-frame mass-weighted sums of the slab (the centre of mass needs every atom of a frame before
//                      any displacement can be wrapped: msd.py:235-237 precedes trajectory.py:285-303), deterministic;
//   k_msd_slab_commit  centre-of-mass shift, displacement wrap (P8, the cell of the earlier frame), running sum along
//                      time with a per-atom carry from the previous slab, and the transposition into the atom-major
//                      store -- as three arrays x[Tp] | y[Tp] | z[Tp] per atom, which is the layout the window kernel
//                      wants in shared memory, so its staging is one bulk copy.
// The running sum starts from 0 instead of the first position (R'_k = R_k - R_0): every difference R_k - R_{k-m} is
// unchanged, and the smaller magnitudes are what lets the window kernel use the autocorrelation form.
//   k_msd_window_soa   |R_k - R_j|^2 = |R_k|^2 + |R_j|^2 - 2 R_k.R_j: the cross term costs 3 FMAs per frame pair (the
//                      difference form 3 subtractions + 3 FMAs), the squares come from one prefix sum per atom.
// =====================================================================================================================

// partial[g][f*3 + c] = sum over the atoms of group g of w[a] * slab[f][a][c]
__global__ void __launch_bounds__(256) k_msd_slab_sums(const double *__restrict__ slab, const double *__restrict__ w, int n, int count,
                                                       double *__restrict__ partial) {
    __shared__ double red[8][3];
    const int f = blockIdx.x, g = blockIdx.y, groups = gridDim.y;
    const int per = (n + groups - 1) / groups;
    const int a_lo = g * per, a_hi = min(n, a_lo + per);
    const double *src = slab + (size_t)f * n * 3;
    double ax = 0.0, ay = 0.0, az = 0.0;
    for (int a = a_lo + threadIdx.x; a < a_hi; a += 256) {
        const double m = w[a];
        const double *p = src + 3 * (size_t)a;
        ax = __fma_rn(m, p[0], ax); ay = __fma_rn(m, p[1], ay); az = __fma_rn(m, p[2], az);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ax += __shfl_xor_sync(0xffffffffu, ax, o); ay += __shfl_xor_sync(0xffffffffu, ay, o); az += __shfl_xor_sync(0xffffffffu, az, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[warp][0] = ax; red[warp][1] = ay; red[warp][2] = az; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0.0;
        for (int q = 0; q < 8; ++q) s += red[q][threadIdx.x];
        partial[(size_t)g * count * 3 + (size_t)f * 3 + threadIdx.x] = s;
    }
}

#ifndef COMMIT_A
#define COMMIT_A 32                       // atoms per block: 768-byte runs of every frame row (64 atoms x 384 threads: 8.86, 32 x 192: 8.53,
                                          // 16 x 96: 8.47, 128 x 384: 11.6 ms of ingest per 100 000 atoms x 5 000 frames)
#endif
#define COMMIT_FS 32                      // frames per round: 256-byte runs of every output row
#define COMMIT_LD (3 * COMMIT_A + 1)      // odd row stride: the column reads of the write-out are conflict-free
#ifndef COMMIT_THREADS
#define COMMIT_THREADS 192                // two frame rows of the block's 96 columns per load pass; 6 blocks per SM
#endif
#define COMMIT_ITEMS ((COMMIT_FS * COMMIT_A + COMMIT_THREADS - 1) / COMMIT_THREADS)
#define COMMIT_SMEM (sizeof(double) * (COMMIT_FS * COMMIT_LD + 6 * COMMIT_A))      // above the 48 KB static limit: opt-in

// carry[a][6]: shifted position of the last committed frame, running sum up to it.
// (A variant that kept the rows of round r+1 in flight with 8-byte cp.async copies into a second buffer was slower:
// 12.1 instead of 10.3 ms per 100 000 atoms x 5 000 frames; the plain loads below already have 16 rows in flight per thread.)
template <int CELL>      // 0 = one cell per frame, 1 = the same cell in every frame, 2 = the same orthorhombic cell
__global__ void __launch_bounds__(COMMIT_THREADS, 1152 / COMMIT_THREADS) k_msd_slab_commit(const double *__restrict__ slab, double *__restrict__ P,
                                                                      const MsdGeom *__restrict__ geom, const double *__restrict__ com,
                                                                      double *__restrict__ carry, int n, int Tp, int first, int count) {
    extern __shared__ __align__(16) double commit_sm[];          // COMMIT_SMEM bytes
    double *tile = commit_sm, *s_prev = tile + COMMIT_FS * COMMIT_LD, *s_run = s_prev + 3 * COMMIT_A;
    const int a0 = blockIdx.x * COMMIT_A, na = min(COMMIT_A, n - a0), ncol = 3 * na;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int c = tid; c < ncol; c += COMMIT_THREADS) {
        const size_t at = (size_t)(a0 + c / 3) * 6 + (size_t)(c % 3);
        s_prev[c] = first > 0 ? carry[at] : 0.0;
        s_run[c] = first > 0 ? carry[at + 3] : 0.0;
    }
    MsdGeom g0;
    if (CELL == 1) g0 = geom[0];
    double i0 = 0.0, i4 = 0.0, i8 = 0.0, c0 = 0.0, c4 = 0.0, c8 = 0.0;
    if (CELL == 2) { i0 = geom[0].inv[0]; i4 = geom[0].inv[4]; i8 = geom[0].inv[8]; c0 = geom[0].cell[0]; c4 = geom[0].cell[4]; c8 = geom[0].cell[8]; }
    const int col = tid % (3 * COMMIT_A), row0 = tid / (3 * COMMIT_A), comp = col % 3;       // load mapping: my column is fixed
    // item mapping of the wrap step: (frame row, atom) = (idx / 64, idx % 64), idx = tid + u * 384
    for (int k0 = 0; k0 < count; k0 += COMMIT_FS) {
        const int nr = min(COMMIT_FS, count - k0);
        __syncthreads();                              // the previous round was written out (and the carry is in place)
        // 1) rows of the slab, shifted by the centre of mass of their frame (translate(-cg), msd.py:237)
        if (col < ncol) {
            const double *src = slab + ((size_t)k0 * n + a0) * 3 + col;
            const double *cm = com + 3 * (size_t)k0 + comp;
            for (int r = row0; r < nr; r += COMMIT_THREADS / (3 * COMMIT_A))
                tile[r * COMMIT_LD + col] = src[(size_t)r * n * 3] - cm[3 * r];
        }
        __syncthreads();
        // 2) wrapped displacement of every (frame, atom) of the round, kept in registers until every position was read
        double dx[COMMIT_ITEMS], dy[COMMIT_ITEMS], dz[COMMIT_ITEMS];
#pragma unroll
        for (int u = 0; u < COMMIT_ITEMS; ++u) {
            const int idx = tid + u * COMMIT_THREADS, r = idx / COMMIT_A, aa = idx % COMMIT_A, k = first + k0 + r;
            dx[u] = dy[u] = dz[u] = 0.0;
            if (r < nr && aa < na && k > 0) {         // delta_0 = 0: the running sum is taken relative to the first frame
                const double *q = tile + r * COMMIT_LD + 3 * aa;
                const double *qp = r > 0 ? q - COMMIT_LD : s_prev + 3 * aa;
                const double ex = q[0] - qp[0], ey = q[1] - qp[1], ez = q[2] - qp[2];
                if (CELL == 2) wrap_disp_diag(i0, i4, i8, c0, c4, c8, ex, ey, ez, dx[u], dy[u], dz[u]);
                else wrap_disp(CELL == 1 ? g0 : geom[k - 1], ex, ey, ez, dx[u], dy[u], dz[u]);     // cell of frame k-1 wraps k-1 -> k
            }
        }
        __syncthreads();
        for (int c = tid; c < ncol; c += COMMIT_THREADS) s_prev[c] = tile[(nr - 1) * COMMIT_LD + c];
        __syncthreads();
#pragma unroll
        for (int u = 0; u < COMMIT_ITEMS; ++u) {
            const int idx = tid + u * COMMIT_THREADS, r = idx / COMMIT_A, aa = idx % COMMIT_A;
            if (r < nr && aa < na) {
                double *q = tile + r * COMMIT_LD + 3 * aa;
                q[0] = dx[u]; q[1] = dy[u]; q[2] = dz[u];
            }
        }
        __syncthreads();
        // 3) running sum along time, one thread per (atom, component), sequential like the reference's r_k += delta_k
        if (tid < ncol) {
            double run = s_run[tid];
            for (int r = 0; r < nr; ++r) { run += tile[r * COMMIT_LD + tid]; tile[r * COMMIT_LD + tid] = run; }
            s_run[tid] = run;
        }
        __syncthreads();
        // 4) write-out: one warp per (atom, component), lanes = frames: 256-byte runs of the atom-major store
        if (lane < nr) {
            double *dst = P + (size_t)a0 * 3 * Tp + (size_t)(first + k0 + lane);
            for (int c = warp; c < ncol; c += COMMIT_THREADS / 32) dst[(size_t)c * Tp] = tile[lane * COMMIT_LD + c];   // (a, comp) rows are consecutive: row a*3 + comp
        }
        // (a warp scan over the 32 frames of a column, fused with the write-out, removes the serial loop and one barrier but was
        // slower: 11.3 against 8.8 ms of ingest per 100 000 atoms x 5 000 frames)
    }
    __syncthreads();
    for (int c = tid; c < ncol; c += COMMIT_THREADS) {
        const size_t at = (size_t)(a0 + c / 3) * 6 + (size_t)(c % 3);
        carry[at] = s_prev[c];
        carry[at + 3] = s_run[c];
    }
}

// (A register version -- a lane per (atom, component) column for the whole slab, the other components by shuffle, carry in registers, one
// barrier per round -- gave the same bits and was slower: 9.1-9.8 against 8.85 ms per 100 000 atoms x 5 000 frames;
// experiments/csrc/msd_commit_reg.cuh.)

// One component of one atom at a time: |R_k - R_j|^2 and R_k . R_j are sums over x, y, z, so the block stages 8*Tp bytes
// instead of 24*Tp, several blocks share an SM, and one block's bulk copy and barriers hide behind the others' tiles.
#ifndef MSD_SOA_THREADS
#define MSD_SOA_THREADS 256
#endif
#ifndef MSD_SOA_MIN_BLOCKS
#define MSD_SOA_MIN_BLOCKS 3
#endif
#ifndef MSD_SOA_KB
#define MSD_SOA_KB 8            // frames a thread keeps in registers per tile (4: 7.09, 6: 6.59, 8: 6.54 ms per 100 000 atoms x 5 000 frames)
#endif

// sm: one component of the series.  DOT: acc[w] += v_k * v_j, else acc[w] += (v_k - v_j)^2.
// EDGE: the walk stops at the first partner before frame 1 (the reference starts at k = m + 1, msd.py:197).
// (Walking the frames with pointer steps instead of b + i*delta, and advancing the task index without a division, measured
// 6 % slower: 6.92 instead of 6.53 ms per 100 000 atoms x 5 000 frames.)
template <int KB, int NWT, bool EDGE, bool DOT>
__device__ __forceinline__ void msd_soa_tile(const double *__restrict__ sm, int b, int delta, int lag0, double (&acc)[NWT]) {
    double kv[KB];
#pragma unroll
    for (int i = 0; i < KB; ++i) kv[i] = sm[b + i * delta];
#pragma unroll
    for (int u = -(KB - 1); u < NWT; ++u) {
        const int j0 = b - lag0 - u * delta;
        if (EDGE && j0 < 1) break;
        const double jv = sm[j0];
#pragma unroll
        for (int i = 0; i < KB; ++i) {
            const int w = u + i;
            if (w >= 0 && w < NWT) {
                if (DOT) acc[w] = __fma_rn(kv[i], jv, acc[w]);
                else { const double d = kv[i] - jv; acc[w] = __fma_rn(d, d, acc[w]); }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Wide register tiles for the autocorrelation form (round 2, second design).  ncu of k_msd_window_soa<8,13,true>: DFMA is 30 % of
// the executed instructions, the shared-memory pipe is busier (49 %) than the FP64 pipe (32 %), 3.4 barrier-stall cycles per
// instruction.  Three changes:
//   * ONE tile covers all the windows of a pass (NWT up to 32 sums per thread) and KB = 6 / 8 / 10 frames: the partners of the
//     windows 0 .. KB-1 ARE the thread's own frames, so a full tile costs NWT - 1 + KB shared-memory reads for KB (NWT - 1)
//     products (10 x 24 products on 34 reads at C5, against 8 x 13 on 28);
//   * no remainder tiles and no per-partner bounds test in the steady state: the series is followed by zeros up to the end of
//     the last super-row (a zero frame adds nothing to a sum of products), tiles whose partners all exist run the unchecked
//     body, the others stop at the first partner before frame 1 (frame 0 is the origin: exactly zero);
//   * two series buffers: the bulk copy of the next component is in flight while this one is worked on, and a block never
//     waits at a barrier for data.
#ifndef MSD_WIDE_PLAIN_LDS
#define MSD_WIDE_PLAIN_LDS 1        // 1: ordinary loads the compiler may schedule early; 0: ld.shared through a 32-bit address, kept in order
#endif
__device__ __forceinline__ double msd_lds(unsigned addr, const char *ptr) {
#if MSD_WIDE_PLAIN_LDS
    return *reinterpret_cast<const double *>(ptr);
#else
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
#endif
}

// sb / sp: shared byte address of / pointer to frame b of the series; d8 = 8 delta; lag8 = 8 wlo delta; partner slot v is frame
// b + (v - wlo) delta and exists iff v >= -e (e = s KB - wlo)
#ifndef MSD_WIDE_AHEAD
#define MSD_WIDE_AHEAD 3            // partner reads issued this many slots before their products
#endif
// EC >= 0: the tile's e is the compile-time EC (no tests in the body); EC < 0: e is only known at run time
template <int KB, int NWT, bool WLO0, int EC>
__device__ __forceinline__ void msd_wide_tile(unsigned sb, const char *sp, int d8, int lag8, int e_rt, double (&acc)[NWT]) {
    constexpr bool RT = EC < 0;
    const int e = RT ? e_rt : EC;
    double kv[KB];
#pragma unroll
    for (int i = 0; i < KB; ++i) kv[i] = msd_lds(sb + (unsigned)(i * d8), sp + i * d8);
    constexpr int V0 = WLO0 ? -1 : KB - 1;                  // first partner slot that is read from shared memory
    double ring[MSD_WIDE_AHEAD];
#pragma unroll
    for (int a = 0; a < MSD_WIDE_AHEAD; ++a) {
        const int v = V0 - a;
        ring[a] = (v > -NWT && v >= -e) ? msd_lds(sb + (unsigned)(v * d8 - lag8), sp + (v * d8 - lag8)) : 0.0;
    }
#pragma unroll
    for (int v = KB - 1; v > -NWT; --v) {
        if ((!WLO0 || v < 0) && v < -e) break;
        double pv;
        if (WLO0 && v >= 0) pv = kv[v];
        else {
            const int slot = (V0 - v) % MSD_WIDE_AHEAD;    // compile-time after unrolling
            pv = ring[slot];
            const int vn = v - MSD_WIDE_AHEAD;
            ring[slot] = (vn > -NWT && vn >= -e) ? msd_lds(sb + (unsigned)(vn * d8 - lag8), sp + (vn * d8 - lag8)) : 0.0;
        }
#pragma unroll
        for (int i = 0; i < KB; ++i) {
            const int t = i - v;
            if (t >= (WLO0 ? 1 : 0) && t < NWT) acc[t] = __fma_rn(kv[i], pv, acc[t]);      // window 0 is returned as exactly zero
        }
    }
}

// first pass (windows from 0): e = s KB is one of a few values below NWT - 1, each with its own test-free body
template <int KB, int NWT, int LVL>
__device__ __forceinline__ void msd_wide_first(int s, unsigned sb, const char *sp, int d8, double (&acc)[NWT]) {
    if constexpr (LVL * KB >= NWT - 1) {
        msd_wide_tile<KB, NWT, true, NWT - 1>(sb, sp, d8, 0, 0, acc);
    } else {
        if (s == LVL) msd_wide_tile<KB, NWT, true, LVL * KB>(sb, sp, d8, 0, 0, acc);
        else msd_wide_first<KB, NWT, LVL + 1>(s, sb, sp, d8, acc);
    }
}

#define MSD_WIDE_THREADS 512       // resident threads per SM the kernel is built for: THREADS = 256 -> two blocks
#ifndef MSD_WIDE_WAIT_NS
#define MSD_WIDE_WAIT_NS 4000       // suspend-time hint of the series wait (0: plain try_wait spin)
#endif
#define MSD_WIDE_NWT_MAX 32
#define MSD_WIDE_MAX_BUF 4

// partial[block][2][S][nw] as k_msd_window_soa<.., true>: [0] sum of R_k . R_{k-m}, [1] sum of |R_k|^2 + |R_{k-m}|^2.
// rowcap: doubles per series buffer (tp + the zero tail, even); nbuf series buffers (2 .. MSD_WIDE_MAX_BUF).
// pieces[pass][threads + 1]: frame ranges for the sums of squares of a pass -- thread t adds up the frames [pieces[t], pieces[t+1]) of
// every series; no range straddles m + 1 or T - m for a window length m of the pass, so
// sum_{k=m+1}^{T-1} v_k^2 + sum_{j=1}^{T-1-m} v_j^2  is a sum of whole ranges.
// Dynamic shared memory: nbuf rowcap | 2 S nw | nwarp NWT | threads.
//
// One block per SM and no block-wide barrier in the steady state: a warp waits for a series on its buffer's mbarrier, works
// through its tiles and counts itself out of the buffer; the last warp out issues the bulk copy of the series nbuf further on.
// Warps drift up to nbuf - 1 series apart.  The tiles of a series are not equally long (those near frame 0 have few partners), so
// the thread -> tile map moves on by one super-row with every series: over nsr series every thread has had every length.
template <int KB, int NWT, int THREADS>
__global__ void __launch_bounds__(THREADS, MSD_WIDE_THREADS / THREADS) k_msd_window_wide(const double *__restrict__ P, const uint8_t *__restrict__ species,
                                                                         const int *__restrict__ perm, const int *__restrict__ pieces,
                                                                         int n, int T, int tp, int delta, int nw, int S, int rowcap, int nbuf,
                                                                         double *__restrict__ partial) {
    extern __shared__ __align__(16) double sm[];
    double *s_acc = sm + (size_t)nbuf * rowcap;             // [S][nw]
    double *s_ss = s_acc + (size_t)S * nw;                  // [S][nw]
    double *s_red = s_ss + (size_t)S * nw;                  // [nwarp][NWT]
    double *s_part = s_red + (size_t)NWT * (THREADS / 32);      // [threads]: sums of squares per range
    __shared__ __align__(8) unsigned long long s_mbar[MSD_WIDE_MAX_BUF];
    __shared__ unsigned s_done[MSD_WIDE_MAX_BUF];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int i = threadIdx.x; i < 2 * S * nw; i += blockDim.x) s_acc[i] = 0.0;      // s_acc and s_ss are adjacent
    for (int b = 0; b < nbuf; ++b)                                                  // the zero tails (never written again)
        for (int i = tp + threadIdx.x; i < rowcap; i += blockDim.x) sm[(size_t)b * rowcap + i] = 0.0;
    const unsigned mbar0 = (unsigned)__cvta_generic_to_shared(&s_mbar[0]);
    const unsigned sm_addr = (unsigned)__cvta_generic_to_shared(sm);
    if ((int)threadIdx.x < nbuf) { mbar_init(mbar0 + 8u * threadIdx.x, 1); s_done[threadIdx.x] = 0u; }
    __syncthreads();
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int a_lo = min(n, blockIdx.x * per), a_hi = min(n, a_lo + per);
    const int nrow = 3 * (a_hi - a_lo);
    const int span = KB * delta;
    const int nsr = (T - 2 + span) / span;                  // super-rows over k = 1 .. T-1, the last one zero-filled
    const int ntask = nsr * delta;
    const unsigned row_bytes = 8u * (unsigned)tp;
    const int d8 = 8 * delta;
    // this thread's tiles on series 0: q = tid, tid + blockDim, ...: (super-row s, residue r), stepped without a division
    const int s_first = (int)threadIdx.x / delta, r_first = (int)threadIdx.x - s_first * delta;
    const int s_step = (int)blockDim.x / delta, r_step = (int)blockDim.x - s_step * delta;
    auto issue = [&](int row, int b) {                      // one thread: bulk copy of series `row` of the block into buffer b
        const int a = perm[a_lo + row / 3];
        const char *src = reinterpret_cast<const char *>(P + ((size_t)a * 3 + (size_t)(row % 3)) * tp);
        const unsigned dst = sm_addr + (unsigned)b * 8u * (unsigned)rowcap, mb = mbar0 + 8u * (unsigned)b;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_expect_tx(mb, row_bytes);
        for (unsigned o = 0; o < row_bytes; o += 32768u) bulk_g2s(dst + o, src + o, min(32768u, row_bytes - o), mb);
    };
    for (int w0 = 0, pass = 0; w0 < nw; w0 += NWT, ++pass) {
        const int lag8 = w0 * d8;
        const int pk0 = pieces[(size_t)pass * (blockDim.x + 1) + threadIdx.x], pk1 = pieces[(size_t)pass * (blockDim.x + 1) + threadIdx.x + 1];
        double acc[NWT];
#pragma unroll
        for (int w = 0; w < NWT; ++w) acc[w] = 0.0;
        double sq = 0.0;
        int cur_sp = -1;
        if ((int)threadIdx.x < nbuf && (int)threadIdx.x < nrow) issue(threadIdx.x, threadIdx.x);   // all buffers are free: the previous pass ended with a barrier
        int b = 0, gen = 0;                                 // buffer of the series, and how many series of this pass it has held before
        int rot = 0;                                        // super-rows the tile map has moved on (series index modulo nsr)
        for (int row = 0; row <= nrow; ++row) {
            if (row % 3 == 0) {
                // the block's atoms are visited in species order (perm: a stable sort of [a_lo, a_hi) by species), so the register
                // sums are reduced at most S times per pass.  Every thread gets here at the same series: the barriers match.
                const int sp = row < nrow ? (int)species[perm[a_lo + row / 3]] : -2;
                if (sp != cur_sp) {
                    if (cur_sp >= 0) {
#pragma unroll
                        for (int w = 0; w < NWT; ++w) {
                            double v = acc[w];
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                            if (lane == 0) s_red[warp * NWT + w] = v;
                            acc[w] = 0.0;
                        }
                        s_part[threadIdx.x] = sq;
                        sq = 0.0;
                        __syncthreads();
                        if ((int)threadIdx.x < NWT && w0 + (int)threadIdx.x < nw) {
                            const int wi = w0 + (int)threadIdx.x, m = wi * delta;
                            double t = 0.0;
                            for (int q = 0; q < nwarp; ++q) t += s_red[q * NWT + threadIdx.x];
                            s_acc[cur_sp * nw + wi] += t;
                            if (wi > 0 && m < T - 1) {
                                double ss = 0.0;            // whole ranges: k >= m + 1 (the later frame of a pair), k <= T - 1 - m (the earlier one)
                                const int *pk = pieces + (size_t)pass * (blockDim.x + 1);
                                for (int q = 0; q < (int)blockDim.x; ++q) {
                                    const int k0 = pk[q], k1 = pk[q + 1];
                                    if (k1 > k0) {
                                        const double v = s_part[q];
                                        if (k0 >= m + 1) ss += v;
                                        if (k1 <= T - m) ss += v;
                                    }
                                }
                                s_ss[cur_sp * nw + wi] += ss;
                            }
                        }
                        __syncthreads();
                    }
                    cur_sp = sp;
                }
            }
            if (row >= nrow) break;
            {   // completed phases of buffer b before this series: those of this pass, plus all of the earlier passes
                const int before = gen + pass * ((nrow - b + nbuf - 1) / nbuf);
#if MSD_WIDE_WAIT_NS > 0
                mbar_wait_hint(mbar0 + 8u * (unsigned)b, (unsigned)(before & 1), MSD_WIDE_WAIT_NS);
#else
                mbar_wait(mbar0 + 8u * (unsigned)b, (unsigned)(before & 1));
#endif
            }
            const unsigned xb = sm_addr + (unsigned)b * 8u * (unsigned)rowcap;
            const double *x = sm + (size_t)b * rowcap;
            {   // four partial sums: one chain of dependent FMAs would cost its full latency per frame
                double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
                int k = pk0;
                for (; k + 3 < pk1; k += 4) {
                    q0 = __fma_rn(x[k], x[k], q0); q1 = __fma_rn(x[k + 1], x[k + 1], q1);
                    q2 = __fma_rn(x[k + 2], x[k + 2], q2); q3 = __fma_rn(x[k + 3], x[k + 3], q3);
                }
                for (; k < pk1; ++k) q0 = __fma_rn(x[k], x[k], q0);
                sq += (q0 + q1) + (q2 + q3);
            }
            int sq_ = s_first + rot, r = r_first;           // (s + rot) mod nsr: the wrap is applied per tile below
            for (int q = threadIdx.x; q < ntask; q += blockDim.x) {
                const int sw = sq_ >= nsr ? sq_ - nsr : sq_;
                const int e = sw * KB - w0;
                const int off = 8 * (1 + sw * span + r);
                const unsigned sb = xb + (unsigned)off;
                const char *sp = reinterpret_cast<const char *>(x) + off;
                if (w0 == 0) msd_wide_first<KB, NWT, 0>(sw, sb, sp, d8, acc);
                else if (e >= NWT - 1) msd_wide_tile<KB, NWT, false, NWT - 1>(sb, sp, d8, lag8, 0, acc);
                else if (e > -KB) msd_wide_tile<KB, NWT, false, -1>(sb, sp, d8, lag8, e, acc);
                sq_ += s_step; r += r_step;
                if (r >= delta) { r -= delta; ++sq_; }
            }
            // this warp has left buffer b; the last warp out refills it with the series nbuf further on
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                if (atomicAdd(&s_done[b], 1u) == (unsigned)(nwarp - 1)) {
                    s_done[b] = 0u;
                    __threadfence_block();
                    if (row + nbuf < nrow) issue(row + nbuf, b);
                }
            }
            if (++b == nbuf) { b = 0; ++gen; }
            if (++rot == nsr) rot = 0;
        }
        __syncthreads();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * S * nw; i += blockDim.x) partial[(size_t)blockIdx.x * 2 * S * nw + i] = s_acc[i];
}

// P: prepared series, atom-major, three arrays of tp doubles per atom.  Window lengths 0, delta, 2 delta, ...
// partial[block][2][S][nw]: [0] the pair sums (cross terms when DOT, squared differences otherwise), [1] when DOT the
// sums of |R_k|^2 + |R_{k-m}|^2 over the same pairs; the caller forms [1] - 2 [0].
template <int KB, int NWT, bool DOT>
__global__ void __launch_bounds__(MSD_SOA_THREADS, MSD_SOA_MIN_BLOCKS) k_msd_window_soa(const double *__restrict__ P, const uint8_t *__restrict__ species,
                                                                                      const int *__restrict__ perm, int n, int T,
                                                                                      int tp, int delta, int nw, int ng, int S, double *__restrict__ partial) {
    extern __shared__ __align__(16) double sm[];
    double *s_acc = sm + (size_t)tp;                        // [S][nw]
    double *s_ss = s_acc + (size_t)S * nw;                  // [S][nw]
    double *s_red = s_ss + (size_t)S * nw;                  // [nwarp][NWT]
    double *s_q = s_red + (size_t)MSD_AP_NWT_MAX * (MSD_SOA_THREADS / 32);      // [blockDim + 32]: exclusive prefix of the per-thread sums of squares
    __shared__ __align__(8) unsigned long long s_mbar;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int grp = warp % ng, wpg = nwarp / ng;            // my window group; warps per group
    const int tg = (warp / ng) * 32 + lane, gs = wpg * 32;  // my index among the group's threads
    for (int i = threadIdx.x; i < 2 * S * nw; i += blockDim.x) s_acc[i] = 0.0;      // s_acc and s_ss are adjacent
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(&s_mbar);
    const unsigned sm_addr = (unsigned)__cvta_generic_to_shared(sm);
    if (threadIdx.x == 0) mbar_init(mbar, 1);
    unsigned phase = 0;
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int a_lo = blockIdx.x * per, a_hi = min(n, a_lo + per);
    const int span = KB * delta;
    const int nsr = (T - 1) / span;                         // full super-rows over k = 1 .. T-1
    const int ntask = nsr * delta;
    const int k_rem = 1 + nsr * span;                       // first frame of the remainder
    const int C = ((T + (int)blockDim.x - 1) / (int)blockDim.x) | 1;     // frames per thread in the prefix of squares (odd: conflict-free)
    const unsigned row_bytes = 8u * (unsigned)tp;
    for (int w0 = 0; w0 < nw; w0 += ng * NWT) {
        const int wlo = w0 + grp * NWT;                     // first window of my group in this pass
        const bool mine = wlo < nw;
        const int lag0 = wlo * delta;
        const int lag_far = lag0 + (NWT - 1) * delta;       // the longest lag of my group
        double acc[NWT];
#pragma unroll
        for (int w = 0; w < NWT; ++w) acc[w] = 0.0;
        int cur_sp = -1;
        for (int ai = a_lo; ai <= a_hi; ++ai) {
            // the block's atoms are visited in species order (perm: a stable sort of [a_lo, a_hi) by species), so the
            // register sums are reduced at most S times per pass instead of at every change along the atom order
            const int a = ai < a_hi ? perm[ai] : -1;
            const int sp = ai < a_hi ? (int)species[a] : -2;
            if (sp != cur_sp) {
                if (cur_sp >= 0) {                          // species changed (or done): reduce the register sums
#pragma unroll
                    for (int w = 0; w < NWT; ++w) {
                        double v = acc[w];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                        if (lane == 0) s_red[warp * NWT + w] = v;
                        acc[w] = 0.0;
                    }
                    __syncthreads();
                    if (threadIdx.x < ng * NWT) {           // thread -> (group, window of the group); warps of a group in order
                        const int g = threadIdx.x / NWT, w = threadIdx.x - g * NWT, wi = w0 + g * NWT + w;
                        if (wi < nw) {
                            double t = 0.0;
                            for (int q = 0; q < wpg; ++q) t += s_red[(q * ng + g) * NWT + w];
                            s_acc[cur_sp * nw + wi] += t;
                        }
                    }
                    __syncthreads();
                }
                cur_sp = sp;
            }
            if (ai >= a_hi) break;
            const char *rec = reinterpret_cast<const char *>(P + (size_t)a * 3 * tp);
            for (int comp = 0; comp < 3; ++comp) {
                __syncthreads();                            // everyone finished the previous component
                if (threadIdx.x == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive_expect_tx(mbar, row_bytes);
                    const char *src = rec + (size_t)comp * row_bytes;
                    for (unsigned o = 0; o < row_bytes; o += 32768u) bulk_g2s(sm_addr + o, src + o, min(32768u, row_bytes - o), mbar);
                    mbar_wait(mbar, phase);
                }
                phase ^= 1u;
                {                                           // pull what comes next into L2 while this component is worked on
                    const char *nx = comp < 2 ? rec + (size_t)(comp + 1) * row_bytes
                                              : (ai + 1 < a_hi ? reinterpret_cast<const char *>(P + (size_t)perm[ai + 1] * 3 * tp) : nullptr);
                    if (nx)
                        for (int i = threadIdx.x; i < (int)((row_bytes + 127u) >> 7); i += blockDim.x)
                            asm volatile("prefetch.global.L2 [%0];" :: "l"(nx + ((size_t)i << 7)));
                }
                __syncthreads();
                if (DOT && w0 == 0) {
                    // Q(k) = sum_{j<=k} v_j^2 (v_0 = 0): per-thread sums over C consecutive frames, block-wide exclusive prefix,
                    // then one thread per window adds  sum_{k=m+1}^{T-1} v_k^2 + sum_{j=1}^{T-1-m} v_j^2 = (Q(T-1) - Q(m)) + Q(T-1-m)
                    const int k0 = threadIdx.x * C, k1 = min(T, k0 + C);
                    double q = 0.0;
                    for (int k = k0; k < k1; ++k) q = __fma_rn(sm[k], sm[k], q);
                    const double iq = warp_incl_scan(q, lane);
                    if (lane == 31) s_q[blockDim.x + warp] = iq;
                    __syncthreads();
                    const double tq = lane < nwarp ? s_q[blockDim.x + lane] : 0.0;
                    const double wq = warp_incl_scan(tq, lane);
                    s_q[threadIdx.x] = __shfl_sync(0xffffffffu, wq - tq, warp) + (iq - q);
                    __syncthreads();
                    for (int wi = threadIdx.x; wi < nw; wi += blockDim.x) {
                        const int m = wi * delta;
                        auto Q = [&](int k) {               // inclusive prefix at frame k
                            const int t = k / C;
                            double v = s_q[t];
                            for (int j = t * C; j <= k; ++j) v = __fma_rn(sm[j], sm[j], v);
                            return v;
                        };
                        if (wi > 0 && m < T - 1) s_ss[cur_sp * nw + wi] += (Q(T - 1) - Q(m)) + Q(T - 1 - m);
                    }
                }
                if (mine) {
                    for (int q = tg; q < ntask; q += gs) {
                        const int s = q / delta, r = q - s * delta;
                        const int b = 1 + s * span + r;
                        if (b - lag_far >= 1) msd_soa_tile<KB, NWT, false, DOT>(sm, b, delta, lag0, acc);
                        else msd_soa_tile<KB, NWT, true, DOT>(sm, b, delta, lag0, acc);
                    }
                    for (int k = k_rem + tg; k < T; k += gs) {
                        if (k - lag_far >= 1) msd_soa_tile<1, NWT, false, DOT>(sm, k, delta, lag0, acc);
                        else msd_soa_tile<1, NWT, true, DOT>(sm, k, delta, lag0, acc);
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * S * nw; i += blockDim.x) partial[(size_t)blockIdx.x * 2 * S * nw + i] = s_acc[i];
}
