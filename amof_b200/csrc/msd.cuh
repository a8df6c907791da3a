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

__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// one warp per atom; PREPARE subtracts com[k] first (translate(-cg)), UNWRAP does not
template <bool PREPARE>
__global__ void __launch_bounds__(256) k_msd_scan(double *__restrict__ P, const MsdGeom *__restrict__ geom,
                                                  const double *__restrict__ com, int n, int T) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long a = warp; a < n; a += nwarps) {
        double *p = P + (size_t)a * T * 3;
        double cx = 0.0, cy = 0.0, cz = 0.0;      // running sum carried between 32-frame chunks
        double lx = 0.0, ly = 0.0, lz = 0.0;      // original (shifted) position of the last frame of the previous chunk
        for (int k0 = 0; k0 < T; k0 += 32) {
            const int k = k0 + lane;
            double x = 0.0, y = 0.0, z = 0.0;
            if (k < T) {
                x = p[3 * (size_t)k]; y = p[3 * (size_t)k + 1]; z = p[3 * (size_t)k + 2];
                if (PREPARE) { x -= com[3 * k]; y -= com[3 * k + 1]; z -= com[3 * k + 2]; }
            }
            double px = __shfl_up_sync(0xffffffffu, x, 1), py = __shfl_up_sync(0xffffffffu, y, 1), pz = __shfl_up_sync(0xffffffffu, z, 1);
            if (lane == 0) { px = lx; py = ly; pz = lz; }
            double dx = 0.0, dy = 0.0, dz = 0.0;
            if (k < T) {
                if (k == 0) { dx = x; dy = y; dz = z; }                       // delta_0 = first positions
                else wrap_disp(geom[k - 1], x - px, y - py, z - pz, dx, dy, dz);   // cell of frame k-1 wraps step k-1 -> k
            }
            double sx = warp_incl_scan(dx, lane) + cx, sy = warp_incl_scan(dy, lane) + cy, sz = warp_incl_scan(dz, lane) + cz;
            if (k < T) { p[3 * (size_t)k] = sx; p[3 * (size_t)k + 1] = sy; p[3 * (size_t)k + 2] = sz; }
            const int last = min(31, T - 1 - k0);
            cx = __shfl_sync(0xffffffffu, sx, last); cy = __shfl_sync(0xffffffffu, sy, last); cz = __shfl_sync(0xffffffffu, sz, last);
            lx = __shfl_sync(0xffffffffu, x, last); ly = __shfl_sync(0xffffffffu, y, last); lz = __shfl_sync(0xffffffffu, z, last);
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
// partial[block][S][nw] += sum over the block's atoms of species s, over k = m+1..T-1, of |R_k - R_{k-m}|^2
template <bool SMEM>
__global__ void __launch_bounds__(512) k_msd_window(const double *__restrict__ P, const uint8_t *__restrict__ species, int n, int T,
                                                    const int *__restrict__ window, int nw, int S, double *__restrict__ partial) {
    extern __shared__ double sm[];
    double *sx = sm, *sy = sm + (SMEM ? T : 0), *sz = sm + (SMEM ? 2 * T : 0);
    double *s_acc = sm + (SMEM ? 3 * (size_t)T : 0);    // [S][nw] block accumulators
    double *s_red = s_acc + (size_t)S * nw;             // [warps]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int i = threadIdx.x; i < S * nw; i += blockDim.x) s_acc[i] = 0.0;
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int a_lo = blockIdx.x * per, a_hi = min(n, a_lo + per);
    for (int a = a_lo; a < a_hi; ++a) {
        const double *p = P + (size_t)a * T * 3;
        __syncthreads();
        if (SMEM)
            for (int i = threadIdx.x; i < 3 * T; i += blockDim.x) {
                const double v = p[i];
                const int k = i / 3, c = i - 3 * k;
                (c == 0 ? sx : c == 1 ? sy : sz)[k] = v;
            }
        __syncthreads();
        const int sp = species[a];
        for (int w = 0; w < nw; ++w) {
            const int m = window[w];
            double acc = 0.0;
            if (m >= 0 && m < T)
                for (int k = m + 1 + threadIdx.x; k < T; k += blockDim.x) {
                    double dx, dy, dz;
                    if (SMEM) { dx = sx[k] - sx[k - m]; dy = sy[k] - sy[k - m]; dz = sz[k] - sz[k - m]; }
                    else {
                        dx = p[3 * (size_t)k] - p[3 * (size_t)(k - m)];
                        dy = p[3 * (size_t)k + 1] - p[3 * (size_t)(k - m) + 1];
                        dz = p[3 * (size_t)k + 2] - p[3 * (size_t)(k - m) + 2];
                    }
                    acc += __fma_rn(dz, dz, __fma_rn(dy, dy, dx * dx));
                }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
            if (lane == 0) s_red[w * nwarp + warp] = acc;
        }
        __syncthreads();
        for (int w = threadIdx.x; w < nw; w += blockDim.x) {
            double s = 0.0;
            for (int q = 0; q < nwarp; ++q) s += s_red[w * nwarp + q];
            s_acc[sp * nw + w] += s;
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
