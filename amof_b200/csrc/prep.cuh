// prep.cuh -- GPU-built linked-cell list for a batch of frames (kernel K1 of SURVEY.md 2.1).
//
// Three launches per batch, all O(atoms):
//   k_cell_assign   wrap every atom into its cell (P1, P2), find its linked cell, take a slot in it
//   k_cell_scan     one block per frame: exclusive scan of the cell populations
//   k_cell_scatter  write the wrapped atoms, ordered by cell, as 32-byte records {x, y, z, species}
// The order of atoms inside a cell depends on atomic scheduling; every consumer only counts, so the
// integers it produces do not.
#pragma once
#include "common.cuh"

struct __align__(16) SAtom {
    double x, y, z;
    long long s;   // species | linked-cell coordinates (8 bytes wide so one atom is two 16-byte loads)
};

// one sorted atom = two read-only 16-byte loads
__device__ __forceinline__ SAtom load_satom(const SAtom *p) {
    const double2 *q = reinterpret_cast<const double2 *>(p);
    double2 a = __ldg(q), b = __ldg(q + 1);
    SAtom s;
    s.x = a.x; s.y = a.y; s.z = b.x; s.s = __double_as_longlong(b.y);
    return s;
}

struct PrepArgs {
    const double *raw;        // [F][N][3]
    const FrameGeom *geom;    // [F]
    const uint8_t *species;   // [N]
    uint32_t *cell_count;     // [sum(ncell+1)]
    uint32_t *cell_start;     // [sum(ncell+1)]
    uint32_t *cid;            // [F*N]
    uint32_t *rank;           // [F*N]
    SAtom *sorted;            // [F*N]
    int n_atoms;
    int n_frames;
    // optional filter (bond angles): only atoms with species_keep[species] != 0 enter the cell list; the others can
    // neither be a centre nor a neighbour under the cutoff matrix, so the sorted frame holds cell_start[ncell] atoms
    const uint8_t *species_keep;   // [n_species] or nullptr
    const int *keep_idx;           // with the filter: [n_keep] original indices of the atoms that pass it (the kernels then run
    int n_keep;                    //   over n_frames * n_keep items instead of testing every atom of every frame)
    uint32_t *orig;                // optional [F*N]: original atom index of every sorted atom (explicit neighbour lists)
    int *wraps;                    // optional [F*N][3]: cell translations removed from every sorted atom by P2
    uint32_t *slot;                // optional [F*N]: position of every atom in its frame's sorted order (pair lists)
    // optional (bond angles): a compact list of the atoms that can be the centre of an angle, so that the search kernel has a
    // thread per CENTRE instead of one per kept atom: centre_rank[n_keep] = rank among the centres of a frame or -1,
    // centre_list[F * n_centres] receives frame * n_keep + position in the sorted frame
    const int *centre_rank;
    unsigned *centre_list;
    int n_centres;
    // optional (bond angles): one cell list PER SPECIES, laid end to end: list_of[species] = which list an atom goes to; the cell
    // index becomes list * ncell + cell and cell_start holds n_lists * ncell + 1 entries per frame.  A search then only visits
    // the lists of the species it has a positive cutoff with.  n_lists = 1 and list_of = 0 everywhere else.
    int n_lists;
    uint8_t list_of[AMOFB_MAX_SPECIES];
    // raw holds only the kept atoms of every frame, [F][n_keep][3] in keep_idx order (gathered on the host before the copy)
    int raw_compact;
};

// wn (optional): the integer cell translations w_k = floor(f_k) that P2 removed
__device__ __forceinline__ void wrap_atom(const FrameGeom &g, const double *__restrict__ p, double *pw, int *c, int *wn = nullptr) {
    double px = p[0], py = p[1], pz = p[2];
    double f[3], w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        f[k] = (px * g.inv[0 + k] + py * g.inv[3 + k]) + pz * g.inv[6 + k];   // P1
        w[k] = floor(f[k]);
    }
    pw[0] = px - ((w[0] * g.cell[0] + w[1] * g.cell[3]) + w[2] * g.cell[6]);  // P2
    pw[1] = py - ((w[0] * g.cell[1] + w[1] * g.cell[4]) + w[2] * g.cell[7]);
    pw[2] = pz - ((w[0] * g.cell[2] + w[1] * g.cell[5]) + w[2] * g.cell[8]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double gk = f[k] - w[k];                      // in [0,1) for finite input
        int ck = (gk >= 0.0) ? (int)(gk * (double)g.nc[k]) : 0;
        if (ck > g.nc[k] - 1) ck = g.nc[k] - 1;
        c[k] = ck;
    }
    if (wn) { wn[0] = (int)w[0]; wn[1] = (int)w[1]; wn[2] = (int)w[2]; }
}

__global__ void __launch_bounds__(256) k_cell_assign(PrepArgs a) {
    const int per = a.keep_idx ? a.n_keep : a.n_atoms;              // items per frame
    long long total = (long long)a.n_frames * per;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int f = (int)(idx / per);
        int i = (int)(idx - (long long)f * per);
        if (a.keep_idx) i = a.keep_idx[i];
        else if (a.species_keep && !a.species_keep[a.species[i]]) { a.cid[idx] = 0xffffffffu; continue; }
        const FrameGeom &g = a.geom[f];
        double pw[3];
        int c[3];
        wrap_atom(g, a.raw + 3 * (a.raw_compact ? (long long)f * per + (idx - (long long)f * per) : (long long)f * a.n_atoms + i), pw, c);
        uint32_t cid = (uint32_t)((c[0] * g.nc[1] + c[1]) * g.nc[2] + c[2]);
        if (a.n_lists > 1) cid += (uint32_t)a.list_of[a.species[i]] * (uint32_t)g.ncell;
        a.cid[idx] = cid;
        a.rank[idx] = atomicAdd(&a.cell_count[g.cs_off + cid], 1u);
    }
}

// one block per frame; writes ncell+1 entries (last = number of atoms).  Each thread scans SCAN_PER consecutive
// cells serially, every warp publishes its total, and after ONE barrier per round each warp scans the 32 warp totals
// itself (redundantly), so the running carry lives in a register of every thread: no second barrier, no shared carry.
#define SCAN_PER 8
__global__ void __launch_bounds__(1024) k_cell_scan(PrepArgs a) {
    __shared__ uint32_t warp_sums[2][32];      // double-buffered: round r+1 may publish while a slow warp still reads round r
    int f = blockIdx.x;
    const FrameGeom &g = a.geom[f];
    const uint32_t *cnt = a.cell_count + g.cs_off;
    uint32_t *out = a.cell_start + g.cs_off;
    int n = g.ncell * a.n_lists;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = blockDim.x >> 5;
    uint32_t carry = 0;
    int buf = 0;
    for (int base = 0; base < n; base += blockDim.x * SCAN_PER, buf ^= 1) {
        const int i0 = base + threadIdx.x * SCAN_PER;
        uint32_t loc[SCAN_PER];
        uint32_t v = 0;
        if (i0 + SCAN_PER <= n && (reinterpret_cast<uintptr_t>(cnt + i0) & 15) == 0) {
            const uint4 q0 = *reinterpret_cast<const uint4 *>(cnt + i0), q1 = *reinterpret_cast<const uint4 *>(cnt + i0 + 4);
            loc[0] = q0.x; loc[1] = q0.y; loc[2] = q0.z; loc[3] = q0.w; loc[4] = q1.x; loc[5] = q1.y; loc[6] = q1.z; loc[7] = q1.w;
        } else {
#pragma unroll
            for (int k = 0; k < SCAN_PER; ++k) loc[k] = (i0 + k < n) ? cnt[i0 + k] : 0u;
        }
#pragma unroll
        for (int k = 0; k < SCAN_PER; ++k) v += loc[k];
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[buf][warp] = incl;
        __syncthreads();
        const uint32_t ws = lane < nwarp ? warp_sums[buf][lane] : 0u;
        uint32_t wi = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        const uint32_t before = __shfl_sync(0xffffffffu, wi - ws, warp);      // total of the warps before mine
        const uint32_t round_total = __shfl_sync(0xffffffffu, wi, 31);
        uint32_t run = carry + before + incl - v;
        if (i0 + SCAN_PER <= n && (reinterpret_cast<uintptr_t>(out + i0) & 15) == 0) {
            uint4 q0, q1;
            q0.x = run; run += loc[0]; q0.y = run; run += loc[1]; q0.z = run; run += loc[2]; q0.w = run; run += loc[3];
            q1.x = run; run += loc[4]; q1.y = run; run += loc[5]; q1.z = run; run += loc[6]; q1.w = run;
            *reinterpret_cast<uint4 *>(out + i0) = q0;
            *reinterpret_cast<uint4 *>(out + i0 + 4) = q1;
        } else {
#pragma unroll
            for (int k = 0; k < SCAN_PER; ++k) {
                if (i0 + k < n) out[i0 + k] = run;
                run += loc[k];
            }
        }
        carry += round_total;
    }
    if (threadIdx.x == 0) out[n] = carry;
}

__global__ void __launch_bounds__(256) k_cell_scatter(PrepArgs a) {
    const int per = a.keep_idx ? a.n_keep : a.n_atoms;
    long long total = (long long)a.n_frames * per;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int f = (int)(idx / per);
        int i = (int)(idx - (long long)f * per);
        const int kk = i;
        if (a.keep_idx) i = a.keep_idx[i];
        if (a.cid[idx] == 0xffffffffu) continue;          // filtered out by species_keep
        const FrameGeom &g = a.geom[f];
        double pw[3];
        int c[3], wn[3];
        wrap_atom(g, a.raw + 3 * (a.raw_compact ? (long long)f * per + kk : (long long)f * a.n_atoms + i), pw, c, wn);
        uint32_t dst = a.cell_start[g.cs_off + a.cid[idx]] + a.rank[idx];
        if (a.centre_list) {
            const int cr = a.centre_rank[kk];
            if (cr >= 0) a.centre_list[(long long)f * a.n_centres + cr] = (unsigned)((long long)f * per + dst);
        }
        SAtom s;
        s.x = pw[0]; s.y = pw[1]; s.z = pw[2];
        // species | c0 << 8 | c1 << 20 | c2 << 32   (nc <= 1024 per axis)
        s.s = (long long)a.species[i] | ((long long)c[0] << 8) | ((long long)c[1] << 20) | ((long long)c[2] << 32);
        a.sorted[(long long)f * a.n_atoms + dst] = s;
        if (a.orig) a.orig[(long long)f * a.n_atoms + dst] = (uint32_t)i;
        if (a.slot) a.slot[(long long)f * a.n_atoms + i] = dst;
        if (a.wraps) {
            int *wp = a.wraps + 3 * ((long long)f * a.n_atoms + dst);
            wp[0] = wn[0]; wp[1] = wn[1]; wp[2] = wn[2];
        }
    }
}
