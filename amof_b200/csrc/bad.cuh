// bad.cuh -- bond-angle triplet kernels (K4 of SURVEY.md 2.1).
//
// Two kernels per batch of frames:
//   k_bad_search  one THREAD per atom of the species-filtered, cell-sorted frame: walks the full stencil and keeps every
//                 neighbour under the pair cutoffs (P5).  An atom that can be the centre of an angle (>= 2 neighbours,
//                 centre of some requested triple) appends its neighbours' unit vectors (P6) as one contiguous run of
//                 a batch-wide pool and one record to a compact centre list.  No shared memory, 64 registers: the walk is
//                 latency-bound (cell_start -> candidates), so occupancy is what matters.
//   k_bad_angles  one thread per CENTRE of the compact list (on a ZIF 'Zn-N' analysis four atoms out of five are N with
//                 a single Zn neighbour and never get here, so every lane has the same handful of pairs to do): all
//                 unordered pairs of its neighbours, x = u_p . u_q, exact bin, histogram increment.  The threshold table
//                 lives in shared memory and so do the histogram rows in use: a block owns BAD_SLOTS rows [nbins] u32,
//                 claimed on first use for a (triple, coordination number) pair and merged into the global u64 histogram
//                 once per launch; rows beyond that (rare coordination numbers) fall through to global atomics.
//
// The angle itself is never formed on the device: x is computed in fp64 in the oracle's operation order (P6), clipped to
// [-1, 1] as ase.geometry.get_angles does, and located in a table of thresholds on -x that the host bisected with its own
// libm acos and the np.histogram edge rule (P7), so the bin is the one the CPU path takes, bit for bit.
#pragma once
#include <cooperative_groups.h>
#include <cooperative_groups/scan.h>
#include "prep.cuh"
namespace cg = cooperative_groups;

#define BAD_NB_MAX 64        // neighbours of any species kept per centre
#define BAD_MAX_TRIPLES 64
#define BAD_SLOTS 4          // histogram rows a block keeps in shared memory
#define BAD_RANGES 18        // candidate ranges per atom on the fast path: 3 x 3 rows, at most two z runs

struct __align__(16) BadNb {          // one neighbour of a centre: image vector (unit vector once the angle kernel has normalised it) + species
    double ux, uy, uz;
    long long sp;
};

struct BadCentre {
    unsigned off;                     // first neighbour in the pool
    unsigned short nn;                // neighbours
    unsigned char sp, pad;            // species of the centre
};

struct BadArgs {
    const SAtom *sorted;
    const FrameGeom *geom;
    const uint32_t *cell_start;
    const double *cn_thr2;        // [nkeys]
    const uint16_t *keyidx;       // [S*S]
    const int2 *triples;          // [n_triples] (A, B), -1 = any
    const double *tthr;           // [nbins+2]: tthr[0] = -2, tthr[k] = smallest -x in bin >= k, tthr[nbins+1] = +2
    unsigned long long *hist;     // [n_triples][AMOFB_BAD_MAX_CN+1][nbins]
    unsigned long long *dropped;  // [n_triples]
    int *flags;                   // bit 0: neighbour overflow, bit 1: cn > AMOFB_BAD_MAX_CN, bit 2: neighbour pool overflow
    BadNb *pool;                  // [pool_cap]
    BadCentre *centres;           // [n_frames * n_keep]
    unsigned *counters;           // [0] centres, [1] pool entries
    unsigned pool_cap;
    int n_slots;                  // histogram rows a block keeps in shared memory (0: none fit)
    int tthr_smem;                // the threshold table fits in shared memory
    int n_keep;                   // atoms per frame in the (species-filtered) cell list
    const unsigned long long *centre_mask;               // [AMOFB_MAX_SPECIES] triples whose A matches this species
    const unsigned *centre_list;                         // optional [n_frames * n_centres]: frame * n_keep + sorted position of every possible centre
    int n_centres;
    int n_lists;                                         // cell lists per frame: one per kept species (1: a single list of all kept atoms)
    uint8_t list_of[AMOFB_MAX_SPECIES];                  // species -> its list
    uint16_t partner_mask[AMOFB_MAX_SPECIES];            // species with a positive cutoff to this one
    double r2search;
    float inv_dtheta_f;
    int n_atoms, n_frames, n_species, nkeys, n_triples, nbins;
};

// index K in [0, nbins] of t = -x: K < nbins is the histogram bin, K == nbins means "beyond the last edge"
__device__ __forceinline__ int bad_bin(double t, const double *__restrict__ tthr, float inv_dtheta_f, int nbins) {
    float xf = fminf(fmaxf((float)(-t), -1.0f), 1.0f);
    int k = (int)(acosf(xf) * 57.29577951308232f * inv_dtheta_f);
    k = k < 0 ? 0 : (k > nbins ? nbins : k);
    while (t < tthr[k]) --k;          // tthr[0] = -2 stops it
    while (t >= tthr[k + 1]) ++k;     // tthr[nbins+1] = +2 stops it
    return k;
}

#ifndef BAD_MIN_BLOCKS
#define BAD_MIN_BLOCKS 8      // 64 registers
#endif
__global__ void __launch_bounds__(128, BAD_MIN_BLOCKS) k_bad_search(BadArgs a) {
    long long t_id = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (a.centre_list) {            // every lane is a possible centre (on a ZIF 'Zn-N' analysis one kept atom in five is)
        if (t_id >= (long long)a.n_frames * a.n_centres) return;
        t_id = (long long)a.centre_list[t_id];
    }
    if (t_id >= (long long)a.n_frames * a.n_keep) return;
    const int f = (int)(t_id / a.n_keep);
    const int i = (int)(t_id - (long long)f * a.n_keep);
    const SAtom *fr = a.sorted + (long long)f * a.n_atoms;
    const SAtom me = load_satom(fr + i);
    const int si = (int)(me.s & 0xff);
    if (!__ldg(a.centre_mask + si)) return;
    const FrameGeom &G = a.geom[f];
    const uint32_t *cs = a.cell_start + G.cs_off;
    const int S = a.n_species;
    const int c0 = (int)((me.s >> 8) & 0xfff), c1 = (int)((me.s >> 20) & 0xfff), c2 = (int)((me.s >> 32) & 0xfff);
    const int nc0 = G.nc[0], nc1 = G.nc[1], nc2 = G.nc[2];
    const int m0 = G.m[0], m1 = G.m[1], m2 = G.m[2];

    // neighbours found: index in the sorted frame | image code << 24 (s in {-1,0,1}^3, 13 = home image); the vector is
    // re-formed from the two records when the centre turns out to need it
    unsigned nb[BAD_NB_MAX];
    int nn = 0;
    bool overflow = false;
    const double mex = me.x, mey = me.y, mez = me.z;
    const uint16_t *krow = a.keyidx + si * S;
    // the z window [c2 - m2, c2 + m2] is the same for every row of the stencil: split it once into its contiguous
    // runs (one per wrap of the column; more than three only happen in boxes narrower than the stencil)
    int zq[3], zl[3], zs[3], nz = 0;
    bool many_wraps = false;
    for (int d2 = -m2; d2 <= m2;) {
        int s2, q2;
        wrap_cell(c2 + d2, nc2, s2, q2);
        const int len = min(m2 - d2, nc2 - 1 - q2) + 1;
        if (nz < 3) { zq[nz] = q2; zl[nz] = len; zs[nz] = s2; ++nz; } else many_wraps = true;
        d2 += len;
    }
    // One cell list per species (n_lists > 1): only the lists of the species this atom has a positive cutoff with are visited --
    // on the ZIF 'Zn-N' analysis an N atom looks at the Zn list alone (one candidate in five of the mixed list).
    const int ncell = nc0 * nc1 * nc2;
    const unsigned pmask = a.n_lists > 1 ? (unsigned)a.partner_mask[si] : 1u;
    __shared__ int2 s_rng[BAD_RANGES][128];          // [range][thread]: first candidate, end | image code << 24
    for (unsigned left = pmask; left; left &= left - 1) {
        const int sj_list = __ffs(left) - 1;             // partner species (or 0: the single list)
        const uint32_t *csl = a.n_lists > 1 ? cs + (int)a.list_of[sj_list] * ncell : cs;
        const double thr_list = a.n_lists > 1 ? __ldg(a.cn_thr2 + krow[sj_list]) : 0.0;
        auto test = [&](int j, unsigned code) {
            const SAtom o = load_satom(fr + j);
            double dx = o.x - mex, dy = o.y - mey, dz = o.z - mez;
            if (code != (13u << 24)) {                  // P3: (pj - pi) + T, T = (s0*a + s1*b) + s2*c (x + 0.0 == x: the home image skips the adds)
                const int c = (int)(code >> 24), s0 = c % 3 - 1, s1 = (c / 3) % 3 - 1, s2 = c / 9 - 1;
                const double fs0 = (double)s0, fs1 = (double)s1, fs2 = (double)s2;
                dx += (fs0 * G.cell[0] + fs1 * G.cell[3]) + fs2 * G.cell[6];
                dy += (fs0 * G.cell[1] + fs1 * G.cell[4]) + fs2 * G.cell[7];
                dz += (fs0 * G.cell[2] + fs1 * G.cell[5]) + fs2 * G.cell[8];
            }
            const double dd = (dx * dx + dy * dy) + dz * dz;
            bool hit;
            if (a.n_lists > 1) hit = dd < thr_list;      // every atom of the list has the partner species
            else hit = dd < a.r2search && dd < __ldg(a.cn_thr2 + krow[(int)(o.s & 0xff)]);
            if (hit) {
                if (nn < BAD_NB_MAX) nb[nn++] = (unsigned)j | code;
                else overflow = true;
            }
        };
        if (!many_wraps && m0 == 1 && m1 == 1 && nz <= 2) {
            // Two steps.  A) every lane lists its candidate ranges -- (row, z run) -> [first, end) | image code -- in shared memory,
            // in lockstep (nine rows, one or two z runs each).  B) ONE flat loop over the candidates of all its ranges: a lane
            // opens its next range (two shared-memory words) when the current one is used up, so a warp runs for as long as its
            // busiest lane has candidates.  With nested per-row loops it ran, row by row, for the longest run of any lane (ncu:
            // 12.5 of 32 lanes active, 77 candidate iterations per warp for ~10 candidates per lane).
            const int tid = threadIdx.x;
            int cnt = 0;                                  // ranges that hold at least one atom (half of them are empty at ~0.3 atoms per cell)
#pragma unroll
            for (int row = 0; row < 9; ++row) {
                int s0, q0, s1, q1;
                wrap_cell(c0 + row / 3 - 1, nc0, s0, q0);
                wrap_cell(c1 + row % 3 - 1, nc1, s1, q1);
                const int rowbase = (q0 * nc1 + q1) * nc2;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    if (k < nz) {
                        const int first = (int)csl[rowbase + zq[k]], end = (int)csl[rowbase + zq[k] + zl[k]];
                        if (end > first) {
                            s_rng[cnt][tid] = make_int2(first, end | (((s0 + 1) + 3 * (s1 + 1) + 9 * (zs[k] + 1)) << 24));
                            ++cnt;
                        }
                    }
                }
            }
            int rg = 0, j = 0, je = 0;
            unsigned code = 0;
            for (;;) {
                if (j >= je) {                            // open the next range (never empty)
                    if (rg >= cnt) break;
                    const int2 e = s_rng[rg++][tid];
                    j = e.x; je = e.y & 0xffffff; code = (unsigned)e.y & 0xff000000u;
                }
                if (!(code == (13u << 24) && j == i)) test(j, code);      // skip the zero-shift self pair
                ++j;
            }
        } else {
            for (int d0 = -m0; d0 <= m0; ++d0) {
                int s0, q0;
                wrap_cell(c0 + d0, nc0, s0, q0);
                for (int d1 = -m1; d1 <= m1; ++d1) {
                    int s1, q1;
                    wrap_cell(c1 + d1, nc1, s1, q1);
                    const int rowbase = (q0 * nc1 + q1) * nc2;
                    for (int d2 = -m2; d2 <= m2;) {
                        int s2, q2;
                        wrap_cell(c2 + d2, nc2, s2, q2);
                        const int len = min(m2 - d2, nc2 - 1 - q2) + 1;
                        const unsigned code = (unsigned)((s0 + 1) + 3 * (s1 + 1) + 9 * (s2 + 1)) << 24;
                        for (int j = (int)csl[rowbase + q2]; j < (int)csl[rowbase + q2 + len]; ++j)
                            if (!(code == (13u << 24) && j == i)) test(j, code);
                        d2 += len;
                    }
                }
            }
        }
    }
    if (overflow) { atomicOr(a.flags, 1); return; }
    if (nn < 2) return;
    // a centre: reserve a run of the pool and write the P3 image vectors (re-formed with the same operations on the same
    // operands: the same doubles as in the walk); they are normalised by the angle kernel, where every lane is a centre
    // one reservation per warp: the centres of a warp take consecutive runs of the pool and consecutive records of the centre list
    // (two same-address atomics per CENTRE -- 1.7 M per batch on C4 -- would be the kernel's longest queue)
    cg::coalesced_group grp = cg::coalesced_threads();
    const unsigned before = cg::exclusive_scan(grp, (unsigned)nn);
    unsigned off = 0, ci = 0;
    if (grp.thread_rank() == grp.size() - 1) {
        off = atomicAdd(a.counters + 1, before + (unsigned)nn);
        ci = atomicAdd(a.counters, (unsigned)grp.size());
    }
    off = grp.shfl(off, grp.size() - 1) + before;
    ci = grp.shfl(ci, grp.size() - 1) + grp.thread_rank();
    if (off + (unsigned)nn > a.pool_cap) { atomicOr(a.flags, 4); nn = 0; off = 0; }      // the record below stays valid (no neighbours); finish reports the overflow
    for (int p = 0; p < nn; ++p) {
        const int j = (int)(nb[p] & 0xffffffu), code = (int)(nb[p] >> 24);
        const SAtom o = load_satom(fr + j);
        double dx = o.x - mex, dy = o.y - mey, dz = o.z - mez;
        if (code != 13) {
            const int s0 = code % 3 - 1, s1 = (code / 3) % 3 - 1, s2 = code / 9 - 1;
            const double fs0 = (double)s0, fs1 = (double)s1, fs2 = (double)s2;
            dx += (fs0 * G.cell[0] + fs1 * G.cell[3]) + fs2 * G.cell[6];
            dy += (fs0 * G.cell[1] + fs1 * G.cell[4]) + fs2 * G.cell[7];
            dz += (fs0 * G.cell[2] + fs1 * G.cell[5]) + fs2 * G.cell[8];
        }
        BadNb r;
        r.ux = dx; r.uy = dy; r.uz = dz; r.sp = (long long)(o.s & 0xff);
        a.pool[off + p] = r;
    }
    BadCentre c;
    c.off = off; c.nn = (unsigned short)nn; c.sp = (unsigned char)si; c.pad = 0;
    a.centres[ci] = c;
}

__global__ void __launch_bounds__(256) k_bad_angles(BadArgs a) {
    extern __shared__ __align__(16) unsigned char bad_sm[];
    double *s_tthr = reinterpret_cast<double *>(bad_sm);                                  // [nbins + 2] when it fits
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(s_tthr + (a.tthr_smem ? a.nbins + 2 : 0));     // [n_slots][nbins]
    __shared__ int s_slot_key[BAD_SLOTS];          // (triple * (MAX_CN+1) + cn) of the row a slot holds, -1 = free
    const int n_slots = a.n_slots;
    if (a.tthr_smem)
        for (int k = threadIdx.x; k < a.nbins + 2; k += blockDim.x) s_tthr[k] = a.tthr[k];
    for (int k = threadIdx.x; k < n_slots * a.nbins; k += blockDim.x) s_hist[k] = 0u;
    if (threadIdx.x < BAD_SLOTS) s_slot_key[threadIdx.x] = -1;
    __syncthreads();
    const double *tthr = a.tthr_smem ? s_tthr : a.tthr;
    const unsigned n_cent = a.counters[0];
    for (unsigned ci = blockIdx.x * blockDim.x + threadIdx.x; ci < n_cent; ci += gridDim.x * blockDim.x) {
        const BadCentre c = a.centres[ci];
        const unsigned long long mine = __ldg(a.centre_mask + c.sp);
        BadNb *nb = a.pool + c.off;
        const int nn = c.nn;
        // P6: u = v / |v| componentwise, once per neighbour, in place (the run belongs to this thread alone)
        for (int p = 0; p < nn; ++p) {
            BadNb v = nb[p];
            const double n = sqrt((v.ux * v.ux + v.uy * v.uy) + v.uz * v.uz);
            v.ux = v.ux / n; v.uy = v.uy / n; v.uz = v.uz / n;
            nb[p] = v;
        }
        for (int t = 0; t < a.n_triples; ++t) {
            if (!((mine >> t) & 1ull)) continue;
            const int B = a.triples[t].y;
            int cn = 0;
            for (int p = 0; p < nn; ++p) cn += (B < 0 || (int)nb[p].sp == B);
            if (cn < 2) continue;
            if (cn > AMOFB_BAD_MAX_CN) { atomicOr(a.flags, 2); continue; }
            // the shared-memory row of (t, cn), claimed on first use; -1: all slots taken by other rows -> global atomics
            const int key = t * (AMOFB_BAD_MAX_CN + 1) + cn;
            int slot = -1;
            for (int s = 0; s < n_slots && slot < 0; ++s) {
                int cur = s_slot_key[s];
                if (cur == -1) cur = atomicCAS(&s_slot_key[s], -1, key) == -1 ? key : s_slot_key[s];
                if (cur == key) slot = s;
            }
            uint32_t *srow = slot >= 0 ? s_hist + slot * a.nbins : nullptr;
            unsigned long long *grow = a.hist + (size_t)key * a.nbins;
            for (int p = 0; p < nn; ++p) {
                const BadNb up = nb[p];
                if (!(B < 0 || (int)up.sp == B)) continue;
                for (int q = p + 1; q < nn; ++q) {
                    const BadNb uq = nb[q];
                    if (!(B < 0 || (int)uq.sp == B)) continue;
                    double x = (up.ux * uq.ux + up.uy * uq.uy) + up.uz * uq.uz;
                    if (x != x) { atomicAdd(a.dropped + t, 1ull); continue; }   // NaN (coincident atoms): np.histogram drops it
                    x = x < -1.0 ? -1.0 : (x > 1.0 ? 1.0 : x);                   // ase.geometry.get_angles clips 1+2e-16 away before arccos
                    const int k = bad_bin(-x, tthr, a.inv_dtheta_f, a.nbins);
                    if (k >= a.nbins) atomicAdd(a.dropped + t, 1ull);
                    else if (srow) atomicAdd(srow + k, 1u);
                    else atomicAdd(grow + k, 1ull);
                }
            }
        }
    }
    __syncthreads();
    for (int s = 0; s < n_slots; ++s) {
        const int key = s_slot_key[s];
        if (key < 0) continue;
        unsigned long long *grow = a.hist + (size_t)key * a.nbins;
        for (int k = threadIdx.x; k < a.nbins; k += blockDim.x) {
            const uint32_t v = s_hist[s * a.nbins + k];
            if (v) atomicAdd(grow + k, (unsigned long long)v);
        }
    }
}
