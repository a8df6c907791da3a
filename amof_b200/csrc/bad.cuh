// bad.cuh -- bond-angle triplet kernel (K4 of SURVEY.md 2.1).
//
// One thread owns one centre atom of the cell-sorted frame: it walks the FULL stencil, keeps the unit vectors of every
// neighbour under the pair cutoffs (P5), then for each requested (A, B) triple enumerates the unordered
// pairs of its B-neighbours.  The angle itself is never formed on the device: x = u_p . u_q is computed in fp64
// in the oracle's operation order (P6) and located in a table of thresholds on -x that the host bisected with its
// own libm acos and the np.histogram edge rule (P7), so the bin is the one the CPU path takes, bit for bit,
// without depending on CUDA's acos.
#pragma once
#include "prep.cuh"

#define BAD_NB_MAX 64        // neighbours of any species kept per centre
#define BAD_MAX_TRIPLES 64

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
    int *flags;                   // bit 0: neighbour overflow, bit 1: cn > AMOFB_BAD_MAX_CN
    int n_keep;                   // atoms per frame in the (species-filtered) cell list
    unsigned long long centre_mask[AMOFB_MAX_SPECIES];   // triples whose A matches this species
    double r2search;
    float inv_dtheta_f;
    int n_atoms, n_frames, n_species, nkeys, n_triples, nbins;
};

// index K in [0, nbins] of t = -x: K < nbins is the histogram bin, K == nbins means "beyond the last edge"
__device__ __forceinline__ int bad_bin(double t, const double *__restrict__ tthr, float inv_dtheta_f, int nbins) {
    float xf = fminf(fmaxf((float)(-t), -1.0f), 1.0f);
    int k = (int)(acosf(xf) * 57.29577951308232f * inv_dtheta_f);
    k = k < 0 ? 0 : (k > nbins ? nbins : k);
    while (t < __ldg(tthr + k)) --k;          // tthr[0] = -2 stops it
    while (t >= __ldg(tthr + k + 1)) ++k;     // tthr[nbins+1] = +2 stops it
    return k;
}

// One THREAD per atom of the cell-sorted frame, in sorted order: the cell list only holds the species that appear in
// the cutoff matrix (PrepArgs::species_keep), so nearly every thread is a centre, and neighbouring threads sit in the
// same or adjacent cells -- their cell_start[] and candidate reads hit the same lines.
#ifndef BAD_MIN_BLOCKS
#define BAD_MIN_BLOCKS 8      // 64 registers: measured best on C4 (tools/sweep_bad.sh)
#endif
__global__ void __launch_bounds__(128, BAD_MIN_BLOCKS) k_bad(BadArgs a) {
    const long long t_id = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t_id >= (long long)a.n_frames * a.n_keep) return;
    const int f = (int)(t_id / a.n_keep);
    const int i = (int)(t_id - (long long)f * a.n_keep);
    const SAtom *fr = a.sorted + (long long)f * a.n_atoms;
    const SAtom me = load_satom(fr + i);
    const int si = (int)(me.s & 0xff);
    const unsigned long long mine = a.centre_mask[si];
    if (!mine) return;
    const FrameGeom &G = a.geom[f];
    const uint32_t *cs = a.cell_start + G.cs_off;
    const int S = a.n_species;
    const int c0 = (int)((me.s >> 8) & 0xfff), c1 = (int)((me.s >> 20) & 0xfff), c2 = (int)((me.s >> 32) & 0xfff);
    const int nc0 = G.nc[0], nc1 = G.nc[1], nc2 = G.nc[2];
    const int m0 = G.m[0], m1 = G.m[1], m2 = G.m[2];

    double ux[BAD_NB_MAX], uy[BAD_NB_MAX], uz[BAD_NB_MAX];
    unsigned char sp[BAD_NB_MAX];
    int nn = 0;
    bool overflow = false;
    const double mex = me.x, mey = me.y, mez = me.z;
    const uint16_t *krow = a.keyidx + si * S;
    // the z window [c2 - m2, c2 + m2] is the same for every row of the stencil: split it once into its contiguous
    // runs (one per wrap of the column; a third wrap only happens in boxes narrower than the stencil)
    int zq[3], zl[3], zs[3], nz = 0;
    bool many_wraps = false;
    for (int d2 = -m2; d2 <= m2;) {
        int s2, q2;
        wrap_cell(c2 + d2, nc2, s2, q2);
        const int len = min(m2 - d2, nc2 - 1 - q2) + 1;
        if (nz < 3) { zq[nz] = q2; zl[nz] = len; zs[nz] = s2; ++nz; } else many_wraps = true;
        d2 += len;
    }
    auto scan_run = [&](int rowbase, int s01, double Rx, double Ry, double Rz, int q2, int len, int s2) {
        const int jb = (int)cs[rowbase + q2], je = (int)cs[rowbase + q2 + len];
        if (je <= jb) return;
        const bool self_image = ((s01 | s2) == 0);
        double Tx = Rx, Ty = Ry, Tz = Rz;          // (0 + 0) + s2*c == s2*c and x + 0.0 == x: same bits as the full P3 sum
        if (s2 != 0) { const double fs2 = (double)s2; Tx = Rx + fs2 * G.cell[6]; Ty = Ry + fs2 * G.cell[7]; Tz = Rz + fs2 * G.cell[8]; }
        for (int j = jb; j < je; ++j) {
            if (self_image && j == i) continue;
            const SAtom o = load_satom(fr + j);
            const double dx = (o.x - mex) + Tx;
            const double dy = (o.y - mey) + Ty;
            const double dz = (o.z - mez) + Tz;
            const double dd = (dx * dx + dy * dy) + dz * dz;
            if (dd < a.r2search) {
                const int sj = (int)(o.s & 0xff);
                if (dd < __ldg(a.cn_thr2 + krow[sj])) {
                    if (nn < BAD_NB_MAX) {                  // keep the raw image vector; it is normalised after the walk
                        ux[nn] = dx; uy[nn] = dy; uz[nn] = dz;
                        sp[nn] = (unsigned char)sj;
                        ++nn;
                    } else overflow = true;
                }
            }
        }
    };
    for (int d0 = -m0; d0 <= m0; ++d0) {
        int s0, q0;
        wrap_cell(c0 + d0, nc0, s0, q0);
        for (int d1 = -m1; d1 <= m1; ++d1) {
            int s1, q1;
            wrap_cell(c1 + d1, nc1, s1, q1);
            const int rowbase = (q0 * nc1 + q1) * nc2;
            // P3 image shift, (s0*a + s1*b) part: zero for most rows
            double Rx = 0.0, Ry = 0.0, Rz = 0.0;
            if ((s0 | s1) != 0) {
                const double fs0 = (double)s0, fs1 = (double)s1;
                Rx = fs0 * G.cell[0] + fs1 * G.cell[3]; Ry = fs0 * G.cell[1] + fs1 * G.cell[4]; Rz = fs0 * G.cell[2] + fs1 * G.cell[5];
            }
            if (!many_wraps) {
                for (int k = 0; k < nz; ++k) scan_run(rowbase, s0 | s1, Rx, Ry, Rz, zq[k], zl[k], zs[k]);
            } else {
                for (int d2 = -m2; d2 <= m2;) {
                    int s2, q2;
                    wrap_cell(c2 + d2, nc2, s2, q2);
                    const int len = min(m2 - d2, nc2 - 1 - q2) + 1;
                    scan_run(rowbase, s0 | s1, Rx, Ry, Rz, q2, len, s2);
                    d2 += len;
                }
            }
        }
    }
    if (overflow) { atomicOr(a.flags, 1); return; }
    if (nn < 2) return;
    // P6: u = v / |v|, componentwise.  Done here, once per kept neighbour and only for centres that can form an angle:
    // inside the divergent candidate loop the sqrt and the three divisions (~160 instructions) ran whenever ANY lane
    // of the warp had a hit.  |v|^2 is re-formed from the stored components in the same order, so n is the same double.
    for (int p = 0; p < nn; ++p) {
        const double dx = ux[p], dy = uy[p], dz = uz[p];
        const double n = sqrt((dx * dx + dy * dy) + dz * dz);
        ux[p] = dx / n; uy[p] = dy / n; uz[p] = dz / n;
    }
    for (int t = 0; t < a.n_triples; ++t) {
        if (!((mine >> t) & 1ull)) continue;
        const int B = a.triples[t].y;
        int cn = 0;
        for (int p = 0; p < nn; ++p) cn += (B < 0 || sp[p] == B);
        if (cn < 2) continue;
        if (cn > AMOFB_BAD_MAX_CN) { atomicOr(a.flags, 2); continue; }
        unsigned long long *row = a.hist + ((size_t)t * (AMOFB_BAD_MAX_CN + 1) + cn) * a.nbins;
        for (int p = 0; p < nn; ++p) {
            if (!(B < 0 || sp[p] == B)) continue;
            for (int q = p + 1; q < nn; ++q) {
                if (!(B < 0 || sp[q] == B)) continue;
                double x = (ux[p] * ux[q] + uy[p] * uy[q]) + uz[p] * uz[q];
                if (x != x) { atomicAdd(a.dropped + t, 1ull); continue; }   // NaN (coincident atoms): np.histogram drops it
                x = x < -1.0 ? -1.0 : (x > 1.0 ? 1.0 : x);                   // ase.geometry.get_angles clips 1+2e-16 away before arccos
                const int k = bad_bin(-x, a.tthr, a.inv_dtheta_f, a.nbins);
                if (k >= a.nbins) atomicAdd(a.dropped + t, 1ull);
                else atomicAdd(row + k, 1ull);
            }
        }
    }
}
