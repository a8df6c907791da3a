// neigh.cuh -- explicit neighbour list of one frame (amof.atom.get_neighborlist, /root/reference/amof/atom.py:72-87:
// ase.neighborlist.neighbor_list('ij', atoms, cutoff_dict) regrouped per atom).
//
// The analyses of this library never materialise the list (counting and angle enumeration are fused with the search);
// this kernel exists for the callers that need the list itself (amof.ring and amof.coordination hand it to graph
// code).  Same search as k_bad: the cell list holds only the species that appear in the cutoff
// matrix, one thread owns one atom of the sorted frame and walks the FULL stencil, a pair (i, j, image) is kept iff
// d2 < cn_thr2[key] (P5: the bisected d2 threshold of `sqrt(d2) < cutoff`), the zero-shift self pair is skipped and the
// same j under several periodic images appears once per image, as ase lists it.
//   FILL = false: count[i] = number of neighbours of atom i (original index)
//   FILL = true : the thread writes its row nbr[offset[i] ...] with ORIGINAL indices (optionally the distance and the
//                 image shift of every pair) and sorts it in place by (j, S)
#pragma once
#include "prep.cuh"

struct NeighArgs {
    const SAtom *sorted;
    const FrameGeom *geom;
    const uint32_t *cell_start;
    const uint32_t *orig;         // [n_keep] sorted position -> original atom index
    const double *cn_thr2;        // [nkeys]
    const uint16_t *keyidx;       // [S*S]
    int *count;                   // [n_atoms]
    const long long *offset;      // [n_atoms + 1]
    int *nbr;                     // [offset[n_atoms]]
    const int *wraps;             // [n_keep][3] cell translations P2 removed from every sorted atom
    double *dist;                 // optional [offset[n_atoms]]: sqrt(d2) of every listed pair (ase's 'd')
    int *shifts;                  // optional [offset[n_atoms]][3]: S with D = p_j - p_i + S.cell for the ORIGINAL positions ('S')
    double r2search;
    int n_atoms, n_keep, n_species;
};

template <bool FILL>
__global__ void __launch_bounds__(128) k_neigh(NeighArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_keep) return;
    const SAtom *fr = a.sorted;
    const SAtom me = load_satom(fr + i);
    const int si = (int)(me.s & 0xff);
    const FrameGeom &G = a.geom[0];
    const uint32_t *cs = a.cell_start + G.cs_off;
    const int c0 = (int)((me.s >> 8) & 0xfff), c1 = (int)((me.s >> 20) & 0xfff), c2 = (int)((me.s >> 32) & 0xfff);
    const int nc0 = G.nc[0], nc1 = G.nc[1], nc2 = G.nc[2];
    const int m0 = G.m[0], m1 = G.m[1], m2 = G.m[2];
    const uint16_t *krow = a.keyidx + si * a.n_species;
    const int io = (int)a.orig[i];
    int *row = nullptr;
    long long base = 0;
    if (FILL) { base = a.offset[io]; row = a.nbr + base; }
    int nn = 0;
    for (int d0 = -m0; d0 <= m0; ++d0) {
        int s0, q0;
        wrap_cell(c0 + d0, nc0, s0, q0);
        for (int d1 = -m1; d1 <= m1; ++d1) {
            int s1, q1;
            wrap_cell(c1 + d1, nc1, s1, q1);
            const int rowbase = (q0 * nc1 + q1) * nc2;
            int d2 = -m2;
            while (d2 <= m2) {
                int s2, q2;
                wrap_cell(c2 + d2, nc2, s2, q2);
                const int len = min(m2 - d2, nc2 - 1 - q2) + 1;
                const int jb = (int)cs[rowbase + q2], je = (int)cs[rowbase + q2 + len];
                d2 += len;
                if (je <= jb) continue;
                const bool self_image = ((s0 | s1 | s2) == 0);
                const double fs0 = (double)s0, fs1 = (double)s1, fs2 = (double)s2;       // P3 image shift
                const double Tx = (fs0 * G.cell[0] + fs1 * G.cell[3]) + fs2 * G.cell[6];
                const double Ty = (fs0 * G.cell[1] + fs1 * G.cell[4]) + fs2 * G.cell[7];
                const double Tz = (fs0 * G.cell[2] + fs1 * G.cell[5]) + fs2 * G.cell[8];
                for (int j = jb; j < je; ++j) {
                    if (self_image && j == i) continue;
                    const SAtom o = load_satom(fr + j);
                    const double dx = (o.x - me.x) + Tx;
                    const double dy = (o.y - me.y) + Ty;
                    const double dz = (o.z - me.z) + Tz;
                    const double dd = (dx * dx + dy * dy) + dz * dz;
                    if (dd < a.r2search && dd < __ldg(a.cn_thr2 + krow[(int)(o.s & 0xff)])) {
                        if (FILL) {
                            row[nn] = (int)a.orig[j];
                            if (a.dist) a.dist[base + nn] = sqrt(dd);
                            if (a.shifts) {
                                // the search shifts the WRAPPED positions pw = p - w.cell:  dv = (p_j - p_i) + (S + w_i - w_j).cell
                                int *sp = a.shifts + 3 * (base + nn);
                                sp[0] = s0 + a.wraps[3 * i] - a.wraps[3 * j];
                                sp[1] = s1 + a.wraps[3 * i + 1] - a.wraps[3 * j + 1];
                                sp[2] = s2 + a.wraps[3 * i + 2] - a.wraps[3 * j + 2];
                            }
                        }
                        ++nn;
                    }
                }
            }
        }
    }
    if (!FILL) { a.count[io] = nn; return; }
    // rows are a handful of entries: insertion sort by the owning thread, keyed by (j, S) so that the several images of
    // one partner come out in a defined order; distances and shifts move with their pair
    auto before = [&](int vj, int v0, int v1, int v2, long long q) {      // is (vj, v) < entry q ?
        if (vj != row[q]) return vj < row[q];
        if (!a.shifts) return false;
        const int *sq = a.shifts + 3 * (base + q);
        if (v0 != sq[0]) return v0 < sq[0];
        if (v1 != sq[1]) return v1 < sq[1];
        return v2 < sq[2];
    };
    for (int p = 1; p < nn; ++p) {
        const int vj = row[p];
        const double vd = a.dist ? a.dist[base + p] : 0.0;
        int v0 = 0, v1 = 0, v2 = 0;
        if (a.shifts) { const int *sp = a.shifts + 3 * (base + p); v0 = sp[0]; v1 = sp[1]; v2 = sp[2]; }
        int q = p - 1;
        while (q >= 0 && before(vj, v0, v1, v2, q)) {
            row[q + 1] = row[q];
            if (a.dist) a.dist[base + q + 1] = a.dist[base + q];
            if (a.shifts) { int *d = a.shifts + 3 * (base + q + 1); const int *s = d - 3; d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; }
            --q;
        }
        row[q + 1] = vj;
        if (a.dist) a.dist[base + q + 1] = vd;
        if (a.shifts) { int *d = a.shifts + 3 * (base + q + 1); d[0] = v0; d[1] = v1; d[2] = v2; }
    }
}
