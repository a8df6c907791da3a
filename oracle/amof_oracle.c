/*
 * amof_oracle.c -- CPU restatement of aMOF's frame-parallel structural-analysis hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (amof_b200/) never links or calls anything in oracle/.
 *
 * PARITY STATUS: **parity unpinned**.  The arithmetic of this path lives in third-party
 * packages that are not on disk here (asap3==3.12.8, ase==3.20.1; /root/reference/requirements.txt:1-2)
 * and the reference ships no tests or golden vectors (/root/reference/amof/tests/__init__.py:1-4).
 * The oracle is therefore a restatement of the published algorithms, anchored on the reference's
 * call sites and on the ZIF-4 known answers of BASELINE.md section 4 (derived by brute force
 * from /root/reference/examples/files/ZIF-4.xyz).  Every place where upstream behaviour had to
 * be decided is a numbered pin below (P1..P9).
 *
 * Reference call sites restated here:
 *   RDF   amof/rdf.py:67-114      (asap3 RadialDistributionFunction(atoms, rMax, nBins).update())
 *   CN    amof/cn.py:48-82        + amof/atom.py:72-87 (ase.neighborlist.neighbor_list('ij', atoms, cutoff_dict))
 *   NL    amof/atom.py:72-87      (the same neighbour list, returned as pairs: orc_neighbour_pairs)
 *   BAD   amof/bad.py:70-160,192-300 (Atoms.get_angles(idx, mic=True), np.histogram(edges))
 *   MSD   amof/msd.py:186-268     + amof/trajectory.py:285-303 (ase wrap_positions(center=(0,0,0)))
 *
 * All bin-deciding arithmetic is IEEE-754 binary64 with a fixed operation order and NO fused
 * multiply-add (compile with -ffp-contract=off; never -ffast-math).
 *
 * Pins (each is mirrored, operation for operation, by the CUDA path):
 *  P1  inverse cell by cofactors / determinant, see orc_cell_inverse; fractional coordinate
 *      f_k = (p0*inv[0][k] + p1*inv[1][k]) + p2*inv[2][k].
 *  P2  atoms are wrapped into the cell first: w_k = floor(f_k);
 *      pw_c = p_c - ((w0*C[0][c] + w1*C[1][c]) + w2*C[2][c]).  Atoms already inside keep their bits.
 *  P3  a pair is (i, j, S) with S an integer image vector, excluding only (i == j, S == 0);
 *      dv_c = (pw_j[c] - pw_i[c]) + T_c(S),  T_c(S) = (s0*C[0][c] + s1*C[1][c]) + s2*C[2][c];
 *      d2 = (dvx*dvx + dvy*dvy) + dvz*dvz.   dv(j,i,-S) == -dv(i,j,S) exactly, so directed and
 *      unordered enumerations give identical integers.
 *  P4  RDF bin (U1): dr = rMax/nBins; d = sqrt(d2); q = d/dr; counted iff q < nBins, bin = (int)q.
 *      Directed pairs (U2): hist[Zi][Zj][bin] += 1 for every ordered (i, j, S).
 *  P5  neighbour criterion (U5): kept iff cutoff[Zi][Zj] > 0 and sqrt(d2) < cutoff[Zi][Zj] (strict).
 *  P6  angle (U6): the two neighbour image vectors v0, v1 of P3 are used (not a second find_mic);
 *      n = sqrt(d2); u = v/n componentwise; x = (u0x*u1x + u0y*u1y) + u0z*u1z;
 *      x is clipped to [-1, 1] (ase.geometry.get_angles: "we just normalized the vectors, but in some cases
 *      we can get bad things like 1+2e-16.  These we clip away"; NaN stays NaN, as with np.clip);
 *      theta = acos(x) * (180.0/pi)  [glibc acos of this container].  Precondition: cutoff below
 *      half the smallest perpendicular cell height, else the call is refused.
 *  P7  np.histogram with explicit edges e_k = k*dtheta (k = 0..nbins): bin b iff e_b <= theta < e_{b+1},
 *      last bin closed on the right, NaN or out-of-range dropped (amof/bad.py:142-160).
 *  P8  wrap_positions(center=0): g = f - shift, shift = (0.0 - 0.5) - 1e-7; g = numpy-remainder(g, 1.0);
 *      g += shift; result_c = (g0*C[0][c] + g1*C[1][c]) + g2*C[2][c]; the cell of frame k wraps
 *      the displacement k -> k+1 (amof/trajectory.py:300-302).
 *  P9  centre of mass = (sum_i m_i p_i) / (sum_i m_i), sequential sums (BLAS order upstream is
 *      unknown; MSD is compared at 1e-12 relative, not bit-exactly).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- named conventions: what asap3 / ase really evaluate cannot be read here, so the
 * choices that decide a last-ulp bin are switches with the pin as default.  Process-wide; tests set and restore them.
 *   bin_rule 0 (pin U1): bin = (int)(d / (rMax/nBins))        1: bin = (int)(d * (nBins/rMax))
 *   dv_rule  0 (pin P2/P3): atoms are wrapped into the cell first, dv = (pw_j - pw_i) + T(S)
 *            1: ase.neighborlist's form on the positions as given, dv = (p_j - p_i) + T(S + w_i - w_j), w = floor(f)
 *               (identical bits whenever every atom already lies inside the cell; brute-force enumeration only) */
static int g_bin_rule = 0, g_dv_rule = 0;
void orc_set_conventions(int bin_rule, int dv_rule) { g_bin_rule = bin_rule; g_dv_rule = dv_rule; }

#define ORC_OK 0
#define ORC_ERR_ARG -1
#define ORC_ERR_GEOM -4
#define ORC_ERR_MEM -5

/* ------------------------------------------------------------------ geometry */

/* P1 */
int orc_cell_inverse(const double *c, double *inv) {
    double m00 = c[4] * c[8] - c[5] * c[7];
    double m01 = c[3] * c[8] - c[5] * c[6];
    double m02 = c[3] * c[7] - c[4] * c[6];
    double det = (c[0] * m00 - c[1] * m01) + c[2] * m02;
    if (!(det != 0.0) || !isfinite(det)) return ORC_ERR_GEOM;
    inv[0] = m00 / det;
    inv[1] = (c[2] * c[7] - c[1] * c[8]) / det;
    inv[2] = (c[1] * c[5] - c[2] * c[4]) / det;
    inv[3] = (c[5] * c[6] - c[3] * c[8]) / det;
    inv[4] = (c[0] * c[8] - c[2] * c[6]) / det;
    inv[5] = (c[2] * c[3] - c[0] * c[5]) / det;
    inv[6] = m02 / det;
    inv[7] = (c[1] * c[6] - c[0] * c[7]) / det;
    inv[8] = (c[0] * c[4] - c[1] * c[3]) / det;
    return ORC_OK;
}

double orc_cell_volume(const double *c) {
    double m00 = c[4] * c[8] - c[5] * c[7];
    double m01 = c[3] * c[8] - c[5] * c[6];
    double m02 = c[3] * c[7] - c[4] * c[6];
    return fabs((c[0] * m00 - c[1] * m01) + c[2] * m02);
}

/* perpendicular heights h_k = 1 / |column k of inv| */
static void cell_heights(const double *inv, double *h) {
    for (int k = 0; k < 3; ++k) {
        double s = (inv[0 + k] * inv[0 + k] + inv[3 + k] * inv[3 + k]) + inv[6 + k] * inv[6 + k];
        h[k] = 1.0 / sqrt(s);
    }
}

static inline void frac_of(const double *p, const double *inv, double *f) {
    for (int k = 0; k < 3; ++k) f[k] = (p[0] * inv[0 + k] + p[1] * inv[3 + k]) + p[2] * inv[6 + k];
}

/* P2.  wrapped[n][3]; frac_out (optional) receives f - floor(f) in [0,1). */
int orc_wrap_positions(int n, const double *pos, const double *cell, double *wrapped, double *frac_out) {
    double inv[9];
    int rc = orc_cell_inverse(cell, inv);
    if (rc) return rc;
    for (int i = 0; i < n; ++i) {
        double f[3], w[3];
        frac_of(pos + 3 * i, inv, f);
        for (int k = 0; k < 3; ++k) w[k] = floor(f[k]);
        for (int c = 0; c < 3; ++c)
            wrapped[3 * i + c] = pos[3 * i + c] - ((w[0] * cell[0 + c] + w[1] * cell[3 + c]) + w[2] * cell[6 + c]);
        if (frac_out)
            for (int k = 0; k < 3; ++k) frac_out[3 * i + k] = f[k] - w[k];
    }
    return ORC_OK;
}

/* P3 */
static inline double pair_d2(const double *pi, const double *pj, const double *T, double *dv) {
    dv[0] = (pj[0] - pi[0]) + T[0];
    dv[1] = (pj[1] - pi[1]) + T[1];
    dv[2] = (pj[2] - pi[2]) + T[2];
    return (dv[0] * dv[0] + dv[1] * dv[1]) + dv[2] * dv[2];
}

static inline void image_shift(const double *cell, int s0, int s1, int s2, double *T) {
    for (int c = 0; c < 3; ++c)
        T[c] = ((double)s0 * cell[0 + c] + (double)s1 * cell[3 + c]) + (double)s2 * cell[6 + c];
}

/* ------------------------------------------------------------------ pair visitor
 * Two independent enumerations of the same (i, j, S) set:
 *   method 0: brute force, every j and every image in a box of images that provably covers rcut
 *   method 1: linked cells of width >= rcut, directed 27-neighbourhood (generalised to small boxes)
 * The callback receives directed pairs.
 */
typedef void (*pair_cb)(void *ctx, int i, int j, const double *dv, double d2);

static int visit_pairs_brute(int n, const double *pw, const double *cell, double rcut, pair_cb cb, void *ctx,
                             const double *raw) {
    double inv[9], h[3];
    int rc = orc_cell_inverse(cell, inv);
    if (rc) return rc;
    cell_heights(inv, h);
    int smax[3];
    for (int k = 0; k < 3; ++k) smax[k] = (int)ceil(rcut / h[k]) + 1;
    double r2pad = rcut * rcut * (1.0 + 1e-9) + 1e-300;
    for (int s0 = -smax[0]; s0 <= smax[0]; ++s0)
        for (int s1 = -smax[1]; s1 <= smax[1]; ++s1)
            for (int s2 = -smax[2]; s2 <= smax[2]; ++s2) {
                double T[3];
                image_shift(cell, s0, s1, s2, T);
                int self_image = (s0 == 0 && s1 == 0 && s2 == 0);
                for (int i = 0; i < n; ++i)
                    for (int j = 0; j < n; ++j) {
                        if (self_image && i == j) continue;
                        double dv[3];
                        double d2 = pair_d2(pw + 3 * i, pw + 3 * j, T, dv);
                        if (raw && d2 <= r2pad) {
                            /* dv_rule 1: same (i, j, image), evaluated on the positions as given */
                            double fi[3], fj[3], Tr[3];
                            frac_of(raw + 3 * i, inv, fi);
                            frac_of(raw + 3 * j, inv, fj);
                            image_shift(cell, s0 + (int)floor(fi[0]) - (int)floor(fj[0]), s1 + (int)floor(fi[1]) - (int)floor(fj[1]),
                                        s2 + (int)floor(fi[2]) - (int)floor(fj[2]), Tr);
                            d2 = pair_d2(raw + 3 * i, raw + 3 * j, Tr, dv);
                            cb(ctx, i, j, dv, d2);         /* callbacks re-test d2 against their own thresholds */
                        } else if (d2 <= r2pad) cb(ctx, i, j, dv, d2);
                    }
            }
    return ORC_OK;
}

static int visit_pairs_cells(int n, const double *pw, const double *frac, const double *cell, double rcut,
                             pair_cb cb, void *ctx) {
    double inv[9], h[3];
    int rc = orc_cell_inverse(cell, inv);
    if (rc) return rc;
    cell_heights(inv, h);
    int nc[3], m[3];
    double rpad = rcut * (1.0 + 1e-9) + 1e-300;
    for (int k = 0; k < 3; ++k) {
        nc[k] = (int)floor(h[k] / rpad);
        if (nc[k] < 1) nc[k] = 1;
        if (nc[k] > 64) nc[k] = 64;
        m[k] = (int)ceil(rpad / (h[k] / nc[k]));
    }
    int ncell = nc[0] * nc[1] * nc[2];
    int *head = (int *)malloc(sizeof(int) * (size_t)(ncell + 1));
    int *cid = (int *)malloc(sizeof(int) * (size_t)n);
    int *order = (int *)malloc(sizeof(int) * (size_t)n);
    if (!head || !cid || !order) { free(head); free(cid); free(order); return ORC_ERR_MEM; }
    memset(head, 0, sizeof(int) * (size_t)(ncell + 1));
    for (int i = 0; i < n; ++i) {
        int c[3];
        for (int k = 0; k < 3; ++k) {
            c[k] = (int)(frac[3 * i + k] * nc[k]);
            if (c[k] > nc[k] - 1) c[k] = nc[k] - 1;
            if (c[k] < 0) c[k] = 0;
        }
        cid[i] = (c[0] * nc[1] + c[1]) * nc[2] + c[2];
        head[cid[i] + 1]++;
    }
    for (int c = 0; c < ncell; ++c) head[c + 1] += head[c];
    int *fill = (int *)calloc((size_t)ncell, sizeof(int));
    if (!fill) { free(head); free(cid); free(order); return ORC_ERR_MEM; }
    for (int i = 0; i < n; ++i) order[head[cid[i]] + fill[cid[i]]++] = i;
    free(fill);
    double r2pad = rcut * rcut * (1.0 + 1e-9) + 1e-300;
    for (int c0 = 0; c0 < nc[0]; ++c0)
        for (int c1 = 0; c1 < nc[1]; ++c1)
            for (int c2 = 0; c2 < nc[2]; ++c2) {
                int home = (c0 * nc[1] + c1) * nc[2] + c2;
                if (head[home] == head[home + 1]) continue;
                for (int d0 = -m[0]; d0 <= m[0]; ++d0)
                    for (int d1 = -m[1]; d1 <= m[1]; ++d1)
                        for (int d2i = -m[2]; d2i <= m[2]; ++d2i) {
                            int t[3] = {c0 + d0, c1 + d1, c2 + d2i}, s[3], q[3];
                            for (int k = 0; k < 3; ++k) {
                                s[k] = (int)floor((double)t[k] / nc[k]);
                                q[k] = t[k] - s[k] * nc[k];
                            }
                            int nb = (q[0] * nc[1] + q[1]) * nc[2] + q[2];
                            double T[3];
                            image_shift(cell, s[0], s[1], s[2], T);
                            int self_image = (s[0] == 0 && s[1] == 0 && s[2] == 0);
                            for (int a = head[home]; a < head[home + 1]; ++a) {
                                int i = order[a];
                                for (int b = head[nb]; b < head[nb + 1]; ++b) {
                                    int j = order[b];
                                    if (self_image && i == j) continue;
                                    double dv[3];
                                    double dd = pair_d2(pw + 3 * i, pw + 3 * j, T, dv);
                                    if (dd <= r2pad) cb(ctx, i, j, dv, dd);
                                }
                            }
                        }
            }
    free(head); free(cid); free(order);
    return ORC_OK;
}

static int visit_pairs(int method, int n, const double *pos, const double *cell, double rcut, pair_cb cb, void *ctx) {
    double *pw = (double *)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1));
    double *fr = (double *)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1));
    if (!pw || !fr) { free(pw); free(fr); return ORC_ERR_MEM; }
    int rc = orc_wrap_positions(n, pos, cell, pw, fr);
    if (!rc && g_dv_rule != 0 && method != 0) rc = ORC_ERR_ARG;          /* dv_rule 1 exists for the brute-force form only */
    if (!rc) rc = method == 0 ? visit_pairs_brute(n, pw, cell, rcut, cb, ctx, g_dv_rule ? pos : NULL)
                              : visit_pairs_cells(n, pw, fr, cell, rcut, cb, ctx);
    free(pw); free(fr);
    return rc;
}

/* ------------------------------------------------------------------ RDF (P4) */
typedef struct {
    const uint8_t *spec; int nspec; int nbins; double dr, inv_dr; uint64_t *hist;
} rdf_ctx;

static void rdf_cb(void *vctx, int i, int j, const double *dv, double d2) {
    (void)dv;
    rdf_ctx *c = (rdf_ctx *)vctx;
    double d = sqrt(d2);
    double q = g_bin_rule == 0 ? d / c->dr : d * c->inv_dr;
    if (q < (double)c->nbins) {
        int b = (int)q;
        c->hist[((size_t)c->spec[i] * c->nspec + c->spec[j]) * c->nbins + b] += 1;
    }
}

/* hist[nspec][nspec][nbins] is ACCUMULATED into (caller zeroes it). */
int orc_rdf_frame(int n, const double *pos, const double *cell, const uint8_t *spec, int nspec,
                  double rmax, int nbins, int method, uint64_t *hist) {
    if (n < 0 || nspec < 1 || nbins < 1 || !(rmax > 0.0)) return ORC_ERR_ARG;
    rdf_ctx c = {spec, nspec, nbins, rmax / nbins, (double)nbins / rmax, hist};
    return visit_pairs(method, n, pos, cell, rmax, rdf_cb, &c);
}

/* Whole trajectory, frames split across OpenMP threads (threads <= 1: serial, like amof/rdf.py:88-93). */
int orc_rdf_traj(int nframes, int n, const double *pos, const double *cell, const uint8_t *spec, int nspec,
                 double rmax, int nbins, int method, int threads, uint64_t *hist, double *volume_sum) {
    size_t hs = (size_t)nspec * nspec * nbins;
    int rc_all = ORC_OK;
    double vs = 0.0;
    for (int f = 0; f < nframes; ++f) vs += orc_cell_volume(cell + 9 * (size_t)f);
    if (volume_sum) *volume_sum = vs;
#ifdef _OPENMP
    if (threads > 1) {
#pragma omp parallel num_threads(threads)
        {
            uint64_t *loc = (uint64_t *)calloc(hs, sizeof(uint64_t));
#pragma omp for schedule(dynamic, 1)
            for (int f = 0; f < nframes; ++f) {
                int rc = orc_rdf_frame(n, pos + 3 * (size_t)n * f, cell + 9 * (size_t)f, spec, nspec, rmax, nbins, method, loc);
                if (rc) {
#pragma omp critical
                    rc_all = rc;
                }
            }
#pragma omp critical
            for (size_t k = 0; k < hs; ++k) hist[k] += loc[k];
            free(loc);
        }
        return rc_all;
    }
#endif
    (void)threads;
    for (int f = 0; f < nframes; ++f) {
        int rc = orc_rdf_frame(n, pos + 3 * (size_t)n * f, cell + 9 * (size_t)f, spec, nspec, rmax, nbins, method, hist);
        if (rc) return rc;
    }
    return rc_all;
}

/* ------------------------------------------------------------------ CN (P5) */
typedef struct {
    const uint8_t *spec; int nspec; const double *cutoff; uint64_t *counts;
} cn_ctx;

static void cn_cb(void *vctx, int i, int j, const double *dv, double d2) {
    (void)dv;
    cn_ctx *c = (cn_ctx *)vctx;
    double cut = c->cutoff[c->spec[i] * c->nspec + c->spec[j]];
    if (cut > 0.0 && sqrt(d2) < cut) c->counts[c->spec[i] * c->nspec + c->spec[j]] += 1;
}

static double max_cutoff(const double *cutoff, int nspec) {
    double m = 0.0;
    for (int k = 0; k < nspec * nspec; ++k) if (cutoff[k] > m) m = cutoff[k];
    return m;
}

/* counts[nspec][nspec]: directed neighbour pairs i(species a) -> j(species b) within cutoff[a][b]. */
int orc_cn_frame(int n, const double *pos, const double *cell, const uint8_t *spec, int nspec,
                 const double *cutoff, int method, uint64_t *counts) {
    if (n < 0 || nspec < 1) return ORC_ERR_ARG;
    double rc = max_cutoff(cutoff, nspec);
    if (!(rc > 0.0)) return ORC_OK;
    cn_ctx c = {spec, nspec, cutoff, counts};
    return visit_pairs(method, n, pos, cell, rc, cn_cb, &c);
}

/* counts[nframes][nspec][nspec]; frames split across OpenMP threads like joblib splits them (amof/cn.py:78-80). */
int orc_cn_traj(int nframes, int n, const double *pos, const double *cell, const uint8_t *spec, int nspec,
                const double *cutoff, int method, int threads, uint64_t *counts) {
    int rc_all = ORC_OK;
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 1 ? threads : 1)
#endif
    for (int f = 0; f < nframes; ++f) {
        int rc = orc_cn_frame(n, pos + 3 * (size_t)n * f, cell + 9 * (size_t)f, spec, nspec, cutoff, method,
                              counts + (size_t)f * nspec * nspec);
        if (rc) {
#ifdef _OPENMP
#pragma omp critical
#endif
            rc_all = rc;
        }
    }
    return rc_all;
}

/* ------------------------------------------------------------------ explicit neighbour list (P5) */
/* amof/atom.py:72-87: nl_i, nl_j = ase.neighborlist.neighbor_list('ij', atom, cutoff_dict).  Directed pairs (i, j),
 * one per periodic image with sqrt(d2) < cutoff[Zi][Zj]; the first `capacity` of them are stored, *count gets the total.
 * Optionally also ase's quantities 'd' and 'S' (what pymatgen's get_neighbor_list hands amof.coordination):
 *   dist   = sqrt(d2) of P3
 *   shifts = the integer image S with D = p_j - p_i + S.cell for the ORIGINAL (unwrapped) positions, recovered from the
 *            wrapped-image vector dv of P3 as round(((dv - (p_j - p_i)) . inv)) -- an integer up to rounding noise. */
typedef struct {
    const uint8_t *spec; int nspec; const double *cutoff; int64_t capacity, count; int32_t *pi, *pj;
    const double *pos; double inv[9]; double *dist; int32_t *shifts;
} nl_ctx;

static void nl_cb(void *vctx, int i, int j, const double *dv, double d2) {
    nl_ctx *c = (nl_ctx *)vctx;
    double cut = c->cutoff[c->spec[i] * c->nspec + c->spec[j]];
    if (cut > 0.0 && sqrt(d2) < cut) {
        if (c->count < c->capacity) {
            c->pi[c->count] = i; c->pj[c->count] = j;
            if (c->dist) c->dist[c->count] = sqrt(d2);
            if (c->shifts) {
                double t[3];
                for (int k = 0; k < 3; ++k) t[k] = dv[k] - (c->pos[3 * j + k] - c->pos[3 * i + k]);
                for (int k = 0; k < 3; ++k)
                    c->shifts[3 * c->count + k] = (int32_t)lround((t[0] * c->inv[0 + k] + t[1] * c->inv[3 + k]) + t[2] * c->inv[6 + k]);
            }
        }
        c->count += 1;
    }
}

int orc_neighbour_pairs(int n, const double *pos, const double *cell, const uint8_t *spec, int nspec,
                        const double *cutoff, int method, int64_t capacity, int32_t *pi, int32_t *pj,
                        double *dist, int32_t *shifts, int64_t *count) {
    if (n < 0 || nspec < 1 || !count) return ORC_ERR_ARG;
    *count = 0;
    double rc = max_cutoff(cutoff, nspec);
    if (!(rc > 0.0)) return ORC_OK;
    nl_ctx c;
    memset(&c, 0, sizeof c);
    c.spec = spec; c.nspec = nspec; c.cutoff = cutoff; c.capacity = capacity; c.pi = pi; c.pj = pj;
    c.pos = pos; c.dist = dist; c.shifts = shifts;
    int r0 = orc_cell_inverse(cell, c.inv);
    if (r0) return r0;
    int r = visit_pairs(method, n, pos, cell, rc, nl_cb, &c);
    *count = c.count;
    return r;
}

/* ------------------------------------------------------------------ BAD (P6, P7) */
#define ORC_MAX_NB 64
typedef struct {
    const uint8_t *spec; int nspec; const double *cutoff; int A, B;
    int *nnb;          /* [n] */
    double *vec;       /* [n][ORC_MAX_NB][4] = dv, d2 */
    int overflow;
} bad_ctx;

static void bad_cb(void *vctx, int i, int j, const double *dv, double d2) {
    bad_ctx *c = (bad_ctx *)vctx;
    if (c->A >= 0 && c->spec[i] != c->A) return;
    double cut = c->cutoff[c->spec[i] * c->nspec + c->spec[j]];
    if (!(cut > 0.0 && sqrt(d2) < cut)) return;       /* j is in nl[i]            (amof/atom.py:82)   */
    if (c->B >= 0 && c->spec[j] != c->B) return;      /* ... and is a B neighbour (amof/bad.py:89)    */
    int k = c->nnb[i];
    if (k >= ORC_MAX_NB) { c->overflow = 1; return; }
    double *v = c->vec + ((size_t)i * ORC_MAX_NB + k) * 4;
    v[0] = dv[0]; v[1] = dv[1]; v[2] = dv[2]; v[3] = d2;
    c->nnb[i] = k + 1;
}

/* degrees(arccos(clip(x, -1, 1))): collinear neighbours land on 0 / 180 degrees, NaN stays NaN (np.clip) */
double orc_angle_of_cosine(double x) {
    if (x < -1.0) x = -1.0;
    if (x > 1.0) x = 1.0;
    return acos(x) * (180.0 / 3.14159265358979323846);
}

double orc_angle_deg(const double *v0, double d20, const double *v1, double d21) {
    double n0 = sqrt(d20), n1 = sqrt(d21);
    double u0x = v0[0] / n0, u0y = v0[1] / n0, u0z = v0[2] / n0;
    double u1x = v1[0] / n1, u1y = v1[1] / n1, u1z = v1[2] / n1;
    double x = (u0x * u1x + u0y * u1y) + u0z * u1z;
    return orc_angle_of_cosine(x);
}

/* P7: index of the np.histogram bin for edges e_k = k*dtheta, k = 0..nbins; -1 if dropped. */
int orc_theta_bin(double theta, double dtheta, int nbins) {
    if (!(theta >= 0.0)) return -1;                       /* NaN or negative */
    double last = (double)nbins * dtheta;
    if (theta > last) return -1;
    if (theta == last) return nbins - 1;
    long b = (long)(theta / dtheta);
    if (b > nbins - 1) b = nbins - 1;
    if (b < 0) b = 0;
    while (b > 0 && theta < (double)b * dtheta) --b;
    while (b < nbins - 1 && theta >= (double)(b + 1) * dtheta) ++b;
    return (int)b;
}

/*
 * hist[max_cn+1][nbins] ACCUMULATED: row cn holds the B-A-B angles of centres with exactly cn B-neighbours.
 * A, B: species indices, -1 means "X" (any species).  dropped: NaN / out-of-range angles.
 * Returns ORC_ERR_GEOM when the cutoff precondition of P6 fails or a centre has > max_cn neighbours.
 */
int orc_bad_frame(int n, const double *pos, const double *cell, const uint8_t *spec, int nspec,
                  const double *cutoff, int A, int B, double dtheta, int nbins, int max_cn, int method,
                  uint64_t *hist, uint64_t *dropped) {
    if (n < 0 || nspec < 1 || nbins < 1 || max_cn < 2 || max_cn > ORC_MAX_NB) return ORC_ERR_ARG;
    double rcut = max_cutoff(cutoff, nspec);
    if (!(rcut > 0.0)) return ORC_OK;
    double inv[9], h[3];
    int rc = orc_cell_inverse(cell, inv);
    if (rc) return rc;
    cell_heights(inv, h);
    for (int k = 0; k < 3; ++k) if (!(rcut < 0.5 * h[k])) return ORC_ERR_GEOM;
    bad_ctx c = {spec, nspec, cutoff, A, B, NULL, NULL, 0};
    c.nnb = (int *)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
    c.vec = (double *)malloc(sizeof(double) * 4 * ORC_MAX_NB * (size_t)(n > 0 ? n : 1));
    if (!c.nnb || !c.vec) { free(c.nnb); free(c.vec); return ORC_ERR_MEM; }
    rc = visit_pairs(method, n, pos, cell, rcut, bad_cb, &c);
    if (!rc && c.overflow) rc = ORC_ERR_GEOM;
    for (int a = 0; a < n && !rc; ++a) {
        int cn = c.nnb[a];
        if (cn < 2) continue;
        if (cn > max_cn) { rc = ORC_ERR_GEOM; break; }
        const double *v = c.vec + (size_t)a * ORC_MAX_NB * 4;
        for (int p = 0; p < cn; ++p)
            for (int q = p + 1; q < cn; ++q) {
                double th = orc_angle_deg(v + 4 * p, v[4 * p + 3], v + 4 * q, v[4 * q + 3]);
                int b = orc_theta_bin(th, dtheta, nbins);
                if (b < 0) { if (dropped) *dropped += 1; }
                else hist[(size_t)cn * nbins + b] += 1;
            }
    }
    free(c.nnb); free(c.vec);
    return rc;
}

/* Raw angle list for one frame (small-case tests): writes up to cap angles, returns count or <0. */
long orc_bad_angles(int n, const double *pos, const double *cell, const uint8_t *spec, int nspec,
                    const double *cutoff, int A, int B, int method, double *angles, long cap) {
    double rcut = max_cutoff(cutoff, nspec);
    if (!(rcut > 0.0)) return 0;
    bad_ctx c = {spec, nspec, cutoff, A, B, NULL, NULL, 0};
    c.nnb = (int *)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
    c.vec = (double *)malloc(sizeof(double) * 4 * ORC_MAX_NB * (size_t)(n > 0 ? n : 1));
    if (!c.nnb || !c.vec) { free(c.nnb); free(c.vec); return ORC_ERR_MEM; }
    long cnt = visit_pairs(method, n, pos, cell, rcut, bad_cb, &c);
    if (!cnt && c.overflow) cnt = ORC_ERR_GEOM;
    for (int a = 0; a < n && cnt >= 0; ++a) {
        int cn = c.nnb[a];
        const double *v = c.vec + (size_t)a * ORC_MAX_NB * 4;
        for (int p = 0; p < cn; ++p)
            for (int q = p + 1; q < cn; ++q) {
                if (cnt < cap) angles[cnt] = orc_angle_deg(v + 4 * p, v[4 * p + 3], v + 4 * q, v[4 * q + 3]);
                ++cnt;
            }
    }
    free(c.nnb); free(c.vec);
    return cnt;
}

/* ------------------------------------------------------------------ MSD (P8, P9) */

/* numpy remainder(g, 1.0) for float64 */
static inline double np_mod1(double g) {
    double r = fmod(g, 1.0);
    if (r != 0.0) { if (r < 0.0) r += 1.0; }
    else r = copysign(0.0, 1.0);
    return r;
}

/* P8: wrap one displacement with the given cell/inverse. */
static inline void wrap_disp(const double *d, const double *cell, const double *inv, double *out) {
    const double shift = (0.0 - 0.5) - 1e-7;
    double g[3];
    frac_of(d, inv, g);
    for (int k = 0; k < 3; ++k) {
        double t = g[k] - shift;
        t = np_mod1(t);
        g[k] = t + shift;
    }
    for (int c = 0; c < 3; ++c) out[c] = (g[0] * cell[0 + c] + g[1] * cell[3 + c]) + g[2] * cell[6 + c];
}

/*
 * delta[T][n][3] from pos[T][n][3] and cell[T][9]  (amof/trajectory.py:285-303)
 */
int orc_delta_pos(int T, int n, const double *pos, const double *cell, double *delta) {
    if (T < 1) return ORC_ERR_ARG;
    memcpy(delta, pos, sizeof(double) * 3 * (size_t)n);
    for (int k = 0; k + 1 < T; ++k) {
        double inv[9];
        int rc = orc_cell_inverse(cell + 9 * (size_t)k, inv);
        if (rc) return rc;
        const double *p0 = pos + 3 * (size_t)n * k, *p1 = pos + 3 * (size_t)n * (k + 1);
        double *o = delta + 3 * (size_t)n * (k + 1);
        for (int i = 0; i < n; ++i) {
            double d[3] = {p1[3 * i] - p0[3 * i], p1[3 * i + 1] - p0[3 * i + 1], p1[3 * i + 2] - p0[3 * i + 2]};
            wrap_disp(d, cell + 9 * (size_t)k, inv, o + 3 * i);
        }
    }
    return ORC_OK;
}

/*
 * WindowMsd.compute_msd (amof/msd.py:207-268) without the pandas assembly.
 *   pos[T][n][3] is MODIFIED in place exactly like the reference mutates the caller's frames (Q7):
 *   optional unwrap (msd.py:222-230), then translate(-COM) per frame (msd.py:235-237).
 *   msd[nspec][nwin] = per-element window MSD with the (T-m-1)/(T-m) quirk Q4.
 */
int orc_msd_window(int T, int n, double *pos, const double *cell, const double *masses,
                   const uint8_t *spec, int nspec, const int *window, int nwin, int unwrap, double *msd) {
    if (T < 1 || n < 1 || nspec < 1 || nwin < 0) return ORC_ERR_ARG;
    size_t fr = 3 * (size_t)n;
    double *delta = (double *)malloc(sizeof(double) * fr * (size_t)T);
    if (!delta) return ORC_ERR_MEM;
    int rc;
    if (unwrap) {
        rc = orc_delta_pos(T, n, pos, cell, delta);
        if (rc) { free(delta); return rc; }
        /* new_pos = positions[0]; new_pos += delta_pos[i]; frame i gets a copy */
        for (int k = 1; k < T; ++k)
            for (size_t a = 0; a < fr; ++a) pos[fr * k + a] = pos[fr * (k - 1) + a] + delta[fr * k + a];
    }
    for (int k = 0; k < T; ++k) {
        double sx = 0, sy = 0, sz = 0, sm = 0;
        double *p = pos + fr * k;
        for (int i = 0; i < n; ++i) {
            sx += masses[i] * p[3 * i]; sy += masses[i] * p[3 * i + 1]; sz += masses[i] * p[3 * i + 2];
            sm += masses[i];
        }
        double cx = sx / sm, cy = sy / sm, cz = sz / sm;
        for (int i = 0; i < n; ++i) { p[3 * i] -= cx; p[3 * i + 1] -= cy; p[3 * i + 2] -= cz; }
    }
    rc = orc_delta_pos(T, n, pos, cell, delta);
    if (rc) { free(delta); return rc; }
    /* unwrapped running positions R_k, in place of delta */
    for (int k = 1; k < T; ++k)
        for (size_t a = 0; a < fr; ++a) delta[fr * k + a] += delta[fr * (k - 1) + a];
    int *cnt = (int *)calloc((size_t)nspec, sizeof(int));
    for (int i = 0; i < n; ++i) cnt[spec[i]]++;
    for (int w = 0; w < nwin; ++w) {
        int m = window[w];
        double *acc = (double *)calloc((size_t)nspec, sizeof(double));
        double *fsum = (double *)malloc(sizeof(double) * (size_t)nspec);
        if (m >= 0 && m < T) {
            for (int k = m + 1; k < T; ++k) {
                const double *rk = delta + fr * k, *rm = delta + fr * (k - m);
                for (int s = 0; s < nspec; ++s) fsum[s] = 0.0;
                for (int i = 0; i < n; ++i) {
                    double dx = rk[3 * i] - rm[3 * i], dy = rk[3 * i + 1] - rm[3 * i + 1], dz = rk[3 * i + 2] - rm[3 * i + 2];
                    fsum[spec[i]] += (dx * dx + dy * dy) + dz * dz;
                }
                for (int s = 0; s < nspec; ++s) if (cnt[s]) acc[s] += fsum[s] / cnt[s];
            }
        }
        for (int s = 0; s < nspec; ++s)
            msd[(size_t)s * nwin + w] = (m >= 0 && m < T && cnt[s]) ? acc[s] / (double)(T - m) : NAN;
        free(acc); free(fsum);
    }
    free(cnt); free(delta);
    return ORC_OK;
}

/*
 * DirectMsd.compute_species_msd (amof/msd.py:81-105), orthogonal cells only:
 *   r_t = r_{t-1} + wrap_box(p_t - (r_{t-1} mod a)), MSD[t] = |r_t - r_0|^2 / N.
 * sel = species index or -1 for all atoms.  msd[T].
 */
int orc_msd_direct(int T, int n, const double *pos, const double *cell, const uint8_t *spec, int sel, double *msd) {
    if (T < 1 || n < 1) return ORC_ERR_ARG;
    size_t fr = 3 * (size_t)n;
    double *r = (double *)malloc(sizeof(double) * fr);
    if (!r) return ORC_ERR_MEM;
    memcpy(r, pos, sizeof(double) * fr);
    int cnt = 0;
    for (int i = 0; i < n; ++i) if (sel < 0 || spec[i] == sel) ++cnt;
    msd[0] = 0.0;
    for (int t = 1; t < T; ++t) {
        const double *p = pos + fr * t;
        const double *c = cell + 9 * (size_t)t;
        double acc = 0.0;
        for (int i = 0; i < n; ++i) {
            if (!(sel < 0 || spec[i] == sel)) continue;
            double s2 = 0.0;
            for (int j = 0; j < 3; ++j) {
                double a = c[4 * j];
                double prev = r[3 * i + j];
                double pm = fmod(prev, a);                    /* python/numpy % : sign of divisor */
                if (pm != 0.0 && ((a < 0.0) != (pm < 0.0))) pm += a;
                double d = p[3 * i + j] - pm;
                if (d > a / 2) d -= a; else if (d < -a / 2) d += a;
                double rt = d + prev;
                r[3 * i + j] = rt;
                double dd = rt - pos[3 * i + j];
                s2 += dd * dd;
            }
            acc += s2;
        }
        msd[t] = cnt ? acc / cnt : NAN;
    }
    free(r);
    return ORC_OK;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
