"""
numpy twin of oracle/amof_oracle.c -- TEST INFRASTRUCTURE ONLY (parity unpinned, see the C header).

An independent second form of the same pinned arithmetic (P1-P8), written as whole-array numpy
expressions over every image of every atom pair.  numpy evaluates each ufunc separately, so there
is no FMA contraction and the results must agree with the C oracle bit for bit on counts.
Only meant for small cases (O(N^2 * images) memory).

Follows the same reference call sites: amof/rdf.py:67-114, amof/atom.py:72-87, amof/cn.py:58-74,
amof/bad.py:70-101, amof/trajectory.py:285-303, amof/msd.py:186-205.
"""
import itertools
import math

import numpy as np


def cell_inverse(c):
    c = np.asarray(c, dtype=np.float64).reshape(3, 3)
    m00 = c[1, 1] * c[2, 2] - c[1, 2] * c[2, 1]
    m01 = c[1, 0] * c[2, 2] - c[1, 2] * c[2, 0]
    m02 = c[1, 0] * c[2, 1] - c[1, 1] * c[2, 0]
    det = (c[0, 0] * m00 - c[0, 1] * m01) + c[0, 2] * m02
    inv = np.empty((3, 3))
    inv[0, 0] = m00 / det
    inv[0, 1] = (c[0, 2] * c[2, 1] - c[0, 1] * c[2, 2]) / det
    inv[0, 2] = (c[0, 1] * c[1, 2] - c[0, 2] * c[1, 1]) / det
    inv[1, 0] = (c[1, 2] * c[2, 0] - c[1, 0] * c[2, 2]) / det
    inv[1, 1] = (c[0, 0] * c[2, 2] - c[0, 2] * c[2, 0]) / det
    inv[1, 2] = (c[0, 2] * c[1, 0] - c[0, 0] * c[1, 2]) / det
    inv[2, 0] = m02 / det
    inv[2, 1] = (c[0, 1] * c[2, 0] - c[0, 0] * c[2, 1]) / det
    inv[2, 2] = (c[0, 0] * c[1, 1] - c[0, 1] * c[1, 0]) / det
    return inv


def fractional(p, inv):
    p = np.asarray(p, dtype=np.float64)
    return np.stack([(p[:, 0] * inv[0, k] + p[:, 1] * inv[1, k]) + p[:, 2] * inv[2, k] for k in range(3)], axis=1)


def wrap(pos, cell):
    """P2"""
    cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    f = fractional(pos, cell_inverse(cell))
    w = np.floor(f)
    t = np.stack([(w[:, 0] * cell[0, c] + w[:, 1] * cell[1, c]) + w[:, 2] * cell[2, c] for c in range(3)], axis=1)
    return np.asarray(pos, dtype=np.float64) - t


def heights(cell):
    inv = cell_inverse(cell)
    return 1.0 / np.sqrt((inv[0] * inv[0] + inv[1] * inv[1]) + inv[2] * inv[2])


def directed_pairs(pos, cell, rcut):
    """All directed (i, j, S) pairs with d2 <= rcut^2 (padded); returns i, j, dv[:,3], d2.  P3."""
    cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    pw = wrap(pos, cell)
    n = len(pw)
    h = heights(cell)
    smax = [int(np.ceil(rcut / h[k])) + 1 for k in range(3)]
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing='ij')
    out_i, out_j, out_dv, out_d2 = [], [], [], []
    r2pad = rcut * rcut * (1.0 + 1e-9) + 1e-300
    for s in itertools.product(*[range(-m, m + 1) for m in smax]):
        T = [(float(s[0]) * cell[0, c] + float(s[1]) * cell[1, c]) + float(s[2]) * cell[2, c] for c in range(3)]
        dv = np.stack([(pw[None, :, c] - pw[:, None, c]) + T[c] for c in range(3)], axis=-1)
        d2 = (dv[..., 0] * dv[..., 0] + dv[..., 1] * dv[..., 1]) + dv[..., 2] * dv[..., 2]
        keep = d2 <= r2pad
        if s == (0, 0, 0):
            keep &= ii != jj
        out_i.append(ii[keep]); out_j.append(jj[keep]); out_dv.append(dv[keep]); out_d2.append(d2[keep])
    return (np.concatenate(out_i), np.concatenate(out_j), np.concatenate(out_dv), np.concatenate(out_d2))


def rdf_hist(pos, cell, spec, nspec, rmax, nbins):
    """P4: uint64 hist[nspec][nspec][nbins] of directed pairs."""
    i, j, _, d2 = directed_pairs(pos, cell, rmax)
    spec = np.asarray(spec)
    dr = rmax / nbins
    q = np.sqrt(d2) / dr
    ok = q < float(nbins)
    b = q[ok].astype(np.int64)
    flat = (spec[i[ok]].astype(np.int64) * nspec + spec[j[ok]]) * nbins + b
    return np.bincount(flat, minlength=nspec * nspec * nbins).astype(np.uint64).reshape(nspec, nspec, nbins)


def neighbour_pairs(pos, cell, spec, cutoff):
    """P5: directed neighbour pairs under a symmetric per-species-pair cutoff matrix."""
    cutoff = np.asarray(cutoff, dtype=np.float64)
    spec = np.asarray(spec)
    rc = float(cutoff.max())
    if not rc > 0:
        e = np.empty(0, dtype=np.int64)
        return e, e, np.empty((0, 3)), np.empty(0)
    i, j, dv, d2 = directed_pairs(pos, cell, rc)
    cut = cutoff[spec[i], spec[j]]
    ok = (cut > 0.0) & (np.sqrt(d2) < cut)
    return i[ok], j[ok], dv[ok], d2[ok]


def cn_counts(pos, cell, spec, nspec, cutoff):
    i, j, _, _ = neighbour_pairs(pos, cell, spec, cutoff)
    spec = np.asarray(spec)
    flat = spec[i].astype(np.int64) * nspec + spec[j]
    return np.bincount(flat, minlength=nspec * nspec).astype(np.uint64).reshape(nspec, nspec)


def bad_angles(pos, cell, spec, cutoff, A, B):
    """P6: dict cn -> list of B-A-B angles (degrees); A/B species index or -1 for any."""
    i, j, dv, d2 = neighbour_pairs(pos, cell, spec, cutoff)
    spec = np.asarray(spec)
    out = {}
    for a in range(len(spec)):
        if A >= 0 and spec[a] != A:
            continue
        sel = (i == a)
        if B >= 0:
            sel &= spec[j] == B
        v, dd = dv[sel], d2[sel]
        cn = len(v)
        for p, q in itertools.combinations(range(cn), 2):
            n0, n1 = np.sqrt(dd[p]), np.sqrt(dd[q])
            u0, u1 = v[p] / n0, v[q] / n1
            x = (u0[0] * u1[0] + u0[1] * u1[1]) + u0[2] * u1[2]
            # libm acos (math.acos), NOT np.arccos: numpy >= 1.22 ships its own SIMD arccos that differs from
            # glibc in the last ulp; the reference pins numpy 1.21.2, whose float64 arccos is libm's.
            x = min(max(float(x), -1.0), 1.0) if x == x else x       # ase clips to [-1, 1] before arccos; NaN stays NaN
            th = math.acos(x) if x == x else float('nan')
            out.setdefault(cn, []).append(th * (180.0 / math.pi))
    return out


def wrap_displacement(d, cell):
    """P8 (ase.geometry.wrap_positions with center=(0,0,0), eps=1e-7)."""
    cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    shift = (0.0 - 0.5) - 1e-7
    g = fractional(d, cell_inverse(cell)) - shift
    g = np.remainder(g, 1.0)
    g = g + shift
    return np.stack([(g[:, 0] * cell[0, c] + g[:, 1] * cell[1, c]) + g[:, 2] * cell[2, c] for c in range(3)], axis=1)


def delta_pos(pos, cell):
    """amof/trajectory.py:285-303"""
    out = [np.array(pos[0], dtype=np.float64)]
    for k in range(len(pos) - 1):
        out.append(wrap_displacement(np.asarray(pos[k + 1]) - np.asarray(pos[k]), cell[k]))
    return out


def msd_of_m(delta, m):
    """amof/msd.py:186-205, without the aliasing drift Q5 (mathematically identical)."""
    T = len(delta)
    part = np.zeros(T - m)
    r_km = np.array(delta[0], dtype=np.float64)
    r_k = r_km * 0
    for k in range(0, m + 1):
        r_k = r_k + delta[k]
    for k in range(m + 1, T):
        r_k = r_k + delta[k]
        r_km = r_km + delta[k - m]
        part[k - m] = np.linalg.norm(r_k - r_km) ** 2 / len(r_km)
    return float(np.mean(part))
