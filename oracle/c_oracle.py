"""ctypes binding of liboracle.so.  TEST INFRASTRUCTURE ONLY (see oracle/amof_oracle.c)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "amof_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None
_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)
_ip = C.POINTER(C.c_int)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_rdf_frame.argtypes = [C.c_int, _dp, _dp, _u8p, C.c_int, C.c_double, C.c_int, C.c_int, _u64p]
        L.orc_rdf_traj.argtypes = [C.c_int, C.c_int, _dp, _dp, _u8p, C.c_int, C.c_double, C.c_int, C.c_int,
                                   C.c_int, _u64p, _dp]
        L.orc_cn_frame.argtypes = [C.c_int, _dp, _dp, _u8p, C.c_int, _dp, C.c_int, _u64p]
        L.orc_cn_traj.argtypes = [C.c_int, C.c_int, _dp, _dp, _u8p, C.c_int, _dp, C.c_int, C.c_int, _u64p]
        L.orc_neighbour_pairs.argtypes = [C.c_int, _dp, _dp, _u8p, C.c_int, _dp, C.c_int, C.c_int64, C.POINTER(C.c_int32),
                                          C.POINTER(C.c_int32), _dp, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
        L.orc_bad_frame.argtypes = [C.c_int, _dp, _dp, _u8p, C.c_int, _dp, C.c_int, C.c_int, C.c_double, C.c_int,
                                    C.c_int, C.c_int, _u64p, _u64p]
        L.orc_bad_angles.argtypes = [C.c_int, _dp, _dp, _u8p, C.c_int, _dp, C.c_int, C.c_int, C.c_int, _dp, C.c_long]
        L.orc_bad_angles.restype = C.c_long
        L.orc_delta_pos.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp]
        L.orc_msd_window.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _u8p, C.c_int, _ip, C.c_int, C.c_int, _dp]
        L.orc_msd_direct.argtypes = [C.c_int, C.c_int, _dp, _dp, _u8p, C.c_int, _dp]
        L.orc_wrap_positions.argtypes = [C.c_int, _dp, _dp, _dp, _dp]
        L.orc_cell_inverse.argtypes = [_dp, _dp]
        L.orc_cell_volume.argtypes = [_dp]
        L.orc_cell_volume.restype = C.c_double
        L.orc_theta_bin.argtypes = [C.c_double, C.c_double, C.c_int]
        L.orc_max_threads.restype = C.c_int
        L.orc_angle_of_cosine.argtypes = [C.c_double]
        L.orc_set_conventions.argtypes = [C.c_int, C.c_int]
        L.orc_set_conventions.restype = None
        L.orc_angle_of_cosine.restype = C.c_double
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _check(rc, what):
    if rc < 0:
        raise RuntimeError("oracle %s failed with code %d" % (what, rc))


def rdf_hist(pos, cell, spec, nspec, rmax, nbins, method=1):
    """One frame: uint64 hist[nspec][nspec][nbins] of directed pairs."""
    pos, pp = _d(pos); cell, cp = _d(cell)
    spec = np.ascontiguousarray(spec, dtype=np.uint8)
    hist = np.zeros((nspec, nspec, nbins), dtype=np.uint64)
    _check(lib().orc_rdf_frame(len(spec), pp, cp, spec.ctypes.data_as(_u8p), nspec, float(rmax), int(nbins),
                               method, hist.ctypes.data_as(_u64p)), "rdf_frame")
    return hist


def rdf_traj(pos, cell, spec, nspec, rmax, nbins, method=1, threads=1):
    """pos[T][n][3], cell[T][3][3] -> (hist, volume_sum)."""
    pos, pp = _d(pos); cell, cp = _d(cell)
    spec = np.ascontiguousarray(spec, dtype=np.uint8)
    T, n = pos.shape[0], pos.shape[1]
    hist = np.zeros((nspec, nspec, nbins), dtype=np.uint64)
    vs = C.c_double(0.0)
    _check(lib().orc_rdf_traj(T, n, pp, cp, spec.ctypes.data_as(_u8p), nspec, float(rmax), int(nbins), method,
                              int(threads), hist.ctypes.data_as(_u64p), C.byref(vs)), "rdf_traj")
    return hist, vs.value


def cn_counts(pos, cell, spec, nspec, cutoff, method=1):
    pos, pp = _d(pos); cell, cp = _d(cell); cutoff, kp = _d(cutoff)
    spec = np.ascontiguousarray(spec, dtype=np.uint8)
    counts = np.zeros((nspec, nspec), dtype=np.uint64)
    _check(lib().orc_cn_frame(len(spec), pp, cp, spec.ctypes.data_as(_u8p), nspec, kp, method,
                              counts.ctypes.data_as(_u64p)), "cn_frame")
    return counts


def cn_traj(pos, cell, spec, nspec, cutoff, method=1, threads=1):
    """pos[T][n][3] -> uint64 counts[T][nspec][nspec]"""
    pos, pp = _d(pos); cell, cp = _d(cell); cutoff, kp = _d(cutoff)
    spec = np.ascontiguousarray(spec, dtype=np.uint8)
    T, n = pos.shape[0], pos.shape[1]
    counts = np.zeros((T, nspec, nspec), dtype=np.uint64)
    _check(lib().orc_cn_traj(T, n, pp, cp, spec.ctypes.data_as(_u8p), nspec, kp, method, int(threads),
                             counts.ctypes.data_as(_u64p)), "cn_traj")
    return counts


def neighbour_pairs(pos, cell, spec, nspec, cutoff, method=1, quantities=False):
    """One frame -> directed neighbour pairs (i, j) as ase.neighbor_list('ij', ...) would list them (amof/atom.py:82),
    sorted by (i, j, S); a pair appears once per periodic image under the cutoff.  With ``quantities`` also the
    distances and the integer image shifts S (D = p_j - p_i + S.cell) of ase's 'd' and 'S'."""
    pos, pp = _d(pos); cell, cp = _d(cell); cutoff, kp = _d(cutoff)
    spec = np.ascontiguousarray(spec, dtype=np.uint8)
    n = pos.shape[0]
    count = C.c_int64(0)
    i32p = C.POINTER(C.c_int32)
    _check(lib().orc_neighbour_pairs(n, pp, cp, spec.ctypes.data_as(_u8p), nspec, kp, method, 0, None, None, None, None,
                                     C.byref(count)), "neighbour_pairs")
    pi = np.zeros(count.value, dtype=np.int32)
    pj = np.zeros(count.value, dtype=np.int32)
    dist = np.zeros(count.value, dtype=np.float64)
    shifts = np.zeros((count.value, 3), dtype=np.int32)
    if count.value:
        _check(lib().orc_neighbour_pairs(n, pp, cp, spec.ctypes.data_as(_u8p), nspec, kp, method, count.value,
                                         pi.ctypes.data_as(i32p), pj.ctypes.data_as(i32p), dist.ctypes.data_as(_dp),
                                         shifts.ctypes.data_as(i32p), C.byref(count)), "neighbour_pairs")
    order = np.lexsort((shifts[:, 2], shifts[:, 1], shifts[:, 0], pj, pi))
    if quantities:
        return pi[order], pj[order], dist[order], shifts[order]
    return pi[order], pj[order]


def bad_hist(pos, cell, spec, nspec, cutoff, A, B, dtheta, nbins, max_cn=32, method=1, hist=None):
    """One frame: (uint64 hist[max_cn+1][nbins], dropped)."""
    pos, pp = _d(pos); cell, cp = _d(cell); cutoff, kp = _d(cutoff)
    spec = np.ascontiguousarray(spec, dtype=np.uint8)
    if hist is None:
        hist = np.zeros((max_cn + 1, nbins), dtype=np.uint64)
    dropped = C.c_uint64(0)
    _check(lib().orc_bad_frame(len(spec), pp, cp, spec.ctypes.data_as(_u8p), nspec, kp, int(A), int(B),
                               float(dtheta), int(nbins), int(max_cn), method, hist.ctypes.data_as(_u64p),
                               C.byref(dropped)), "bad_frame")
    return hist, dropped.value


def set_conventions(bin_rule=0, dv_rule=0):
    """Process-wide switches of the C oracle (see the header of amof_oracle.c); call again with no arguments to restore the pins."""
    lib().orc_set_conventions(int(bin_rule), int(dv_rule))


def angle_of_cosine(x):
    """degrees(arccos(clip(x, -1, 1))) with the container's libm (pin P6)."""
    return float(lib().orc_angle_of_cosine(float(x)))


def bad_angles(pos, cell, spec, nspec, cutoff, A, B, method=1, cap=1 << 20):
    pos, pp = _d(pos); cell, cp = _d(cell); cutoff, kp = _d(cutoff)
    spec = np.ascontiguousarray(spec, dtype=np.uint8)
    out = np.zeros(cap, dtype=np.float64)
    n = lib().orc_bad_angles(len(spec), pp, cp, spec.ctypes.data_as(_u8p), nspec, kp, int(A), int(B), method,
                             out.ctypes.data_as(_dp), cap)
    _check(n, "bad_angles")
    return out[:min(n, cap)]


def delta_pos(pos, cell):
    pos, pp = _d(pos); cell, cp = _d(cell)
    out = np.empty_like(pos)
    _check(lib().orc_delta_pos(pos.shape[0], pos.shape[1], pp, cp, out.ctypes.data_as(_dp)), "delta_pos")
    return out


def msd_window(pos, cell, masses, spec, nspec, window, unwrap=False):
    """Returns (msd[nspec][nwin], mutated positions) -- the reference mutates the caller's frames (Q7)."""
    pos = np.array(pos, dtype=np.float64, order='C', copy=True)
    cell, cp = _d(cell); masses, mp = _d(masses)
    spec = np.ascontiguousarray(spec, dtype=np.uint8)
    window = np.ascontiguousarray(window, dtype=np.int32)
    msd = np.zeros((nspec, len(window)), dtype=np.float64)
    _check(lib().orc_msd_window(pos.shape[0], pos.shape[1], pos.ctypes.data_as(_dp), cp, mp,
                                spec.ctypes.data_as(_u8p), nspec, window.ctypes.data_as(_ip), len(window),
                                int(bool(unwrap)), msd.ctypes.data_as(_dp)), "msd_window")
    return msd, pos


def msd_direct(pos, cell, spec, sel):
    pos, pp = _d(pos); cell, cp = _d(cell)
    spec = np.ascontiguousarray(spec, dtype=np.uint8)
    out = np.zeros(pos.shape[0], dtype=np.float64)
    _check(lib().orc_msd_direct(pos.shape[0], pos.shape[1], pp, cp, spec.ctypes.data_as(_u8p), int(sel),
                                out.ctypes.data_as(_dp)), "msd_direct")
    return out


def theta_bin(theta, dtheta, nbins):
    return lib().orc_theta_bin(float(theta), float(dtheta), int(nbins))


def max_threads():
    return lib().orc_max_threads()
