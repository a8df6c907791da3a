"""
ctypes binding of ``libamofb.so`` (include/amofb.h) and the :class:`GpuBackend` the analysis classes drive.

There is NO CPU fallback: when the shared library is missing, or no CUDA device can be opened, every
entry point raises.  Nothing in this package imports ``oracle/``.

The backend interface (three methods, numpy in / numpy out) is what ``amof_b200.rdf``/``cn``/``bad``/``msd``
call; it mirrors the accumulator trios of the C ABI one to one:

    pair_counts   amofb_pair_begin / _push / _finish   (asap3 RDF + ase neighbour counts, amof/rdf.py:87-93, amof/cn.py:58-74)
    bad_counts    amofb_bad_begin / _push / _finish    (amof/bad.py:70-114)
    msd_*         amofb_msd_*                          (amof/msd.py:186-268)
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libamofb.so")
if os.environ.get("AMOFB_LIB"):          # another build of the same library (tools/build_variants.sh), for side-by-side timing
    _SO = os.path.abspath(os.environ["AMOFB_LIB"])

AMOFB_MAX_SPECIES = 16
AMOFB_BAD_MAX_CN = 32
AMOFB_OPT_RDF_BIN_RULE = 1

_ERRORS = {-1: ValueError, -2: RuntimeError, -3: RuntimeError, -4: ValueError, -5: MemoryError}

_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p

# every symbol include/amofb.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "amofb_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "amofb_destroy": (C.c_int, [_vp]),
    "amofb_last_error": (C.c_char_p, [_vp]),
    "amofb_version": (C.c_char_p, []),
    "amofb_sync": (C.c_int, [_vp]),
    "amofb_sync_copies": (C.c_int, [_vp]),
    "amofb_launch_count": (C.c_int64, [_vp]),
    "amofb_set_option": (C.c_int, [_vp, C.c_int, C.c_int]),
    "amofb_set_profiling": (C.c_int, [_vp, C.c_int]),
    "amofb_pair_kernel_time": (C.c_int, [_vp, _dp, _i64p, C.c_int]),
    "amofb_timer_mark": (C.c_int, [_vp, C.c_int]),
    "amofb_timer_elapsed": (C.c_int, [_vp, C.c_int, C.c_int, _dp]),
    "amofb_host_alloc": (C.c_int, [_vp, C.c_uint64, C.POINTER(_vp)]),
    "amofb_host_free": (C.c_int, [_vp, _vp]),
    "amofb_device_alloc": (C.c_int, [_vp, C.c_uint64, C.POINTER(_vp)]),
    "amofb_device_free": (C.c_int, [_vp, _vp]),
    "amofb_memcpy_h2d": (C.c_int, [_vp, _vp, _vp, C.c_uint64]),
    "amofb_memcpy_d2h": (C.c_int, [_vp, _vp, _vp, C.c_uint64]),
    "amofb_pair_begin": (C.c_int, [_vp, C.c_int, C.c_int, _u8p, C.c_double, C.c_int, _dp]),
    "amofb_pair_push": (C.c_int, [_vp, C.c_int, _vp, _dp]),
    "amofb_pair_push_device": (C.c_int, [_vp, C.c_int, _vp, _dp]),
    "amofb_pair_finish": (C.c_int, [_vp, _u64p, _u64p, C.c_int64, _i64p, _dp]),
    "amofb_pair_take": (C.c_int, [_vp, _u64p, _u64p, C.c_int64, _i64p, _dp]),
    "amofb_rdf_begin": (C.c_int, [_vp, C.c_int, C.c_int, _u8p, C.c_double, C.c_int]),
    "amofb_rdf_push": (C.c_int, [_vp, C.c_int, _vp, _dp]),
    "amofb_rdf_finish": (C.c_int, [_vp, _u64p, _i64p, _dp]),
    "amofb_cn_begin": (C.c_int, [_vp, C.c_int, C.c_int, _u8p, _dp]),
    "amofb_cn_push": (C.c_int, [_vp, C.c_int, _vp, _dp]),
    "amofb_cn_finish": (C.c_int, [_vp, _u64p, C.c_int64]),
    "amofb_bad_begin": (C.c_int, [_vp, C.c_int, C.c_int, _u8p, _dp, C.c_int, _ip, C.c_double, C.c_int]),
    "amofb_bad_push": (C.c_int, [_vp, C.c_int, _vp, _dp]),
    "amofb_bad_push_device": (C.c_int, [_vp, C.c_int, _vp, _dp]),
    "amofb_bad_finish": (C.c_int, [_vp, _u64p, _u64p, _i64p]),
    "amofb_neigh_count": (C.c_int, [_vp, C.c_int, C.c_int, _u8p, _dp, _dp, _dp, _i64p]),
    "amofb_neigh_fill": (C.c_int, [_vp, C.POINTER(C.c_int32), C.c_int64]),
    "amofb_neigh_fill_ex": (C.c_int, [_vp, C.POINTER(C.c_int32), _dp, C.POINTER(C.c_int32), C.c_int64]),
    "amofb_msd_begin": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _u8p, C.c_int, _dp]),
    "amofb_msd_load": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "amofb_msd_load_device": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "amofb_msd_unwrap": (C.c_int, [_vp]),
    "amofb_msd_com_sums": (C.c_int, [_vp, _dp]),
    "amofb_msd_set_com": (C.c_int, [_vp, _dp]),
    "amofb_msd_slab_frames": (C.c_int, [_vp, _ip]),
    "amofb_msd_slab_sums": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _dp]),
    "amofb_msd_slab_sums_device": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _dp]),
    "amofb_msd_slab_sums_begin": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "amofb_msd_slab_sums_begin_strided": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int64]),
    "amofb_msd_slab_sums_begin_device": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "amofb_msd_slab_sums_wait": (C.c_int, [_vp, _dp]),
    "amofb_msd_slab_commit": (C.c_int, [_vp, _dp]),
    "amofb_msd_window": (C.c_int, [_vp, C.c_int, _ip, _dp]),
    "amofb_msd_direct": (C.c_int, [_vp, _dp]),
    "amofb_msd_get_positions": (C.c_int, [_vp, _dp]),
    "amofb_msd_end": (C.c_int, [_vp]),
    "amofb_guard_violations": (C.c_int64, [_vp]),
    "amofb_xyz_index": (C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _i64p, C.c_int64, _i64p, _i64p]),
    "amofb_xyz_parse": (C.c_int, [_vp, _i64p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int, _dp, C.c_int, _ip]),
}

_lib = None
_lib_lock = threading.Lock()


def library_path():
    return _SO


def load_library():
    """dlopen libamofb.so and type every entry point.  Raises if the library has not been built."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(_SO):
                raise ImportError(
                    "amof_b200: %s is missing. Build it with `python -m amof_b200.build` (needs nvcc); "
                    "there is no CPU fallback." % _SO)
            lib = C.CDLL(_SO)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)     # AttributeError here = header and library disagree
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def _ptr(a, typ):
    return a.ctypes.data_as(typ)


def _text_view(text):
    """bytes, a memory map or a uint8 array -> (uint8 array sharing the memory, its address): nothing is copied"""
    arr = text if isinstance(text, np.ndarray) else np.frombuffer(text, dtype=np.uint8)
    return arr, arr.ctypes.data


def xyz_index(text, lines_before, period, base):
    """amofb_xyz_index on one block of a file (bytes, mmap or uint8 array): (file offsets of the frames starting after a newline of
    this block, newlines seen)"""
    lib = load_library()
    text, addr = _text_view(text)
    cap = len(text) // max(1, 2 * int(period)) + 2      # a line is at least 2 bytes
    starts = np.empty(cap, dtype=np.int64)
    n, lines = C.c_int64(0), C.c_int64(0)
    rc = lib.amofb_xyz_index(addr, len(text), int(lines_before), int(period), int(base), _ptr(starts, _i64p), cap, C.byref(n), C.byref(lines))
    if rc != 0:
        raise ValueError("amofb_xyz_index failed (%d)" % rc)
    return starts[:n.value], lines.value


def xyz_parse(text, frame_off, n_atoms, pos_col, symbols, symbols_known, out, threads=0):
    """amofb_xyz_parse (host code of the library: needs no CUDA device).  ``text`` bytes / mmap / uint8 array, ``frame_off`` int64[F + 1] offsets into it,
    ``symbols`` a writable bytearray of 8 * n_atoms, ``out`` float64[>= F][n_atoms][3] C-contiguous."""
    lib = load_library()
    text, addr = _text_view(text)
    frame_off = np.ascontiguousarray(frame_off, dtype=np.int64)
    F = len(frame_off) - 1
    assert out.dtype == np.float64 and out.flags.c_contiguous and out.size >= F * n_atoms * 3
    bad = C.c_int(-1)
    sym = (C.c_char * len(symbols)).from_buffer(symbols)
    rc = lib.amofb_xyz_parse(addr, _ptr(frame_off, _i64p), F, int(n_atoms), int(pos_col), sym, int(bool(symbols_known)),
                             _ptr(out, _dp), int(threads), C.byref(bad))
    if rc != 0:
        raise ValueError("XYZ text: frame %d of the block is truncated or malformed, or its atom order differs from the first frame's"
                         % bad.value)


class Context:
    """One amofb_ctx (one per GPU).  Not thread-safe; ctypes releases the GIL during calls."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = _vp()
        rc = self.lib.amofb_create(int(device), C.byref(h))
        if rc != 0:
            raise RuntimeError(
                "amof_b200: cannot open CUDA device %d (amofb_create -> %d). This package computes on the GPU only; "
                "there is no CPU fallback." % (device, rc))
        self.h = h
        self.device = int(device)
        self._pinned = []
        self._scratch = {}

    def check(self, rc):
        if rc != 0:
            msg = self.lib.amofb_last_error(self.h)
            msg = msg.decode("utf-8", "replace") if msg else "error %d" % rc
            raise _ERRORS.get(rc, RuntimeError)("amofb: " + msg)

    def close(self):
        if self.h:
            for p in self._pinned:
                self.lib.amofb_host_free(self.h, p)
            self._pinned = []
            self._scratch = {}
            self.lib.amofb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- memory -----------------------------------------------------------------------------------
    def pinned_empty(self, shape, dtype=np.float64):
        """numpy array backed by page-locked host memory owned by this context."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = _vp()
        self.check(self.lib.amofb_host_alloc(self.h, max(n, 1), C.byref(p)))
        self._pinned.append(p)
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def scratch(self, name, shape, dtype=np.float64):
        """Reusable page-locked array: one allocation per name, grown on demand (staging buffers of frames.py)."""
        need = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ent = self._scratch.get(name)
        if ent is None or ent[1] < need:
            # a grown buffer gets a NEW block; the old one stays allocated until close(): a suspended generator of
            # frames.iter_chunks (or any caller) may still hold a numpy view of it
            p = _vp()
            self.check(self.lib.amofb_host_alloc(self.h, max(need, 1), C.byref(p)))
            self._pinned.append(p)
            ent = (p, max(need, 1))
            self._scratch[name] = ent
        buf = (C.c_char * ent[1]).from_address(ent[0].value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def pinned_free(self, arr):
        """Give back a block obtained from :meth:`pinned_empty` (the array must not be used afterwards)."""
        addr = arr.ctypes.data
        for q in self._pinned:
            if q.value == addr:
                self._pinned.remove(q)
                self.check(self.lib.amofb_host_free(self.h, q))
                return
        raise ValueError("not a pinned_empty block of this context")

    def device_alloc(self, nbytes):
        p = _vp()
        self.check(self.lib.amofb_device_alloc(self.h, int(nbytes), C.byref(p)))
        return p

    def device_free(self, p):
        self.check(self.lib.amofb_device_free(self.h, p))

    def h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        self.check(self.lib.amofb_memcpy_h2d(self.h, dptr, arr.ctypes.data, arr.nbytes))

    def d2h(self, arr, dptr):
        self.check(self.lib.amofb_memcpy_d2h(self.h, arr.ctypes.data, dptr, arr.nbytes))

    def sync(self):
        self.check(self.lib.amofb_sync(self.h))

    def sync_copies(self):
        self.check(self.lib.amofb_sync_copies(self.h))

    def launch_count(self):
        return int(self.lib.amofb_launch_count(self.h))

    def guard_violations(self):
        """pooled device blocks found written out of bounds (AMOFB_GUARD=1 at context creation), -1 when the guard is off"""
        return int(self.lib.amofb_guard_violations(self.h))

    def set_option(self, option, value):
        """amofb_set_option: option names are the AMOFB_OPT_* constants of include/amofb.h"""
        self.check(self.lib.amofb_set_option(self.h, int(option), int(value)))

    def set_profiling(self, on):
        self.check(self.lib.amofb_set_profiling(self.h, 1 if on else 0))

    def timer_mark(self, slot):
        self.check(self.lib.amofb_timer_mark(self.h, int(slot)))

    def timer_elapsed(self, a, b):
        ms = C.c_double(0.0)
        self.check(self.lib.amofb_timer_elapsed(self.h, int(a), int(b), C.byref(ms)))
        return ms.value

    def pair_kernel_time(self, reset=True):
        ms, n = C.c_double(0.0), C.c_int64(0)
        self.check(self.lib.amofb_pair_kernel_time(self.h, C.byref(ms), C.byref(n), 1 if reset else 0))
        return ms.value, n.value


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class GpuBackend:
    """numpy-level driver of one :class:`Context`.

    ``chunks`` arguments are iterables of ``(pos, cell)`` with ``pos`` float64 ``[F][N][3]`` and ``cell`` float64
    ``[F][3][3]``.  A ``pos`` that is an int/``c_void_p`` is taken as a DEVICE pointer holding F frames
    (``cell`` then carries F).  Host arrays are passed through as they are: page-locked arrays (from
    :meth:`Context.pinned_empty`) are copied asynchronously while the previous batch computes.
    """

    name = "libamofb/sm_100a"

    def __init__(self, device=0):
        self.ctx = Context(device)

    # -- pair -------------------------------------------------------------------------------------
    def pair_counts(self, species, n_species, chunks, rmax=0.0, nbins=0, cn_cutoff=None):
        """-> dict(hist uint64[S][S][nbins] | None, cn uint64[T][S][S] | None, n_frames, volume_sum)"""
        ctx, lib = self.ctx, self.ctx.lib
        species = np.ascontiguousarray(species, dtype=np.uint8)
        S = int(n_species)
        cut = None if cn_cutoff is None else _f64(cn_cutoff).reshape(S, S)
        ctx.check(lib.amofb_pair_begin(ctx.h, len(species), S, _ptr(species, _u8p), float(rmax), int(nbins),
                                       None if cut is None else _ptr(cut, _dp)))
        try:
            total = self._push_all(lib.amofb_pair_push, lib.amofb_pair_push_device, chunks, len(species))
        except BaseException:
            lib.amofb_pair_finish(ctx.h, None, None, 0, None, None)
            raise
        hist = np.zeros((S, S, int(nbins)), dtype=np.uint64) if nbins > 0 else None
        cn = np.zeros((total, S, S), dtype=np.uint64) if cut is not None else None
        nf, vs = C.c_int64(0), C.c_double(0.0)
        ctx.check(lib.amofb_pair_finish(ctx.h, None if hist is None else _ptr(hist, _u64p),
                                        None if cn is None else _ptr(cn, _u64p), total, C.byref(nf), C.byref(vs)))
        return {"hist": hist, "cn": cn, "n_frames": int(nf.value), "volume_sum": float(vs.value)}

    def pair_counts_each(self, species, n_species, chunks, rmax, nbins):
        """One RDF histogram PER CHUNK (generator of pair_counts-like dicts): one analysis stays open and is emptied after each
        chunk (amofb_pair_take), instead of a begin/finish pair per frame."""
        ctx, lib = self.ctx, self.ctx.lib
        species = np.ascontiguousarray(species, dtype=np.uint8)
        S = int(n_species)
        ctx.check(lib.amofb_pair_begin(ctx.h, len(species), S, _ptr(species, _u8p), float(rmax), int(nbins), None))
        try:
            for chunk in chunks:
                self._push_all(lib.amofb_pair_push, lib.amofb_pair_push_device, [chunk], len(species))
                hist = np.zeros((S, S, int(nbins)), dtype=np.uint64)
                nf, vs = C.c_int64(0), C.c_double(0.0)
                ctx.check(lib.amofb_pair_take(ctx.h, _ptr(hist, _u64p), None, 0, C.byref(nf), C.byref(vs)))
                yield {"hist": hist, "cn": None, "n_frames": int(nf.value), "volume_sum": float(vs.value)}
        finally:
            lib.amofb_pair_finish(ctx.h, None, None, 0, None, None)

    def _push_all(self, push_host, push_device, chunks, n_atoms):
        ctx = self.ctx
        total = 0
        keep = []   # host arrays must stay alive until their copies were issued
        for pos, cell in chunks:
            cell = _f64(cell).reshape(-1, 3, 3)
            F = cell.shape[0]
            if isinstance(pos, (int, C.c_void_p)):
                ctx.check(push_device(ctx.h, F, pos, _ptr(cell, _dp)))
            else:
                pos = _f64(pos)
                if pos.shape != (F, n_atoms, 3):
                    raise ValueError("positions chunk has shape %r, expected %r" % (pos.shape, (F, n_atoms, 3)))
                ctx.check(push_host(ctx.h, F, pos.ctypes.data, _ptr(cell, _dp)))
                keep.append(pos)
                if len(keep) > 4:
                    keep.pop(0)
            total += F
        ctx.sync_copies()
        return total

    # -- bond angles ------------------------------------------------------------------------------
    def bad_counts(self, species, n_species, chunks, cutoff, triples, dtheta, nbins):
        """-> (hist uint64[n_triples][AMOFB_BAD_MAX_CN+1][nbins], dropped uint64[n_triples], n_frames)"""
        ctx, lib = self.ctx, self.ctx.lib
        species = np.ascontiguousarray(species, dtype=np.uint8)
        S = int(n_species)
        cut = _f64(cutoff).reshape(S, S)
        tr = np.ascontiguousarray(triples, dtype=np.int32).reshape(-1, 2)
        ctx.check(lib.amofb_bad_begin(ctx.h, len(species), S, _ptr(species, _u8p), _ptr(cut, _dp), len(tr),
                                      _ptr(tr, _ip), float(dtheta), int(nbins)))
        try:
            self._push_all(lib.amofb_bad_push, lib.amofb_bad_push_device, chunks, len(species))
        except BaseException:
            lib.amofb_bad_finish(ctx.h, None, None, None)
            raise
        hist = np.zeros((len(tr), AMOFB_BAD_MAX_CN + 1, int(nbins)), dtype=np.uint64)
        dropped = np.zeros(len(tr), dtype=np.uint64)
        nf = C.c_int64(0)
        ctx.check(lib.amofb_bad_finish(ctx.h, _ptr(hist, _u64p), _ptr(dropped, _u64p), C.byref(nf)))
        return hist, dropped, int(nf.value)

    # -- explicit neighbour list --------------------------------------------------------------------
    def neighbour_list(self, species, n_species, positions, cell, cutoff, quantities=False):
        """One frame -> (offsets int64[n+1], neighbours int32[offsets[n]]): row i = neighbours[offsets[i]:offsets[i+1]],
        ascending original indices, one entry per periodic image under cutoff[Zi][Zj] (amof/atom.py:72-87).
        ``quantities=True`` adds (distances float64[...], shifts int32[...][3]): ase's 'd' and 'S' of every pair."""
        ctx, lib = self.ctx, self.ctx.lib
        species = np.ascontiguousarray(species, dtype=np.uint8)
        S = int(n_species)
        cut = _f64(cutoff).reshape(S, S)
        pos = _f64(positions).reshape(len(species), 3)
        cell = _f64(cell).reshape(3, 3)
        offsets = np.zeros(len(species) + 1, dtype=np.int64)
        ctx.check(lib.amofb_neigh_count(ctx.h, len(species), S, _ptr(species, _u8p), _ptr(cut, _dp), _ptr(pos, _dp),
                                        _ptr(cell, _dp), _ptr(offsets, _i64p)))
        total = int(offsets[-1])
        nbr = np.zeros(total, dtype=np.int32)
        i32p = C.POINTER(C.c_int32)
        if not quantities:
            ctx.check(lib.amofb_neigh_fill(ctx.h, _ptr(nbr, i32p) if total else None, total))
            return offsets, nbr
        dist = np.zeros(total, dtype=np.float64)
        shifts = np.zeros((total, 3), dtype=np.int32)
        ctx.check(lib.amofb_neigh_fill_ex(ctx.h, _ptr(nbr, i32p) if total else None, _ptr(dist, _dp) if total else None,
                                          _ptr(shifts, i32p) if total else None, total))
        return offsets, nbr, dist, shifts

    # -- MSD --------------------------------------------------------------------------------------
    def msd_open(self, n_frames, masses, species, n_species, cells):
        ctx, lib = self.ctx, self.ctx.lib
        masses = _f64(masses)
        species = np.ascontiguousarray(species, dtype=np.uint8)
        cells = _f64(cells).reshape(int(n_frames), 3, 3)
        ctx.check(lib.amofb_msd_begin(ctx.h, int(n_frames), len(species), _ptr(masses, _dp), _ptr(species, _u8p),
                                      int(n_species), _ptr(cells, _dp)))
        return _MsdSession(self, int(n_frames), len(species), int(n_species))


class _MsdSession:
    """amofb_msd_* state machine: load -> [unwrap] -> com_sums -> set_com -> window | direct -> close,
    or the streaming path: (slab_sums -> slab_commit) per slab of frames, in order -> window -> close."""

    def __init__(self, backend, T, n, S):
        self.b, self.T, self.n, self.S = backend, T, n, S
        self.open = True

    def load(self, first, pos):
        ctx = self.b.ctx
        if isinstance(pos, tuple):        # (device pointer, count)
            ctx.check(ctx.lib.amofb_msd_load_device(ctx.h, int(first), int(pos[1]), pos[0]))
            return
        pos = _f64(pos)
        if pos.ndim != 3 or pos.shape[1:] != (self.n, 3):
            raise ValueError("positions chunk has shape %r, expected (F, %d, 3)" % (pos.shape, self.n))
        ctx.check(ctx.lib.amofb_msd_load(ctx.h, int(first), pos.shape[0], pos.ctypes.data))
        ctx.sync_copies()

    def slab_frames(self):
        """frames a host slab may hold (amofb_msd_slab_frames)"""
        ctx = self.b.ctx
        n = C.c_int(0)
        ctx.check(ctx.lib.amofb_msd_slab_frames(ctx.h, C.byref(n)))
        return int(n.value)

    def slab_sums(self, first, pos):
        """-> float64[count][4] = (sum m x, sum m y, sum m z, sum m) of the local atoms of every frame of the slab"""
        ctx = self.b.ctx
        if isinstance(pos, tuple):        # (device pointer, count)
            count = int(pos[1])
            out = np.zeros((count, 4), dtype=np.float64)
            ctx.check(ctx.lib.amofb_msd_slab_sums_device(ctx.h, int(first), count, pos[0], _ptr(out, _dp)))
            return out
        pos = _f64(pos)
        if pos.ndim != 3 or pos.shape[1:] != (self.n, 3):
            raise ValueError("positions slab has shape %r, expected (F, %d, 3)" % (pos.shape, self.n))
        out = np.zeros((pos.shape[0], 4), dtype=np.float64)
        ctx.check(ctx.lib.amofb_msd_slab_sums(ctx.h, int(first), pos.shape[0], pos.ctypes.data, _ptr(out, _dp)))
        return out

    def slab_sums_begin(self, first, pos):
        """enqueue the sums of a slab (at most two slabs may await their commit); returns the number of frames"""
        ctx = self.b.ctx
        if isinstance(pos, tuple):        # (device pointer, count)
            ctx.check(ctx.lib.amofb_msd_slab_sums_begin_device(ctx.h, int(first), int(pos[1]), pos[0]))
            return int(pos[1])
        if (isinstance(pos, np.ndarray) and pos.dtype == np.float64 and pos.ndim == 3 and pos.shape[1:] == (self.n, 3)
                and pos.strides[1:] == (24, 8) and pos.strides[0] > 24 * self.n and pos.strides[0] % 8 == 0):
            # a column block of wider frames (positions[a:b, lo:hi]): strided copy straight from the caller's array
            ctx.check(ctx.lib.amofb_msd_slab_sums_begin_strided(ctx.h, int(first), pos.shape[0], pos.ctypes.data, pos.strides[0] // 8))
            self._keep = pos
            return pos.shape[0]
        pos = _f64(pos)
        if pos.ndim != 3 or pos.shape[1:] != (self.n, 3):
            raise ValueError("positions slab has shape %r, expected (F, %d, 3)" % (pos.shape, self.n))
        ctx.check(ctx.lib.amofb_msd_slab_sums_begin(ctx.h, int(first), pos.shape[0], pos.ctypes.data))
        self._keep = pos                  # the copy may still be in flight
        return pos.shape[0]

    def slab_sums_wait(self, count):
        ctx = self.b.ctx
        out = np.zeros((int(count), 4), dtype=np.float64)
        ctx.check(ctx.lib.amofb_msd_slab_sums_wait(ctx.h, _ptr(out, _dp)))
        return out

    def slab_commit(self, com):
        ctx = self.b.ctx
        com = _f64(com).reshape(-1, 3)
        ctx.check(ctx.lib.amofb_msd_slab_commit(ctx.h, _ptr(com, _dp)))

    def unwrap(self):
        ctx = self.b.ctx
        ctx.check(ctx.lib.amofb_msd_unwrap(ctx.h))

    def com_sums(self):
        ctx = self.b.ctx
        out = np.zeros((self.T, 4), dtype=np.float64)
        ctx.check(ctx.lib.amofb_msd_com_sums(ctx.h, _ptr(out, _dp)))
        return out

    def set_com(self, com):
        ctx = self.b.ctx
        com = _f64(com).reshape(self.T, 3)
        ctx.check(ctx.lib.amofb_msd_set_com(ctx.h, _ptr(com, _dp)))

    def window(self, window):
        ctx = self.b.ctx
        window = np.ascontiguousarray(window, dtype=np.int32)
        out = np.zeros((self.S, len(window)), dtype=np.float64)
        ctx.check(ctx.lib.amofb_msd_window(ctx.h, len(window), _ptr(window, _ip), _ptr(out, _dp)))
        return out

    def direct(self):
        ctx = self.b.ctx
        out = np.zeros((self.S, self.T), dtype=np.float64)
        ctx.check(ctx.lib.amofb_msd_direct(ctx.h, _ptr(out, _dp)))
        return out

    def get_positions(self):
        ctx = self.b.ctx
        out = np.zeros((self.T, self.n, 3), dtype=np.float64)
        ctx.check(ctx.lib.amofb_msd_get_positions(ctx.h, _ptr(out, _dp)))
        return out

    def close(self):
        if self.open:
            ctx = self.b.ctx
            ctx.lib.amofb_msd_end(ctx.h)
            self.open = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


# ---- the process-wide backend -----------------------------------------------------------------------
_backend = None
_backend_lock = threading.Lock()


def default_device():
    for var in ("AMOFB_DEVICE", "LOCAL_RANK"):
        v = os.environ.get(var)
        if v not in (None, ""):
            return int(v)
    return 0


def get_backend():
    """The GPU backend of this process (created on first use).  Raises without a CUDA device."""
    global _backend
    with _backend_lock:
        if _backend is None:
            _backend = GpuBackend(default_device())
        return _backend


def _set_backend_for_tests(backend):
    """Test hook: lets tests/ exercise the host-side logic (schemas, sharding, normalisation) on a machine
    without a GPU by substituting an object with the :class:`GpuBackend` interface.  Never called by the
    package itself; passing ``None`` restores the GPU backend."""
    global _backend
    with _backend_lock:
        old, _backend = _backend, backend
    return old
