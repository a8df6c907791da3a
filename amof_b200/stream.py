"""
Streaming trajectory ingest: XYZ-family text files -> page-locked ``positions[F][N][3]`` chunks, parsed while the GPU works.

aMOF reads trajectories with ``ase.io.read`` into a list of Atoms (/root/reference/amof/trajectory.py:37-60) and offers
``read_lammps_traj`` (xyz + optional cell array) and ``read_cp2k_traj`` (xyz + CP2K ``.cell`` file, optionally gzipped:
trajectory.py:193-228).  End to end that list is the bottleneck, so :class:`XyzStream` is a LAZY trajectory:

  * a valid aMOF trajectory -- ``len()``, indexing (a frame is parsed when asked for), iteration -- whose cells are known
    up front (extended-XYZ ``Lattice=`` headers, a constant cell, a cell array, or the CP2K cell file);
  * and a chunk source for the analyses: :func:`amof_b200.frames.iter_chunks` asks it for ``(positions, cells)`` blocks, which
    the library's native parser (``amofb_xyz_parse``, all host cores) fills into a ring of page-locked buffers ahead of the consumer,
    so chunk k+1.. are being parsed while chunk k is copied and counted.

The file is memory-mapped; one parallel pass (``amofb_xyz_index``) finds where every frame starts; gzipped files are inflated
to a temporary file first, as ``Trajectory.from_traj(unzip=True)`` does.
"""
import gzip
import logging
import os
import shutil
import tempfile
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .atoms import _LATTICE, _PROPS, Atoms
from .elements import atomic_numbers

logger = logging.getLogger(__name__)


def _map_file(path):
    """the file as a read-only uint8 array (memory-mapped: the index and the parser read the page cache in place)"""
    size = os.path.getsize(path)
    if size == 0:
        return np.zeros(0, dtype=np.uint8)
    return np.memmap(path, dtype=np.uint8, mode="r")


def _frame_offsets(data, period):
    """byte offset of the first line of every frame (a frame = ``period`` lines) of the mapped file, and its size"""
    from . import _lib
    size = len(data)
    while size > 0 and data[size - 1] in (10, 13, 32, 9):      # blank space after the last atom line is not part of a frame
        size -= 1
    if size == 0:
        return np.zeros(0, dtype=np.int64), 0
    found, lines = _lib.xyz_index(data[:size], 0, period, 0)
    if (lines + 1) % period != 0:
        raise ValueError("truncated XYZ file: %d lines are not a whole number of frames of %d lines" % (lines + 1, period))
    offs = np.concatenate([np.zeros(1, dtype=np.int64), found])
    return offs, size


class XyzStream(object):
    """A lazily read XYZ-family trajectory (see the module docstring).

    Args:
        path: ``.xyz`` / extended-XYZ file, optionally ``.gz``
        cell: None (extended-XYZ ``Lattice=`` headers), one 3x3 cell, or an array [T][3][3] / [T][9]
        index: a slice selecting frames (the reference also takes ase's 'a:b:c' strings; a slice is what
            ``read_cp2k_traj`` supports, trajectory.py:213-214)
        unzip: inflate a gzipped file to a temporary file first (implied by a ``.gz`` suffix)
    """

    def __init__(self, path, cell=None, index=None, unzip=False, chunk_bytes=32 << 20, threads=None):
        self._tmp = None
        self._symbytes = None
        if unzip or str(path).endswith(".gz"):
            logger.info("Unzip trajectory file")
            self._tmp = tempfile.NamedTemporaryFile(suffix=".xyz")
            with gzip.open(path, "rb") as f_in:
                shutil.copyfileobj(f_in, self._tmp)
            self._tmp.flush()
            path = self._tmp.name
        self.path = str(path)
        with open(self.path, "rb") as fh:
            self.n_atoms = int(fh.readline().split()[0])
            header = fh.readline().decode("utf-8", "replace")
        self._period = self.n_atoms + 2
        self._pos_col = 1
        pm = _PROPS.search(header)
        if pm is not None:
            fields = pm.group(1).split(':')
            col = 0
            for name, _kind, width in zip(fields[0::3], fields[1::3], fields[2::3]):
                if name == 'pos':
                    self._pos_col = col
                col += int(width)
        self._data = _map_file(self.path)
        offs, size = _frame_offsets(self._data, self._period)
        self._offs_all = np.append(offs, size)
        n_file = len(offs)
        sel = np.arange(n_file)
        if index is not None:
            if not isinstance(index, slice):
                raise TypeError("index must be a slice")
            sel = sel[index]
        self._sel = sel
        self.n_frames = len(sel)
        # cells
        if cell is None:
            cells = np.empty((n_file, 9))
            with open(self.path, "rb") as fh:
                for k in sel:
                    fh.seek(int(offs[k]))
                    fh.readline()
                    m = _LATTICE.search(fh.readline().decode("utf-8", "replace"))
                    if m is None:
                        raise ValueError("frame %d: no Lattice= in the comment line and no cell given" % k)
                    cells[k] = [float(x) for x in m.group(1).split()]
            self.cells = np.ascontiguousarray(cells[sel].reshape(-1, 3, 3))
        else:
            cell = np.asarray(cell, dtype=np.float64)
            if cell.shape in ((3, 3), (9,)):
                self.cells = np.ascontiguousarray(np.broadcast_to(cell.reshape(3, 3), (self.n_frames, 3, 3)))
            else:
                cell = cell.reshape(len(cell), 3, 3)
                if len(cell) == n_file and index is not None:
                    cell = cell[index]
                if len(cell) != self.n_frames:
                    # Trajectory.set_cell(fit_size=True) trims the longer of the two (trajectory.py:99-109)
                    logger.warning("Mismatch in file sizes; traj: %s vs cell: %s", self.n_frames, len(cell))
                    m = min(len(cell), self.n_frames)
                    cell, self._sel, self.n_frames = cell[:m], self._sel[:m], m
                self.cells = np.ascontiguousarray(cell)
        first = self._parse_frames(self._sel[:1]) if self.n_frames else (np.empty((0, self.n_atoms, 3)), [])
        self._symbols = first[1]
        self.numbers = np.array([atomic_numbers[s] for s in self._symbols], dtype=np.int64)
        self._first_positions = first[0][0] if self.n_frames else None
        frame_bytes = max(1, int((self._offs_all[-1]) // max(n_file, 1)))
        self._chunk_frames = max(1, int(chunk_bytes // frame_bytes))
        self._threads = threads or max(1, min(16, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 2)))

    # ---- parsing ---------------------------------------------------------------------------------------
    def _parse_frames(self, frame_ids, out=None, threads=1):
        """positions [len(frame_ids)][N][3] of the given file frames (+ the symbols of the first one), through the library's
        native parser (amofb_xyz_parse: correctly rounded decimal -> binary64, like float(); ctypes releases the GIL)"""
        from . import _lib
        ids = np.asarray(frame_ids, dtype=np.int64)
        n = self.n_atoms
        if out is None:
            out = np.empty((len(ids), n, 3))
        known = getattr(self, "_symbytes", None) is not None
        symbytes = self._symbytes if known else bytearray(8 * n)
        i = 0
        while i < len(ids):                       # consecutive file frames are one call on the mapped file
            j = i
            while j + 1 < len(ids) and ids[j + 1] == ids[j] + 1:
                j += 1
            a, b = int(ids[i]), int(ids[j])
            base, end = int(self._offs_all[a]), int(self._offs_all[b + 1])
            try:
                _lib.xyz_parse(self._data[base:end], self._offs_all[a:b + 2] - base, n, self._pos_col, symbytes, known, out[i:j + 1],
                               threads=threads)
            except ValueError as err:
                raise ValueError("%s: %s (block starting at file frame %d)" % (self.path, err, a))
            known = True
            i = j + 1
        if getattr(self, "_symbytes", None) is None:
            self._symbytes = symbytes
        symbols = [bytes(symbytes[8 * k:8 * k + 8]).rstrip(b"\0").decode() for k in range(n)] if len(ids) else []
        return out[:len(ids)], symbols

    # ---- the aMOF trajectory surface -------------------------------------------------------------------
    def __len__(self):
        return self.n_frames

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(self.n_frames))]
        k = int(k)
        if k < 0:
            k += self.n_frames
        if not 0 <= k < self.n_frames:
            raise IndexError(k)
        pos = self._first_positions if k == 0 else self._parse_frames(self._sel[k:k + 1])[0][0]
        return Atoms(numbers=self.numbers, positions=pos, cell=self.cells[k])

    def __iter__(self):
        return (self[k] for k in range(self.n_frames))

    # ---- chunk source of the analyses ------------------------------------------------------------------
    def stream_chunks(self, lo, hi, backend=None):
        """Yield (positions[F][N][3], cells[F][3][3]) for the frames [lo, hi): parser threads run ahead of the consumer into a
        ring of buffers (page-locked when the backend owns a context); a buffer is refilled only after the copies that read it
        have finished."""
        if hi <= lo:
            return
        ctx = getattr(backend, "ctx", None)
        F = min(self._chunk_frames, hi - lo)
        workers = 2 if self._threads > 1 else 1            # one block being parsed while the next is being read
        native = max(1, self._threads // workers)
        nslot = workers + 2
        if ctx is not None:
            bufs = [ctx.scratch("xyzstream%d" % i, (F, self.n_atoms, 3)) for i in range(nslot)]
        else:
            bufs = [np.empty((F, self.n_atoms, 3)) for _ in range(nslot)]
        spans = [(a, min(hi, a + F)) for a in range(lo, hi, F)]
        with ThreadPoolExecutor(max_workers=workers) as pool:
            pending = []
            nxt = 0

            def submit():
                nonlocal nxt
                a, b = spans[nxt]
                pending.append((a, b, pool.submit(self._parse_frames, self._sel[a:b], bufs[nxt % nslot], native)))
                nxt += 1

            while nxt < len(spans) and len(pending) < workers:
                submit()
            done = 0
            while pending:
                a, b, fut = pending.pop(0)
                pos, _ = fut.result()
                yield pos, self.cells[a:b]
                done += 1
                if nxt < len(spans):
                    # slot nxt % nslot was last handed out nslot chunks ago: its copy must have finished before it is refilled
                    if ctx is not None:
                        ctx.sync_copies()
                    submit()

    def close(self):
        self._data = None                       # drops the memory map
        if self._tmp is not None:
            self._tmp.close()
            self._tmp = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def read_cp2k_cell(path_to_cell, index=None):
    """cells [T][3][3] from a CP2K ``.cell`` file: columns 2..10 of every row (Step, Time, Ax .. Cz, Volume), as
    ``read_cp2k_traj`` takes them (trajectory.py:216-225)"""
    cell = np.genfromtxt(path_to_cell)
    if cell.ndim == 1:
        cell = cell[None, :]
    cell = cell[:, 2:-1]
    if index is not None:
        cell = cell[index]
    return np.array([c.reshape(3, 3) for c in np.atleast_2d(cell)])
