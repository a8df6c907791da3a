"""
Streaming trajectory ingest: XYZ-family text files -> page-locked ``positions[F][N][3]`` chunks, parsed while the GPU works.

aMOF reads trajectories with ``ase.io.read`` into a list of Atoms (/root/reference/amof/trajectory.py:37-60) and offers
``read_lammps_traj`` (xyz + optional cell array) and ``read_cp2k_traj`` (xyz + CP2K ``.cell`` file, optionally gzipped:
trajectory.py:193-228).  End to end that list is the bottleneck, so :class:`XyzStream` is a LAZY trajectory:

  * a valid aMOF trajectory -- ``len()``, indexing (a frame is parsed when asked for), iteration -- whose cells are known
    up front (extended-XYZ ``Lattice=`` headers, a constant cell, a cell array, or the CP2K cell file);
  * and a chunk source for the analyses: :func:`amof_b200.frames.iter_chunks` asks it for ``(positions, cells)`` blocks, which
    a few parser threads (pandas' C tokenizer releases the GIL) fill into a ring of page-locked buffers ahead of the consumer,
    so chunk k+1.. are being parsed while chunk k is copied and counted.

One pass over the file finds where every frame starts (newline counting on raw bytes, vectorised); gzipped files are
inflated to a temporary file first, as ``Trajectory.from_traj(unzip=True)`` does.
"""
import gzip
import io
import logging
import os
import shutil
import tempfile
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .atoms import _LATTICE, _PROPS, Atoms
from .elements import atomic_numbers

logger = logging.getLogger(__name__)


def _frame_offsets(path, period):
    """byte offset of the first line of every frame (a frame = ``period`` lines), and the file size"""
    size = os.path.getsize(path)
    starts = [np.zeros(1, dtype=np.int64)]
    lines_before = 0
    with open(path, "rb") as fh:
        base = 0
        while True:
            buf = fh.read(64 << 20)
            if not buf:
                break
            nl = np.flatnonzero(np.frombuffer(buf, dtype=np.uint8) == 10)
            # newline i of this block ends line (lines_before + i); the line after it starts a frame when its number is a
            # multiple of the period
            ends = lines_before + 1 + np.arange(len(nl), dtype=np.int64)
            hit = ends % period == 0
            starts.append(base + nl[hit].astype(np.int64) + 1)
            lines_before += len(nl)
            base += len(buf)
    offs = np.concatenate(starts)
    offs = offs[offs < size]                       # a trailing newline does not start a frame
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

    def __init__(self, path, cell=None, index=None, unzip=False, chunk_bytes=48 << 20, threads=None):
        self._tmp = None
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
        offs, size = _frame_offsets(self.path, self._period)
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
        self._threads = threads or max(1, min(8, (os.cpu_count() or 2) // 2))

    # ---- parsing ---------------------------------------------------------------------------------------
    def _parse_frames(self, frame_ids, out=None):
        """positions [len(frame_ids)][N][3] of the given file frames (+ the symbols of the first one)"""
        import pandas as pd
        bodies = []
        with open(self.path, "rb") as fh:
            runs = []
            ids = np.asarray(frame_ids)
            i = 0
            while i < len(ids):                   # consecutive file frames are one read
                j = i
                while j + 1 < len(ids) and ids[j + 1] == ids[j] + 1:
                    j += 1
                runs.append((ids[i], ids[j]))
                i = j + 1
            for a, b in runs:
                fh.seek(int(self._offs_all[a]))
                buf = fh.read(int(self._offs_all[b + 1] - self._offs_all[a]))
                base = int(self._offs_all[a])
                for k in range(a, b + 1):
                    lo = int(self._offs_all[k]) - base
                    hi = int(self._offs_all[k + 1]) - base
                    p = buf.find(b"\n", lo) + 1
                    p = buf.find(b"\n", p) + 1                      # skip the count and the comment line
                    bodies.append(buf[p:hi])
        pc = self._pos_col
        body = pd.read_csv(io.BytesIO(b"".join(bodies)), sep=r"\s+", header=None, usecols=[0, pc, pc + 1, pc + 2], engine="c",
                           names=["s", "x", "y", "z"], float_precision="round_trip")      # bit-exact decimal -> binary64, like float()
        n = self.n_atoms
        if len(body) != len(frame_ids) * n:
            raise ValueError("truncated XYZ file: %d atom lines for %d frames of %d atoms" % (len(body), len(frame_ids), n))
        sym = body["s"].to_numpy()
        symbols = [str(s) for s in sym[:n]]
        if getattr(self, "_symbols", None):
            want = np.array(self._symbols, dtype=object)
            if not (sym.reshape(len(frame_ids), n) == want[None, :]).all():
                raise ValueError("atom order changes between frames")
        pos = body[["x", "y", "z"]].to_numpy(dtype=np.float64).reshape(len(frame_ids), n, 3)
        if out is not None:
            out[:len(frame_ids)] = pos
            pos = out[:len(frame_ids)]
        return pos, symbols

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
        nslot = self._threads + 2
        if ctx is not None:
            bufs = [ctx.scratch("xyzstream%d" % i, (F, self.n_atoms, 3)) for i in range(nslot)]
        else:
            bufs = [np.empty((F, self.n_atoms, 3)) for _ in range(nslot)]
        spans = [(a, min(hi, a + F)) for a in range(lo, hi, F)]
        with ThreadPoolExecutor(max_workers=self._threads) as pool:
            pending = []
            nxt = 0

            def submit():
                nonlocal nxt
                a, b = spans[nxt]
                pending.append((a, b, pool.submit(self._parse_frames, self._sel[a:b], bufs[nxt % nslot])))
                nxt += 1

            while nxt < len(spans) and len(pending) < self._threads:
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
