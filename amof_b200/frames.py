"""
Trajectory ingest: ``list[ase.Atoms]`` (or anything indexable with ``len``) -> contiguous float64 chunks.

aMOF's only trajectory contract is "an indexable, re-iterable sequence of ase.Atoms"
(/root/reference/amof/rdf.py:71-88, amof/cn.py:77, amof/bad.py:149, amof/msd.py:218-237).  The GPU path wants
``positions[F][N][3]`` + ``cell[F][3][3]`` blocks in page-locked memory, so frames are packed chunk by chunk into
two alternating pinned buffers: while the GPU works on chunk k the interpreter packs chunk k+1.

:class:`ArrayTrajectory` is an array-backed trajectory (still a valid aMOF trajectory: indexing yields Atoms)
whose chunks are handed to the GPU without any per-frame Python work.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _dist
from .atoms import Atoms

_POOL = None
_POOL_THREADS = max(1, min(8, (os.cpu_count() or 2) // 2))


def _parallel_copy(dst, sources):
    """dst[i] = sources[i] for every frame, split over a few threads: numpy releases the GIL inside large copies, and
    one core's memcpy (~10 GB/s) is slower than the PCIe link the chunk is about to cross."""
    global _POOL
    n = len(sources)
    if n == 0:
        return
    if _POOL_THREADS == 1 or n < 2 * _POOL_THREADS:
        for i, src in enumerate(sources):
            dst[i] = src
        return
    if _POOL is None:
        _POOL = ThreadPoolExecutor(max_workers=_POOL_THREADS)

    def work(lo, hi):
        for i in range(lo, hi):
            dst[i] = sources[i]
    cuts = [(n * t) // _POOL_THREADS for t in range(_POOL_THREADS + 1)]
    list(_POOL.map(lambda ab: work(*ab), zip(cuts[:-1], cuts[1:])))


class ArrayTrajectory:
    """A trajectory stored as arrays: ``numbers[N]``, ``positions[T][N][3]``, ``cells[T][3][3]`` (or one [3][3]).

    A rank of a frame-sharded job may hold only ITS block of frames: ``positions`` then covers the frames
    ``[first_frame, first_frame + len(positions))`` of a trajectory of ``n_frames`` frames (``cells`` always covers all of
    them).  ``len()`` is the length of the whole trajectory, so the analyses shard it exactly as they shard a fully
    resident one (:func:`frame_range`); indexing a frame that is not resident yields an Atoms that carries the atomic
    numbers, masses and cell but NaN positions.

    A subclass that serves its frames from a bounded buffer (``block(a, b)``) may set ``block_frames``: WindowMsd then asks
    for slabs of at most that many frames."""

    block_frames = None

    def __init__(self, numbers, positions, cells, masses=None, pinned=False, first_frame=0, n_frames=None):
        self.pinned = bool(pinned)       # positions live in page-locked memory (Context.pinned_empty): copied as they are
        self.numbers = np.asarray(numbers, dtype=np.int64)
        self.positions = np.asarray(positions, dtype=np.float64)
        if self.positions.ndim != 3 or self.positions.shape[1:] != (len(self.numbers), 3):
            raise ValueError("positions must be [T][N][3]")
        self.first_frame = int(first_frame)
        self.n_frames = self.first_frame + self.positions.shape[0] if n_frames is None else int(n_frames)
        if self.first_frame < 0 or self.first_frame + self.positions.shape[0] > self.n_frames:
            raise ValueError("resident frames [%d, %d) do not fit a trajectory of %d frames"
                             % (self.first_frame, self.first_frame + self.positions.shape[0], self.n_frames))
        cells = np.asarray(cells, dtype=np.float64)
        if cells.shape == (3, 3):
            cells = np.broadcast_to(cells, (self.n_frames, 3, 3))
        self.cells = np.ascontiguousarray(cells).reshape(self.n_frames, 3, 3)
        self.masses = None if masses is None else np.asarray(masses, dtype=np.float64)

    def __len__(self):
        return self.n_frames

    def resident(self, a, b):
        return self.first_frame <= a and b <= self.first_frame + self.positions.shape[0]

    def block(self, a, b):
        """positions of the frames [a, b) (a view); they must be resident"""
        if not self.resident(a, b):
            raise IndexError("frames [%d, %d) are not resident (this rank holds [%d, %d))"
                             % (a, b, self.first_frame, self.first_frame + self.positions.shape[0]))
        return self.positions[a - self.first_frame:b - self.first_frame]

    def __getitem__(self, k):
        if isinstance(k, slice):
            a, b, step = k.indices(self.n_frames)
            if step != 1:
                raise IndexError("ArrayTrajectory slices must be contiguous")
            return ArrayTrajectory(self.numbers, self.block(a, max(a, b)), self.cells[a:max(a, b)], self.masses, self.pinned)
        k = int(k)
        if k < 0:
            k += self.n_frames
        if not 0 <= k < self.n_frames:
            raise IndexError(k)
        if self.resident(k, k + 1):
            a = Atoms(numbers=self.numbers, positions=self.positions[k - self.first_frame], cell=self.cells[k], masses=self.masses)
        else:
            a = Atoms(numbers=self.numbers, positions=np.full((len(self.numbers), 3), np.nan), cell=self.cells[k], masses=self.masses)
        a._parent = (self, k)
        return a

    def __iter__(self):
        return (self[k] for k in range(len(self)))


def _positions_of(atoms):
    p = getattr(atoms, "positions", None)
    return atoms.get_positions() if p is None else p


def _cell_of(atoms):
    return np.asarray(atoms.get_cell(), dtype=np.float64).reshape(3, 3)


def species_index(numbers):
    """atomic numbers -> (sorted unique Z, uint8 index per atom).  The index order is internal; output column
    order is decided by the callers exactly as the reference does (``list(set(Z))``)."""
    numbers = np.asarray(numbers)
    zs = np.unique(numbers)
    if len(zs) > 16:
        raise ValueError("more than 16 chemical species are not supported by libamofb (AMOFB_MAX_SPECIES)")
    return [int(z) for z in zs], np.searchsorted(zs, numbers).astype(np.uint8)


def frame_range(n_frames, distributed=None):
    """Contiguous block of frames this rank owns (frames shard naturally)."""
    rank, world = _dist.rank_world(distributed)
    lo = (n_frames * rank) // world
    hi = (n_frames * (rank + 1)) // world
    return lo, hi


def _numbers_of(atoms):
    z = getattr(atoms, "numbers", None)          # ase.Atoms.numbers / our shim: the array itself, no copy
    return atoms.get_atomic_numbers() if z is None else z


def check_same_atoms(trajectory, numbers, lo, hi):
    """The C ABI takes one species vector per analysis; aMOF trajectories keep atom order fixed."""
    if isinstance(trajectory, ArrayTrajectory) or hasattr(trajectory, "stream_chunks"):
        return                                   # one species vector by construction (a stream checks the symbols as it parses)
    for k in range(lo, hi):
        z = _numbers_of(trajectory[k])
        if z is numbers:
            continue
        if len(z) != len(numbers) or not np.array_equal(z, numbers):
            raise ValueError("frame %d has different atoms (count or order) than frame 0; "
                             "amof_b200 needs a fixed atom order over the trajectory" % k)


def gather_cells(trajectory, lo=0, hi=None):
    """cells[hi-lo][3][3] of the frames [lo, hi)"""
    hi = len(trajectory) if hi is None else hi
    if isinstance(trajectory, ArrayTrajectory) or hasattr(trajectory, "stream_chunks"):
        return trajectory.cells[lo:hi]
    out = np.empty((max(hi - lo, 0), 3, 3))
    for k in range(lo, hi):
        out[k - lo] = _cell_of(trajectory[k])
    return out


def iter_chunks(trajectory, lo, hi, backend, target_bytes=192 << 20):
    """Yield (positions[F][N][3], cell[F][3][3]) for frames [lo, hi).

    ArrayTrajectory: zero-copy slices.  Otherwise frames are packed into two alternating page-locked buffers
    obtained from the backend's context (plain numpy buffers when the backend has none); the per-frame work in the
    interpreter is two attribute reads, the copies of a chunk are one ``np.concatenate`` into the buffer."""
    if hi <= lo:
        return
    if hasattr(trajectory, "stream_chunks"):     # a lazily read file (amof_b200.stream.XyzStream): parsed ahead by its own threads
        yield from trajectory.stream_chunks(lo, hi, backend)
        return
    ctx = getattr(backend, "ctx", None)
    is_array = isinstance(trajectory, ArrayTrajectory)
    n = len(trajectory.numbers) if is_array else len(trajectory[lo])
    step = max(1, min(hi - lo, int(target_bytes // (24 * max(n, 1)))))
    if is_array and (trajectory.pinned or ctx is None):
        for a in range(lo, hi, step):
            b = min(hi, a + step)
            yield trajectory.block(a, b), trajectory.cells[a:b]
        return
    if ctx is not None:
        bufs = [ctx.scratch("chunks%d" % i, (step, n, 3)) for i in range(2)]        # own names: msd._load_local stages too
    else:
        bufs = [np.empty((step, n, 3)) for _ in range(2)]
    which = 0
    for a in range(lo, hi, step):
        b = min(hi, a + step)
        buf = bufs[which]
        which ^= 1
        if ctx is not None:
            ctx.sync_copies()            # the copy that last read this buffer has finished
        if is_array:                     # pageable array: stage it ourselves, threaded, instead of the driver's bounce
            blk = trajectory.block(a, b)
            _parallel_copy(buf, [blk[k] for k in range(b - a)])
            cells = trajectory.cells[a:b]
        else:
            frames_ = [trajectory[k] for k in range(a, b)]
            if n > 0:
                _parallel_copy(buf, [_positions_of(fr) for fr in frames_])
            cells = np.empty((b - a, 3, 3))
            for i, fr in enumerate(frames_):
                cells[i] = _cell_of(fr)
        yield buf[:b - a], cells
