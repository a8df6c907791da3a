"""
Mean-squared displacement on the GPU behind the API of ``amof.msd`` (/root/reference/amof/msd.py).

WindowMsd (msd.py:139-268), per element and in total:
    MSD(m) = mean_k |R_k - R_{k-m}|^2 / N_el,  R = positions rebuilt from per-step displacements wrapped into the
    cell of the earlier frame (trajectory.get_delta_pos, trajectory.py:285-303), after removing the centre of mass
    of every frame (msd.py:235-237) and, optionally, unwrapping first (msd.py:222-230).
The whole trajectory is resident on the device, atom-major, so every window length is evaluated from one read of
HBM (amofb_msd_* in include/amofb.h).  Across GPUs the ATOMS are sharded; the only exchanges are the per-frame
mass-weighted sums (centre of mass) and the final per-element sums.

Reference quirks kept: Q4 the mean runs over T-m slots of which the first is never written, so
MSD(m) carries a factor (T-m-1)/(T-m); Q7 the caller's frames are translated by -COM (and overwritten with the
unwrapped positions when ``unwrap=True``).  Q5 (in-place drift of delta_pos[0]) only changes last-ulp rounding and
is not reproduced.
"""
import logging

import numpy as np
import pandas as pd

from . import _dist, _lib, frames
from .elements import chemical_symbols
from .files import path as _path
from .trajectory import construct_step

logger = logging.getLogger(__name__)


def _atom_range(n_atoms, distributed):
    rank, world = _dist.rank_world(distributed)
    return (n_atoms * rank) // world, (n_atoms * (rank + 1)) // world


def _cells_of(trajectory):
    if isinstance(trajectory, frames.ArrayTrajectory) or hasattr(trajectory, "stream_chunks"):
        return trajectory.cells
    return np.array([np.asarray(a.get_cell(), dtype=np.float64).reshape(3, 3) for a in trajectory])


def _load_local(session, trajectory, lo, hi, backend, target_bytes=64 << 20):
    """Stream the atoms [lo, hi) of every frame to the device."""
    T = len(trajectory)
    n = hi - lo
    step = max(1, min(T, int(target_bytes // (24 * max(n, 1)))))
    if isinstance(trajectory, frames.ArrayTrajectory):
        whole = lo == 0 and hi == trajectory.positions.shape[1]
        for a in range(0, T, step):
            b = min(T, a + step)
            blk = trajectory.block(a, b)
            session.load(a, blk if whole else blk[:, lo:hi])
        return
    ctx = getattr(backend, "ctx", None)
    bufs = [ctx.scratch("msdload%d" % i, (step, n, 3)) if ctx is not None else np.empty((step, n, 3)) for i in range(2)]
    which = 0
    for a in range(0, T, step):
        b = min(T, a + step)
        buf = bufs[which]
        which ^= 1
        for k in range(a, b):
            buf[k - a] = frames._positions_of(trajectory[k])[lo:hi]
        session.load(a, buf[:b - a])       # returns once the copy has been issued and completed


def _streamable(window, T):
    """window lengths 0, D, 2D, ... with 4 D < T: what the streaming path of libamofb takes (amofb_msd_slab_*)"""
    if len(window) < 2 or window[0] != 0 or window[1] <= 0:
        return False
    d = int(window[1])
    return bool(np.all(window == np.arange(len(window)) * d)) and 4 * d < T


def _stream_slabs(session, trajectory, lo, hi, backend, distributed):
    """Feed the atoms [lo, hi) of every frame slab by slab; returns the centre of mass [T][3].

    Two slabs are in flight: the sums of slab i+1 are enqueued before the (all-reduced) sums of slab i are turned into its
    centre of mass, so the device always has work queued."""
    T = len(trajectory)
    n = hi - lo
    step = max(1, min(T, session.slab_frames()))
    if getattr(trajectory, "block_frames", None):       # a trajectory that serves its frames from a bounded buffer says how many at a time
        step = max(1, min(step, int(trajectory.block_frames)))
    com = np.empty((T, 3), dtype=np.float64)
    is_array = isinstance(trajectory, frames.ArrayTrajectory)
    whole = is_array and lo == 0 and hi == trajectory.positions.shape[1]
    ctx = getattr(backend, "ctx", None)
    bufs = None
    if not is_array:    # three staging buffers: a buffer is refilled only after its slab was committed
        bufs = [ctx.scratch("msdslab%d" % i, (step, n, 3)) if ctx is not None else np.empty((step, n, 3)) for i in range(3)]
    begun = []

    def finish(a, b):
        sums = _dist.allreduce_sum(session.slab_sums_wait(b - a), distributed)      # [count][4]: sum m*x, m*y, m*z, m
        com[a:b] = sums[:, 0:3] / sums[:, 3:4]
        session.slab_commit(com[a:b])

    for i, a in enumerate(range(0, T, step)):
        b = min(T, a + step)
        if is_array:
            blk = trajectory.block(a, b)         # a view; an atom shard is a column block the library copies strided
            if not whole:
                blk = blk[:, lo:hi]
        else:
            blk = bufs[i % 3][:b - a]
            for k in range(a, b):
                blk[k - a] = frames._positions_of(trajectory[k])[lo:hi]
        session.slab_sums_begin(a, blk)
        begun.append((a, b))
        if len(begun) == 2:
            finish(*begun.pop(0))
    while begun:
        finish(*begun.pop(0))
    return com


def _open_session(trajectory, distributed, backend):
    backend = backend or _lib.get_backend()
    first = trajectory[0]
    numbers = np.asarray(first.get_atomic_numbers())
    zs, spec = frames.species_index(numbers)
    frames.check_same_atoms(trajectory, numbers, 0, len(trajectory))
    lo, hi = _atom_range(len(numbers), distributed)
    if hi <= lo:
        raise ValueError("more ranks than atoms")
    masses = np.asarray(first.get_masses(), dtype=np.float64)
    cells = _cells_of(trajectory)
    session = backend.msd_open(len(trajectory), masses[lo:hi], spec[lo:hi], len(zs), cells)
    return backend, session, zs, spec, lo, hi


class Msd(object):
    """File round-trips shared by the MSD classes (msd.py:25-51)."""

    def write_to_file(self, path_to_output):
        path_to_output = _path.append_suffix(path_to_output, 'msd')
        self.data.to_feather(path_to_output)

    @classmethod
    def from_msd(cls, *args):
        logger.exception('from_msd is deprecated, use from_file instead')

    @classmethod
    def from_file(cls, path_to_msd):
        msd_class = cls()
        msd_class.read_msd_file(path_to_msd)
        return msd_class

    def read_msd_file(self, path_to_data):
        path_to_data = _path.append_suffix(path_to_data, 'msd')
        self.data = pd.read_feather(path_to_data)


class WindowMsd(Msd):
    """Drop-in for ``amof.msd.WindowMsd``; ``.data`` has ``Time``, one column per element, and ``X``."""

    def __init__(self):
        self.data = pd.DataFrame({"Time": np.empty([0])})

    @classmethod
    def from_trajectory(cls, trajectory, delta_time=100, max_time="half", timestep=1, parallel=False, unwrap=False,
                        distributed=None, mutate=True):
        """
        Args:
            trajectory: sequence of ase.Atoms (or ArrayTrajectory)
            delta_time: int, time between two computed values of the MSD, in fs
            max_time: int or "half" (upper limit; capped at half of the simulation)
            timestep: int, time between two frames, in fs
            parallel: accepted for compatibility
            unwrap: unwrap the trajectory before removing the centre of mass
            distributed: None/True/False, see amof_b200._dist (atoms are sharded over ranks)
            mutate: reproduce the reference's in-place translation of the caller's frames
        """
        msd_class = cls()
        half_time = (len(trajectory) // 2) * timestep
        if max_time == "half" or max_time > half_time:
            max_time = half_time
        if delta_time < timestep:
            logger.exception("Delta_time should be larger than timestep")
        delta_m = delta_time // timestep
        window = np.arange(0, max_time // timestep, delta_m)
        time = timestep * window
        msd_class.compute_msd(trajectory, window, time, parallel, unwrap, distributed=distributed, mutate=mutate)
        return msd_class

    def compute_msd(self, trajectory, window, time, parallel, unwrap, distributed=None, mutate=True, backend=None):
        elements = list(set(trajectory[0].get_atomic_numbers()))          # column order, SURVEY.md Q3
        T = len(trajectory)
        backend, session, zs, spec, lo, hi = _open_session(trajectory, distributed, backend)
        w = np.asarray(window, dtype=np.int64)
        with session:
            new_positions = None
            if not unwrap and _streamable(w, T) and hasattr(session, "slab_sums_begin"):
                # every pass over the positions happens on the way in: mass sums per slab of frames, (all-reduced) centre of
                # mass back, and the shift / wrap / running sum are fused into the transposition
                logger.info("Start computing msd at %s times on a trajectory of %s frames", len(window), T)
                com = _stream_slabs(session, trajectory, lo, hi, backend, distributed)
            else:
                _load_local(session, trajectory, lo, hi, backend)
                if unwrap:
                    logger.info("Unwrap trajectory before computing msd")
                    session.unwrap()
                    if mutate:
                        new_positions = session.get_positions()
                        if _dist.active(distributed):       # every rank mutates whole frames: gather the atom shards
                            _, world = _dist.rank_world(distributed)
                            n_all = len(spec)
                            counts = [(n_all * (r + 1)) // world - (n_all * r) // world for r in range(world)]
                            new_positions = np.ascontiguousarray(
                                _dist.allgather_rows(np.ascontiguousarray(new_positions.transpose(1, 0, 2)), counts, distributed).transpose(1, 0, 2))
                logger.info("Start computing msd at %s times on a trajectory of %s frames", len(window), T)
                sums = _dist.allreduce_sum(session.com_sums(), distributed)      # [T][4]: sum m*x, m*y, m*z, m
                com = sums[:, 0:3] / sums[:, 3:4]
                session.set_com(com)
            raw = _dist.allreduce_sum(session.window(w), distributed) if len(w) else np.zeros((len(zs), 0))
        if mutate:
            self._mutate_frames(trajectory, com, new_positions)
        n_of = np.bincount(spec, minlength=len(zs)).astype(np.float64)
        idx = {z: k for k, z in enumerate(zs)}
        columns = {"Time": time}
        for x in elements:
            # mean over the T-m slots of MSD_partial (slot 0 stays 0, Q4), each slot a sum over atoms / N_el
            columns[chemical_symbols[x]] = raw[idx[int(x)]] / n_of[idx[int(x)]] / (T - w).astype(np.float64)
        formula_dict = trajectory[0].symbols.formula._count
        total = sum(formula_dict.values())
        columns["X"] = np.sum([columns[k] * v for k, v in formula_dict.items()], axis=0) / total
        self.data = pd.DataFrame(columns)
        self.com = com

    @staticmethod
    def _mutate_frames(trajectory, com, new_positions):
        if isinstance(trajectory, frames.ArrayTrajectory):
            if new_positions is not None:
                trajectory.positions[...] = new_positions
            trajectory.positions -= com[:, None, :]
            return
        for k in range(len(trajectory)):
            atom = trajectory[k]
            if new_positions is not None:
                atom.set_positions(new_positions[k])
            atom.translate(-com[k])


class DirectMsd(Msd):
    """Drop-in for the deprecated ``amof.msd.DirectMsd`` (orthogonal cells only, msd.py:54-137)."""

    def __init__(self):
        self.data = pd.DataFrame({"Step": np.empty([0])})
        logger.warning('DirectMsd is deprecated and not suitable for non-orthogonal cells, use WindowMsd instead')

    @classmethod
    def from_trajectory(cls, trajectory, delta_Step=1, first_frame=0, parallel=False, distributed=None):
        msd_class = cls()
        step = construct_step(delta_Step=delta_Step, first_frame=first_frame, number_of_frames=len(trajectory))
        msd_class.compute_msd(trajectory, step, parallel, distributed=distributed)
        return msd_class

    def compute_msd(self, trajectory, step, parallel, distributed=None, backend=None):
        logger.info("Start computing msd for %s frames", len(trajectory))
        elements = list(set(trajectory[0].get_atomic_numbers()))
        backend, session, zs, spec, lo, hi = _open_session(trajectory, distributed, backend)
        with session:
            _load_local(session, trajectory, lo, hi, backend)
            raw = _dist.allreduce_sum(session.direct(), distributed)          # [S][T] sums of |r_t - r_0|^2
        n_of = np.bincount(spec, minlength=len(zs)).astype(np.float64)
        idx = {z: k for k, z in enumerate(zs)}
        columns = {"Step": step, "X": raw.sum(axis=0) / float(len(spec))}
        for x in elements:
            columns[chemical_symbols[x]] = raw[idx[int(x)]] / n_of[idx[int(x)]]
        self.data = pd.DataFrame(columns)
