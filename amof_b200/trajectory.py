"""
Trajectory-level helpers with the reference's names (/root/reference/amof/trajectory.py:230-303).

``read_lammps_traj`` / ``read_cp2k_traj`` keep the reference's names and arguments (trajectory.py:193-228) but return a
lazily read :class:`amof_b200.stream.XyzStream` instead of a list of Atoms: still a valid aMOF trajectory, and the analyses
parse it in chunks while the GPU works.  ``get_delta_pos`` lives on the GPU inside the MSD
analysis (amofb_msd_* in include/amofb.h).
"""
import logging

import numpy as np

from . import atom as amatom
from .frames import ArrayTrajectory  # noqa: F401  (array-backed trajectories are part of the public surface)

logger = logging.getLogger(__name__)


def read_extxyz_trajectory(path):
    """Read a multi-frame extended-XYZ file (``Lattice="..."`` headers, constant atom count and order) straight into an
    :class:`ArrayTrajectory`, i.e. into the contiguous ``positions[T][N][3]`` / ``cells[T][3][3]`` blocks the GPU path
    streams, without building one Atoms object per frame.  The per-atom lines go through pandas' C parser; only the
    T header lines are handled in Python.  (The reference ingests trajectories with ase.io.read into a list of Atoms,
    amof/trajectory.py:37-60; that list works here too, this is the fast path for large files -- SURVEY.md 8(f) rank 1.)
    """
    import pandas as pd

    from .atoms import _LATTICE, _PROPS
    from .elements import atomic_numbers
    with open(path) as fh:
        n = int(fh.readline().split()[0])
        header = fh.readline()
    pos_col = 1
    pm = _PROPS.search(header)
    if pm is not None:
        fields = pm.group(1).split(':')
        col = 0
        for name, _kind, width in zip(fields[0::3], fields[1::3], fields[2::3]):
            if name == 'pos':
                pos_col = col
            col += int(width)
    period = n + 2
    cells = []
    with open(path) as fh:
        for i, line in enumerate(fh):
            if i % period == 1:
                m = _LATTICE.search(line)
                if m is None:
                    raise ValueError("frame %d: extended-XYZ header without Lattice=" % (i // period))
                cells.append([float(x) for x in m.group(1).split()])
            elif i % period == 0 and line.strip() and int(line.split()[0]) != n:
                raise ValueError("frame %d has %s atoms, expected %d" % (i // period, line.split()[0], n))
    T = len(cells)
    body = pd.read_csv(path, sep=r"\s+", header=None, usecols=[0, pos_col, pos_col + 1, pos_col + 2], engine="c",
                       skiprows=lambda i: i % period < 2, names=["s", "x", "y", "z"], skip_blank_lines=False,
                       nrows=T * n, float_precision="round_trip")      # bit-exact decimal -> binary64, like float()
    if len(body) != T * n:
        raise ValueError("truncated extended-XYZ file: %d atom lines for %d frames of %d atoms" % (len(body), T, n))
    symbols = body["s"].to_numpy()[:n]
    if not (body["s"].to_numpy().reshape(T, n) == symbols[None, :]).all():
        raise ValueError("atom order changes between frames")
    numbers = np.array([atomic_numbers[str(s)] for s in symbols], dtype=np.int64)
    positions = np.ascontiguousarray(body[["x", "y", "z"]].to_numpy(dtype=np.float64).reshape(T, n, 3))
    return ArrayTrajectory(numbers, positions, np.array(cells, dtype=np.float64).reshape(T, 3, 3))


def read_lammps_traj(path_to_xyz, index=None, cell=None, unzip_xyz=False):
    """xyz trajectory (optionally gzipped) with an optional cell -- one 3x3 cell or one per frame (trajectory.py:193-206).
    Without ``cell`` the file must carry extended-XYZ ``Lattice=`` headers."""
    from .stream import XyzStream
    return XyzStream(path_to_xyz, cell=cell, index=index, unzip=unzip_xyz)


def read_cp2k_traj(path_to_xyz, path_to_cell, index=None, unzip_xyz=False):
    """CP2K xyz trajectory + ``.cell`` file (trajectory.py:208-228); ``index`` is a slice."""
    from .stream import XyzStream, read_cp2k_cell
    return XyzStream(path_to_xyz, cell=read_cp2k_cell(path_to_cell, index), index=index, unzip=unzip_xyz)


def apply_to_traj(trajectory, function, how):
    if how == 'mean':
        return np.mean([function(atom) for atom in trajectory])


def get_density(trajectory, how='mean'):
    return apply_to_traj(trajectory, amatom.get_density, how)


def get_number_density(trajectory, how='mean'):
    return apply_to_traj(trajectory, amatom.get_number_density, how)


def construct_step(**kwargs):
    """Simulation-step axis from the same keyword combinations as the reference (trajectory.py:244-283)."""
    delta_Step = kwargs.get('delta_Step', None)
    first_frame = kwargs.get('first_frame', None)
    last_frame = kwargs.get('last_frame', None)
    number_of_frames = kwargs.get('number_of_frames', None)
    step = kwargs.get('step', None)
    try:
        if step is not None:
            if isinstance(step, slice):
                return np.array(list(range(step.start or 0, step.stop, step.step or 1)))
            return np.array(step)
        if delta_Step is not None:
            if first_frame is not None and last_frame is not None:
                return np.arange(first_frame, last_frame, delta_Step)
            if number_of_frames is not None:
                if first_frame is None and last_frame is not None:
                    first_frame = last_frame - number_of_frames * delta_Step
                if first_frame is not None:
                    return np.arange(first_frame, first_frame + number_of_frames * delta_Step, delta_Step)
        elif number_of_frames is not None:
            if first_frame is not None and last_frame is not None:
                return np.linspace(first_frame, last_frame, number_of_frames)
    except Exception:
        logger.exception("Cannot construct step from provided args")
        raise ValueError
    return None
