"""
Trajectory-level helpers with the reference's names (/root/reference/amof/trajectory.py:230-303).

File readers (``read_cp2k_traj`` ...) are ase.io's job and stay with the reference;
``get_delta_pos`` lives on the GPU inside the MSD analysis (amofb_msd_* in include/amofb.h).
"""
import logging

import numpy as np

from . import atom as amatom
from .frames import ArrayTrajectory  # noqa: F401  (array-backed trajectories are part of the public surface)

logger = logging.getLogger(__name__)


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
