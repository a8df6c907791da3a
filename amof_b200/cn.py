"""
Coordination numbers from a cutoff neighbour search on the GPU, behind the API of ``amof.cn``
(/root/reference/amof/cn.py).

Per frame and per neighbour set ``'A-B'`` the reference builds ase's full neighbour list under the dict cutoffs
(cn.py:65 -> atom.py:72-87), then averages over the atoms of species A the number of their neighbours of species B
(cn.py:67-73).  The integer behind that mean is the number of directed A->B pairs with ``d < cutoff(A, B)``
(strict); libamofb counts exactly those per frame (amofb_pair_* with a cutoff matrix) and the mean is formed here.
"""
import logging

import numpy as np
import pandas as pd

from . import atom as amatom
from . import rdf as _rdf
from .elements import atomic_numbers
from .files import path as _path
from .trajectory import construct_step

logger = logging.getLogger(__name__)


class CoordinationNumber(object):
    """Drop-in for ``amof.cn.CoordinationNumber``."""

    def __init__(self):
        self.data = pd.DataFrame({"Step": np.empty([0])})

    @classmethod
    def from_trajectory(cls, trajectory, nb_set_and_cutoff, delta_Step=1, first_frame=0, parallel=False, distributed=None):
        """
        Args:
            nb_set_and_cutoff: dict, keys are str naming a pair of neighbours ('Zn-N'), values cutoffs in Angstrom
            delta_Step, first_frame: build the ``Step`` column (trajectory.construct_step)
            parallel: accepted for compatibility with the reference (joblib workers); frames are batched on the
                GPU whatever its value
            distributed: None/True/False, see amof_b200._dist
        """
        cn_class = cls()
        step = construct_step(delta_Step=delta_Step, first_frame=first_frame, number_of_frames=len(trajectory))
        cn_class.compute_cn(trajectory, nb_set_and_cutoff, step, parallel, distributed=distributed)
        return cn_class

    def compute_cn(self, trajectory, nb_set_and_cutoff, step, parallel, distributed=None):
        logger.info("Start computing coordination number for %s frames", len(trajectory))
        cutoff_dict = amatom.format_cutoff(nb_set_and_cutoff)
        zs, spec, res = _rdf.pair_histograms(trajectory, 0.0, 0, cn_cutoff=cutoff_dict, distributed=distributed)
        self._assemble(nb_set_and_cutoff, step, zs, spec, res["cn"], len(trajectory))

    def _assemble(self, nb_set_and_cutoff, step, zs, spec, counts, n_frames):
        """counts uint64 [T][S][S] of directed neighbour pairs -> the reference's DataFrame (cn.py:67-82)."""
        n_of = np.bincount(spec, minlength=len(zs))
        idx = {z: k for k, z in enumerate(zs)}
        columns = {"Step": step}
        for nb_set in nb_set_and_cutoff.keys():
            a, b = tuple(atomic_numbers[i] for i in nb_set.split('-'))
            if a in idx and b in idx:
                columns[nb_set] = counts[:, idx[a], idx[b]].astype(np.float64) / float(n_of[idx[a]])
            elif a in idx:
                columns[nb_set] = np.zeros(n_frames)            # A atoms exist, none of their neighbours is B
            else:
                columns[nb_set] = np.full(n_frames, np.nan)     # np.mean([]) in the reference
        self.data = pd.DataFrame(columns)
        self.counts = counts
        self.species = zs

    @classmethod
    def from_file(cls, filename):
        cn_class = cls()
        cn_class.read_cn_file(filename)
        return cn_class

    def read_cn_file(self, filename):
        filename = _path.append_suffix(filename, 'cn')
        self.data = pd.read_feather(filename)

    def write_to_file(self, filename):
        filename = _path.append_suffix(filename, 'cn')
        self.data.to_feather(filename)
