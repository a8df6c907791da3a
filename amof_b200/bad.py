"""
Bond-angle distributions on the GPU behind the API of ``amof.bad`` (/root/reference/amof/bad.py).

For every requested centre species A and ligand species B the reference gathers, frame by frame, the angle
B-A-B of every unordered pair of B-neighbours of every A atom (bad.py:70-114; neighbours from ase's dict-cutoff
neighbour list, angles from ``Atoms.get_angles(mic=True)``), concatenates the angles of the whole trajectory in
Python lists and histograms them at the end (bad.py:154-160).  libamofb accumulates the integer histogram directly
(amofb_bad_*), split by the coordination number of the centre, so memory does not grow with the trajectory;
``np.histogram(..., density=True)`` is then reproduced on the counts.

Deviation: when the neighbour-set keys cover every species of the frame, the reference appends the
pseudo-species "X" and then crashes on ``ase.data.chemical_symbols["X"]`` (bad.py:111,127).  Here the X columns
("X-A-X": any neighbours around A; "X-X-X": any neighbours around any atom) are computed as the docstring of
``bad_BAB`` describes instead of raising.

Deviation (minimum image): ``get_angles(mic=True)`` re-derives the two bond vectors from the atom INDICES with the minimum
image convention, so a neighbour found through a farther periodic image is measured along its nearest image instead
(bad.py:113, ase/atoms.py get_angles).  libamofb measures the image the neighbour list found.  The two agree whenever
every cutoff is below half the smallest perpendicular cell height -- then the nearest image is the only one under the
cutoff --, and ``amofb_bad_begin`` refuses any other request (AMOFB_ERR_GEOMETRY -> ValueError) rather than differ silently.
"""
import logging

import numpy as np
import pandas as pd

from . import _dist, _lib, frames
from . import atom as amatom
from .elements import atomic_numbers, chemical_symbols
from .files import path as _path

logger = logging.getLogger(__name__)

MAX_CN = _lib.AMOFB_BAD_MAX_CN


def _symbol(c):
    return "X" if isinstance(c, str) else chemical_symbols[c]


def _elements(trajectory, nb_set_and_cutoff):
    """(A, B) pairs exactly as bad.py:119-133 enumerates them."""
    atomic_numbers_unique = list(set(trajectory[0].get_atomic_numbers()))
    elements_present_unique = list(set([atomic_numbers[i] for nb_set in nb_set_and_cutoff.keys() for i in nb_set.split('-')]))
    if len(elements_present_unique) == len(atomic_numbers_unique):
        elements_present_unique.append("X")
    return [(a, b) for b in elements_present_unique for a in elements_present_unique
            if (a not in [b, "X"] or ((a, b) == ("X", "X")))]


def _theta_axis(dtheta):
    bins = int(180 // dtheta)                       # 180 // 0.05 == 3599
    theta_bins = np.arange(bins + 2) * dtheta       # bins + 2 edges -> bins + 1 histogram bins
    theta = np.arange(bins + 1) * dtheta + dtheta / 2
    return theta_bins, theta


def _density(counts, theta_bins):
    """np.histogram(..., density=True) applied to integer counts (numpy: n / db / n.sum())."""
    n = np.asarray(counts).astype(np.int64)
    db = np.array(np.diff(theta_bins), float)
    return n / db / n.sum()


def angle_histograms(trajectory, nb_set_and_cutoff, dtheta, distributed=None, backend=None):
    """-> (elements, names, theta_bins, theta, hist uint64[n_triples][MAX_CN+1][nbins]) summed over frames and ranks."""
    backend = backend or _lib.get_backend()
    elements = _elements(trajectory, nb_set_and_cutoff)
    names = ["-".join([_symbol(c) for c in [b, a, b]]) for a, b in elements]
    theta_bins, theta = _theta_axis(dtheta)
    nbins = len(theta_bins) - 1
    if not elements:
        return elements, names, theta_bins, theta, np.zeros((0, MAX_CN + 1, nbins), dtype=np.uint64)
    numbers = np.asarray(trajectory[0].get_atomic_numbers())
    zs, spec = frames.species_index(numbers)
    idx = {z: k for k, z in enumerate(zs)}
    cut = amatom.cutoff_matrix(amatom.format_cutoff(nb_set_and_cutoff), zs)

    def sidx(c):
        if isinstance(c, str):
            return -1
        return idx.get(c, None)

    triples, live = [], []
    for t, (a, b) in enumerate(elements):
        ia, ib = sidx(a), sidx(b)
        if ia is None or ib is None:        # a species named in the cutoffs that the frames do not contain
            continue
        triples.append((ia, ib))
        live.append(t)
    hist = np.zeros((len(elements), MAX_CN + 1, nbins), dtype=np.uint64)
    if triples:
        T = len(trajectory)
        lo, hi = frames.frame_range(T, distributed)
        frames.check_same_atoms(trajectory, numbers, lo, hi)
        h, _dropped, _nf = backend.bad_counts(spec, len(zs), frames.iter_chunks(trajectory, lo, hi, backend), cut,
                                              triples, float(dtheta), nbins)
        h = _dist.allreduce_sum(h, distributed)
        if int(np.sum(_dropped)):
            # only NaN angles get here (a neighbour on top of its centre): np.histogram drops them silently too (bad.py:160)
            logger.warning("%d angles of this rank's frames were undefined (zero-length bond) and are not counted", int(np.sum(_dropped)))
        hist[live] = h
    return elements, names, theta_bins, theta, hist


class CoreBad(object):
    """Constructors shared by Bad and BadByCn (bad.py:33-59)."""

    @classmethod
    def from_trajectory(cls, trajectory, nb_set_and_cutoff, dtheta=0.05, normalization='total', parallel=False,
                        distributed=None):
        """
        Args:
            nb_set_and_cutoff: dict, keys are str naming a pair of neighbours, values cutoffs in Angstrom
            dtheta: float, degrees
            normalization: 'total' or 'partial' (only BadByCn uses it)
            parallel: accepted for compatibility; the GPU batches frames itself
            distributed: None/True/False, see amof_b200._dist
        """
        bad_class = cls()
        bad_class.compute_bad(trajectory, nb_set_and_cutoff, dtheta, normalization, parallel, distributed=distributed)
        return bad_class

    @classmethod
    def from_file(cls, filename):
        bad_class = cls()
        bad_class.read_bad_file(filename)
        return bad_class


class Bad(CoreBad):
    """Drop-in for ``amof.bad.Bad``: ``.data`` has ``theta`` plus one density column per B-A-B triple that occurs."""

    def __init__(self):
        self.data = pd.DataFrame({"theta": np.empty([0])})

    def compute_bad(self, trajectory, nb_set_and_cutoff, dtheta, normalization, parallel, distributed=None):
        logger.info("Start computing bad for %s frames with dtheta = %s", len(trajectory), dtheta)
        elements, names, theta_bins, theta, hist = angle_histograms(trajectory, nb_set_and_cutoff, dtheta, distributed)
        columns = {"theta": theta}
        self.counts = {}
        for t, name in enumerate(names):
            total = hist[t].sum(axis=0)
            if total.sum() > 0:                     # the reference adds a column only when angles exist (bad.py:159)
                columns[name] = _density(total, theta_bins)
                self.counts[name] = total
        self.data = pd.DataFrame(columns)

    def write_to_file(self, filename):
        filename = _path.append_suffix(filename, 'bad')
        self.data.to_feather(filename)

    def read_bad_file(self, path_to_data):
        path_to_data = _path.append_suffix(path_to_data, 'bad')
        self.data = pd.read_feather(path_to_data)


class BadByCn(CoreBad):
    """Drop-in for ``amof.bad.BadByCn``: the distribution of each triple split by the number of B-neighbours of
    the centre.  ``.data`` is an ``xarray.Dataset`` with variable ``bad`` over (atom_triple, cn, theta) when xarray
    is importable; ``.by_cn`` always holds the same numbers as ``{triple: {cn: density}}``."""

    def __init__(self):
        self.by_cn = {}
        self.theta = np.empty([0])
        self.data = None

    def compute_bad(self, trajectory, nb_set_and_cutoff, dtheta, normalisation, parallel, distributed=None):
        logger.info("Start computing bad for %s frames with dtheta = %s", len(trajectory), dtheta)
        elements, names, theta_bins, theta, hist = angle_histograms(trajectory, nb_set_and_cutoff, dtheta, distributed)
        self.theta = theta
        self.by_cn = {}
        self.counts = {}
        for t, name in enumerate(names):
            per_cn = {cn: hist[t, cn] for cn in range(2, MAX_CN + 1) if hist[t, cn].sum() > 0}
            if not per_cn:
                continue
            n_all = sum(int(h.sum()) for h in per_cn.values())
            self.by_cn[name] = {}
            self.counts[name] = per_cn
            for cn, h in per_cn.items():
                ratio = int(h.sum()) / n_all if normalisation == 'partial' else 1
                self.by_cn[name][cn] = ratio * _density(h, theta_bins)
        self.data = self._to_xarray()

    def _to_xarray(self):
        try:
            import xarray as xr
        except ImportError:
            logger.warning("xarray is not installed: BadByCn.data is None, use BadByCn.by_cn")
            return None
        dic = {name: xr.DataArray([d[cn] for cn in d], coords={"cn": list(d), "theta": self.theta}, dims=("cn", "theta"))
               for name, d in self.by_cn.items()}
        xa = xr.Dataset(dic).to_array("atom_triple")
        return xr.Dataset({'bad': xa})

    def write_to_file(self, filename):
        """netCDF like the reference's ``xr.Dataset.to_netcdf`` (bad.py:303-305): variable ``bad`` over (atom_triple, cn, theta)
        with its three coordinates.  Without xarray the same classic-format file is written through scipy.io.netcdf_file (what
        xarray's own scipy engine writes: fixed-width character array for the string coordinate, NaN where a triple has no
        centre of that coordination number)."""
        filename = _path.append_suffix(filename, 'bad')
        if self.data is not None:
            self.data.to_netcdf(filename)
            return
        from scipy.io import netcdf_file
        names = list(self.by_cn)
        cns = sorted({cn for d in self.by_cn.values() for cn in d})
        width = max([len(n) for n in names] + [1])
        with netcdf_file(filename, 'w') as nc:
            nc.createDimension('atom_triple', len(names))
            nc.createDimension('cn', len(cns))
            nc.createDimension('theta', len(self.theta))
            nc.createDimension('string%d' % width, width)
            v = nc.createVariable('atom_triple', 'c', ('atom_triple', 'string%d' % width))
            for i, n in enumerate(names):
                v[i] = np.array(list(n.ljust(width, '\0')), dtype='S1')
            v = nc.createVariable('cn', 'i', ('cn',))
            v[:] = np.array(cns, dtype=np.int32)
            v = nc.createVariable('theta', 'd', ('theta',))
            v[:] = self.theta
            v = nc.createVariable('bad', 'd', ('atom_triple', 'cn', 'theta'))
            v._FillValue = np.nan
            block = np.full((len(names), len(cns), len(self.theta)), np.nan)
            for i, n in enumerate(names):
                for cn, dens in self.by_cn[n].items():
                    block[i, cns.index(cn)] = dens
            v[:] = block

    def read_bad_file(self, filename):
        filename = _path.append_suffix(filename, 'bad')
        try:
            import xarray as xr
        except ImportError:
            xr = None
        if xr is not None:
            self.data = xr.open_dataset(filename)
            names = [str(n) for n in self.data['atom_triple'].values]
            cns = [int(c) for c in self.data['cn'].values]
            self.theta = np.asarray(self.data['theta'].values, dtype=np.float64)
            block = np.asarray(self.data['bad'].values, dtype=np.float64)
        else:
            from scipy.io import netcdf_file
            with netcdf_file(filename, 'r', mmap=False) as nc:
                raw = nc.variables['atom_triple'][:]
                names = [b"".join(row).decode().rstrip("\0").rstrip() for row in raw]
                cns = [int(c) for c in nc.variables['cn'][:]]
                self.theta = np.array(nc.variables['theta'][:], dtype=np.float64)
                block = np.array(nc.variables['bad'][:], dtype=np.float64)
        self.by_cn = {n: {cn: block[i, j] for j, cn in enumerate(cns) if not np.all(np.isnan(block[i, j]))} for i, n in enumerate(names)}
