"""
Radial distribution functions on the GPU behind the API of ``amof.rdf`` (/root/reference/amof/rdf.py).

``Rdf.from_trajectory(trajectory, dr=0.01, rmax='half_cell')`` returns an object whose ``.data`` DataFrame has the
reference's columns in the reference's order: ``r``, ``X-X``, every ordered pair ``A-B``, then ``A-X``.

What runs where
  * pair counting (asap3's C++ ``RadialDistributionFunction.update()``, rdf.py:87-93) -> libamofb pair kernel,
    integer histograms of directed pairs per ordered species pair, accumulated over frames;
  * normalisation (asap3's ``get_rdf``, rdf.py:96,109) -> a few numpy lines below, on the host in fp64.

asap3 is not on disk, so its conventions are pinned here explicitly (same pins as the oracle) -- and, because
they cannot be checked here, every one of them is a NAMED SWITCH in :data:`CONVENTIONS` with the pin as default:
  U1  ``bin_rule``      'divide': bin = int(d / (rMax/nBins)) | 'multiply': bin = int(d * (nBins/rMax)); counted iff < nBins.
                        Decides last-ulp bins; evaluated by the library (amofb_set_option AMOFB_OPT_RDF_BIN_RULE).
  U3  ``shell_volume``  'exact': 4*pi/3 * ((i+1)^3 - i^3) * dr^3 | 'midpoint': 4*pi * ((i + 1/2) dr)^2 * dr.
  U4  ``volume``        'mean': mean cell volume over the accumulated frames | 'first': the volume of the first frame.
  a3  ``partial_norm``  'centre_species': g_ab = count_ab / (N_a * frames * shell * N/V): partials are normalised with the
                        TOTAL number density, which is what makes ``A-X = sum_B A-B`` (rdf.py:111-114) an RDF that tends to 1
                        | 'sum_to_global': g_ab = count_ab / (N * frames * shell * N/V), so that sum_ab g_ab is the global RDF.
The first box with asap3 settles them: tests/golden/make_golden.py writes reference outputs when it can import the
packages, and tests/test_golden_reference.py compares against them under the defaults.
"""
import logging

import numpy as np
import pandas as pd

from . import _dist, _lib, frames
from .elements import atomic_numbers as _atomic_numbers
from .elements import chemical_symbols
from .files import path as _path
from .trajectory import construct_step

logger = logging.getLogger(__name__)

#: conventions of asap3's RDF that cannot be read off its source here (module docstring); change with set_conventions
CONVENTIONS = {"bin_rule": "divide", "shell_volume": "exact", "volume": "mean", "partial_norm": "centre_species"}
_CHOICES = {"bin_rule": ("divide", "multiply"), "shell_volume": ("exact", "midpoint"), "volume": ("mean", "first"),
            "partial_norm": ("centre_species", "sum_to_global")}


def set_conventions(**kwargs):
    """Change (and return the previous values of) the named conventions, e.g. ``set_conventions(bin_rule='multiply')``."""
    old = dict(CONVENTIONS)
    for k, v in kwargs.items():
        if k not in _CHOICES or v not in _CHOICES[k]:
            raise ValueError("unknown convention %s=%r (choices: %s)" % (k, v, _CHOICES.get(k)))
        CONVENTIONS[k] = v
    return old


def _half_cell_rmax(trajectory):
    """min over frames and axes of the cell LENGTHS / 2 (rdf.py:74; lengths, not perpendicular heights).
    One vectorised pass over the stacked cells instead of ``get_cell_lengths_and_angles()`` per frame (which also
    evaluates three arccos per frame that the reference then discards)."""
    cells = frames.gather_cells(trajectory)
    return float(np.min(np.sqrt((cells ** 2).sum(axis=2))) / 2)


def normalise_counts(counts, n_centres, n_frames, n_atoms, volume, rmax):
    """counts[bins] of directed pairs -> g(r) (pins U3, a3; ``volume`` is the volume pin U4 selected)."""
    bins = len(counts)
    dr = rmax / bins
    i = np.arange(bins, dtype=np.float64)
    if CONVENTIONS["shell_volume"] == "exact":
        shell = 4.0 * np.pi / 3.0 * (((i + 1.0) * dr) ** 3 - (i * dr) ** 3)
    else:
        shell = 4.0 * np.pi * ((i + 0.5) * dr) ** 2 * dr
    norm = shell * (float(n_centres) * float(n_frames)) * (float(n_atoms) / volume)
    return np.asarray(counts, dtype=np.float64) / norm


def pair_histograms(trajectory, rmax, bins, cn_cutoff=None, distributed=None, backend=None):
    """Run the pair analysis over the frames of this rank and combine ranks.

    Returns (zs, spec, result) with result = dict(hist, cn, n_frames, volume_sum) as in GpuBackend.pair_counts;
    ``hist`` is summed over ranks (integer all-reduce), ``cn`` rows are gathered in frame order."""
    backend = backend or _lib.get_backend()
    numbers = np.asarray(frames._numbers_of(trajectory[0]))
    zs, spec = frames.species_index(numbers)
    T = len(trajectory)
    lo, hi = frames.frame_range(T, distributed)
    frames.check_same_atoms(trajectory, numbers, lo, hi)
    cut = None
    if cn_cutoff is not None:
        from .atom import cutoff_matrix
        cut = cutoff_matrix(cn_cutoff, zs)
    ctx = getattr(backend, "ctx", None)
    rule = 1 if CONVENTIONS["bin_rule"] == "multiply" else 0
    if ctx is not None:
        ctx.set_option(_lib.AMOFB_OPT_RDF_BIN_RULE, rule)
    elif rule:
        raise NotImplementedError("this backend only evaluates the default bin rule")
    try:
        res = backend.pair_counts(spec, len(zs), frames.iter_chunks(trajectory, lo, hi, backend), rmax=rmax, nbins=bins,
                                  cn_cutoff=cut)
    finally:
        if ctx is not None and rule:
            ctx.set_option(_lib.AMOFB_OPT_RDF_BIN_RULE, 0)
    first_cell = np.asarray(frames.gather_cells(trajectory, 0, 1)[0], dtype=np.float64) if T > 0 else np.eye(3)
    res["volume_first"] = float(abs(np.linalg.det(first_cell)))
    if _dist.active(distributed):
        _, world = _dist.rank_world(distributed)
        if res["hist"] is not None:
            res["hist"] = _dist.allreduce_sum(res["hist"], distributed)
        if res["cn"] is not None:
            counts = [(T * (r + 1)) // world - (T * r) // world for r in range(world)]
            res["cn"] = _dist.allgather_rows(res["cn"], counts, distributed)
        # the float volume sums are gathered and added in rank order, so g(r) does not depend on the reduction tree
        rank, _ = _dist.rank_world(distributed)
        per_rank = _dist.allgather_rows(np.array([[res["volume_sum"], float(res["n_frames"])]]), [1] * world, distributed)
        res["volume_sum"] = float(np.sum(per_rank[:, 0]))        # np.sum over <= 8 addends: sequential
        res["n_frames"] = int(round(float(np.sum(per_rank[:, 1]))))
    return zs, spec, res


class Rdf(object):
    """Total and partial g(r) of a trajectory; drop-in for ``amof.rdf.Rdf``."""

    def __init__(self):
        self.data = pd.DataFrame({"r": np.empty([0])})

    @classmethod
    def from_trajectory(cls, trajectory, dr=0.01, rmax='half_cell', distributed=None):
        """
        Args:
            trajectory: sequence of ase.Atoms (or an ArrayTrajectory)
            dr, rmax: floats in Angstrom; ``rmax='half_cell'`` uses half of the smallest cell length over the
                trajectory, and a larger request is clamped to it (rdf.py:74-79)
            distributed: None/True/False, see amof_b200._dist (frames are sharded over ranks)
        """
        rdf_class = cls()
        rdf_class.compute_rdf(trajectory, dr, rmax, distributed=distributed)
        return rdf_class

    @classmethod
    def from_rdf(cls, *args):
        logger.exception('from_rdf is deprecated, use from_file instead')

    @classmethod
    def from_file(cls, path_to_rdf):
        rdf_class = cls()
        rdf_class.read_rdf_file(path_to_rdf)
        return rdf_class

    def compute_rdf(self, trajectory, dr, rmax, distributed=None):
        atomic_numbers_unique = list(set(trajectory[0].get_atomic_numbers()))   # column order, SURVEY.md Q3

        rmax, bins, r = self._axis(trajectory, dr, rmax)      # half-cell clamp; int(rmax // dr) bins (Q1); r = arange(bins)*dr
        logger.info("Start computing rdf for %s frames with dr = %s and rmax = %s", len(trajectory), dr, rmax)
        zs, spec, res = pair_histograms(trajectory, float(rmax), bins, distributed=distributed)
        self._assemble(atomic_numbers_unique, zs, spec, res, r, rmax)

    @staticmethod
    def _axis(trajectory, dr, rmax):
        """(rmax after the half-cell rule, bins, r) exactly as compute_rdf derives them (rdf.py:74-83)."""
        rmax_half_cell = _half_cell_rmax(trajectory)
        if isinstance(rmax, str):
            if rmax != 'half_cell':
                raise ValueError("rmax must be a float or 'half_cell'")
            rmax = rmax_half_cell
        elif rmax > rmax_half_cell:
            logger.info("Specified rmax %s is larger than half cell; will use half_cell rmax", rmax)
            rmax = rmax_half_cell
        bins = int(rmax // dr)
        if bins < 1:
            raise ValueError("rmax // dr gives no bin (rmax = %s, dr = %s)" % (rmax, dr))
        return rmax, bins, np.arange(bins) * dr

    def _assemble(self, atomic_numbers_unique, zs, spec, res, r, rmax):
        """counts -> the reference's DataFrame (rdf.py:95-114)."""
        hist, n_frames = res["hist"], res["n_frames"]
        n_atoms = len(spec)
        volume_mean = res["volume_sum"] / n_frames if CONVENTIONS["volume"] == "mean" else res["volume_first"]
        n_of = np.bincount(spec, minlength=len(zs))
        idx = {z: k for k, z in enumerate(zs)}

        columns = {"r": r}
        columns["X-X"] = normalise_counts(hist.sum(axis=(0, 1)), n_atoms, n_frames, n_atoms, volume_mean, rmax)
        partial = {}
        for x in atomic_numbers_unique:
            for y in atomic_numbers_unique:
                n_centres = n_of[idx[int(x)]] if CONVENTIONS["partial_norm"] == "centre_species" else n_atoms
                g = normalise_counts(hist[idx[int(x)], idx[int(y)]], n_centres, n_frames, n_atoms, volume_mean, rmax)
                partial[(x, y)] = g
                columns[chemical_symbols[x] + "-" + chemical_symbols[y]] = g
        for x in atomic_numbers_unique:
            columns[chemical_symbols[x] + "-X"] = sum([partial[(x, y)] for y in atomic_numbers_unique])
        self.data = pd.DataFrame(columns)
        self.counts = hist                           # raw directed-pair histogram [S][S][bins], sorted-Z order
        self.species = zs
        self.n_frames = n_frames

    def write_to_file(self, filename):
        filename = _path.append_suffix(filename, 'rdf')
        self.data.to_feather(filename)

    def read_rdf_file(self, path_to_data):
        path_to_data = _path.append_suffix(path_to_data, 'rdf')
        self.data = pd.read_feather(path_to_data)

    def get_coordination_number(self, nn_set, cutoff, density):
        """coordination number of the pair ``nn_set`` (e.g. 'Zn-N') by integrating g(r) up to ``cutoff``"""
        return get_coordination_number(self.data['r'], self.data[nn_set], cutoff, density)


def rdf_and_cn(trajectory, nb_set_and_cutoff, dr=0.01, rmax='half_cell', delta_Step=1, first_frame=0, distributed=None):
    """One pass over the trajectory for both ``Rdf.from_trajectory(trajectory, dr, rmax)`` and
    ``amof_b200.cn.CoordinationNumber.from_trajectory(trajectory, nb_set_and_cutoff, delta_Step, first_frame)``:
    the pair kernel fills the RDF histograms and the per-frame cutoff counts from the same distances, so the frames
    cross PCIe once.  Returns (Rdf, CoordinationNumber) identical to the two separate calls."""
    from . import atom as amatom
    from . import cn as _cn
    rdf_class = Rdf()
    atomic_numbers_unique = list(set(trajectory[0].get_atomic_numbers()))
    rmax, bins, r = Rdf._axis(trajectory, dr, rmax)
    cutoff_dict = amatom.format_cutoff(nb_set_and_cutoff)
    zs, spec, res = pair_histograms(trajectory, float(rmax), bins, cn_cutoff=cutoff_dict, distributed=distributed)
    rdf_class._assemble(atomic_numbers_unique, zs, spec, res, r, rmax)
    cn_class = _cn.CoordinationNumber()
    step = construct_step(delta_Step=delta_Step, first_frame=first_frame, number_of_frames=len(trajectory))
    cn_class._assemble(nb_set_and_cutoff, step, zs, spec, res["cn"], len(trajectory))
    return rdf_class, cn_class


def _simpson_avg(y, x):
    """Composite Simpson on samples, ``even='avg'`` for an even number of points -- the rule
    ``scipy.integrate.simps`` applied in scipy 1.7.1 (the reference's pin; ``simps`` no longer exists)."""
    y = np.asarray(y, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    n = len(y)
    if n != len(x):
        raise ValueError("x and y must have the same length")
    if n < 2:
        return 0.0

    def basic(lo, hi):          # Simpson over points lo..hi inclusive, (hi - lo) even, unequal spacing allowed
        if hi - lo < 2:
            return 0.0
        i0 = np.arange(lo, hi - 1, 2)
        h0 = x[i0 + 1] - x[i0]
        h1 = x[i0 + 2] - x[i0 + 1]
        hsum, hprod, hdiv = h0 + h1, h0 * h1, h0 / h1
        t = hsum / 6.0 * (y[i0] * (2.0 - 1.0 / hdiv) + y[i0 + 1] * (hsum * hsum / hprod) + y[i0 + 2] * (2.0 - hdiv))
        return float(np.sum(t))

    if n % 2 == 1:
        return basic(0, n - 1)
    first = basic(0, n - 2) + 0.5 * (x[n - 1] - x[n - 2]) * (y[n - 1] + y[n - 2])      # trapezoid on the last interval
    last = basic(1, n - 1) + 0.5 * (x[1] - x[0]) * (y[1] + y[0])                      # trapezoid on the first interval
    return 0.5 * (first + last)


def get_coordination_number(r, rdf, cutoff, density):
    """4 pi rho * integral_0^cutoff g(r) r^2 dr over the samples with 0 < r < cutoff (rdf.py:216-227)."""
    r = np.asarray(r, dtype=np.float64)
    rdf = np.asarray(rdf, dtype=np.float64)
    mask = (r > 0) & (r < cutoff)
    r = r[mask]
    rdf = rdf[mask]
    integral = _simpson_avg(rdf * (r ** 2), r)
    return 4 * np.pi * density * integral


class CoordinationNumber(object):
    """Coordination numbers by integrating per-frame partial RDFs (``amof.rdf.CoordinationNumber``,
    rdf.py:135-214).  Subject to integration error; prefer :class:`amof_b200.cn.CoordinationNumber`."""

    def __init__(self):
        logger.warning('Compute CoordinationNumber from RDF, best to use amof.cn.CoordinationNumber')
        self.data = pd.DataFrame({"Step": np.empty([0])})

    @classmethod
    def from_trajectory(cls, trajectory, nb_set_and_cutoff, delta_Step=1, first_frame=0, dr=0.0001, parallel=False):
        cn_class = cls()
        step = construct_step(delta_Step=delta_Step, first_frame=first_frame, number_of_frames=len(trajectory))
        cn_class.compute_cn(trajectory, nb_set_and_cutoff, step, dr, parallel)
        return cn_class

    def compute_cn(self, trajectory, nb_set_and_cutoff, step, dr, parallel):
        """``parallel`` is accepted for signature compatibility; frames are batched on the GPU instead."""
        rmax = np.max(list(nb_set_and_cutoff.values()))
        logger.info("Start computing coordination number for %s frames with dr = %s and rmax = %s", len(trajectory), dr, rmax)
        bins = int(rmax // dr)
        r = np.arange(bins) * dr
        rows = []
        T = len(trajectory)
        if T == 0:
            self.data = pd.DataFrame(rows)
            return
        backend = _lib.get_backend()
        numbers = np.asarray(frames._numbers_of(trajectory[0]))
        zs, spec = frames.species_index(numbers)
        frames.check_same_atoms(trajectory, numbers, 0, T)
        n_of = np.bincount(spec, minlength=len(zs))
        idx = {z: k for k, z in enumerate(zs)}
        ctx = getattr(backend, "ctx", None)
        if ctx is not None:
            ctx.set_option(_lib.AMOFB_OPT_RDF_BIN_RULE, 1 if CONVENTIONS["bin_rule"] == "multiply" else 0)

        def single_frames():
            for i in range(T):
                atom = trajectory[i]
                yield (np.asarray(atom.get_positions(), dtype=np.float64)[None], np.asarray(atom.get_cell(), dtype=np.float64)[None])

        # a fresh histogram per frame, like the reference's per-frame RadialDistributionFunction (rdf.py:181); ONE analysis stays
        # open on the GPU and is emptied after every frame
        each = backend.pair_counts_each(spec, len(zs), single_frames(), float(rmax), bins)
        try:
            for i, res in enumerate(each):
                atom = trajectory[i]
                density = len(atom) / atom.get_volume()
                dic = {'Step': step[i]}
                for nn_set, cutoff in nb_set_and_cutoff.items():
                    a, b = tuple(_atomic_numbers[s] for s in nn_set.split('-'))
                    g = normalise_counts(res["hist"][idx[a], idx[b]], n_of[idx[a]], 1, len(spec), res["volume_sum"], rmax)
                    dic[nn_set] = get_coordination_number(r, g, cutoff, density)
                rows.append(dic)
        finally:
            each.close()            # ends the open analysis if the loop was left early
            if ctx is not None:
                ctx.set_option(_lib.AMOFB_OPT_RDF_BIN_RULE, 0)
        self.data = pd.DataFrame(rows)

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
