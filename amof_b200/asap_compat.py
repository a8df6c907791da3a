"""
The object protocol of ``asap3.analysis.rdf.RadialDistributionFunction`` on top of libamofb -- the narrowest drop-in.

``amof.rdf`` drives ASAP through exactly this surface (/root/reference/amof/rdf.py:87-96,109 and :181-185)::

    RDFobj = asap3.analysis.rdf.RadialDistributionFunction(atoms, rmax, bins)     # first frame
    RDFobj.atoms = atoms ; RDFobj.update()                                        # every frame
    RDFobj.get_rdf(groups=0)                                                      # total g(r)
    RDFobj.get_rdf(elements=(Za, Zb), groups=0)                                   # partial g(r)

so replacing the import (``import amof_b200.asap_compat as asap3_rdf`` / ``asap3.analysis.rdf = amof_b200.asap_compat``)
makes the UNMODIFIED ``amof.rdf`` count on the GPU.  Frames handed to :meth:`update` are copied and counted in batches;
the integer histograms accumulate over updates like ASAP's.  Normalisation follows the named conventions of
:mod:`amof_b200.rdf` (what asap3 really does cannot be read here: SURVEY.md 8(c) U1-U4).

Only what aMOF uses is implemented: ``groups`` other than None/0, ``interval``/``average`` other than 1, ``autoclear`` and
``output_file`` raise NotImplementedError.
"""
import numpy as np

from . import _lib, frames
from . import rdf as _rdf

_FLUSH_FRAMES = 64


class RadialDistributionFunction(object):
    def __init__(self, atoms, rMax, nBins, groups=None, interval=1, average=1, autoclear=False, verbose=False):
        if groups is not None:
            raise NotImplementedError("amof_b200.asap_compat: atom groups are not supported (aMOF never passes them)")
        if interval != 1 or average != 1 or autoclear:
            raise NotImplementedError("amof_b200.asap_compat: interval/average/autoclear are not supported (aMOF never passes them)")
        self.atoms = atoms
        self.rMax = float(rMax)
        self.nBins = int(nBins)
        if not (self.rMax > 0.0) or self.nBins < 1:
            raise ValueError("rMax must be > 0 and nBins >= 1")
        self.dr = self.rMax / self.nBins
        self.verbose = verbose
        self.countRDF = 0                 # frames accumulated (asap3's name)
        self._numbers = None
        self._zs = self._spec = None
        self._hist = None                 # uint64 [S][S][nBins], sorted-Z order
        self._volume_sum = 0.0
        self._volume_first = None
        self._pending_pos, self._pending_cell = [], []
        self._backend = None

    # -- accumulation ---------------------------------------------------------------------------
    def update(self, atoms=None):
        """Count the pairs of ``atoms`` (default: ``self.atoms``) into the accumulated histograms."""
        if atoms is not None:
            self.atoms = atoms
        a = self.atoms
        numbers = np.asarray(a.get_atomic_numbers())
        if self._numbers is None:
            self._numbers = numbers.copy()
            self._zs, self._spec = frames.species_index(numbers)
            self._hist = np.zeros((len(self._zs), len(self._zs), self.nBins), dtype=np.uint64)
        elif len(numbers) != len(self._numbers) or not np.array_equal(numbers, self._numbers):
            raise ValueError("amof_b200.asap_compat: the atoms changed (count or order) between updates")
        cell = np.asarray(a.get_cell(), dtype=np.float64).reshape(3, 3)
        self._pending_pos.append(np.array(a.get_positions(), dtype=np.float64))      # a copy: the caller may move the atoms on
        self._pending_cell.append(cell.copy())
        if self._volume_first is None:
            self._volume_first = float(abs(np.linalg.det(cell)))
        if len(self._pending_pos) >= _FLUSH_FRAMES:
            self._flush()

    def _flush(self):
        if not self._pending_pos:
            return
        if self._backend is None:
            self._backend = _lib.get_backend()
        ctx = getattr(self._backend, "ctx", None)
        rule = 1 if _rdf.CONVENTIONS["bin_rule"] == "multiply" else 0
        if ctx is not None:
            ctx.set_option(_lib.AMOFB_OPT_RDF_BIN_RULE, rule)
        try:
            res = self._backend.pair_counts(self._spec, len(self._zs), [(np.stack(self._pending_pos), np.stack(self._pending_cell))],
                                            rmax=self.rMax, nbins=self.nBins)
        finally:
            if ctx is not None and rule:
                ctx.set_option(_lib.AMOFB_OPT_RDF_BIN_RULE, 0)
        self._hist += res["hist"]
        self._volume_sum += res["volume_sum"]
        self.countRDF += res["n_frames"]
        self._pending_pos, self._pending_cell = [], []

    def clear(self):
        self._pending_pos, self._pending_cell = [], []
        if self._hist is not None:
            self._hist[...] = 0
        self._volume_sum, self._volume_first, self.countRDF = 0.0, None, 0

    # -- results ----------------------------------------------------------------------------------
    def get_rdf(self, groups=None, elements=None):
        """g(r) over the accumulated frames: the global RDF, or with ``elements=(Za, Zb)`` the partial one of Zb around Za.
        On an object that never saw :meth:`update`, the frame it was built with is counted first (what aMOF's
        rdf.CoordinationNumber relies on, rdf.py:181-185)."""
        if groups not in (None, 0):
            raise NotImplementedError("amof_b200.asap_compat: atom groups are not supported")
        if self._hist is None and not self._pending_pos:
            self.update()
        self._flush()
        n_atoms = len(self._spec)
        volume = self._volume_sum / self.countRDF if _rdf.CONVENTIONS["volume"] == "mean" else self._volume_first
        if elements is None:
            return _rdf.normalise_counts(self._hist.sum(axis=(0, 1)), n_atoms, self.countRDF, n_atoms, volume, self.rMax)
        za, zb = int(elements[0]), int(elements[1])
        if za not in self._zs or zb not in self._zs:
            return np.zeros(self.nBins)
        ia, ib = self._zs.index(za), self._zs.index(zb)
        n_a = int(np.count_nonzero(self._spec == ia))
        n_centres = n_a if _rdf.CONVENTIONS["partial_norm"] == "centre_species" else n_atoms
        return _rdf.normalise_counts(self._hist[ia, ib], n_centres, self.countRDF, n_atoms, volume, self.rMax)

    def get_counts(self):
        """raw directed-pair histograms uint64 [S][S][nBins] (species in ascending Z) and the species list"""
        self._flush()
        return self._hist.copy(), list(self._zs)

    def output_file(self, prefix):
        raise NotImplementedError("amof_b200.asap_compat: output_file is not supported")
