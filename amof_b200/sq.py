"""
Structure factors from the accumulated partial pair histograms (not in aMOF itself -- the natural
next observable once the GPU has counted the pairs: nothing here touches the device again).

For species a, b with concentrations c_a = N_a / N and number density rho = N / V, the Faber-Ziman partial structure factor is

    S_ab(q) = 1 + 4 pi rho * integral_0^rmax r^2 (g_ab(r) - 1) sin(q r) / (q r) W(r) dr ,

with g_ab normalised to tend to 1 (count_ab / (N_a frames shell rho c_b)), W a window that tames the truncation at rmax
('lorch': sin(pi r / rmax) / (pi r / rmax), or None), and the total (number-weighted, i.e. all scattering lengths equal)

    S(q) = sum_ab c_a c_b S_ab(q) = 1 + 4 pi rho * integral r^2 (g(r) - 1) sin(q r) / (q r) W(r) dr .

The integrals are midpoint sums over the histogram bins (width rmax / bins, the bins the pairs were counted in -- not the
``r`` labels of ``Rdf.data``, which carry the caller's dr: SURVEY.md Q1).
"""
import numpy as np
import pandas as pd

from .elements import chemical_symbols


def _window(r, rmax, window):
    if window is None:
        return np.ones_like(r)
    if window == "lorch":
        x = np.pi * r / rmax
        return np.sinc(x / np.pi)           # np.sinc(t) = sin(pi t) / (pi t)
    raise ValueError("window must be 'lorch' or None")


def _transform(h, r, dr, q, rho, window, rmax):
    """1 + 4 pi rho sum_i r_i^2 h_i sin(q r_i)/(q r_i) W(r_i) dr for every q"""
    w = _window(r, rmax, window)
    kernel = np.sinc(np.outer(q, r) / np.pi)                 # sin(q r) / (q r), 1 at q r = 0
    return 1.0 + 4.0 * np.pi * rho * (kernel * (r * r * h * w)[None, :]).sum(axis=1) * dr


class StructureFactor(object):
    """``.data``: DataFrame with ``q``, the total ``X-X`` and one column ``A-B`` per unordered species pair (A <= B by atomic
    number; S_ab = S_ba)."""

    def __init__(self):
        self.data = pd.DataFrame({"q": np.empty([0])})

    @classmethod
    def from_rdf(cls, rdf, rmax, n_atoms_by_species, volume, q=None, window="lorch"):
        """
        Args:
            rdf: an :class:`amof_b200.rdf.Rdf` computed by ``from_trajectory`` (carries ``.counts``, ``.species``, ``.n_frames``)
            rmax: the rmax the histograms were counted with (after the half-cell rule: ``Rdf._axis``)
            n_atoms_by_species: {Z: number of atoms}
            volume: mean cell volume in Angstrom^3
            q: wave numbers in 1/Angstrom (default: 2 pi / rmax .. 25 in steps of 0.02)
        """
        self = cls()
        counts = np.asarray(rdf.counts, dtype=np.float64)     # [S][S][bins], directed pairs, ascending Z
        zs = list(rdf.species)
        bins = counts.shape[2]
        dr = float(rmax) / bins
        r = (np.arange(bins) + 0.5) * dr
        i = np.arange(bins, dtype=np.float64)
        shell = 4.0 * np.pi / 3.0 * (((i + 1.0) * dr) ** 3 - (i * dr) ** 3)
        n_of = np.array([float(n_atoms_by_species[z]) for z in zs])
        n = n_of.sum()
        rho = n / float(volume)
        if q is None:
            q = np.arange(2.0 * np.pi / rmax, 25.0, 0.02)
        q = np.asarray(q, dtype=np.float64)
        columns = {"q": q}
        g_tot = counts.sum(axis=(0, 1)) / (n * rdf.n_frames * shell * rho)
        columns["X-X"] = _transform(g_tot - 1.0, r, dr, q, rho, window, rmax)
        for a in range(len(zs)):
            for b in range(a, len(zs)):
                g = counts[a, b] / (n_of[a] * rdf.n_frames * shell * rho * (n_of[b] / n))
                columns[chemical_symbols[zs[a]] + "-" + chemical_symbols[zs[b]]] = _transform(g - 1.0, r, dr, q, rho, window, rmax)
        self.data = pd.DataFrame(columns)
        self.concentrations = {z: n_of[k] / n for k, z in enumerate(zs)}
        return self

    @classmethod
    def from_trajectory(cls, trajectory, dr=0.01, rmax='half_cell', q=None, window="lorch", distributed=None):
        """g(r) on the GPU (:class:`amof_b200.rdf.Rdf`), then the transforms above on the host."""
        from . import frames, rdf as _rdf
        r = _rdf.Rdf.from_trajectory(trajectory, dr=dr, rmax=rmax, distributed=distributed)
        rmax_used, _, _ = _rdf.Rdf._axis(trajectory, dr, rmax)
        numbers = np.asarray(trajectory[0].get_atomic_numbers())
        n_by = {int(z): int((numbers == z).sum()) for z in set(numbers.tolist())}
        cells = frames.gather_cells(trajectory)
        volume = float(np.mean(np.abs(np.linalg.det(cells))))
        return cls.from_rdf(r, rmax_used, n_by, volume, q=q, window=window)
