"""
Minimal duck-typed stand-in for ``ase.Atoms``.

aMOF's "trajectory" is any indexable, re-iterable sequence of ``ase.Atoms``
(/root/reference/amof/rdf.py:71-88, amof/cn.py:77, amof/bad.py:149, amof/msd.py:218-237).
The analysis classes in this package only touch the small surface listed in SURVEY.md 8(b); real
``ase.Atoms`` objects are accepted unchanged.  This shim exists because ASE is not installed in
the build container; it also carries the extended-XYZ reader needed for the ZIF-4 example frame.
"""
import copy as _copy
import re

import numpy as np

from .elements import atomic_masses, atomic_numbers, chemical_symbols


class _Formula:
    """Only ``._count`` is used (amof/msd.py:263): symbol -> number of atoms, in first-seen order."""

    def __init__(self, symbols):
        self._count = {}
        for s in symbols:
            self._count[s] = self._count.get(s, 0) + 1


class _Symbols:
    def __init__(self, numbers):
        self._numbers = numbers

    @property
    def formula(self):
        return _Formula(chemical_symbols[int(z)] for z in self._numbers)

    def __iter__(self):
        return (chemical_symbols[int(z)] for z in self._numbers)

    def __len__(self):
        return len(self._numbers)


class Atoms:
    """positions in Angstrom, cell rows are the lattice vectors, fully periodic."""

    def __init__(self, numbers=None, positions=None, cell=None, symbols=None, masses=None, pbc=True):
        if numbers is None:
            if symbols is None:
                raise ValueError("numbers or symbols required")
            numbers = [atomic_numbers[s] for s in symbols]
        self.numbers = np.array(numbers, dtype=np.int64)
        self.positions = np.array(positions, dtype=np.float64).reshape(len(self.numbers), 3)
        cell = np.array(cell, dtype=np.float64)
        if cell.shape == (3,):
            cell = np.diag(cell)
        self.cell = cell.reshape(3, 3)
        self.pbc = np.array([bool(pbc)] * 3) if np.isscalar(pbc) else np.array(pbc, dtype=bool)
        self._masses = None if masses is None else np.array(masses, dtype=np.float64)

    # -- the surface aMOF touches -------------------------------------------------------------
    def __len__(self):
        return len(self.numbers)

    def get_global_number_of_atoms(self):
        return len(self.numbers)

    def get_atomic_numbers(self):
        return self.numbers.copy()

    def get_positions(self):
        return self.positions.copy()

    def set_positions(self, positions):
        self.positions = np.array(positions, dtype=np.float64).reshape(len(self.numbers), 3)

    def get_cell(self):
        return self.cell.copy()

    def get_pbc(self):
        return self.pbc.copy()

    def get_volume(self):
        return float(abs(np.linalg.det(self.cell)))

    def get_cell_lengths_and_angles(self):
        lengths = np.sqrt((self.cell ** 2).sum(axis=1))
        angles = []
        for i, j in ((1, 2), (0, 2), (0, 1)):
            ll = lengths[i] * lengths[j]
            angles.append(np.degrees(np.arccos(np.dot(self.cell[i], self.cell[j]) / ll)) if ll > 1e-16 else 90.0)
        return np.array(list(lengths) + angles)

    def get_masses(self):
        if self._masses is not None:
            return self._masses.copy()
        return np.array([atomic_masses[int(z)] for z in self.numbers], dtype=np.float64)

    def get_center_of_mass(self):
        m = self.get_masses()
        return np.dot(m, self.positions) / m.sum()

    def translate(self, displacement):
        self.positions = self.positions + np.asarray(displacement, dtype=np.float64)

    def get_chemical_symbols(self):
        return [chemical_symbols[int(z)] for z in self.numbers]

    @property
    def symbols(self):
        return _Symbols(self.numbers)

    def copy(self):
        return _copy.deepcopy(self)

    def rattle(self, stdev=0.001, seed=None, rng=None):
        """Gaussian displacement of every coordinate.  ASE's own default draws from a fixed-seed
        RandomState(42); a counter-based generator is used here instead."""
        if rng is None:
            rng = np.random.Generator(np.random.Philox(42 if seed is None else seed))
        self.positions = self.positions + rng.normal(scale=stdev, size=self.positions.shape)

    def repeat(self, rep):
        rep = (rep,) * 3 if np.isscalar(rep) else tuple(rep)
        n = len(self)
        pos, num = [], []
        for i in range(rep[0]):
            for j in range(rep[1]):
                for k in range(rep[2]):
                    pos.append(self.positions + np.dot((i, j, k), self.cell))
                    num.append(self.numbers)
        cell = self.cell * np.array(rep, dtype=np.float64)[:, None]
        out = Atoms(numbers=np.concatenate(num), positions=np.concatenate(pos), cell=cell)
        assert len(out) == n * rep[0] * rep[1] * rep[2]
        return out


_LATTICE = re.compile(r'Lattice="([^"]*)"')
_PROPS = re.compile(r'Properties=(\S+)')


def read_extxyz(path, index=None):
    """Read (all frames of) an extended-XYZ file with a ``Lattice="..."`` header, e.g.
    /root/reference/examples/files/ZIF-4.xyz:1-3.  Returns a list of :class:`Atoms`, or one
    frame when ``index`` is an int."""
    frames = []
    with open(path) as fh:
        lines = fh.read().split('\n')
    k = 0
    while k < len(lines) and lines[k].strip():
        n = int(lines[k].split()[0])
        header = lines[k + 1]
        m = _LATTICE.search(header)
        if m is None:
            raise ValueError("extended-XYZ header without Lattice= : %r" % header[:80])
        cell = np.array([float(x) for x in m.group(1).split()]).reshape(3, 3)
        pos_col = 1
        pm = _PROPS.search(header)
        if pm is not None:
            fields = pm.group(1).split(':')
            col = 0
            for name, _kind, width in zip(fields[0::3], fields[1::3], fields[2::3]):
                if name == 'pos':
                    pos_col = col
                col += int(width)
        symbols, pos = [], []
        for line in lines[k + 2:k + 2 + n]:
            tok = line.split()
            symbols.append(tok[0])
            pos.append([float(x) for x in tok[pos_col:pos_col + 3]])
        frames.append(Atoms(symbols=symbols, positions=pos, cell=cell))
        k += 2 + n
    return frames if index is None else frames[index]
