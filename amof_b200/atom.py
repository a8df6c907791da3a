"""
Helpers on single frames, same names and meaning as /root/reference/amof/atom.py.

The analyses (amof_b200.cn, amof_b200.bad) never build a neighbour list: the search is fused with the counting and
the angle enumeration on the device.  ``get_neighborlist`` is kept for the callers that want the list itself
(amof.ring, amof.coordination) and runs the same GPU search through ``amofb_neigh_count`` / ``amofb_neigh_fill``.
"""
import numpy as np

from .elements import atomic_numbers


def get_density(atom):
    """mass density in kg/L (atom.py:12-16)"""
    conversion_factor = 1.66053906660
    return conversion_factor * get_total_mass(atom) / atom.get_volume()


def get_number_density(atom):
    """number density in Angstrom^-3 (atom.py:18-22)"""
    return len(atom) / atom.get_volume()


def get_total_mass(atom):
    return np.sum(atom.get_masses())


def select_species_positions(atom, atomic_number):
    """positions of the atoms of one species, all atoms when ``atomic_number`` is None (atom.py:29-42)"""
    if atomic_number is None:
        return atom.get_positions()
    return atom.get_positions()[atom.get_atomic_numbers() == atomic_number]


def get_atomic_numbers_unique(atom):
    """atomic numbers present, in the order of ``list(set(...))`` like the reference (atom.py:44-46)"""
    return list(set(atom.get_atomic_numbers()))


def format_cutoff(nb_set_and_cutoff, format='ase', sort_pair=False):
    """{'Zn-N': 2.5} -> {(30, 7): 2.5}  (atom.py:48-70)"""
    if format == 'ase':
        cutoff_dict = {}
        for nn_set, cutoff in nb_set_and_cutoff.items():
            xx = tuple(atomic_numbers[i] for i in nn_set.split('-'))
            if sort_pair:
                xx = tuple(sorted(xx))
            cutoff_dict[xx] = cutoff
        return cutoff_dict


def get_neighborlist_csr(atom, cutoff_dict, backend=None, quantities=False):
    """Neighbour list of one frame in CSR form: ``(offsets int64[n+1], neighbours int32[offsets[n]])``.

    Same pairs as ``ase.neighborlist.neighbor_list('ij', atom, cutoff_dict)`` (amof/atom.py:82): j is a neighbour of
    i iff d < cutoff[(Zi, Zj)] (both key orders, unlisted pairs never), once per periodic image, without the
    zero-shift self pair.  Inside a row the pairs are ordered by (j, S) (ase leaves that order unspecified).
    ``quantities=True`` appends ``(distances, shifts)``: ase's 'd' and 'S' (D = p_j - p_i + S.cell), the quantities
    ``pymatgen.Structure.get_neighbor_list`` returns to amof.coordination (coordination/core.py:62,181)."""
    from . import _lib, frames
    backend = backend or _lib.get_backend()
    numbers = np.asarray(atom.get_atomic_numbers())
    zs, spec = frames.species_index(numbers)
    cut = cutoff_matrix(cutoff_dict, zs)
    return backend.neighbour_list(spec, len(zs), atom.get_positions(), np.asarray(atom.get_cell(), dtype=np.float64), cut,
                                  quantities=quantities)


def get_neighborlist(atom, cutoff_dict, backend=None):
    """list (one entry per atom) of lists of neighbour indices, as amof.atom.get_neighborlist (atom.py:72-87)"""
    offsets, nbr = get_neighborlist_csr(atom, cutoff_dict, backend)[:2]
    flat = nbr.tolist()
    off = offsets.tolist()
    return [flat[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def cutoff_matrix(cutoff_dict, zs):
    """Dict cutoffs -> symmetric [S][S] matrix over the sorted species list ``zs``.

    Mirrors how ase.neighborlist.neighbor_list applies a dict (amof/atom.py:82): every key acts on both
    orientations of the pair, later keys overwrite earlier ones, unlisted pairs get 0 (never neighbours).
    Keys naming a species that is absent from the frame are ignored."""
    index = {z: k for k, z in enumerate(zs)}
    m = np.zeros((len(zs), len(zs)), dtype=np.float64)
    for (za, zb), c in cutoff_dict.items():
        if za in index and zb in index:
            m[index[za], index[zb]] = m[index[zb], index[za]] = float(c)
    return m
