"""
Helpers on single frames, same names and meaning as /root/reference/amof/atom.py.

``get_neighborlist`` is deliberately NOT provided as a Python list-of-lists (atom.py:72-87): on the GPU path the
neighbour search is fused with the counting (amof_b200.cn) and the angle enumeration (amof_b200.bad).
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
