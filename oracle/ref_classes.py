"""
Restatement of the reference's DRIVER code (the Python around the third-party numerics) on top of the C oracle.
TEST INFRASTRUCTURE ONLY -- parity unpinned (asap3 / ase are not installable here; see amof_oracle.c).

Each function follows the control flow of the reference method it names and produces the same DataFrame, so that
tests can compare ``amof_b200`` objects column by column:

    rdf_dataframe   amof/rdf.py:67-114   (asap3 object replaced by orc.rdf_traj + the normalisation pins U3/U4/a3)
    cn_dataframe    amof/cn.py:48-82     (ase neighbour list replaced by orc.cn_counts)
    bad_dataframe   amof/bad.py:116-160  (angle LISTS from orc.bad_angles, then np.histogram(density=True) as there)
    wmsd_dataframe  amof/msd.py:157-268  (get_delta_pos -> orc.delta_pos; compute_msd_of_m restated with numpy)
"""
import numpy as np
import pandas as pd

from amof_b200.elements import atomic_numbers, chemical_symbols
from oracle import c_oracle as orc


def _species(numbers):
    zs = sorted(set(int(z) for z in numbers))
    return zs, np.array([zs.index(int(z)) for z in numbers], dtype=np.uint8)


def _arrays(trajectory):
    pos = np.array([a.get_positions() for a in trajectory])
    cell = np.array([np.asarray(a.get_cell()) for a in trajectory]).reshape(len(pos), 3, 3)
    return pos, cell


def rdf_dataframe(trajectory, dr=0.01, rmax='half_cell'):
    atomic_numbers_unique = list(set(trajectory[0].get_atomic_numbers()))
    rmax_half_cell = np.min([a for t in trajectory for a in t.get_cell_lengths_and_angles()[0:3]]) / 2
    if rmax == 'half_cell':
        rmax = rmax_half_cell
    elif rmax > rmax_half_cell:
        rmax = rmax_half_cell
    bins = int(rmax // dr)
    data = pd.DataFrame({"r": np.arange(bins) * dr})
    zs, spec = _species(trajectory[0].get_atomic_numbers())
    pos, cell = _arrays(trajectory)
    hist, vsum = orc.rdf_traj(pos, cell, spec, len(zs), float(rmax), bins, method=1, threads=1)
    T, N = len(trajectory), len(spec)
    volume = vsum / T                                                     # U4
    d = rmax / bins
    i = np.arange(bins)
    shell = 4 * np.pi / 3 * (((i + 1) * d) ** 3 - (i * d) ** 3)        # U3

    def get_rdf(counts, ncentre):                                         # a3
        return counts / (shell * (ncentre * T) * (N / volume))
    data["X-X"] = get_rdf(hist.sum(axis=(0, 1)).astype(float), N)
    partial = {}
    for x in atomic_numbers_unique:
        for y in atomic_numbers_unique:
            a, b = zs.index(int(x)), zs.index(int(y))
            partial[(x, y)] = get_rdf(hist[a, b].astype(float), int((spec == a).sum()))
            data[chemical_symbols[x] + "-" + chemical_symbols[y]] = partial[(x, y)]
    for x in atomic_numbers_unique:
        data[chemical_symbols[x] + "-X"] = sum([partial[(x, y)] for y in atomic_numbers_unique])
    return data


def _cutoff_matrix(nb_set_and_cutoff, zs):
    m = np.zeros((len(zs), len(zs)))
    for nn_set, c in nb_set_and_cutoff.items():
        a, b = (atomic_numbers[s] for s in nn_set.split('-'))
        if a in zs and b in zs:
            m[zs.index(a), zs.index(b)] = m[zs.index(b), zs.index(a)] = c
    return m


def cn_dataframe(trajectory, nb_set_and_cutoff, delta_Step=1, first_frame=0):
    zs, spec = _species(trajectory[0].get_atomic_numbers())
    cut = _cutoff_matrix(nb_set_and_cutoff, zs)
    rows = []
    for i, atom in enumerate(trajectory):
        dic = {'Step': first_frame + i * delta_Step}
        counts = orc.cn_counts(atom.get_positions(), np.asarray(atom.get_cell()), spec, len(zs), cut)
        for nb_set in nb_set_and_cutoff:
            a, b = (atomic_numbers[s] for s in nb_set.split('-'))
            na = int((spec == zs.index(a)).sum())
            dic[nb_set] = counts[zs.index(a), zs.index(b)] / na          # np.mean of the per-atom neighbour counts
        rows.append(dic)
    return pd.DataFrame(rows)


def bad_dataframe(trajectory, nb_set_and_cutoff, dtheta=0.05):
    atomic_numbers_unique = list(set(trajectory[0].get_atomic_numbers()))
    elements_present_unique = list(set([atomic_numbers[i] for nb_set in nb_set_and_cutoff.keys() for i in nb_set.split('-')]))
    if len(elements_present_unique) == len(atomic_numbers_unique):
        elements_present_unique.append("X")
    elements = [(a, b) for b in elements_present_unique for a in elements_present_unique
                if (a not in [b, "X"] or ((a, b) == ("X", "X")))]
    bins = int(180 // dtheta)
    theta_bins = np.arange(bins + 2) * dtheta
    theta = np.arange(bins + 1) * dtheta + dtheta / 2
    data = pd.DataFrame({"theta": theta})
    zs, spec = _species(trajectory[0].get_atomic_numbers())
    cut = _cutoff_matrix(nb_set_and_cutoff, zs)

    def sym(c):
        return "X" if c == "X" else chemical_symbols[c]
    for A, B in elements:
        angles = []
        for atom in trajectory:
            ia = -1 if A == "X" else zs.index(A)
            ib = -1 if B == "X" else zs.index(B)
            angles += list(orc.bad_angles(atom.get_positions(), np.asarray(atom.get_cell()), spec, len(zs), cut, ia, ib))
        if angles != []:
            data["-".join([sym(B), sym(A), sym(B)])] = np.histogram(angles, bins=theta_bins, density=True)[0]
    return data


def msd_of_m(delta_pos, m):
    """amof/msd.py:186-205 (without the aliasing drift Q5)"""
    MSD_partial = np.zeros(len(delta_pos) - m)
    r_k_minus_m = np.array(delta_pos[0])
    r_k = r_k_minus_m * 0
    for k in range(0, m + 1):
        r_k = r_k + delta_pos[k]
    for k in range(m + 1, len(delta_pos)):
        r_k = r_k + delta_pos[k]
        r_k_minus_m = r_k_minus_m + delta_pos[k - m]
        MSD_partial[k - m] = np.linalg.norm(r_k - r_k_minus_m) ** 2 / len(r_k_minus_m)
    return np.mean(MSD_partial)


def wmsd_dataframe(trajectory, delta_time=100, max_time="half", timestep=1, unwrap=False):
    """Works on copies: the caller's frames are left alone."""
    half_time = (len(trajectory) // 2) * timestep
    if max_time == "half" or max_time > half_time:
        max_time = half_time
    window = np.arange(0, max_time // timestep, delta_time // timestep)
    time = timestep * window
    elements = list(set(trajectory[0].get_atomic_numbers()))
    pos, cell = _arrays(trajectory)
    numbers = np.asarray(trajectory[0].get_atomic_numbers())
    masses = np.asarray(trajectory[0].get_masses())
    if unwrap:
        pos = np.cumsum(orc.delta_pos(pos, cell), axis=0)
    com = (masses[None, :, None] * pos).sum(axis=1) / masses.sum()
    pos = pos - com[:, None, :]
    data = pd.DataFrame({"Time": time})
    for x in elements:
        sel = numbers == x
        delta = orc.delta_pos(np.ascontiguousarray(pos[:, sel]), cell)
        data[chemical_symbols[x]] = [msd_of_m(delta, int(m)) for m in window]
    formula = {}
    for z in numbers:
        formula[chemical_symbols[int(z)]] = formula.get(chemical_symbols[int(z)], 0) + 1
    data['X'] = np.sum([data[k].to_numpy() * v for k, v in formula.items()], axis=0) / sum(formula.values())
    return data
