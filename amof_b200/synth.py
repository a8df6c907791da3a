"""
Deterministic synthetic amorphous-ZIF trajectories: the workloads C1-C5 of SURVEY.md 8(d) / BASELINE.md 3.

Every configuration is built from the 272-atom ZIF-4 cell shipped in ``amof_b200/data/zif4_unit_cell.json``
(extracted from the reference's example frame by tests/golden/make_golden.py): a supercell, Gaussian
amorphisation, optional shear to a triclinic cell, then a random walk in time.  Random numbers come from a
counter-based generator (Philox) seeded with 20261018 + configuration index, so any rank can generate any frame
range without generating the ones before it: frame t = frame 0 + cumulative sum of per-frame increments, where
the increments of frame t use the Philox stream keyed by (seed, t).
"""
import json
import os

import numpy as np

from .elements import atomic_numbers
from .frames import ArrayTrajectory

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "zif4_unit_cell.json")
SEED0 = 20261018

CONFIGS = {
    #        supercell      n_frames  shear  sigma0  step   wrap   description
    "c1": dict(rep=(1, 1, 1), frames=11, shear=False, sigma0=0.0, step=0.5, wrap=False, index=1),
    "c2": dict(rep=(3, 4, 3), frames=10000, shear=False, sigma0=0.30, step=0.05, wrap=True, index=2),
    "c3": dict(rep=(8, 8, 6), frames=2000, shear=True, sigma0=0.30, step=0.05, wrap=True, index=3),
    "c4": dict(rep=(6, 6, 5), frames=5000, shear=False, sigma0=0.30, step=0.05, wrap=True, index=4),
    "c5": dict(rep=(15, 16, 15), frames=5000, shear=False, sigma0=0.30, step=0.05, wrap=False, index=5),
}


def zif4_unit():
    """(numbers[272], positions[272][3], cell[3][3]) of the reference's ZIF-4 example frame."""
    with open(_DATA) as fh:
        d = json.load(fh)
    numbers = np.array([atomic_numbers[s] for s in d["symbols"]], dtype=np.int64)
    positions = np.array([[float(x) for x in row] for row in d["positions"]], dtype=np.float64)
    cell = np.array([[float(x) for x in row] for row in d["cell"]], dtype=np.float64)
    return numbers, positions, cell


def supercell(rep):
    numbers, pos, cell = zif4_unit()
    shifts = np.array([(i, j, k) for i in range(rep[0]) for j in range(rep[1]) for k in range(rep[2])], dtype=np.float64)
    allpos = (pos[None, :, :] + (shifts @ cell)[:, None, :]).reshape(-1, 3)
    allnum = np.tile(numbers, len(shifts))
    return allnum, allpos, cell * np.array(rep, dtype=np.float64)[:, None]


def _rng(seed, stream):
    return np.random.Generator(np.random.Philox(key=[seed, stream]))


def _wrap(pos, cell, inv):
    f = pos @ inv
    f -= np.floor(f)
    return f @ cell


def base_frame(name):
    """(numbers, positions of frame 0, cell) of a configuration."""
    cfg = CONFIGS[name]
    seed = SEED0 + cfg["index"]
    numbers, pos, cell = supercell(cfg["rep"])
    if cfg["sigma0"] > 0:
        pos = pos + _rng(seed, 0).normal(scale=cfg["sigma0"], size=pos.shape)
    if cfg["shear"]:
        frac = pos @ np.linalg.inv(cell)
        cell = cell.copy()
        a, b = cell[0].copy(), cell[1].copy()
        cell[1] = b + 0.10 * a
        cell[2] = cell[2] + 0.05 * a - 0.08 * b
        pos = frac @ cell
    if cfg["wrap"]:
        pos = _wrap(pos, cell, np.linalg.inv(cell))
    return numbers, pos, cell


def fill_frames(name, first, count, out, drift=None):
    """Write frames [first, first+count) of configuration ``name`` into ``out[count][N][3]``.

    Frame t = base + sum_{s=1..t} increment_s; the running sum up to ``first`` is recomputed from the per-frame
    streams (cheap next to the analysis and exactly reproducible on any rank)."""
    cfg = CONFIGS[name]
    seed = SEED0 + cfg["index"]
    numbers, pos, cell = base_frame(name)
    inv = np.linalg.inv(cell)
    cur = pos.copy()
    if first == 0 and count > 0:
        out[0] = pos
    for t in range(1, first + count):
        cur += _rng(seed, t).normal(scale=cfg["step"], size=pos.shape)
        if drift is not None:
            cur += drift
        if t >= first:
            out[t - first] = _wrap(cur, cell, inv) if cfg["wrap"] else cur
    return numbers, cell


def make_trajectory(name, n_frames=None, out=None):
    """ArrayTrajectory of the first ``n_frames`` frames of a configuration (all of them by default)."""
    cfg = CONFIGS[name]
    T = cfg["frames"] if n_frames is None else int(n_frames)
    numbers, pos, cell = base_frame(name)
    if out is None:
        out = np.empty((T, len(numbers), 3), dtype=np.float64)
    drift = np.array([1e-4, -2e-4, 1.5e-4]) if name == "c5" else None
    fill_frames(name, 0, T, out, drift=drift)
    return ArrayTrajectory(numbers, out, cell)


def reduced_network(name="c4", n_frames=None):
    """Zn + imidazolate-centroid pseudo-atoms (Fr stands for 'Im', cf. amof/symbols.py:15-18): per ZIF-4 cell 16 Zn
    and 32 ring centroids, each centroid the mean of the 2 N of one N-C-N bridge.  Used for the Zn-Im-Zn angles."""
    traj = make_trajectory(name, n_frames)
    numbers = traj.numbers
    base = base_frame(name)[1]
    cell = traj.cells[0]
    inv = np.linalg.inv(cell)
    zn = np.where(numbers == 30)[0]
    n_idx = np.where(numbers == 7)[0]
    # pair up the two N of each imidazolate in frame 0: nearest N-N partner at ~2.2 A (same ring)
    pn = base[n_idx]
    partner = np.full(len(n_idx), -1)
    order = np.argsort(pn[:, 0])
    from math import inf
    for ii in range(len(n_idx)):
        if partner[ii] >= 0:
            continue
        d = pn - pn[ii]
        f = d @ inv
        f -= np.round(f)
        d = f @ cell
        r2 = (d * d).sum(axis=1)
        r2[ii] = inf
        r2[partner >= 0] = inf
        jj = int(np.argmin(r2))
        partner[ii], partner[jj] = jj, ii
    del order
    firsts = np.array([i for i in range(len(n_idx)) if i < partner[i]])
    seconds = partner[firsts]
    T = len(traj)
    pos = np.empty((T, len(zn) + len(firsts), 3))
    for t in range(T):
        p = traj.positions[t]
        a, b = p[n_idx[firsts]], p[n_idx[seconds]]
        d = b - a
        f = d @ inv
        f -= np.round(f)
        mid = a + 0.5 * (f @ cell)
        pos[t, :len(zn)] = p[zn]
        pos[t, len(zn):] = _wrap(mid, cell, inv)
    nums = np.concatenate([np.full(len(zn), 30), np.full(len(firsts), atomic_numbers["Fr"])])
    return ArrayTrajectory(nums, pos, cell)
