"""GPU parity of the explicit neighbour list (amof.atom.get_neighborlist, atom.py:72-87) against the CPU oracle."""
import json
import os

import numpy as np
import pytest

from conftest import random_box
from oracle import c_oracle as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "zif4_known_answers.json")


def _rows(offsets, nbr):
    return [nbr[offsets[i]:offsets[i + 1]].tolist() for i in range(len(offsets) - 1)]


def _oracle_rows(pos, cell, spec, S, cut):
    i, j = orc.neighbour_pairs(pos, cell, spec, S, cut)
    rows = [[] for _ in range(len(pos))]
    for a, b in zip(i.tolist(), j.tolist()):
        rows[a].append(b)
    return rows


def test_zif4_golden(backend, zif4):
    gold = json.load(open(GOLD))
    order = gold["species_order"]
    spec = np.array([order.index(int(z)) for z in zif4.numbers], dtype=np.uint8)
    cut = np.zeros((4, 4))
    cut[3, 2] = cut[2, 3] = 2.5
    off, nbr = backend.neighbour_list(spec, 4, zif4.positions, zif4.cell, cut)
    rows = _rows(off, nbr)
    assert off[-1] == 128                                     # 64 Zn->N + 64 N->Zn
    assert all(len(rows[i]) == 4 and all(spec[j] == 2 for j in rows[i]) for i in np.where(spec == 3)[0])
    assert all(len(rows[i]) == 1 and spec[rows[i][0]] == 3 for i in np.where(spec == 2)[0])
    assert all(len(rows[i]) == 0 for i in np.where(spec < 2)[0])
    assert rows == _oracle_rows(zif4.positions, zif4.cell, spec, 4, cut)


@pytest.mark.parametrize("seed,n,tri,size", [(1, 300, False, 11.0), (2, 500, True, 13.0), (3, 40, True, 5.0), (4, 1, True, 4.0)])
def test_random_boxes(backend, seed, n, tri, size):
    """small boxes included: the same j several times (one entry per periodic image) and self images"""
    S = 3
    pos, cell, spec = random_box(seed, n, S, tri, size, scale_pos=2.0)
    cut = np.array([[2.6, 3.0, 0.0], [3.0, 0.0, 2.8], [0.0, 2.8, 2.4]])
    if n <= 40:
        cut = cut * 2.2                                        # beyond half the box: multiple images
    off, nbr = backend.neighbour_list(spec, S, pos, cell, cut)
    want = _oracle_rows(pos, cell, spec, S, cut)
    assert off[-1] == sum(len(r) for r in want)
    assert _rows(off, nbr) == want


@pytest.mark.parametrize("seed,n,tri,size,scale", [(11, 250, True, 10.0, 1.0), (12, 40, True, 5.0, 2.2), (13, 60, False, 6.0, 1.6)])
def test_distances_and_shifts(backend, seed, n, tri, size, scale):
    """ase's 'd' and 'S' per pair: bit-exact distances, integer image shifts for the positions AS GIVEN (they spill
    outside the cell on purpose), rows ordered by (j, S); D = p_j - p_i + S.cell reproduces the distance."""
    S = 3
    pos, cell, spec = random_box(seed, n, S, tri, size, scale_pos=3.0)
    cut = np.array([[2.6, 3.0, 0.0], [3.0, 0.0, 2.8], [0.0, 2.8, 2.4]]) * scale
    off, nbr, dist, shifts = backend.neighbour_list(spec, S, pos, cell, cut, quantities=True)
    wi, wj, wd, ws = orc.neighbour_pairs(pos, cell, spec, S, cut, quantities=True)
    owner = np.repeat(np.arange(n), np.diff(off))
    assert len(nbr) == len(wi) > 50
    assert np.array_equal(owner, wi) and np.array_equal(nbr, wj)
    assert np.array_equal(shifts, ws)
    assert np.array_equal(dist, wd)                              # same d2 (P3), IEEE sqrt on both sides
    D = pos[nbr] - pos[owner] + shifts @ cell
    np.testing.assert_allclose(np.linalg.norm(D, axis=1), dist, rtol=0, atol=1e-12)
    lim = cut[spec[owner], spec[nbr]]
    assert np.all(dist < lim)


def test_counts_match_the_fused_cn_analysis(backend):
    """the list and the fused counting kernel see the same pairs"""
    S = 3
    pos, cell, spec = random_box(9, 700, S, True, 15.0, scale_pos=2.0)
    cut = np.array([[2.6, 3.0, 0.0], [3.0, 0.0, 2.8], [0.0, 2.8, 2.4]])
    off, nbr = backend.neighbour_list(spec, S, pos, cell, cut)
    got = np.zeros((S, S), dtype=np.uint64)
    owner = np.repeat(np.arange(len(spec)), np.diff(off))
    np.add.at(got, (spec[owner], spec[nbr]), 1)
    res = backend.pair_counts(spec, S, [(pos[None], cell[None])], cn_cutoff=cut)
    assert np.array_equal(got, res["cn"][0])


def test_empty_cases(backend):
    pos, cell, spec = random_box(5, 50, 2, False, 8.0)
    off, nbr = backend.neighbour_list(spec, 2, pos, cell, np.zeros((2, 2)))
    assert off.tolist() == [0] * 51 and len(nbr) == 0
    cut = np.array([[0.0, 1e-6], [1e-6, 0.0]])                 # listed pair, nobody that close
    off, nbr = backend.neighbour_list(spec, 2, pos, cell, cut)
    assert off[-1] == 0 and len(nbr) == 0
    with pytest.raises(ValueError):
        backend.neighbour_list(spec, 2, pos, cell, np.array([[0.0, 1.0], [2.0, 0.0]]))      # not symmetric


def test_get_neighborlist_class_level(zif4, backend):
    from amof_b200 import atom as amatom
    frame = zif4
    nl = amatom.get_neighborlist(frame, amatom.format_cutoff({'Zn-N': 2.5}))
    numbers = np.asarray(frame.get_atomic_numbers())
    assert len(nl) == len(numbers) and isinstance(nl[0], list)
    assert all(len(nl[i]) == 4 for i in np.where(numbers == 30)[0])
    assert sum(len(r) for r in nl) == 128
