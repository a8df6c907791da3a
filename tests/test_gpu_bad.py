"""GPU parity of the bond-angle analysis against the CPU oracle, through the C ABI (bit-exact histograms)."""
import json
import os

import numpy as np
import pytest

from conftest import random_box
from oracle import c_oracle as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "zif4_known_answers.json")
MAXCN = 32


def _oracle(pos, cell, spec, S, cut, triples, dtheta, nbins):
    hist = np.zeros((len(triples), MAXCN + 1, nbins), dtype=np.uint64)
    dropped = np.zeros(len(triples), dtype=np.uint64)
    for f in range(len(pos)):
        for t, (A, B) in enumerate(triples):
            _, d = orc.bad_hist(pos[f], cell[f], spec, S, cut, A, B, dtheta, nbins, max_cn=MAXCN, hist=hist[t])
            dropped[t] += d
    return hist, dropped


def test_zif4_golden(backend, zif4):
    gold = json.load(open(GOLD))
    order = gold["species_order"]
    spec = np.array([order.index(int(z)) for z in zif4.numbers], dtype=np.uint8)
    cut = np.zeros((4, 4))
    cut[3, 2] = cut[2, 3] = 2.5
    nbins = int(180 // 0.05) + 1
    assert nbins == 3600
    hist, dropped, nf = backend.bad_counts(spec, 4, [(zif4.positions[None], zif4.cell[None])], cut, [(3, 2), (2, 3)], 0.05, nbins)
    g = gold["bad_N_Zn_N@2.5"]
    assert int(hist[0].sum()) == g["count"] == 96 and int(hist[0, 4].sum()) == 96      # every Zn has 4 N
    assert int(hist[1].sum()) == 0 and int(dropped.sum()) == 0 and nf == 1
    edges = np.arange(nbins + 1) * 0.05
    want = np.histogram(g["angles_sorted"], bins=edges)[0]
    assert np.array_equal(hist[0].sum(axis=0), want.astype(np.uint64))


@pytest.mark.parametrize("seed,n,tri,size,dtheta", [
    (1, 400, False, 14.0, 0.05),
    (2, 700, True, 16.0, 0.05),
    (3, 900, True, 18.0, 1.0),
    (4, 300, True, 12.0, 0.7),      # 180 // 0.7 = 257 -> last edge 180.6 > 180
    (5, 500, False, 15.0, 7.0),     # 180 // 7 = 25 -> last edge 182
])
def test_random_boxes(backend, seed, n, tri, size, dtheta):
    S = 3
    T = 2
    frames = [random_box(seed * 10 + f, n, S, tri, size, scale_pos=2.0) for f in range(T)]
    spec = frames[0][2]
    pos = np.array([f[0] for f in frames])
    cell = np.array([f[1] for f in frames])
    cut = np.array([[2.6, 3.0, 0.0], [3.0, 0.0, 2.8], [0.0, 2.8, 2.4]])
    triples = [(0, 1), (1, 0), (1, 2), (2, 2), (0, -1), (-1, -1)]
    nbins = int(180 // dtheta) + 1
    hist, dropped, nf = backend.bad_counts(spec, S, [(pos, cell)], cut, triples, dtheta, nbins)
    want, wdrop = _oracle(pos, cell, spec, S, cut, triples, dtheta, nbins)
    assert int(want.sum()) > 50
    assert np.array_equal(hist, want)
    assert np.array_equal(dropped, wdrop)
    assert nf == T


@pytest.mark.parametrize("triples", [[(0, 1)], [(0, 1), (2, 1)], [(1, -1)], [(2, 2), (0, 1)]])
@pytest.mark.parametrize("one_list", [False, True])
def test_centre_list_and_species_lists(backend, monkeypatch, triples, one_list):
    """Only some of the kept species are centres: the search then runs over the compact centre list the scatter kernel writes
    (one thread per possible centre), and with one cell list per species it visits the lists of the partner species only.
    Both against the oracle, and the single mixed list (AMOFB_BAD_ONE_LIST) against the same numbers."""
    S, T = 3, 3
    frames = [random_box(77 + f, 600, S, True, 15.0, scale_pos=2.0) for f in range(T)]
    spec = frames[0][2]
    pos = np.array([f[0] for f in frames])
    cell = np.array([f[1] for f in frames])
    cut = np.array([[2.6, 3.0, 0.0], [3.0, 0.0, 2.8], [0.0, 2.8, 2.4]])
    if one_list:
        monkeypatch.setenv("AMOFB_BAD_ONE_LIST", "1")
    hist, dropped, nf = backend.bad_counts(spec, S, [(pos, cell)], cut, triples, 0.5, 361)
    want, wdrop = _oracle(pos, cell, spec, S, cut, triples, 0.5, 361)
    assert int(want.sum()) > 20 and nf == T
    assert np.array_equal(hist, want) and np.array_equal(dropped, wdrop)


@pytest.mark.parametrize("tri", [False, True])
def test_species_without_cutoffs_are_filtered(backend, tri):
    """Species that appear in no cutoff pair never enter the cell list (PrepArgs::species_keep): the counts of the
    listed triples, of 'any neighbour' (B = -1) and of 'any centre' (A = -1) triples must not change."""
    S = 5
    T = 3
    frames = [random_box(70 + f, 800, S, tri, 17.0, scale_pos=2.0) for f in range(T)]
    spec = frames[0][2]
    pos = np.array([f[0] for f in frames])
    cell = np.array([f[1] for f in frames])
    cut = np.zeros((S, S))
    cut[1, 3] = cut[3, 1] = 3.1          # species 0, 2 and 4 have no cutoff at all: 60 % of the atoms drop out
    cut[3, 3] = 2.7
    triples = [(1, 3), (3, 1), (3, 3), (3, -1), (-1, 3), (-1, -1), (0, 2), (2, -1)]
    nbins = int(180 // 0.5) + 1
    hist, dropped, nf = backend.bad_counts(spec, S, [(pos, cell)], cut, triples, 0.5, nbins)
    want, wdrop = _oracle(pos, cell, spec, S, cut, triples, 0.5, nbins)
    assert int(want[:6].sum()) > 100 and int(want[6:].sum()) == 0
    assert np.array_equal(hist, want)
    assert np.array_equal(dropped, wdrop)
    assert nf == T


def test_collinear_and_degenerate_angles(backend):
    """0 and 180 degree angles sit exactly on histogram edges; a short dtheta grid must agree with the oracle."""
    cell = np.diag([20.0, 20.0, 20.0])
    pos = np.array([[10.0, 10.0, 10.0], [11.5, 10.0, 10.0], [8.5, 10.0, 10.0], [10.0, 11.5, 10.0],
                    [10.0, 10.0, 8.2], [12.0, 10.0, 10.0]])
    spec = np.array([0, 1, 1, 1, 1, 1], dtype=np.uint8)
    cut = np.array([[0.0, 2.5], [2.5, 0.0]])
    for dtheta in (0.05, 1.0, 90.0, 45.0):
        nbins = int(180 // dtheta) + 1
        hist, dropped, _ = backend.bad_counts(spec, 2, [(pos[None], cell[None])], cut, [(0, 1)], dtheta, nbins)
        want, wdrop = _oracle(pos[None], cell[None], spec, 2, cut, [(0, 1)], dtheta, nbins)
        assert np.array_equal(hist, want) and np.array_equal(dropped, wdrop)
        assert int(hist.sum() + dropped.sum()) == 10          # C(5, 2) angles around the centre


def test_cutoff_precondition(backend):
    pos, cell, spec = random_box(7, 50, 2, False, 6.0)
    cut = np.full((2, 2), 4.5)      # above half the cell height (cell edges are 6.0 .. 7.8)
    with pytest.raises(ValueError):
        backend.bad_counts(spec, 2, [(pos[None], cell[None])], cut, [(0, 1)], 0.05, 3600)


def test_too_many_neighbours_is_an_error(backend):
    """more than 32 B-neighbours around a centre is refused (AMOFB_BAD_MAX_CN), like the oracle"""
    rng = np.random.default_rng(0)
    cell = np.diag([30.0, 30.0, 30.0])
    pos = np.vstack([[15.0, 15.0, 15.0], 15.0 + rng.normal(scale=0.8, size=(60, 3))])
    spec = np.array([0] + [1] * 60, dtype=np.uint8)
    cut = np.array([[0.0, 6.0], [6.0, 0.0]])
    with pytest.raises(ValueError):
        backend.bad_counts(spec, 2, [(pos[None], cell[None])], cut, [(0, 1)], 1.0, 181)
    ok = backend.bad_counts(spec, 2, [(pos[None], cell[None])], np.array([[0.0, 0.5], [0.5, 0.0]]), [(0, 1)], 1.0, 181)
    assert ok[2] == 1


def test_many_frames_streaming(backend, monkeypatch):
    monkeypatch.setenv("AMOFB_BATCH_ATOMS", "3000")
    T, n = 23, 400
    rng = np.random.default_rng(2)
    base, cell0, spec = random_box(90, n, 2, True, 14.0)
    pos = base[None] + rng.normal(scale=0.2, size=(T, n, 3))
    cell = np.array([cell0 * (1 + 0.001 * f) for f in range(T)])
    cut = np.array([[0.0, 2.7], [2.7, 2.2]])
    triples = [(0, 1), (1, 1), (1, -1)]
    hist, dropped, nf = backend.bad_counts(spec, 2, [(pos[:9], cell[:9]), (pos[9:], cell[9:])], cut, triples, 0.5, 361)
    want, wdrop = _oracle(pos, cell, spec, 2, cut, triples, 0.5, 361)
    assert nf == T and np.array_equal(hist, want) and np.array_equal(dropped, wdrop)


def test_cosine_beyond_one_is_clipped(backend):
    """Exactly collinear neighbours: with v1 = 2 v0 the unit vectors are the same doubles and u.u = 1 + 2^-52, with
    v1 = -2 v0 it is -(1 + 2^-52).  ase clips the cosine before arccos (0 and 180 degrees); nothing may be dropped."""
    cell = np.diag([32.0, 32.0, 32.0])
    c = np.array([16.0, 16.0, 16.0])
    v = np.array([-4.0, -1.0, -2.0]) / 4.0                      # |v| = 1.1456; every sum below is exact in binary64
    n = np.sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2])
    u = v / n
    assert (u[0] * u[0] + u[1] * u[1]) + u[2] * u[2] > 1.0
    spec = np.array([0, 1, 1], dtype=np.uint8)
    cut = np.array([[0.0, 2.5], [2.5, 0.0]])
    for sign, where in ((2.0, 0), (-2.0, -1)):
        pos = np.vstack([c, c + v, c + sign * v])
        for dtheta in (0.05, 1.0):
            nbins = int(180 // dtheta) + 1
            hist, dropped, _ = backend.bad_counts(spec, 2, [(pos[None], cell[None])], cut, [(0, 1)], dtheta, nbins)
            want, wdrop = _oracle(pos[None], cell[None], spec, 2, cut, [(0, 1)], dtheta, nbins)
            assert np.array_equal(hist, want) and int(dropped.sum()) == 0 and int(wdrop.sum()) == 0
            assert int(hist[0, 2].sum()) == 1 and int(hist[0, 2][where]) == 1      # first bin / the closed last bin
