"""The CPU oracle against the committed golden vectors, analytic known answers, and its
own independent second forms (brute force vs linked cells in C, numpy twin).  Runs without a GPU."""
import itertools
import json
import math
import os

import numpy as np
import pytest

from conftest import random_box
from oracle import c_oracle as orc
from oracle import np_oracle as npo

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "zif4_known_answers.json")))


def _spec(zif4):
    order = GOLD["species_order"]
    return np.array([order.index(int(z)) for z in zif4.numbers], dtype=np.uint8)


def test_zif4_composition_and_volume(zif4):
    assert len(zif4) == GOLD["n_atoms"] == 272
    assert {str(z): int((zif4.numbers == z).sum()) for z in GOLD["species_order"]} == GOLD["composition"]
    assert abs(zif4.get_volume() - 4380.485812) < 1e-5
    assert abs(orc.lib().orc_cell_volume(np.ascontiguousarray(zif4.cell).ctypes.data_as(orc._dp)) - GOLD["volume"]) < 1e-9


@pytest.mark.parametrize("method", [0, 1])
def test_zif4_rdf_golden(zif4, method):
    g = GOLD["rdf_default"]
    assert abs(g["rmax"] - 7.7021) < 1e-4 and g["bins"] == int(g["rmax"] // 0.01) == 770
    hist = orc.rdf_hist(zif4.positions, zif4.cell, _spec(zif4), 4, g["rmax"], g["bins"], method=method)
    assert int(hist.sum()) == g["directed_pairs_total"] == 30968
    order = GOLD["species_order"]
    for i, j in itertools.product(range(4), range(4)):
        assert int(hist[i, j].sum()) == g["directed_pairs"]["%d-%d" % (order[i], order[j])]
    assert int(hist[3, 2].sum()) == 472 and int(hist[3, 3].sum()) == 80      # Zn->N, Zn->Zn
    assert [int(x) for x in hist[3, 2]] == g["hist_Zn_N"]
    assert [int(x) for x in hist.sum(axis=(0, 1))] == g["hist_total"]


@pytest.mark.parametrize("method", [0, 1])
def test_zif4_cn_golden(zif4, method):
    spec = _spec(zif4)
    want = {"Zn-N@2.5": (3, 2, 2.5, 64), "N-Zn@2.5": (2, 3, 2.5, 64), "Zn-Zn@7.0": (3, 3, 7.0, 64),
            "C-N@1.728": (1, 2, 1.728, 128), "C-C@1.752": (1, 1, 1.752, 64)}
    for name, (a, b, c, n) in want.items():
        cut = np.zeros((4, 4))
        cut[a, b] = cut[b, a] = c
        counts = orc.cn_counts(zif4.positions, zif4.cell, spec, 4, cut, method=method)
        assert int(counts[a, b]) == n == GOLD["cn_directed_pairs"][name]
    # every Zn has exactly 4 N: mean 4.0; N-Zn 1.0; C-N 4/3
    assert 64 / 16 == 4.0 and 64 / 64 == 1.0 and abs(128 / 96 - 4 / 3) < 1e-15


@pytest.mark.parametrize("method", [0, 1])
def test_zif4_neighbour_list_golden(zif4, method):
    """amof.atom.get_neighborlist on the example frame: 64 Zn->N + 64 N->Zn pairs, every Zn exactly 4 N, every N one
    Zn; C-N@1.728: 128 + 128 directed pairs; the list and the counting form agree."""
    spec = _spec(zif4)
    cut = np.zeros((4, 4))
    cut[3, 2] = cut[2, 3] = 2.5
    i, j = orc.neighbour_pairs(zif4.positions, zif4.cell, spec, 4, cut, method=method)
    assert len(i) == 128 == 2 * GOLD["cn_directed_pairs"]["Zn-N@2.5"]
    deg = np.bincount(i, minlength=272)
    assert np.all(deg[spec == 3] == 4) and np.all(deg[spec == 2] == 1) and np.all(deg[spec < 2] == 0)
    assert np.all(spec[j[spec[i] == 3]] == 2) and np.all(spec[j[spec[i] == 2]] == 3)
    fwd = set(zip(i.tolist(), j.tolist()))
    assert fwd == set(zip(j.tolist(), i.tolist()))             # symmetric cutoffs: (i, j) listed iff (j, i) is
    cut = np.zeros((4, 4))
    cut[1, 2] = cut[2, 1] = 1.728
    i, j = orc.neighbour_pairs(zif4.positions, zif4.cell, spec, 4, cut, method=method)
    counts = orc.cn_counts(zif4.positions, zif4.cell, spec, 4, cut, method=method)
    assert len(i) == 256 and int(counts[1, 2]) == 128 == int((spec[i] == 1).sum())


def test_neighbour_list_images_and_twin():
    """one atom in a unit cube with cutoff 1.1: its 6 face images, each listed as j = 0; sqrt(2) < 1.5 adds the 12 edge
    images; and the numpy twin lists the same pairs on random triclinic boxes (brute force and linked cells alike)"""
    pos = np.array([[0.3, 0.4, 0.5]])
    spec = np.zeros(1, dtype=np.uint8)
    for c, n in ((1.1, 6), (1.5, 18), (0.9, 0)):
        for method in (0, 1):
            i, j = orc.neighbour_pairs(pos, np.eye(3), spec, 1, np.array([[c]]), method=method)
            assert len(i) == n and np.all(i == 0) and np.all(j == 0)
    cut = np.array([[2.6, 3.0, 0.0], [3.0, 0.0, 2.8], [0.0, 2.8, 2.4]])
    for seed, n, size in ((1, 150, 9.0), (2, 30, 4.5)):
        p, cell, sp = random_box(seed, n, 3, True, size, scale_pos=2.0)
        i2, j2, _, _ = npo.neighbour_pairs(p, cell, sp, cut)
        o = np.lexsort((j2, i2))
        for method in (0, 1):
            i, j = orc.neighbour_pairs(p, cell, sp, 3, cut, method=method)
            assert len(i) > 20 and np.array_equal(i, i2[o]) and np.array_equal(j, j2[o])


def test_neighbour_list_distances_and_shifts():
    """ase's 'd' and 'S': for one atom in a unit cube the 6 neighbours at cutoff 1.1 are its own images S = +-e_k at
    distance exactly 1; on random boxes with positions outside the cell, D = p_j - p_i + S.cell reproduces d and both
    enumerations return the same (i, j, S, d)."""
    pos = np.array([[2.3, -0.6, 0.5]])                        # outside the cell on purpose
    i, j, d, S = orc.neighbour_pairs(pos, np.eye(3), np.zeros(1, dtype=np.uint8), 1, np.array([[1.1]]), quantities=True)
    assert sorted(map(tuple, S.tolist())) == sorted([(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)])
    assert np.all(d == 1.0)
    cut = np.array([[2.6, 3.0, 0.0], [3.0, 0.0, 2.8], [0.0, 2.8, 2.4]]) * 1.5
    p, cell, sp = random_box(8, 80, 3, True, 6.0, scale_pos=3.0)
    a = orc.neighbour_pairs(p, cell, sp, 3, cut, method=0, quantities=True)
    b = orc.neighbour_pairs(p, cell, sp, 3, cut, method=1, quantities=True)
    assert len(a[0]) > 200 and all(np.array_equal(x, y) for x, y in zip(a, b))
    i, j, d, S = a
    np.testing.assert_allclose(np.linalg.norm(p[j] - p[i] + S @ cell, axis=1), d, rtol=0, atol=1e-12)
    assert np.abs(S).max() >= 2                               # positions spill over several cells: S absorbs it


def test_zif4_bad_golden(zif4):
    g = GOLD["bad_N_Zn_N@2.5"]
    cut = np.zeros((4, 4))
    cut[3, 2] = cut[2, 3] = 2.5
    ang = np.sort(orc.bad_angles(zif4.positions, zif4.cell, _spec(zif4), 4, cut, 3, 2))
    assert len(ang) == g["count"] == 96
    assert abs(ang[0] - 103.377) < 1e-3 and abs(ang[-1] - 113.190) < 1e-3 and abs(ang.mean() - 109.437) < 1e-3
    assert np.array_equal(ang, np.array(g["angles_sorted"]))            # bit-exact with the numpy twin's fixture
    hist, dropped = orc.bad_hist(zif4.positions, zif4.cell, _spec(zif4), 4, cut, 3, 2, 0.05, 3600)
    assert int(hist[4].sum()) == 96 and int(hist.sum()) == 96 and dropped == 0
    assert np.array_equal(hist.sum(axis=0), np.histogram(ang, bins=np.arange(3601) * 0.05)[0].astype(np.uint64))


@pytest.mark.parametrize("seed,n,tri,size,rmax", [(1, 60, True, 7.0, 9.0), (2, 120, False, 11.0, 5.0), (3, 5, True, 4.0, 10.0)])
def test_three_forms_agree(seed, n, tri, size, rmax):
    pos, cell, spec = random_box(seed, n, 3, tri, size, scale_pos=3.0)
    a = orc.rdf_hist(pos, cell, spec, 3, rmax, 400, method=0)
    b = orc.rdf_hist(pos, cell, spec, 3, rmax, 400, method=1)
    c = npo.rdf_hist(pos, cell, spec, 3, rmax, 400)
    assert np.array_equal(a, b) and np.array_equal(a, c) and a.sum() > 0
    cut = np.array([[2.0, 2.5, 0.0], [2.5, 1.8, 3.0], [0.0, 3.0, 0.0]])
    assert np.array_equal(orc.cn_counts(pos, cell, spec, 3, cut, method=0), npo.cn_counts(pos, cell, spec, 3, cut))
    assert np.array_equal(orc.cn_counts(pos, cell, spec, 3, cut, method=1), npo.cn_counts(pos, cell, spec, 3, cut))


def test_bad_angles_match_numpy_twin():
    pos, cell, spec = random_box(5, 150, 2, True, 10.0)
    cut = np.array([[0.0, 2.4], [2.4, 0.0]])
    got = np.sort(orc.bad_angles(pos, cell, spec, 2, cut, 0, 1))
    twin = npo.bad_angles(pos, cell, spec, cut, 0, 1)
    want = np.sort(np.array(sum(twin.values(), [])))
    assert len(got) == len(want) > 10 and np.array_equal(got, want)


def _lattice(kind, a, reps):
    basis = {"sc": [(0, 0, 0)], "fcc": [(0, 0, 0), (0.5, 0.5, 0), (0.5, 0, 0.5), (0, 0.5, 0.5)]}[kind]
    pos = [(np.array(b) + np.array(s)) * a for s in itertools.product(range(reps), repeat=3) for b in basis]
    return np.array(pos), np.eye(3) * a * reps


@pytest.mark.parametrize("kind,shells", [("sc", [(1.0, 6), (math.sqrt(2), 12), (math.sqrt(3), 8), (2.0, 6), (math.sqrt(5), 24)]),
                                         ("fcc", [(math.sqrt(0.5), 12), (1.0, 6), (math.sqrt(1.5), 24), (math.sqrt(2), 12), (math.sqrt(2.5), 24)])])
def test_lattice_shells(kind, shells):
    a = 2.0
    pos, cell = _lattice(kind, a, 5)
    n = len(pos)
    spec = np.zeros(n, dtype=np.uint8)
    rmax, nbins = 4.8, 4800
    hist = orc.rdf_hist(pos, cell, spec, 1, rmax, nbins)[0, 0]
    dr = rmax / nbins
    for d, mult in shells:
        b = int(round(d * a / dr))
        assert int(hist[b - 2:b + 3].sum()) == mult * n
    assert int(hist[:int(0.5 * a / dr)].sum()) == 0


def test_tetrahedral_and_square_planar_angles():
    cell = np.eye(3) * 30.0
    c = np.array([15.0, 15.0, 15.0])
    tet = np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]]) / math.sqrt(3) * 2.0
    pos = np.vstack([c, c + tet])
    spec = np.array([0, 1, 1, 1, 1], dtype=np.uint8)
    cut = np.array([[0.0, 2.5], [2.5, 0.0]])
    ang = orc.bad_angles(pos, cell, spec, 2, cut, 0, 1)
    assert len(ang) == 6 and np.allclose(ang, 109.47122063449069, atol=1e-10)
    sq = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0]]) * 2.0
    ang = np.sort(orc.bad_angles(np.vstack([c, c + sq]), cell, spec, 2, cut, 0, 1))
    assert np.allclose(ang, [90, 90, 90, 90, 180, 180], atol=1e-10)


def test_theta_bin_is_numpy_histogram():
    rng = np.random.default_rng(0)
    for dtheta in (0.05, 0.7, 1.0, 7.0):
        bins = int(180 // dtheta) + 1
        edges = np.arange(bins + 1) * dtheta
        th = np.concatenate([rng.uniform(0, 180, 2000), edges[edges <= 180.0], [0.0, 180.0, np.nextafter(180.0, 0)]])
        want = np.histogram(th, bins=edges)[0]
        got = np.zeros(bins, dtype=np.int64)
        for t in th:
            b = orc.theta_bin(t, dtheta, bins)
            if b >= 0:
                got[b] += 1
        assert np.array_equal(got, want)
    assert int(180 // 0.05) == 3599 and int(10 // 0.01) == 999 and int(2.5 // 1e-4) == 24999     # quirks Q1, Q2


def test_wrap_and_delta_pos_match_twin():
    rng = np.random.default_rng(3)
    pos0, cell, _ = random_box(9, 40, 2, True, 8.0)
    T = 6
    pos = pos0[None] + np.cumsum(rng.normal(scale=2.0, size=(T, 40, 3)), axis=0)
    cells = np.array([cell * (1 + 0.01 * k) for k in range(T)])
    got = orc.delta_pos(pos, cells)
    want = np.array(npo.delta_pos(list(pos), list(cells)))
    assert np.array_equal(got, want)
    f = np.linalg.solve(cells[0].T, got[1].T).T           # wrapped displacements lie in [-0.5, 0.5) fractional
    assert f.min() > -0.5 - 1e-6 and f.max() < 0.5


def test_window_msd_matches_reference_loop_and_random_walk():
    rng = np.random.default_rng(11)
    T, n, sigma = 60, 400, 0.1
    cell = np.eye(3) * 50.0
    pos = 25.0 + np.cumsum(rng.normal(scale=sigma, size=(T, n, 3)), axis=0)
    cells = np.broadcast_to(cell, (T, 3, 3)).copy()
    spec = np.zeros(n, dtype=np.uint8)
    masses = np.ones(n)
    window = np.arange(0, 30, 5)
    msd, mutated = orc.msd_window(pos, cells, masses, spec, 1, window)
    com = pos.mean(axis=1)
    delta = npo.delta_pos(list(pos - com[:, None, :]), list(cells))
    for w, m in enumerate(window):
        assert abs(msd[0, w] - npo.msd_of_m(delta, int(m))) <= 1e-12 * max(1.0, abs(msd[0, w]))
    assert msd[0, 0] == 0.0                                                           # Q4: MSD(0) is exactly 0
    expect = 3 * sigma ** 2 * window * (T - window - 1) / (T - window)               # random walk incl. the Q4 factor
    assert np.allclose(msd[0, 1:], expect[1:], rtol=0.15)
    assert np.allclose(mutated, pos - com[:, None, :], atol=1e-12)                    # Q7: frames translated by -COM


def test_direct_msd_free_particles():
    T, n = 20, 30
    rng = np.random.default_rng(5)
    cell = np.diag([10.0, 12.0, 14.0])
    true = rng.uniform(0, 10, size=(1, n, 3)) + np.cumsum(rng.normal(scale=0.4, size=(T, n, 3)), axis=0)
    wrapped = true - np.floor(true / np.diag(cell)) * np.diag(cell)
    cells = np.broadcast_to(cell, (T, 3, 3)).copy()
    spec = np.zeros(n, dtype=np.uint8)
    got = orc.msd_direct(wrapped, cells, spec, -1)
    want = ((true - true[0]) ** 2).sum(axis=2).mean(axis=1)
    assert np.allclose(got, want, rtol=1e-10, atol=1e-12)


def test_collinear_neighbours_are_clipped_not_dropped():
    """ase.geometry.get_angles clips the cosine to [-1, 1] before arccos ("we can get bad things like 1+2e-16"): exactly
    collinear neighbours whose normalised dot product rounds past +-1 land on 0 / 180 degrees instead of NaN."""
    cell = np.eye(3) * 30.0
    c = np.array([15.0, 15.0, 15.0])
    v0 = np.array([-1.8913201215548283, -0.39184909715672805, -0.2982765861550717])
    v1 = np.array([-1.5876314148962145, -0.3289300047383363, -0.2503824038621698])
    u0, u1 = v0 / math.sqrt((v0[0] * v0[0] + v0[1] * v0[1]) + v0[2] * v0[2]), v1 / math.sqrt((v1[0] * v1[0] + v1[1] * v1[1]) + v1[2] * v1[2])
    assert (u0[0] * u1[0] + u0[1] * u1[1]) + u0[2] * u1[2] > 1.0          # the case the clip exists for
    spec = np.array([0, 1, 1], dtype=np.uint8)
    cut = np.array([[0.0, 2.5], [2.5, 0.0]])
    for sign, want in ((1.0, 0.0), (-1.0, 180.0)):
        pos = np.vstack([c, c + v0, c + sign * v1])
        # the frame stores c + v, so the vectors are only nearly collinear; what matters is that nothing is dropped
        hist, dropped = orc.bad_hist(pos, cell, spec, 2, cut, 0, 1, 0.05, 3600)
        assert dropped == 0 and int(hist.sum()) == 1
        assert int(hist[2].argmax()) in ((0, 1) if sign > 0 else (3598, 3599))
        tw = npo.bad_angles(pos, cell, spec, cut, 0, 1)[2]
        assert abs(tw[0] - want) < 1e-5
    # the clipped arithmetic itself: x = 1 + 2^-52 and x = -(1 + 2^-52) give 0 and 180 degrees, never NaN
    big = np.nextafter(1.0, 2.0)
    for x, want in ((big, 0.0), (-big, 180.0)):
        th = orc.angle_of_cosine(x)
        assert th == want
