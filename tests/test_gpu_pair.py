"""GPU parity of the pair analysis (RDF histograms + cutoff neighbour counts) against the CPU oracle, through the
C ABI.  Counts must be bit-exact (north_star)."""
import json
import os

import numpy as np
import pytest

from conftest import random_box
from oracle import c_oracle as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "zif4_known_answers.json")


def _oracle_traj(pos, cell, spec, S, rmax, nbins, cut=None):
    hist = np.zeros((S, S, nbins), dtype=np.uint64)
    cn = []
    for f in range(len(pos)):
        if nbins:
            hist += orc.rdf_hist(pos[f], cell[f], spec, S, rmax, nbins)
        if cut is not None:
            cn.append(orc.cn_counts(pos[f], cell[f], spec, S, cut))
    return hist, (np.array(cn) if cut is not None else None)


def test_zif4_golden(backend, zif4):
    gold = json.load(open(GOLD))
    order = gold["species_order"]
    spec = np.array([order.index(int(z)) for z in zif4.numbers], dtype=np.uint8)
    rmax, bins = gold["rdf_default"]["rmax"], gold["rdf_default"]["bins"]
    cut = np.zeros((4, 4))
    cut[3, 2] = cut[2, 3] = 2.5
    res = backend.pair_counts(spec, 4, [(zif4.positions[None], zif4.cell[None])], rmax=rmax, nbins=bins, cn_cutoff=cut)
    assert int(res["hist"].sum()) == gold["rdf_default"]["directed_pairs_total"] == 30968
    assert [int(x) for x in res["hist"][3, 2]] == gold["rdf_default"]["hist_Zn_N"]
    assert [int(x) for x in res["hist"].sum(axis=(0, 1))] == gold["rdf_default"]["hist_total"]
    assert int(res["cn"][0, 3, 2]) == gold["cn_directed_pairs"]["Zn-N@2.5"] == 64
    assert int(res["cn"][0, 2, 3]) == gold["cn_directed_pairs"]["N-Zn@2.5"] == 64
    assert res["n_frames"] == 1
    assert abs(res["volume_sum"] - gold["volume"]) < 1e-9


@pytest.mark.parametrize("seed,n,tri,size,rmax,nbins", [
    (1, 300, False, 14.0, 6.0, 600),
    (2, 500, True, 15.0, 7.0, 333),
    (3, 64, True, 6.0, 9.5, 950),       # cutoff larger than the box: several images per pair, self images
    (4, 2000, True, 30.0, 10.0, 999),
    (5, 1, True, 5.0, 12.0, 100),       # a single atom only sees its own images
    (6, 7, False, 40.0, 3.0, 30),       # nearly empty box
])
def test_rdf_and_cn_random_boxes(backend, seed, n, tri, size, rmax, nbins):
    S = 3
    T = 3
    frames = [random_box(seed * 10 + f, n, S, tri, size, scale_pos=2.5) for f in range(T)]
    spec = frames[0][2]
    pos = np.array([f[0] for f in frames])
    cell = np.array([f[1] for f in frames])
    cut = np.array([[2.9, 3.3, 0.0], [3.3, 0.0, 4.1], [0.0, 4.1, 2.2]])
    res = backend.pair_counts(spec, S, [(pos, cell)], rmax=rmax, nbins=nbins, cn_cutoff=cut)
    hist, cn = _oracle_traj(pos, cell, spec, S, rmax, nbins, cut)
    assert np.array_equal(res["hist"], hist)
    assert np.array_equal(res["cn"], cn)
    # structural identities: symmetric in the species pair, same-species counts even
    assert np.array_equal(res["hist"], res["hist"].transpose(1, 0, 2))
    assert np.all(res["hist"][np.arange(S), np.arange(S)] % 2 == 0)


def test_rdf_only_and_cn_only_agree_with_combined(backend):
    pos, cell, spec = random_box(11, 800, 4, True, 20.0)
    cut = np.zeros((4, 4))
    cut[0, 1] = cut[1, 0] = 3.0
    cut[2, 2] = 2.5
    both = backend.pair_counts(spec, 4, [(pos[None], cell[None])], rmax=8.0, nbins=799, cn_cutoff=cut)
    rdf = backend.pair_counts(spec, 4, [(pos[None], cell[None])], rmax=8.0, nbins=799)
    cn = backend.pair_counts(spec, 4, [(pos[None], cell[None])], cn_cutoff=cut)
    assert np.array_equal(both["hist"], rdf["hist"])
    assert np.array_equal(both["cn"], cn["cn"])
    assert np.array_equal(cn["cn"][0], orc.cn_counts(pos, cell, spec, 4, cut))


def test_many_bins_global_histogram_path(backend):
    """dr = 1e-4 (amof.rdf.CoordinationNumber): 24 999 bins do not fit in shared memory."""
    pos, cell, spec = random_box(21, 400, 2, False, 16.0)
    nbins = int(2.5 // 1e-4)
    assert nbins == 24999
    res = backend.pair_counts(spec, 2, [(pos[None], cell[None])], rmax=2.5, nbins=nbins)
    assert np.array_equal(res["hist"], orc.rdf_hist(pos, cell, spec, 2, 2.5, nbins))


def test_streaming_batches_and_chunking(backend, monkeypatch):
    """Many small frames pushed in uneven chunks = one push; forces several batches per push."""
    monkeypatch.setenv("AMOFB_BATCH_ATOMS", "2000")
    T, n, S = 37, 300, 2
    rng = np.random.default_rng(5)
    base, cell0, spec = random_box(31, n, S, True, 13.0)
    pos = base[None] + rng.normal(scale=0.3, size=(T, n, 3))
    cell = np.array([cell0 * (1.0 + 0.002 * f) for f in range(T)])
    cut = np.full((S, S), 3.1)
    one = backend.pair_counts(spec, S, [(pos, cell)], rmax=6.0, nbins=600, cn_cutoff=cut)
    chunks = [(pos[0:5], cell[0:5]), (pos[5:6], cell[5:6]), (pos[6:30], cell[6:30]), (pos[30:], cell[30:])]
    many = backend.pair_counts(spec, S, chunks, rmax=6.0, nbins=600, cn_cutoff=cut)
    assert np.array_equal(one["hist"], many["hist"]) and np.array_equal(one["cn"], many["cn"])
    hist, cn = _oracle_traj(pos, cell, spec, S, 6.0, 600, cut)
    assert np.array_equal(one["hist"], hist) and np.array_equal(one["cn"], cn)
    assert one["n_frames"] == T
    assert abs(one["volume_sum"] - sum(abs(np.linalg.det(c)) for c in cell)) < 1e-6


def test_invariances(backend):
    """Permutation of the atom order and lattice-vector translations leave every count unchanged."""
    pos, cell, spec = random_box(41, 600, 3, True, 18.0)
    ref = backend.pair_counts(spec, 3, [(pos[None], cell[None])], rmax=7.5, nbins=750)["hist"]
    perm = np.random.default_rng(1).permutation(len(spec))
    again = backend.pair_counts(spec[perm], 3, [(pos[perm][None], cell[None])], rmax=7.5, nbins=750)["hist"]
    assert np.array_equal(ref, again)


def test_empty_and_errors(backend):
    pos, cell, spec = random_box(51, 50, 2, False, 10.0)
    res = backend.pair_counts(spec, 2, [], rmax=4.0, nbins=40)           # no frames at all
    assert res["n_frames"] == 0 and int(res["hist"].sum()) == 0
    with pytest.raises(ValueError):
        backend.pair_counts(spec, 2, [(pos[None], np.zeros((1, 3, 3)))], rmax=4.0, nbins=40)   # singular cell
    with pytest.raises(ValueError):
        backend.pair_counts(spec, 2, [(pos[None], cell[None])], rmax=-1.0, nbins=40)
    with pytest.raises(ValueError):
        backend.pair_counts(spec + 5, 2, [(pos[None], cell[None])], rmax=4.0, nbins=40)       # species out of range
    bad_cut = np.array([[0.0, 1.0], [2.0, 0.0]])
    with pytest.raises(ValueError):
        backend.pair_counts(spec, 2, [(pos[None], cell[None])], cn_cutoff=bad_cut)             # not symmetric
    # the context is still usable afterwards
    ok = backend.pair_counts(spec, 2, [(pos[None], cell[None])], rmax=4.0, nbins=40)
    assert np.array_equal(ok["hist"], orc.rdf_hist(pos, cell, spec, 2, 4.0, 40))


@pytest.mark.parametrize("env", [{"AMOFB_PAIR_GENERIC": "1"}, {"AMOFB_TILE_CAP": "256"}, {"AMOFB_TILE_CAP": "300", "AMOFB_CELL_DIV": "1"},
                                 {"AMOFB_CELL_DIV": "3"}, {"AMOFB_CELL_DIV": "4"}, {"AMOFB_TILE_BLOCKS_PER_SM": "1"}, {"AMOFB_CN_NO_FILTER": "1"}])
def test_kernel_variants_agree(backend, monkeypatch, env):
    """Generic kernel, tiled kernel with tiny staging capacity (row-split tiles and 'hard' cells handed to the
    generic kernel), other cell sizes and occupancies: all must give the oracle's integers."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    S = 3
    # a dense blob inside a dilute box makes some home cells exceed a small staging capacity
    pos, cell, spec = random_box(61, 1500, S, True, 24.0)
    rng = np.random.default_rng(3)
    pos[:500] = pos[0] + rng.normal(scale=1.2, size=(500, 3))
    cut = np.array([[2.9, 3.3, 0.0], [3.3, 0.0, 4.1], [0.0, 4.1, 2.2]])
    res = backend.pair_counts(spec, S, [(pos[None], cell[None])], rmax=9.0, nbins=900, cn_cutoff=cut)
    assert np.array_equal(res["hist"], orc.rdf_hist(pos, cell, spec, S, 9.0, 900))
    assert np.array_equal(res["cn"][0], orc.cn_counts(pos, cell, spec, S, cut))


@pytest.mark.parametrize("S,nbins", [(1, 500), (16, 300), (16, 999), (7, 2000)])
def test_species_counts_and_histogram_placement(backend, S, nbins):
    """1 species; 16 species (136 folded pairs: the histogram leaves shared memory for the global-atomic path at 999
    bins, stays in shared memory at 300); 7 species with 2000 bins."""
    pos, cell, spec = random_box(70 + S, 900, S, True, 17.0)
    cut = np.zeros((S, S))
    cut[0, S - 1] = cut[S - 1, 0] = 3.0
    res = backend.pair_counts(spec, S, [(pos[None], cell[None])], rmax=8.0, nbins=nbins, cn_cutoff=cut)
    assert np.array_equal(res["hist"], orc.rdf_hist(pos, cell, spec, S, 8.0, nbins))
    assert np.array_equal(res["cn"][0], orc.cn_counts(pos, cell, spec, S, cut))
    with pytest.raises(ValueError):
        backend.pair_counts(np.zeros(10, dtype=np.uint8), 17, [(pos[None, :10], cell[None])], rmax=4.0, nbins=10)


def test_cutoffs_beyond_rmax(backend):
    """coordination cutoffs larger than rmax (the CN_WIDE kernel variant): pairs counted for CN but not binned"""
    pos, cell, spec = random_box(81, 1200, 3, True, 20.0)
    cut = np.array([[6.5, 0.0, 5.0], [0.0, 0.0, 7.25], [5.0, 7.25, 3.0]])
    res = backend.pair_counts(spec, 3, [(pos[None], cell[None])], rmax=4.0, nbins=400, cn_cutoff=cut)
    assert np.array_equal(res["hist"], orc.rdf_hist(pos, cell, spec, 3, 4.0, 400))
    assert np.array_equal(res["cn"][0], orc.cn_counts(pos, cell, spec, 3, cut))


def test_repeated_analyses_reuse_pooled_buffers(backend):
    """begin/finish cycles of different shapes on one context (buffers come back from the pool, sizes differ)"""
    for n, S, nb in [(200, 2, 100), (1500, 4, 999), (50, 1, 10), (1500, 4, 999), (700, 3, 50)]:
        pos, cell, spec = random_box(n, n, S, True, 14.0)
        res = backend.pair_counts(spec, S, [(pos[None], cell[None])], rmax=6.0, nbins=nb)
        assert np.array_equal(res["hist"], orc.rdf_hist(pos, cell, spec, S, 6.0, nb))


def test_bin_rule_option(backend):
    """AMOFB_OPT_RDF_BIN_RULE (pin U1 as a switch): bin = int(d * (nBins/rMax)) instead of int(d / (rMax/nBins)); on a lattice
    many distances sit exactly on bin edges, where the two rules can part -- the GPU must follow the oracle under both."""
    from amof_b200 import _lib
    cell = np.eye(3) * 8.0
    g = np.arange(4) * 2.0
    lattice = np.array([[x, y, z] for x in g for y in g for z in g])
    rng = np.random.default_rng(8)
    cases = [(lattice, np.zeros(len(lattice), dtype=np.uint8), 1, 3.9, 39), (lattice, (np.arange(len(lattice)) % 2).astype(np.uint8), 2, 7.3, 999)]
    pos, cell2, spec = random_box(5, 900, 3, True, 16.0)
    try:
        for rule in (1, 0):
            orc.set_conventions(rule, 0)
            backend.ctx.set_option(_lib.AMOFB_OPT_RDF_BIN_RULE, rule)
            for p, sp, S, rmax, nb in cases:
                res = backend.pair_counts(sp, S, [(p[None], cell[None])], rmax=rmax, nbins=nb)
                assert np.array_equal(res["hist"], orc.rdf_hist(p, cell, sp, S, rmax, nb)), (rule, rmax)
            res = backend.pair_counts(spec, 3, [(pos[None], cell2[None])], rmax=7.0, nbins=700)
            assert np.array_equal(res["hist"], orc.rdf_hist(pos, cell2, spec, 3, 7.0, 700))
        with pytest.raises(ValueError):
            backend.ctx.set_option(_lib.AMOFB_OPT_RDF_BIN_RULE, 7)
        with pytest.raises(ValueError):
            backend.ctx.set_option(99, 0)
    finally:
        orc.set_conventions(0, 0)
        backend.ctx.set_option(_lib.AMOFB_OPT_RDF_BIN_RULE, 0)
