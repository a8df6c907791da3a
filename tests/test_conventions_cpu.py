"""The pins that cannot be checked against asap3 / ase here are named switches:
each switch has a test of what it changes and of what it must leave alone.  Also: the asap3 object-protocol shim driven by a
line-for-line restatement of the reference's own loop, and the hand-typed BASELINE.md section 4 literals as a fixture that
does not come out of oracle/."""
import json
import os

import numpy as np
import pytest

import amof_b200
from amof_b200 import _lib, asap_compat, synth
from amof_b200 import rdf as amrdf
from amof_b200.atoms import Atoms
from amof_b200.elements import atomic_numbers, chemical_symbols
from oracle import c_oracle as orc
from oracle import np_oracle as npo
from oracle_backend import OracleBackend
from ref_loops import compute_rdf_with
from test_classes_cpu import small_traj

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(autouse=True)
def oracle_backend():
    old = _lib._set_backend_for_tests(OracleBackend())
    conv = dict(amrdf.CONVENTIONS)
    yield
    amrdf.CONVENTIONS.update(conv)
    orc.set_conventions(0, 0)
    _lib._set_backend_for_tests(old)


# ------------------------------------------------------------------------------------------------ BASELINE.md section 4
def test_baseline_md_literals(zif4):
    lit = json.load(open(os.path.join(HERE, "golden", "baseline_md_section4.json")))
    z = zif4.get_atomic_numbers()
    assert len(z) == lit["atoms"]
    for sym, n in lit["species_counts"].items():
        assert int((z == atomic_numbers[sym]).sum()) == n
    assert abs(zif4.get_volume() - lit["volume"]) < 1e-6
    assert abs(len(z) / zif4.get_volume() - lit["number_density"]) < 1e-7
    zs = sorted(set(int(v) for v in z))
    spec = np.array([zs.index(int(v)) for v in z], dtype=np.uint8)
    idx = {chemical_symbols[v]: k for k, v in enumerate(zs)}
    # coordination numbers through the public class (oracle backend) and raw counts through the C oracle
    for key, pairs in lit["cn_directed_pairs"].items():
        name, cut = key.split("@")
        a, b = name.split("-")
        m = np.zeros((len(zs), len(zs)))
        m[idx[a], idx[b]] = m[idx[b], idx[a]] = float(cut)
        counts = orc.cn_counts(zif4.positions, zif4.cell, spec, len(zs), m)
        assert int(counts[idx[a], idx[b]]) == pairs, key
        cn = amof_b200.cn.CoordinationNumber.from_trajectory([zif4], {name: float(cut)}).data[name][0]
        assert abs(cn - lit["cn_values"][key]) < 1e-12, key
    # distance shells (brute force over the 27 images, independent of the oracle's enumerations)
    shifts = np.array([[i, j, k] for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)]) @ zif4.cell
    for key, a, b in (("zn_n_shell", "Zn", "N"), ("zn_zn_shell", "Zn", "Zn")):
        pa, pb = zif4.positions[z == atomic_numbers[a]], zif4.positions[z == atomic_numbers[b]]
        d = np.linalg.norm(pb[None, :, None, :] + shifts[None, None, :, :] - pa[:, None, None, :], axis=3).ravel()
        d = np.sort(d[d > 1e-9])
        n = lit[key]["count"]
        assert abs(d[0] - lit[key]["min"]) < 5e-5 and abs(d[n - 1] - lit[key]["max"]) < 5e-5 and abs(d[n] - lit[key]["next"]) < 5e-5
    # bond angles
    m = np.zeros((len(zs), len(zs)))
    m[idx["Zn"], idx["N"]] = m[idx["N"], idx["Zn"]] = 2.5
    ang = np.array(orc.bad_angles(zif4.positions, zif4.cell, spec, len(zs), m, idx["Zn"], idx["N"]))
    b = lit["bad_N_Zn_N"]
    assert len(ang) == b["angles"] and abs(ang.min() - b["min"]) < 1e-3 and abs(ang.max() - b["max"]) < 1e-3 and abs(ang.mean() - b["mean"]) < 1e-3
    # RDF defaults
    r = amof_b200.rdf.Rdf.from_trajectory([zif4])
    rd = lit["rdf_default"]
    assert len(r.data) == rd["bins"] and int(r.counts.sum()) == rd["directed_pairs_total"]
    assert int(r.counts[idx["Zn"], idx["N"]].sum()) == rd["Zn_to_N"] and int(r.counts[idx["N"], idx["Zn"]].sum()) == rd["N_to_Zn"]
    assert int(r.counts[idx["Zn"], idx["Zn"]].sum()) == rd["Zn_to_Zn"]
    assert abs(amrdf._half_cell_rmax([zif4]) - rd["rmax"]) < 5e-5
    q = lit["binning_quirks"]
    assert int(10 // 0.01) == q["10//0.01"] and int(180 // 0.05) == q["180//0.05"] and int(2.5 // 1e-4) == q["2.5//1e-4"]


# ------------------------------------------------------------------------------------------------ named switches
def test_partial_norm_switch(zif4):
    """a3: 'sum_to_global' makes the partials add up to the global RDF; 'centre_species' makes A-X an RDF that is the
    N_a-weighted decomposition of it.  Counts are untouched."""
    a = amof_b200.rdf.Rdf.from_trajectory([zif4], dr=0.05)
    amrdf.set_conventions(partial_norm="sum_to_global")
    b = amof_b200.rdf.Rdf.from_trajectory([zif4], dr=0.05)
    assert np.array_equal(a.counts, b.counts)
    sym = [chemical_symbols[z] for z in set(zif4.get_atomic_numbers())]
    total = sum(b.data["%s-%s" % (x, y)].to_numpy() for x in sym for y in sym)
    np.testing.assert_allclose(total, b.data["X-X"].to_numpy(), rtol=1e-13)
    n_of = {s: int((zif4.get_atomic_numbers() == atomic_numbers[s]).sum()) for s in sym}
    for x in sym:
        for y in sym:
            np.testing.assert_allclose(b.data["%s-%s" % (x, y)].to_numpy() * 272 / n_of[x], a.data["%s-%s" % (x, y)].to_numpy(), rtol=1e-13)
    np.testing.assert_allclose(a.data["X-X"].to_numpy(), b.data["X-X"].to_numpy(), rtol=0)


def test_shell_volume_switch(zif4):
    """U3: the two shell volumes differ by dr^2/12 relative to (i+1/2)^2 dr^2, and by nothing else."""
    a = amof_b200.rdf.Rdf.from_trajectory([zif4], dr=0.05).data["X-X"].to_numpy()
    amrdf.set_conventions(shell_volume="midpoint")
    b = amof_b200.rdf.Rdf.from_trajectory([zif4], dr=0.05).data["X-X"].to_numpy()
    i = np.arange(len(a), dtype=np.float64)
    ratio = ((i + 1.0) ** 3 - i ** 3) / (3.0 * (i + 0.5) ** 2)          # exact / midpoint
    nz = a > 0
    np.testing.assert_allclose(b[nz] / a[nz], ratio[nz], rtol=1e-12)


def test_volume_switch():
    """U4: mean over the frames vs the first frame; identical for a constant cell."""
    traj = small_traj(3)                       # the cell grows by 0.2 % per frame
    a = amof_b200.rdf.Rdf.from_trajectory(traj, dr=0.05, rmax=5.0).data["X-X"].to_numpy()
    amrdf.set_conventions(volume="first")
    b = amof_b200.rdf.Rdf.from_trajectory(traj, dr=0.05, rmax=5.0).data["X-X"].to_numpy()
    vols = np.array([t.get_volume() for t in traj])
    nz = a > 0
    np.testing.assert_allclose(b[nz] / a[nz], vols[0] / vols.mean(), rtol=1e-12)
    same = [Atoms(numbers=t.numbers, positions=t.positions, cell=traj[0].cell) for t in traj]
    c = amof_b200.rdf.Rdf.from_trajectory(same, dr=0.05, rmax=5.0).data
    amrdf.set_conventions(volume="mean")
    d = amof_b200.rdf.Rdf.from_trajectory(same, dr=0.05, rmax=5.0).data
    np.testing.assert_allclose(c["X-X"].to_numpy(), d["X-X"].to_numpy(), rtol=1e-14)


def test_bin_rule_switch_in_the_oracle():
    """U1: d / (rMax/nBins) against d * (nBins/rMax).  The two rules give the same histogram up to pairs that sit within an ulp
    of a bin edge; on a lattice, where many distances ARE bin edges, they differ -- which is why it is a switch."""
    cell = np.eye(3) * 8.0
    g = np.arange(4) * 2.0
    pos = np.array([[x, y, z] for x in g for y in g for z in g])      # simple cubic, spacing 2.0: distances on bin edges for dr = 0.1
    spec = np.zeros(len(pos), dtype=np.uint8)
    a = orc.rdf_hist(pos, cell, spec, 1, 3.9, 39, method=0)
    orc.set_conventions(1, 0)
    b = orc.rdf_hist(pos, cell, spec, 1, 3.9, 39, method=0)
    orc.set_conventions(0, 0)
    assert int(a.sum()) == int(b.sum())                                # the same pairs ...
    rng = np.random.default_rng(0)
    pos = rng.uniform(0, 8.0, (300, 3))                                # ... and for generic positions the same bins
    a = orc.rdf_hist(pos, cell, np.zeros(300, dtype=np.uint8), 1, 3.9, 39, method=0)
    orc.set_conventions(1, 0)
    b = orc.rdf_hist(pos, cell, np.zeros(300, dtype=np.uint8), 1, 3.9, 39, method=0)
    assert np.array_equal(a, b)


def test_dv_rule_switch_in_the_oracle():
    """P2/P3 (wrap first) against ase's form on the positions as given: identical bits when every atom is inside the cell;
    for unwrapped positions the pairs are the same and only last-ulp distances move."""
    rng = np.random.default_rng(2)
    cell = np.array([[9.0, 0.0, 0.0], [1.0, 8.0, 0.0], [0.5, -0.7, 10.0]])
    frac = rng.random((120, 3))
    inside = frac @ cell
    spec = rng.integers(0, 2, 120).astype(np.uint8)
    cut = np.array([[3.0, 2.5], [2.5, 0.0]])
    a = orc.neighbour_pairs(inside, cell, spec, 2, cut, quantities=True, method=0)
    orc.set_conventions(0, 1)
    b = orc.neighbour_pairs(inside, cell, spec, 2, cut, quantities=True, method=0)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)                                     # inside the cell: bit-identical
    outside = (frac + rng.integers(-2, 3, (120, 3))) @ cell
    orc.set_conventions(0, 0)
    a = orc.neighbour_pairs(outside, cell, spec, 2, cut, quantities=True, method=0)
    orc.set_conventions(0, 1)
    b = orc.neighbour_pairs(outside, cell, spec, 2, cut, quantities=True, method=0)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])   # same pairs (nothing here sits within an ulp of a cutoff)
    np.testing.assert_allclose(a[2], b[2], rtol=1e-13)                  # distances agree to rounding


# ------------------------------------------------------------------------------------------------ asap3 protocol shim
def test_reference_loop_on_the_shim(zif4):
    """amof.rdf.Rdf.compute_rdf's own loop (rdf.py:87-114, restated in tests/ref_loops.py) driving
    amof_b200.asap_compat.RadialDistributionFunction gives the DataFrame of amof_b200.rdf.Rdf."""
    traj = small_traj(4)
    for dr, rmax in ((0.05, 'half_cell'), (0.02, 5.0)):
        got = compute_rdf_with(asap_compat.RadialDistributionFunction, chemical_symbols, traj, dr, rmax)
        want = amof_b200.rdf.Rdf.from_trajectory(traj, dr=dr, rmax=rmax).data
        assert list(got.columns) == list(want.columns)
        for c in got.columns:
            np.testing.assert_allclose(got[c].to_numpy(), want[c].to_numpy(), rtol=1e-13, atol=0, err_msg=c)


def test_shim_protocol_details(zif4):
    obj = asap_compat.RadialDistributionFunction(zif4, 6.0, 120)
    g0 = obj.get_rdf(elements=(30, 7), groups=0)            # never updated: the construction frame is counted (Q8)
    assert obj.countRDF == 1 and g0.shape == (120,)
    obj.update()                                            # the same frame again: counts double, g(r) stays
    g1 = obj.get_rdf(elements=(30, 7), groups=0)
    np.testing.assert_allclose(g1, g0, rtol=1e-14)
    hist, zs = obj.get_counts()
    assert zs == [1, 6, 7, 30] and int(hist[3, 2].sum()) % 2 == 0
    assert np.all(obj.get_rdf(elements=(30, 8)) == 0.0)     # an element that is not there
    with pytest.raises(NotImplementedError):
        asap_compat.RadialDistributionFunction(zif4, 6.0, 120, groups=[[0, 1]])
    moved = Atoms(numbers=zif4.numbers[::-1], positions=zif4.positions[::-1], cell=zif4.cell)
    obj.atoms = moved
    with pytest.raises(ValueError):
        obj.update()
