"""Host-side logic of the drop-in classes (argument handling, quirks, schemas, files, helpers) with the oracle
standing in for the GPU backend.  The GPU parity of the same classes is in test_gpu_classes.py."""
import numpy as np
import pandas as pd
import pytest

import amof_b200
from amof_b200 import _lib, atom as amatom, frames, synth
from amof_b200.atoms import Atoms
from oracle import ref_classes as ref
from oracle_backend import OracleBackend

RTOL = 1e-12


@pytest.fixture(autouse=True)
def oracle_backend():
    old = _lib._set_backend_for_tests(OracleBackend())
    yield
    _lib._set_backend_for_tests(old)


def small_traj(n_frames=3, seed=0, sigma=0.05):
    numbers, pos, cell = synth.zif4_unit()
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n_frames):
        out.append(Atoms(numbers=numbers, positions=pos + rng.normal(scale=sigma, size=pos.shape) * k, cell=cell * (1 + 0.002 * k)))
    return out


def assert_frames_equal(got, want, rtol=RTOL):
    assert list(got.columns) == list(want.columns)
    assert len(got) == len(want)
    for c in got.columns:
        np.testing.assert_allclose(got[c].to_numpy(dtype=float), want[c].to_numpy(dtype=float), rtol=rtol, atol=1e-300, err_msg=c)


# ------------------------------------------------------------------------------------------------ rdf
def test_rdf_schema_and_values(zif4):
    r = amof_b200.rdf.Rdf.from_trajectory([zif4])
    cols = list(r.data.columns)
    uniq = list(set(zif4.get_atomic_numbers()))
    sym = [amof_b200.elements.chemical_symbols[z] for z in uniq]
    assert cols == ["r", "X-X"] + ["%s-%s" % (a, b) for a in sym for b in sym] + ["%s-X" % a for a in sym]
    assert len(r.data) == 770 and r.data["r"][1] == 0.01                 # r labelled with the caller's dr (Q1)
    assert_frames_equal(r.data, ref.rdf_dataframe([zif4]))
    # A-X is the sum of the partials, X-X their atom-weighted mean
    np.testing.assert_allclose(r.data["Zn-X"], sum(r.data["Zn-" + s] for s in sym), rtol=1e-13)
    w = {s: (zif4.numbers == amof_b200.elements.atomic_numbers[s]).sum() / 272 for s in sym}
    np.testing.assert_allclose(r.data["X-X"], sum(w[s] * r.data[s + "-X"] for s in sym), rtol=1e-12)


def test_rdf_rmax_rules_and_multi_frame():
    traj = small_traj(3)
    a = amof_b200.rdf.Rdf.from_trajectory(traj, dr=0.02, rmax=5.0)
    assert len(a.data) == int(5.0 // 0.02) and a.n_frames == 3
    assert_frames_equal(a.data, ref.rdf_dataframe(traj, dr=0.02, rmax=5.0))
    b = amof_b200.rdf.Rdf.from_trajectory(traj, dr=0.02, rmax=50.0)       # clamped to half the smallest cell length
    c = amof_b200.rdf.Rdf.from_trajectory(traj, dr=0.02)
    assert_frames_equal(b.data, c.data)
    with pytest.raises(ValueError):
        amof_b200.rdf.Rdf.from_trajectory(traj, dr=100.0)


def test_rdf_ideal_gas_tends_to_one():
    rng = np.random.default_rng(1)
    cell = np.eye(3) * 20.0
    traj = [Atoms(numbers=[18] * 2000, positions=rng.uniform(0, 20, (2000, 3)), cell=cell) for _ in range(2)]
    g = amof_b200.rdf.Rdf.from_trajectory(traj, dr=0.5, rmax=9.0).data["X-X"].to_numpy()
    assert np.all(np.abs(g[4:] - 1.0) < 0.05)


def test_rdf_file_roundtrip(tmp_path, zif4):
    r = amof_b200.rdf.Rdf.from_trajectory([zif4], dr=0.05)
    r.write_to_file(tmp_path / "zif4")
    assert (tmp_path / "zif4.rdf").exists()
    back = amof_b200.rdf.Rdf.from_file(tmp_path / "zif4")
    assert np.allclose(back.data, r.data)                                  # the reference's own check (examples:79)
    again = amof_b200.rdf.Rdf.from_file(tmp_path / "zif4.rdf")
    assert list(again.data.columns) == list(r.data.columns)


def test_rdf_integrated_coordination_number(zif4):
    r = amof_b200.rdf.Rdf.from_trajectory([zif4], dr=0.001)
    rho = amatom.get_number_density(zif4)
    # Simpson weights alternate 4/3, 2/3 over a delta-like first peak: the reference warns that this class is "subjected
    # to numerical errors in the integration step" (rdf.py:139-145); the exact count is 4
    cn = r.get_coordination_number("Zn-N", 2.5, rho)
    assert 4.0 * 2 / 3 - 0.1 < cn < 4.0 * 4 / 3 + 0.1
    coarse = amof_b200.rdf.Rdf.from_trajectory([zif4], dr=0.05).get_coordination_number("Zn-N", 3.0, rho)
    assert 2.0 < coarse < 6.0
    obj = amof_b200.rdf.CoordinationNumber.from_trajectory([zif4, zif4], {"Zn-N": 2.5}, dr=0.001, delta_Step=2)
    assert list(obj.data.columns) == ["Step", "Zn-N"] and list(obj.data["Step"]) == [0, 2]
    assert 4.0 * 2 / 3 - 0.1 < obj.data["Zn-N"][0] < 4.0 * 4 / 3 + 0.1 and obj.data["Zn-N"][0] == obj.data["Zn-N"][1]


def test_simpson_rule_matches_known_integrals():
    from amof_b200.rdf import _simpson_avg
    x = np.linspace(0.0, 2.0, 11)
    assert abs(_simpson_avg(x ** 2, x) - 8.0 / 3.0) < 1e-12               # odd number of points: exact for quadratics
    x = np.linspace(0.0, 2.0, 10)
    assert abs(_simpson_avg(x ** 3, x) - 4.0) < 2e-2                      # even: average of first/last rules
    assert _simpson_avg([1.0], [0.0]) == 0.0


# ------------------------------------------------------------------------------------------------ cn
def test_cn_schema_values_and_steps(zif4):
    sets = {"Zn-N": 2.5, "N-Zn": 2.5, "Zn-Zn": 7.0, "C-N": 1.728, "C-C": 1.752}
    c = amof_b200.cn.CoordinationNumber.from_trajectory([zif4, zif4], sets, delta_Step=5, first_frame=10)
    assert list(c.data.columns) == ["Step"] + list(sets)
    assert list(c.data["Step"]) == [10, 15]
    assert list(c.data.iloc[0][1:]) == [4.0, 1.0, 4.0, 128 / 96, 64 / 96]
    assert_frames_equal(c.data, ref.cn_dataframe([zif4, zif4], sets, delta_Step=5, first_frame=10))


def test_cn_later_keys_overwrite_and_missing_species(zif4):
    c = amof_b200.cn.CoordinationNumber.from_trajectory([zif4], {"Zn-N": 2.5, "N-Zn": 1.0})
    assert c.data["Zn-N"][0] == 0.0 and c.data["N-Zn"][0] == 0.0          # ase applies the last cutoff to both orientations
    c = amof_b200.cn.CoordinationNumber.from_trajectory([zif4], {"Zn-O": 2.5, "O-Zn": 2.5})
    assert c.data["Zn-O"][0] == 0.0 and np.isnan(c.data["O-Zn"][0])


def test_cn_file_roundtrip(tmp_path, zif4):
    c = amof_b200.cn.CoordinationNumber.from_trajectory([zif4], {"Zn-N": 2.5})
    c.write_to_file(tmp_path / "x.cn")
    assert amof_b200.cn.CoordinationNumber.from_file(tmp_path / "x").data.equals(c.data)


# ------------------------------------------------------------------------------------------------ bad
def test_bad_schema_and_density(zif4):
    b = amof_b200.bad.Bad.from_trajectory([zif4], {"Zn-N": 2.5})
    assert list(b.data.columns) == ["theta", "N-Zn-N"]                     # Zn-N-Zn has no angle -> no column
    assert len(b.data) == 3600 and abs(b.data["theta"][0] - 0.025) < 1e-15  # Q2: 3600 bins, centres
    widths = np.diff(np.arange(3601) * 0.05)
    assert abs((b.data["N-Zn-N"] * widths).sum() - 1.0) < 1e-12
    peak = b.data["theta"][b.data["N-Zn-N"] > 0]
    assert peak.min() > 100 and peak.max() < 120                            # the example's xlim=(100, 120)
    assert_frames_equal(b.data, ref.bad_dataframe([zif4], {"Zn-N": 2.5}))


def test_bad_multi_frame_and_x_columns():
    traj = small_traj(3)
    sets = {"Zn-N": 2.5, "C-N": 1.728, "C-H": 1.3}                          # covers all four species -> X columns
    b = amof_b200.bad.Bad.from_trajectory(traj, sets, dtheta=0.5)
    want = ref.bad_dataframe(traj, sets, dtheta=0.5)
    assert "X-X-X" in b.data.columns and "X-Zn-X" in b.data.columns
    assert sorted(b.data.columns) == sorted(want.columns)
    assert_frames_equal(b.data[list(want.columns)], want)


def test_bad_by_cn(zif4):
    traj = small_traj(2, sigma=0.2)
    total = amof_b200.bad.BadByCn.from_trajectory(traj, {"Zn-N": 2.5}, dtheta=1.0)
    part = amof_b200.bad.BadByCn.from_trajectory(traj, {"Zn-N": 2.5}, dtheta=1.0, normalization='partial')
    whole = amof_b200.bad.Bad.from_trajectory(traj, {"Zn-N": 2.5}, dtheta=1.0)
    widths = np.diff(np.arange(182) * 1.0)
    for cn, dens in total.by_cn["N-Zn-N"].items():
        assert cn >= 2 and abs((dens * widths).sum() - 1.0) < 1e-12
    summed = sum(part.by_cn["N-Zn-N"].values())
    np.testing.assert_allclose(summed, whole.data["N-Zn-N"], rtol=1e-12, atol=1e-300)


def test_bad_by_cn_netcdf_roundtrip(tmp_path):
    """BadByCn.write_to_file / from_file (bad.py:303-309): a classic netCDF file with variable 'bad' over (atom_triple, cn,
    theta); written through xarray when it is there, through scipy.io.netcdf_file otherwise -- the file is the same."""
    from scipy.io import netcdf_file
    traj = small_traj(3, sigma=0.25)                        # enough disorder for several coordination numbers
    b = amof_b200.bad.BadByCn.from_trajectory(traj, {"Zn-N": 2.5, "C-N": 1.728}, dtheta=1.0, normalization='partial')
    assert len(b.by_cn) >= 2 and any(len(d) > 1 for d in b.by_cn.values())
    b.write_to_file(tmp_path / "bycn")
    assert (tmp_path / "bycn.bad").exists()
    with netcdf_file(str(tmp_path / "bycn.bad"), 'r', mmap=False) as nc:
        assert set(nc.variables) == {"atom_triple", "cn", "theta", "bad"}
        assert nc.variables["bad"].dimensions == ("atom_triple", "cn", "theta")
        assert nc.variables["bad"].shape[2] == len(b.theta)
    back = amof_b200.bad.BadByCn.from_file(tmp_path / "bycn")
    assert sorted(back.by_cn) == sorted(b.by_cn)
    np.testing.assert_array_equal(back.theta, b.theta)
    for name, d in b.by_cn.items():
        assert sorted(back.by_cn[name]) == sorted(d)
        for cn, dens in d.items():
            np.testing.assert_array_equal(back.by_cn[name][cn], dens)


def test_bad_file_roundtrip(tmp_path, zif4):
    b = amof_b200.bad.Bad.from_trajectory([zif4], {"Zn-N": 2.5}, dtheta=1.0)
    b.write_to_file(tmp_path / "a")
    assert (tmp_path / "a.bad").exists()
    assert amof_b200.bad.Bad.from_file(tmp_path / "a").data.equals(b.data)


# ------------------------------------------------------------------------------------------------ msd
def rattled(n_frames=11, seed=3):
    """the example's recipe: cumulative Gaussian rattle of ZIF-4 (examples/Compute structural properties.py:110-114)"""
    numbers, pos, cell = synth.zif4_unit()
    rng = np.random.default_rng(seed)
    out, cur = [], pos.copy()
    for _ in range(n_frames):
        out.append(Atoms(numbers=numbers, positions=cur.copy(), cell=cell))
        cur = cur + rng.normal(scale=0.5, size=cur.shape)
    return out


def test_window_msd_schema_values_and_mutation():
    traj = rattled()
    before = [a.get_positions() for a in traj]
    want = ref.wmsd_dataframe(rattled(), delta_time=1, timestep=1)
    m = amof_b200.msd.WindowMsd.from_trajectory(traj, delta_time=1, timestep=1)
    uniq = list(set(traj[0].get_atomic_numbers()))
    assert list(m.data.columns) == ["Time"] + [amof_b200.elements.chemical_symbols[z] for z in uniq] + ["X"]
    assert list(m.data["Time"]) == [0, 1, 2, 3, 4] and m.data["X"][0] == 0.0
    assert_frames_equal(m.data, want)
    for k, a in enumerate(traj):                                            # Q7: frames are translated by -COM in place
        masses = a.get_masses()
        com = (masses[:, None] * before[k]).sum(axis=0) / masses.sum()
        np.testing.assert_allclose(a.get_positions(), before[k] - com, atol=1e-10)
        assert np.abs(a.get_center_of_mass()).max() < 1e-10


def test_window_msd_window_rules_and_unwrap():
    traj = rattled(21)
    m = amof_b200.msd.WindowMsd.from_trajectory(traj, delta_time=4, max_time=1000, timestep=2, mutate=False)
    assert list(m.data["Time"]) == [0, 4, 8, 12, 16]                       # window = arange(0, 20 // 2, 2) * 2
    wrapped = rattled(9)
    for a in wrapped:
        a.set_positions(a.get_positions() % np.diag(a.get_cell()))
    got = amof_b200.msd.WindowMsd.from_trajectory(wrapped, delta_time=1, timestep=1, unwrap=True, mutate=False)
    assert_frames_equal(got.data, ref.wmsd_dataframe(wrapped, delta_time=1, timestep=1, unwrap=True))


def test_direct_msd_and_files(tmp_path):
    traj = rattled(8)
    ortho = [Atoms(numbers=a.numbers, positions=a.positions, cell=np.diag(np.diag(a.cell))) for a in traj]
    d = amof_b200.msd.DirectMsd.from_trajectory(ortho, delta_Step=2, first_frame=4)
    assert list(d.data.columns)[:2] == ["Step", "X"] and list(d.data["Step"][:3]) == [4, 6, 8]
    assert d.data["X"][0] == 0.0 and np.all(np.diff(d.data["X"]) > -1e-9)
    d.write_to_file(tmp_path / "m")
    assert amof_b200.msd.DirectMsd.from_file(tmp_path / "m.msd").data.equals(d.data)
    w = amof_b200.msd.WindowMsd.from_trajectory(traj, delta_time=1, timestep=1, mutate=False)
    w.write_to_file(tmp_path / "w")
    assert amof_b200.msd.WindowMsd.from_file(tmp_path / "w").data.equals(w.data)


# ------------------------------------------------------------------------------------------------ helpers
def test_helpers_match_the_reference_semantics(zif4):
    assert amatom.format_cutoff({"Zn-N": 2.5, "C-C": 1.7}) == {(30, 7): 2.5, (6, 6): 1.7}
    assert amatom.format_cutoff({"Zn-N": 2.5}, sort_pair=True) == {(7, 30): 2.5}
    m = amatom.cutoff_matrix({(30, 7): 2.5, (7, 30): 3.0, (8, 8): 9.0}, [1, 6, 7, 30])
    assert m[3, 2] == m[2, 3] == 3.0 and m.sum() == 6.0
    assert abs(amatom.get_number_density(zif4) - 0.0620936) < 1e-7
    assert amatom.select_species_positions(zif4, 30).shape == (16, 3)
    assert sorted(amatom.get_atomic_numbers_unique(zif4)) == [1, 6, 7, 30]
    cs = amof_b200.trajectory.construct_step
    assert list(cs(delta_Step=2, first_frame=1, number_of_frames=3)) == [1, 3, 5]
    assert list(cs(step=slice(0, 6, 2))) == [0, 2, 4] and list(cs(delta_Step=3, first_frame=0, last_frame=7)) == [0, 3, 6]
    assert list(cs(delta_Step=1, last_frame=5, number_of_frames=2)) == [3, 4]
    ap = amof_b200.files.path.append_suffix
    assert str(ap("a/b", "rdf")) == "a/b.rdf" and str(ap("a/b.rdf", ".rdf")) == "a/b.rdf" and str(ap("a/b.x", "rdf")) == "a/b.x.rdf"


def test_array_trajectory_and_chunking(zif4):
    traj = small_traj(5)
    arr = frames.ArrayTrajectory(traj[0].numbers, np.array([a.positions for a in traj]), np.array([a.cell for a in traj]))
    assert len(arr) == 5 and np.array_equal(arr[2].get_positions(), traj[2].positions) and len(arr[1:3]) == 2
    a = amof_b200.rdf.Rdf.from_trajectory(traj, dr=0.05).data
    b = amof_b200.rdf.Rdf.from_trajectory(arr, dr=0.05).data
    assert_frames_equal(a, b, rtol=0)
    chunks = list(frames.iter_chunks(traj, 1, 5, OracleBackend(), target_bytes=2 * 272 * 24))
    assert [len(c[0]) for c in chunks] == [2, 2] and np.array_equal(chunks[1][0][1], traj[4].positions)
    with pytest.raises(ValueError):
        bad = list(traj)
        bad[3] = Atoms(numbers=traj[0].numbers[::-1], positions=traj[0].positions, cell=traj[0].cell)
        amof_b200.rdf.Rdf.from_trajectory(bad, dr=0.05)


def test_extxyz_reader_roundtrip(tmp_path, zif4):
    p = tmp_path / "z.xyz"
    lat = " ".join(repr(float(x)) for x in zif4.cell.ravel())
    with open(p, "w") as fh:
        for _ in range(2):
            fh.write("%d\n" % len(zif4))
            fh.write('Lattice="%s" Properties=species:S:1:pos:R:3:occ:R:1\n' % lat)
            for s, r in zip(zif4.get_chemical_symbols(), zif4.positions):
                fh.write("%s %r %r %r 1.0\n" % (s, float(r[0]), float(r[1]), float(r[2])))
    fr = amof_b200.read_extxyz(p)
    assert len(fr) == 2 and np.array_equal(fr[1].positions, zif4.positions) and np.array_equal(fr[0].numbers, zif4.numbers)
    assert np.array_equal(fr[0].cell, zif4.cell)


def test_fast_extxyz_trajectory_reader(tmp_path):
    traj = small_traj(4)
    p = tmp_path / "t.xyz"
    with open(p, "w") as fh:
        for a in traj:
            lat = " ".join(repr(float(x)) for x in a.cell.ravel())
            fh.write("%d\n" % len(a))
            fh.write('Lattice="%s" Properties=species:S:1:pos:R:3:occ:R:1 pbc="T T T"\n' % lat)
            for s, r in zip(a.get_chemical_symbols(), a.positions):
                fh.write("%-2s %r %r %r 1.0\n" % (s, float(r[0]), float(r[1]), float(r[2])))
    arr = amof_b200.trajectory.read_extxyz_trajectory(p)
    assert len(arr) == 4 and np.array_equal(arr.numbers, traj[0].numbers)
    assert np.array_equal(arr.positions, np.array([a.positions for a in traj]))
    assert np.array_equal(arr.cells, np.array([a.cell for a in traj]))
    slow = amof_b200.read_extxyz(p)
    assert np.array_equal(slow[3].positions, arr.positions[3])
    assert_frames_equal(amof_b200.rdf.Rdf.from_trajectory(arr, dr=0.05).data, amof_b200.rdf.Rdf.from_trajectory(traj, dr=0.05).data, rtol=0)


def test_synthetic_configs_are_deterministic():
    a = synth.make_trajectory("c2", 3)
    b = synth.make_trajectory("c2", 3)
    assert np.array_equal(a.positions, b.positions) and a.positions.shape == (3, 9792, 3)
    f = a.positions[2] @ np.linalg.inv(a.cells[0])
    assert f.min() >= 0 and f.max() < 1 and np.abs(a.positions[1] - a.positions[0]).max() > 0
    n3, _, c3 = synth.base_frame("c3")
    assert len(n3) == 104448 and abs(c3[1, 0]) > 1 and abs(c3[2, 1]) > 1                 # sheared: triclinic
    r = synth.reduced_network("c4", 1)
    assert r.positions.shape == (1, 8640, 3)


# ------------------------------------------------------------------------------------------------ neighbour list
def test_get_neighborlist_host_logic(zif4):
    """amof.atom.get_neighborlist (atom.py:72-87): a list with one list of neighbour indices per atom, built from the
    cutoff dictionary in both key orders; the CSR form and the optional ase quantities 'd' and 'S' describe the same pairs."""
    from amof_b200 import atom as amatom
    cut = amatom.format_cutoff({'Zn-N': 2.5, 'C-N': 1.728})
    nl = amatom.get_neighborlist(zif4, cut)
    numbers = np.asarray(zif4.get_atomic_numbers())
    assert isinstance(nl, list) and len(nl) == 272 and all(isinstance(r, list) for r in nl)
    assert all(len(nl[i]) == 4 and all(numbers[j] == 7 for j in nl[i]) for i in np.where(numbers == 30)[0])
    assert all(len(nl[i]) == 0 for i in np.where(numbers == 1)[0])                    # H is in no cutoff pair
    assert all(i in nl[j] for i in range(272) for j in nl[i])                          # symmetric
    assert all(r == sorted(r) for r in nl)
    off, nbr, dist, shifts = amatom.get_neighborlist_csr(zif4, cut, quantities=True)
    assert off[-1] == len(nbr) == sum(len(r) for r in nl) == 128 + 256
    owner = np.repeat(np.arange(272), np.diff(off))
    pos, cell = zif4.get_positions(), np.asarray(zif4.get_cell())
    np.testing.assert_allclose(np.linalg.norm(pos[nbr] - pos[owner] + shifts @ cell, axis=1), dist, rtol=0, atol=1e-12)
    lim = np.where((numbers[owner] == 30) | (numbers[nbr] == 30), 2.5, 1.728)
    assert np.all(dist < lim)


# ------------------------------------------------------------------------------------------------ structure factor
def test_structure_factor_ideal_gas_and_fcc():
    """S(q) from the partial histograms (amof_b200.sq): an ideal gas has S = 1 within the counting noise; an fcc crystal with
    a little thermal noise has Bragg peaks at 2 pi / a * sqrt(3), sqrt(8), sqrt(11) ((200) is a shoulder of (111) at the
    resolution rmax = 12 A gives)."""
    from amof_b200 import sq
    rng = np.random.default_rng(4)
    cell = np.eye(3) * 24.0
    gas = [Atoms(numbers=[18] * 3000, positions=rng.uniform(0, 24, (3000, 3)), cell=cell) for _ in range(3)]
    s = sq.StructureFactor.from_trajectory(gas, dr=0.05, rmax=11.0, q=np.arange(1.0, 10.0, 0.1))
    assert list(s.data.columns) == ["q", "X-X", "Ar-Ar"]
    assert np.all(np.abs(s.data["X-X"].to_numpy() - 1.0) < 0.15)
    np.testing.assert_allclose(s.data["X-X"].to_numpy(), s.data["Ar-Ar"].to_numpy(), rtol=1e-12)
    a = 4.0
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]]) * a
    cells_ = np.array([[i, j, k] for i in range(6) for j in range(6) for k in range(6)]) * a
    pos = (cells_[:, None, :] + base[None, :, :]).reshape(-1, 3)
    numbers = np.where(np.arange(len(pos)) % 2 == 0, 29, 47)            # two species on one lattice: partials share the peaks
    fcc = [Atoms(numbers=numbers, positions=pos + rng.normal(scale=0.05, size=pos.shape), cell=np.eye(3) * 6 * a) for _ in range(2)]
    q = np.arange(1.5, 6.0, 0.01)
    s = sq.StructureFactor.from_trajectory(fcc, dr=0.02, q=q)
    assert list(s.data.columns) == ["q", "X-X", "Cu-Cu", "Cu-Ag", "Ag-Ag"]
    tot = s.data["X-X"].to_numpy()
    peaks = [q[i] for i in range(1, len(q) - 1) if tot[i] > tot[i - 1] and tot[i] > tot[i + 1] and tot[i] > 1.5]
    want = 2 * np.pi / a * np.sqrt([3, 8, 11])
    assert len(peaks) >= 3
    for w in want:
        assert min(abs(p - w) for p in peaks) < 0.08, (w, peaks)
    c = s.concentrations
    mix = sum(c[x] * c[y] * s.data["%s-%s" % (amof_b200.elements.chemical_symbols[min(x, y)], amof_b200.elements.chemical_symbols[max(x, y)])].to_numpy()
              for x in c for y in c)
    np.testing.assert_allclose(mix, tot, rtol=1e-10, atol=1e-10)         # S = sum_ab c_a c_b S_ab


# ------------------------------------------------------------------------------------------------ streaming ingest
def _write_xyz(path, traj, extended=True, gz=False):
    import gzip
    sym = traj[0].get_chemical_symbols()
    lines = []
    for a in traj:
        lines.append("%d" % len(a))
        if extended:
            lines.append('Lattice="%s" Properties=species:S:1:pos:R:3 pbc="T T T"' % " ".join(repr(float(x)) for x in np.asarray(a.get_cell()).ravel()))
        else:
            lines.append(" i = 1, time = 0.5, E = -1.0")
        for s, p in zip(sym, a.get_positions()):
            lines.append("%s %r %r %r" % (s, float(p[0]), float(p[1]), float(p[2])))
    data = ("\n".join(lines) + "\n").encode()
    with (gzip.open(path, "wb") if gz else open(path, "wb")) as fh:
        fh.write(data)


def test_streaming_readers(tmp_path):
    """amof_b200.stream.XyzStream behind the reference's reader names: extended XYZ, xyz + cell array, CP2K xyz + .cell file,
    gzip; lazily indexed frames and streamed chunks are bit-identical to what was written, and the analyses give the same
    DataFrames from the stream as from the list of Atoms."""
    from amof_b200 import stream, trajectory as amtraj
    traj = small_traj(7)
    want_pos = np.array([t.get_positions() for t in traj])
    want_cell = np.array([np.asarray(t.get_cell()) for t in traj])
    p1 = str(tmp_path / "ext.xyz")
    _write_xyz(p1, traj, extended=True)
    s1 = amtraj.read_lammps_traj(p1)
    assert len(s1) == 7 and np.array_equal(s1.numbers, traj[0].get_atomic_numbers())
    assert np.array_equal(s1.cells, want_cell)
    assert np.array_equal(s1[3].get_positions(), want_pos[3]) and np.array_equal(s1[-1].get_positions(), want_pos[6])
    s1._chunk_frames = 2                                      # several chunks, several parser threads
    got = list(s1.stream_chunks(1, 6, None))
    assert [len(c[0]) for c in got] == [2, 2, 1]
    assert np.array_equal(np.concatenate([c[0] for c in got]), want_pos[1:6])
    assert np.array_equal(np.concatenate([c[1] for c in got]), want_cell[1:6])
    # CP2K: plain xyz + cell file (Step Time Ax..Cz Volume), gzipped, with a slice
    p2 = str(tmp_path / "cp2k-pos.xyz.gz")
    _write_xyz(p2, traj, extended=False, gz=True)
    pc = str(tmp_path / "cp2k.cell")
    with open(pc, "w") as fh:
        fh.write("#   Step   Time [fs]       Ax [Angstrom] ...\n")
        for k, c in enumerate(want_cell):
            fh.write("%8d %12.3f " % (k, 0.5 * k) + " ".join(repr(float(x)) for x in c.ravel()) + " %r\n" % float(abs(np.linalg.det(c))))
    s2 = amtraj.read_cp2k_traj(p2, pc, index=slice(1, 7, 2), unzip_xyz=True)
    assert len(s2) == 3 and np.array_equal(s2.cells, want_cell[1:7:2])
    assert np.array_equal(np.concatenate([c[0] for c in s2.stream_chunks(0, 3, None)]), want_pos[1:7:2])
    # xyz + one constant cell; a cell array shorter than the file trims the trajectory like Trajectory.set_cell(fit_size=True)
    p3 = str(tmp_path / "plain.xyz")
    _write_xyz(p3, traj, extended=False)
    s3 = amtraj.read_lammps_traj(p3, cell=want_cell[0])
    assert len(s3) == 7 and np.array_equal(s3.cells[5], want_cell[0])
    s4 = amtraj.read_lammps_traj(p3, cell=want_cell[:5])
    assert len(s4) == 5
    with pytest.raises(ValueError):
        amtraj.read_lammps_traj(p3)                           # no Lattice= and no cell
    # the analyses take the stream as they take the list
    a = amof_b200.rdf.Rdf.from_trajectory(s1, dr=0.05, rmax=5.0)
    b = amof_b200.rdf.Rdf.from_trajectory(traj, dr=0.05, rmax=5.0)
    assert_frames_equal(a.data, b.data)
    assert np.array_equal(a.counts, b.counts)
    assert_frames_equal(amof_b200.cn.CoordinationNumber.from_trajectory(s1, {"Zn-N": 2.5}).data,
                        amof_b200.cn.CoordinationNumber.from_trajectory(traj, {"Zn-N": 2.5}).data)
    assert_frames_equal(amof_b200.bad.Bad.from_trajectory(s1, {"Zn-N": 2.5}, dtheta=1.0).data,
                        amof_b200.bad.Bad.from_trajectory(traj, {"Zn-N": 2.5}, dtheta=1.0).data)
    assert_frames_equal(amof_b200.msd.WindowMsd.from_trajectory(s1, delta_time=1, timestep=1, mutate=False).data,
                        amof_b200.msd.WindowMsd.from_trajectory(traj, delta_time=1, timestep=1, mutate=False).data)


def test_native_xyz_parser_matches_python_float(tmp_path):
    """amofb_xyz_parse (host code of libamofb.so): every decimal string becomes the double Python's float() gives -- including
    17-digit values, halfway cases, subnormals, '+' signs and Fortran D exponents --, extended-XYZ column layouts are honoured,
    and malformed input is refused with the frame named."""
    from amof_b200 import _lib, stream
    rng = np.random.default_rng(11)
    tricky = ["0.1", "-0.30000000000000004", "1e23", "8.5e-324", "2.2250738585072011e-308", "9007199254740993", "+1.5",
              "1.0D+01", "-2.5d-3", "123456789012345678901234567890.5", "0.500000000000000166533453693773481063544750213623046875",
              "1.7976931348623157e308", ".5", "5.", "1E5", "-0.0"]
    n = 40
    vals = [repr(float(x)) for x in rng.standard_normal(3 * n * 3) * 10.0 ** rng.integers(-8, 9, 3 * n * 3)]
    vals[:len(tricky)] = tricky
    frames = []
    for f in range(3):
        lines = ["%d" % n, "frame %d" % f]
        for a in range(n):
            x, y, z = vals[(f * n + a) * 3:(f * n + a) * 3 + 3]
            lines.append(" %s\t%s   %s %s  extra" % ("Zn" if a % 2 else "N", x, y, z))
        frames.append("\n".join(lines) + "\n")
    text = "".join(frames).encode()
    off = np.cumsum([0] + [len(f) for f in frames])
    out = np.zeros((3, n, 3))
    sym = bytearray(8 * n)
    _lib.xyz_parse(text, off, n, 1, sym, False, out, threads=2)
    want = np.array([float(v.replace("D", "e").replace("d", "e")) for v in vals]).reshape(3, n, 3)
    assert np.array_equal(out.view(np.uint64), want.view(np.uint64))
    assert bytes(sym[:8]) == b"N\0\0\0\0\0\0\0" and bytes(sym[8:16]) == b"Zn\0\0\0\0\0\0"
    # a frame whose atom order changes, a truncated frame and a non-number are refused, with the frame named
    for bad_text in (text.replace(b"frame 2\n N", b"frame 2\n O"), text[:-40] + b"\n" * 0, text.replace(b"extra\n", b"\n").replace(b"0.1", b"0.1x")):
        with pytest.raises(ValueError, match="frame"):
            o2 = off.copy()
            o2[-1] = len(bad_text)
            _lib.xyz_parse(bad_text, o2, n, 1, bytearray(sym), True, np.zeros((3, n, 3)), threads=1)
    # extended XYZ with the positions after another column
    p = tmp_path / "cols.xyz"
    with open(p, "w") as fh:
        for f in range(2):
            fh.write('3\nLattice="5 0 0 0 5 0 0 0 5" Properties=species:S:1:tag:I:1:pos:R:3\n')
            for a, s in enumerate(("Zn", "N", "N")):
                fh.write("%s %d %r %r %r\n" % (s, a, 0.1 * a + f, 0.2, 0.3 * a))
    s = stream.XyzStream(str(p))
    assert s._pos_col == 2 and np.array_equal(s[1].get_positions(), [[1.0, 0.2, 0.0], [0.1 * 1 + 1, 0.2, 0.3], [0.1 * 2 + 1, 0.2, 0.3 * 2]])


def test_native_xyz_index_blocks_and_missing_final_newline(tmp_path):
    """amofb_xyz_index finds the frame starts across block boundaries exactly as a line count does, and a file whose last line
    has no newline still parses (the last frame ends at the end of the file)."""
    from amof_b200 import _lib, stream
    n = 5
    lines = []
    for f in range(7):
        lines += ["%d" % n, 'Lattice="6 0 0 0 6 0 0 0 6" Properties=species:S:1:pos:R:3'] + ["Zn %r %r %r" % (0.5 * a + f, 1.0, 2.0 + a) for a in range(n)]
    text = ("\n".join(lines)).encode()            # no final newline
    want = [0]
    for i, ch in enumerate(text):
        if ch == 10 and (text[:i + 1].count(b"\n")) % (n + 2) == 0:
            want.append(i + 1)
    # one block, and three blocks cut in the middle of lines
    starts, nl = _lib.xyz_index(text, 0, n + 2, 0)
    assert [0] + list(starts) == want and nl == text.count(b"\n")
    got, before, base = [0], 0, 0
    for a, b in ((0, 37), (37, 151), (151, len(text))):
        st, k = _lib.xyz_index(text[a:b], before, n + 2, base)
        got += list(st)
        before += k
        base += b - a
    assert got == want
    p = tmp_path / "nonl.xyz"
    p.write_bytes(text)
    s = stream.XyzStream(str(p))
    assert len(s) == 7 and np.array_equal(s[6].get_positions()[:, 0], 0.5 * np.arange(n) + 6)
    blocks = list(s.stream_chunks(0, 7, None))
    assert sum(len(b[0]) for b in blocks) == 7 and blocks[-1][0][-1][4][2] == 6.0


def test_stream_refuses_truncated_files_and_ignores_trailing_blank_lines(tmp_path):
    from amof_b200 import stream
    frame = '2\nLattice="5 0 0 0 5 0 0 0 5"\nZn 0 0 0\nN %d 1 1\n'
    p = tmp_path / "trunc.xyz"
    p.write_text(frame % 1 + '2\nLattice="5 0 0 0 5 0 0 0 5"\nZn 0 0 0\n')
    with pytest.raises(ValueError, match="truncated"):
        stream.XyzStream(str(p))
    q = tmp_path / "blank.xyz"
    q.write_text(frame % 1 + frame % 2 + "\n\n  \n")
    s = stream.XyzStream(str(q))
    assert len(s) == 2 and s[1].get_positions()[1, 0] == 2.0
