"""The drop-in classes on the GPU against the restated reference driver code (oracle/ref_classes.py): same columns,
same values (normalised outputs within 1e-12 relative), plus size-independent properties on a larger workload."""
import os
import numpy as np
import pytest

import amof_b200
from amof_b200 import frames, synth
from oracle import c_oracle as orc
from oracle import ref_classes as ref
from test_classes_cpu import assert_frames_equal, rattled, small_traj

pytestmark = pytest.mark.gpu


def test_backend_is_the_cuda_library(backend):
    assert backend.name.startswith("libamofb") and backend.ctx.launch_count() >= 0


def test_rdf_class(zif4, backend):
    assert_frames_equal(amof_b200.rdf.Rdf.from_trajectory([zif4]).data, ref.rdf_dataframe([zif4]))
    traj = small_traj(4)
    assert_frames_equal(amof_b200.rdf.Rdf.from_trajectory(traj, dr=0.02, rmax=6.5).data, ref.rdf_dataframe(traj, dr=0.02, rmax=6.5))


def test_cn_class(zif4, backend):
    sets = {"Zn-N": 2.5, "N-Zn": 2.5, "Zn-Zn": 7.0, "C-N": 1.728, "C-C": 1.752}
    c = amof_b200.cn.CoordinationNumber.from_trajectory([zif4], sets)
    assert list(c.data.iloc[0][1:]) == [4.0, 1.0, 4.0, 128 / 96, 64 / 96]
    traj = small_traj(4, sigma=0.15)
    assert_frames_equal(amof_b200.cn.CoordinationNumber.from_trajectory(traj, sets, delta_Step=10).data,
                        ref.cn_dataframe(traj, sets, delta_Step=10), rtol=0)


def test_rdf_and_cn_one_pass_equals_two(backend):
    traj = small_traj(3)
    sets = {"Zn-N": 2.5, "C-N": 1.728}
    r, c = amof_b200.rdf.rdf_and_cn(traj, sets, dr=0.05, rmax=7.0)
    assert r.data.equals(amof_b200.rdf.Rdf.from_trajectory(traj, dr=0.05, rmax=7.0).data)
    assert c.data.equals(amof_b200.cn.CoordinationNumber.from_trajectory(traj, sets).data)


def test_bad_classes(zif4, backend):
    assert_frames_equal(amof_b200.bad.Bad.from_trajectory([zif4], {"Zn-N": 2.5}).data, ref.bad_dataframe([zif4], {"Zn-N": 2.5}))
    traj = small_traj(3, sigma=0.1)
    sets = {"Zn-N": 2.5, "C-N": 1.728, "C-H": 1.3}
    got = amof_b200.bad.Bad.from_trajectory(traj, sets, dtheta=0.5).data
    want = ref.bad_dataframe(traj, sets, dtheta=0.5)
    assert sorted(got.columns) == sorted(want.columns)
    assert_frames_equal(got[list(want.columns)], want)
    by = amof_b200.bad.BadByCn.from_trajectory(traj, {"Zn-N": 2.5}, dtheta=1.0, normalization='partial')
    whole = amof_b200.bad.Bad.from_trajectory(traj, {"Zn-N": 2.5}, dtheta=1.0)
    np.testing.assert_allclose(sum(by.by_cn["N-Zn-N"].values()), whole.data["N-Zn-N"], rtol=1e-12, atol=1e-300)


def test_msd_classes(backend):
    want = ref.wmsd_dataframe(rattled(), delta_time=1, timestep=1)
    traj = rattled()
    got = amof_b200.msd.WindowMsd.from_trajectory(traj, delta_time=1, timestep=1)
    assert_frames_equal(got.data, want)
    assert np.abs(traj[3].get_center_of_mass()).max() < 1e-9                       # Q7
    wrapped = rattled(9)
    for a in wrapped:
        a.set_positions(a.get_positions() % np.diag(a.get_cell()))
    got = amof_b200.msd.WindowMsd.from_trajectory(wrapped, delta_time=1, timestep=1, unwrap=True, mutate=False)
    assert_frames_equal(got.data, ref.wmsd_dataframe(wrapped, delta_time=1, timestep=1, unwrap=True))
    ortho = [amof_b200.Atoms(numbers=a.numbers, positions=a.positions, cell=np.diag(np.diag(a.cell))) for a in rattled(8)]
    d = amof_b200.msd.DirectMsd.from_trajectory(ortho).data
    zs, spec = frames.species_index(ortho[0].numbers)
    pos = np.array([a.positions for a in ortho]); cells = np.array([a.cell for a in ortho])
    np.testing.assert_allclose(d["X"], orc.msd_direct(pos, cells, spec, -1), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(d["Zn"], orc.msd_direct(pos, cells, spec, zs.index(30)), rtol=1e-12, atol=1e-14)


# ---- full-size properties: what the oracle cannot check in seconds -------------------------------------------------
def test_c3_triclinic_frames_properties_and_oracle_sample(backend):
    """104 448-atom sheared box (config C3): histogram-sum identities on 6 frames, oracle agreement on 1."""
    traj = synth.make_trajectory("c3", 6)
    zs, spec = frames.species_index(traj.numbers)
    sets = {"Zn-N": 2.5, "C-N": 1.728, "C-C": 1.752}
    r, c = amof_b200.rdf.rdf_and_cn(traj, sets, dr=0.01, rmax=10.0)
    hist = r.counts
    assert hist.shape == (4, 4, 999) and r.n_frames == 6
    assert np.array_equal(hist, hist.transpose(1, 0, 2))                            # i->j pairs mirror j->i pairs
    assert np.all(hist[np.arange(4), np.arange(4)] % 2 == 0)
    one = amof_b200.rdf.Rdf.from_trajectory(traj[0:1], dr=0.01, rmax=10.0)
    rest = amof_b200.rdf.Rdf.from_trajectory(traj[1:6], dr=0.01, rmax=10.0)
    assert np.array_equal(one.counts + rest.counts, hist)                           # additivity over frames
    want, _ = orc.rdf_traj(traj.positions[:1], traj.cells[:1], spec, 4, 10.0, 999, threads=orc.max_threads())
    assert np.array_equal(one.counts, want)
    from amof_b200 import atom as amatom
    cut = amatom.cutoff_matrix(amatom.format_cutoff(sets), zs)
    assert np.array_equal(c.counts[:2], orc.cn_traj(traj.positions[:2], traj.cells[:2], spec, 4, cut, threads=orc.max_threads()))
    g = r.data["X-X"].to_numpy()
    assert abs(g[-200:].mean() - 1.0) < 0.1                                         # g(r) -> 1 at large r (ZIF order persists to 10 A)


def test_c2_translation_and_permutation_invariance(backend):
    traj = synth.make_trajectory("c2", 2)
    base = amof_b200.rdf.Rdf.from_trajectory(traj, dr=0.01, rmax=10.0).counts
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(traj.numbers))
    shuffled = frames.ArrayTrajectory(traj.numbers[perm], traj.positions[:, perm], traj.cells)
    assert np.array_equal(amof_b200.rdf.Rdf.from_trajectory(shuffled, dr=0.01, rmax=10.0).counts, base)


def test_c4_bad_counts_against_oracle(backend):
    traj = synth.make_trajectory("c4", 2)
    zs, spec = frames.species_index(traj.numbers)
    b = amof_b200.bad.Bad.from_trajectory(traj, {"Zn-N": 2.5})
    from amof_b200 import atom as amatom
    cut = amatom.cutoff_matrix(amatom.format_cutoff({"Zn-N": 2.5}), zs)
    want = np.zeros((33, 3600), dtype=np.uint64)
    for f in range(2):
        orc.bad_hist(traj.positions[f], traj.cells[f], spec, 4, cut, zs.index(30), zs.index(7), 0.05, 3600, hist=want)
    assert np.array_equal(b.counts["N-Zn-N"], want.sum(axis=0)) and int(want.sum()) > 20000
    red = synth.reduced_network("c4", 1)
    zb = amof_b200.bad.Bad.from_trajectory(red, {"Zn-Fr": 4.0}, dtheta=0.5)         # Zn-Im-Zn on the reduced network (Q6 path)
    assert "Zn-Fr-Zn" in zb.data.columns and "X-X-X" in zb.data.columns


def test_asap_shim_runs_the_reference_loop(backend):
    """The reference's own RDF loop (amof/rdf.py:87-114, restated in tests/ref_loops.py) on
    amof_b200.asap_compat.RadialDistributionFunction: the unmodified amof.rdf would count on the GPU through this class."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from amof_b200 import asap_compat, synth
    from amof_b200.elements import chemical_symbols
    from ref_loops import compute_rdf_with
    traj = synth.make_trajectory("c2", 5)
    frames_ = [traj[k] for k in range(len(traj))]                      # a list of Atoms, as aMOF holds a trajectory
    got = compute_rdf_with(asap_compat.RadialDistributionFunction, chemical_symbols, frames_, 0.01, 10.0)
    want = amof_b200.rdf.Rdf.from_trajectory(traj, dr=0.01, rmax=10.0)
    assert list(got.columns) == list(want.data.columns)
    for c in got.columns:
        np.testing.assert_allclose(got[c].to_numpy(), want.data[c].to_numpy(), rtol=1e-13, atol=0, err_msg=c)
    obj = asap_compat.RadialDistributionFunction(frames_[0], 10.0, 999)
    for a in frames_:
        obj.atoms = a
        obj.update()
    hist, zs = obj.get_counts()
    assert np.array_equal(hist, want.counts) and zs == want.species


def test_streamed_file_matches_in_memory(backend, tmp_path):
    """An extended-XYZ file analysed through amof_b200.stream.XyzStream (parser threads filling page-locked buffers ahead of
    the GPU) gives the integers of the same frames held in memory."""
    from amof_b200 import stream, synth
    from amof_b200.elements import chemical_symbols
    traj = synth.make_trajectory("c2", 9)
    sym = [chemical_symbols[z] for z in traj.numbers]
    path = str(tmp_path / "c2.xyz")
    with open(path, "w") as fh:
        for k in range(len(traj)):
            fh.write('%d\nLattice="%s" Properties=species:S:1:pos:R:3\n' % (len(sym), " ".join(repr(float(x)) for x in traj.cells[k].ravel())))
            fh.write("\n".join("%s %r %r %r" % (s, float(p[0]), float(p[1]), float(p[2])) for s, p in zip(sym, traj.positions[k])) + "\n")
    s = stream.XyzStream(path, chunk_bytes=1 << 20, threads=3)          # two frames per chunk
    sets = {'Zn-N': 2.5, 'C-N': 1.728}
    r1, c1 = amof_b200.rdf.rdf_and_cn(s, sets, dr=0.01, rmax=10.0)
    r2, c2 = amof_b200.rdf.rdf_and_cn(traj, sets, dr=0.01, rmax=10.0)
    assert np.array_equal(r1.counts, r2.counts) and np.array_equal(c1.counts, c2.counts)
    b1 = amof_b200.bad.Bad.from_trajectory(s, {'Zn-N': 2.5})
    b2 = amof_b200.bad.Bad.from_trajectory(traj, {'Zn-N': 2.5})
    assert np.array_equal(b1.counts["N-Zn-N"], b2.counts["N-Zn-N"])


def test_rdf_coordination_number_per_frame_take(backend):
    """amof.rdf.CoordinationNumber (rdf.py:135-214) integrates ONE histogram per frame: here a single analysis stays open and
    amofb_pair_take empties it after every frame.  Each frame's histogram must equal the one a separate analysis of that frame
    gives, on the default dr = 1e-4 (global-memory histograms) and on a coarse dr (shared-memory histograms)."""
    traj = small_traj(4, sigma=0.12)
    zs, spec = frames.species_index(traj[0].get_atomic_numbers())
    for rmax, bins in ((2.5, 24999), (6.0, 300)):
        chunks = [(t.get_positions()[None], np.asarray(t.get_cell())[None]) for t in traj]
        each = list(backend.pair_counts_each(spec, len(zs), iter(chunks), rmax, bins))
        assert len(each) == 4
        for k, res in enumerate(each):
            want, _ = orc.rdf_traj(chunks[k][0], chunks[k][1], spec, len(zs), rmax, bins)
            assert res["n_frames"] == 1 and np.array_equal(res["hist"], want)
            assert abs(res["volume_sum"] - abs(np.linalg.det(chunks[k][1][0]))) < 1e-12 * res["volume_sum"]
    sets = {"Zn-N": 2.5, "C-N": 1.728}
    got = amof_b200.rdf.CoordinationNumber.from_trajectory(traj, sets, dr=0.001, delta_Step=5).data
    assert list(got.columns) == ["Step", "Zn-N", "C-N"] and list(got["Step"]) == [0, 5, 10, 15]
    for k in range(4):
        one = amof_b200.rdf.CoordinationNumber.from_trajectory([traj[k]], sets, dr=0.001).data
        assert got["Zn-N"][k] == one["Zn-N"][0] and got["C-N"][k] == one["C-N"][0]
    # the analysis was closed: the next one opens normally
    amof_b200.rdf.Rdf.from_trajectory(traj, dr=0.05, rmax=5.0)
