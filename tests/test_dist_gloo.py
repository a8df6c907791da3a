"""The N>1 path on CPU: two gloo ranks shard the frames (RDF/CN/BAD) or the atoms (MSD) and must reproduce the
single-process results -- identical integers for histograms, 1e-12 for MSD.  The oracle stands in
for the GPU backend (tests/oracle_backend.py); the real multi-GPU run uses NCCL through the same code."""
import os
import pickle
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as td
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    import amof_b200
    from amof_b200 import _lib, frames
    from oracle_backend import OracleBackend
    from test_classes_cpu import rattled, small_traj
    _lib._set_backend_for_tests(OracleBackend())
    traj = small_traj(5)
    res = {"range": frames.frame_range(len(traj))}
    rdf, cn = amof_b200.rdf.rdf_and_cn(traj, {"Zn-N": 2.5, "C-N": 1.728}, dr=0.05, rmax=6.0)
    res["rdf_counts"], res["rdf"], res["cn"] = rdf.counts, rdf.data, cn.data
    res["cn_only"] = amof_b200.cn.CoordinationNumber.from_trajectory(traj, {"Zn-N": 2.5}).data
    bad = amof_b200.bad.Bad.from_trajectory(traj, {"Zn-N": 2.5}, dtheta=1.0)
    res["bad_counts"], res["bad"] = bad.counts["N-Zn-N"], bad.data
    res["msd"] = amof_b200.msd.WindowMsd.from_trajectory(rattled(9), delta_time=1, timestep=1, mutate=False).data
    if rank == 0:
        with open(out_path, "wb") as fh:
            pickle.dump(res, fh)
    td.barrier()
    td.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "rank0.pkl")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = pickle.load(open(out, "rb"))
    assert got["range"] == (0, 2)

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import amof_b200
    from amof_b200 import _lib
    from oracle_backend import OracleBackend
    from test_classes_cpu import rattled, small_traj
    old = _lib._set_backend_for_tests(OracleBackend())
    try:
        traj = small_traj(5)
        rdf, cn = amof_b200.rdf.rdf_and_cn(traj, {"Zn-N": 2.5, "C-N": 1.728}, dr=0.05, rmax=6.0)
        assert np.array_equal(got["rdf_counts"], rdf.counts)                    # integer all-reduce: identical
        assert got["rdf"].equals(rdf.data) and got["cn"].equals(cn.data)
        assert got["cn_only"].equals(amof_b200.cn.CoordinationNumber.from_trajectory(traj, {"Zn-N": 2.5}).data)
        bad = amof_b200.bad.Bad.from_trajectory(traj, {"Zn-N": 2.5}, dtheta=1.0)
        assert np.array_equal(got["bad_counts"], bad.counts["N-Zn-N"]) and got["bad"].equals(bad.data)
        msd = amof_b200.msd.WindowMsd.from_trajectory(rattled(9), delta_time=1, timestep=1, mutate=False).data
        assert list(msd.columns) == list(got["msd"].columns)
        np.testing.assert_allclose(got["msd"].to_numpy(), msd.to_numpy(), rtol=1e-12, atol=1e-14)   # atom-sharded fp64 sums
    finally:
        _lib._set_backend_for_tests(old)


def test_frame_and_atom_ranges_cover_everything():
    from amof_b200 import _dist
    assert _dist.rank_world(False) == (0, 1) and not _dist.active(None)
    for total in (1, 5, 16, 2000):
        for world in (1, 2, 3, 8):
            cuts = [(total * r) // world for r in range(world + 1)]
            assert cuts[0] == 0 and cuts[-1] == total and all(b >= a for a, b in zip(cuts, cuts[1:]))
    with pytest.raises(RuntimeError):
        _dist.active(True)
