"""The N>1 path on real GPUs: two NCCL ranks (one process per GPU) shard the frames (RDF/CN/BAD) or the atoms (MSD) through the
public classes with distributed=True and must reproduce the single-GPU results -- identical integers for histograms and
coordination counts, 1e-12 for MSD.  Skipped on a box with one GPU."""
import os
import pickle
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _analyses(distributed):
    import amof_b200
    from amof_b200 import synth
    traj = synth.make_trajectory("c2", 7)
    walk = synth.make_trajectory("c1", 41)
    rdf, cn = amof_b200.rdf.rdf_and_cn(traj, {"Zn-N": 2.5, "C-N": 1.728}, dr=0.01, rmax=10.0, distributed=distributed)
    bad = amof_b200.bad.Bad.from_trajectory(traj, {"Zn-N": 2.5}, dtheta=0.05, distributed=distributed)
    msd = amof_b200.msd.WindowMsd.from_trajectory(walk, delta_time=2, timestep=1, mutate=False, distributed=distributed)
    msd_unwrap = amof_b200.msd.WindowMsd.from_trajectory(walk, delta_time=2, timestep=1, mutate=False, unwrap=True, distributed=distributed)
    return {"rdf_counts": rdf.counts, "rdf": rdf.data.to_numpy(), "cn_counts": cn.counts, "cn": cn.data.to_numpy(),
            "bad_counts": bad.counts["N-Zn-N"], "msd": msd.data.to_numpy(), "msd_unwrap": msd_unwrap.data.to_numpy()}


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["AMOFB_DEVICE"] = str(rank)
    import torch
    import torch.distributed as td
    torch.cuda.set_device(rank)
    td.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    res = _analyses(True)
    if rank == 0:
        with open(out_path, "wb") as fh:
            pickle.dump(res, fh)
    td.barrier()
    td.destroy_process_group()


def test_two_nccl_ranks_match_one_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "rank0.pkl")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = pickle.load(open(out, "rb"))
    want = _analyses(False)
    for k in ("rdf_counts", "cn_counts", "bad_counts"):
        assert np.array_equal(got[k], want[k]), k
    np.testing.assert_allclose(got["rdf"], want["rdf"], rtol=1e-14, atol=0)
    assert np.array_equal(got["cn"], want["cn"])
    np.testing.assert_allclose(got["msd"], want["msd"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(got["msd_unwrap"], want["msd_unwrap"], rtol=1e-12, atol=1e-13)
