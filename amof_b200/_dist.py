"""
Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests).

SURVEY.md 8(e): RDF / CN / BAD shard by frames and need ONE integer all-reduce of the histograms at the end
(bit-identical at any world size); MSD shards by atoms and all-reduces the per-frame centre-of-mass sums and the
final window sums.  Nothing here touches positions: there is no data-path collective.

``distributed`` argument of the analysis classes: ``None`` = automatic (shard iff torch.distributed is initialised
with world_size > 1), ``False`` = never, ``True`` = require an initialised process group.
"""
import numpy as np


def _td():
    try:
        import torch.distributed as td
    except Exception:           # torch is plumbing only; single-process use does not need it
        return None
    return td


def active(distributed=None):
    if distributed is False:
        return False
    td = _td()
    ok = td is not None and td.is_available() and td.is_initialized() and td.get_world_size() > 1
    if distributed is True and not ok:
        raise RuntimeError("distributed=True needs torch.distributed initialised with world_size > 1")
    return ok


def rank_world(distributed=None):
    if not active(distributed):
        return 0, 1
    td = _td()
    return td.get_rank(), td.get_world_size()


def _device_for_backend():
    import torch
    td = _td()
    if td.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def allreduce_sum(arr, distributed=None):
    """Element-wise sum over ranks of a numpy array (uint64 counts or float64 sums); every rank gets the result."""
    arr = np.ascontiguousarray(arr)
    if not active(distributed):
        return arr
    import torch
    td = _td()
    if arr.dtype == np.uint64:
        t = torch.from_numpy(arr.view(np.int64).copy())      # counts stay far below 2^63
    else:
        t = torch.from_numpy(arr.copy())
    dev = _device_for_backend()
    t = t.to(dev)
    td.all_reduce(t, op=td.ReduceOp.SUM)
    out = t.cpu().numpy()
    return out.view(np.uint64) if arr.dtype == np.uint64 else out


def allgather_rows(arr, counts, distributed=None):
    """Concatenate per-rank row blocks (rank r contributes ``counts[r]`` rows) in rank order."""
    arr = np.ascontiguousarray(arr)
    if not active(distributed):
        return arr
    total = int(sum(counts))
    rank, _ = rank_world(distributed)
    full = np.zeros((total,) + arr.shape[1:], dtype=arr.dtype)
    lo = int(sum(counts[:rank]))
    full[lo:lo + arr.shape[0]] = arr
    return allreduce_sum(full, distributed)     # disjoint blocks: the sum is the concatenation
