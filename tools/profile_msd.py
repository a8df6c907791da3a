"""Time the MSD paths on a C5-shaped trajectory generated on the device (unwrapped random walk, orthorhombic box).
    python tools/profile_msd.py [atoms] [frames] [reps] [legacy]
Prints, per repetition, the device time of the ingest (mass sums + commit, or transposition + frame sums for the legacy path)
and of the window kernel, CUDA events on the library's compute stream."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amof_b200 import _lib, frames as fr, synth  # noqa: E402
from amof_b200.elements import atomic_masses  # noqa: E402

if os.environ.get("AMOFB_LIB"):          # a variant built by tools/build_variants.sh
    _lib._SO = os.path.abspath(os.environ["AMOFB_LIB"])

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
legacy = len(sys.argv) > 4 and sys.argv[4] == "legacy"
backend = _lib.get_backend()
ctx = backend.ctx
numbers, pos0, cell = synth.base_frame("c5")
numbers, pos0 = numbers[:n], pos0[:n]
zs, spec = fr.species_index(numbers)
masses = np.array([atomic_masses[z] for z in numbers])
cells = np.broadcast_to(cell, (T, 3, 3)).copy()
window = np.arange(0, T // 2, 100).astype(np.int32)
dev = torch.device("cuda", ctx.device)
base = torch.from_numpy(pos0).to(dev)
slab = max(1, min(T, 256, (3 << 29) // (24 * n)))
if slab >= 32:
    slab -= slab % 32
for rep in range(reps):
    g = torch.Generator(device=dev); g.manual_seed(20261023)
    cur = base.clone()
    t_in = t_win = 0.0
    with backend.msd_open(T, masses, spec, len(zs), cells) as s:
        begun = []

        def finish():
            a_, b_, keep = begun.pop(0)
            sums = s.slab_sums_wait(b_ - a_)
            s.slab_commit(sums[:, :3] / sums[:, 3:4])

        for a in range(0, T, slab):
            b = min(T, a + slab)
            inc = 0.05 * torch.randn((b - a, n, 3), generator=g, device=dev, dtype=torch.float64)
            inc[0] += cur
            blk = inc.cumsum(0)
            cur = blk[-1].clone()
            del inc
            torch.cuda.synchronize()
            ctx.timer_mark(0)
            if legacy:
                s.load(a, (blk.data_ptr(), b - a))
            else:
                s.slab_sums_begin(a, (blk.data_ptr(), b - a))
                begun.append((a, b, blk))
                if len(begun) == 2:
                    finish()
            ctx.timer_mark(1)
            ctx.sync()
            t_in += ctx.timer_elapsed(0, 1)
        ctx.timer_mark(0)
        while begun:
            finish()
        ctx.timer_mark(1)
        ctx.sync()
        t_in += ctx.timer_elapsed(0, 1)
        ctx.timer_mark(2)
        if legacy:
            sums = s.com_sums()
            s.set_com(sums[:, :3] / sums[:, 3:4])
        ctx.timer_mark(3)
        raw = s.window(window)
        ctx.timer_mark(4)
        ctx.sync()
        t_com, t_win = ctx.timer_elapsed(2, 3), ctx.timer_elapsed(3, 4)
    tot = t_in + t_com + t_win
    print("rep %d %s: ingest %.2f ms (%.0f GB/s of 48 B per atom.frame), com %.2f ms, window %.2f ms, total %.2f ms -> %.0f frames/s; msd[0][1] %.6f"
          % (rep, "legacy" if legacy else "stream", t_in, 48.0 * n * T / t_in / 1e6, t_com, t_win, tot, T / tot * 1e3,
             raw[0][1] / max(1, int((spec == 0).sum())) / (T - 100)))
