"""Times the bond-angle analysis on C4 frames resident on the device (no CPU baseline, no e2e leg).
    python tools/profile_bad.py [frames] [repeats]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amof_b200 import _lib, atom as amatom, frames as fr, synth  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
backend = _lib.get_backend()
ctx = backend.ctx
traj = synth.make_trajectory("c4", T)
zs, spec = fr.species_index(traj.numbers)
cut = amatom.cutoff_matrix(amatom.format_cutoff({'Zn-N': 2.5}), zs)
triples = [(zs.index(30), zs.index(7)), (zs.index(7), zs.index(30))]
dev = ctx.device_alloc(traj.positions.nbytes)
ctx.h2d(dev, traj.positions)
for r in range(reps):
    ctx.sync()
    t0 = time.perf_counter()
    hist, dropped, nf = backend.bad_counts(spec, len(zs), [(dev.value, traj.cells)], cut, triples, 0.05, 3600)
    ctx.sync()
    dt = time.perf_counter() - t0
    print("rep %d: %.2f ms for %d frames = %.2f us/frame, %d angles" % (r, dt * 1e3, T, dt * 1e6 / T, int(hist.sum())))
