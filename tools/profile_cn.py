"""Timing of the coordination-number-only path (amof.cn): generic pair kernel with a cutoff matrix, no histogram."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amof_b200 import _lib, atom as amatom, frames as fr, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
backend = _lib.get_backend()
ctx = backend.ctx
traj = synth.make_trajectory(name, T)
zs, spec = fr.species_index(traj.numbers)
cut = amatom.cutoff_matrix(amatom.format_cutoff({'Zn-N': 2.5, 'C-N': 1.728, 'C-C': 1.752}), zs)
dev = ctx.device_alloc(traj.positions.nbytes)
ctx.h2d(dev, traj.positions)
for r in range(3):
    ctx.sync()
    t0 = time.perf_counter()
    res = backend.pair_counts(spec, len(zs), [(dev.value, traj.cells)], cn_cutoff=cut)
    dt = time.perf_counter() - t0
    print("rep %d: %.2f ms for %d frames -> %.0f frames/s; Zn-N pairs frame 0: %d" % (r, dt * 1e3, T, T / dt, int(res["cn"][0, zs.index(30), zs.index(7)])))
