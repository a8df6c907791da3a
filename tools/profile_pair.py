"""Small driver for ncu: a few launches of the pair kernel on C2-like frames already resident on the device.
    python tools/profile_pair.py [workload] [frames] [repeats]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amof_b200 import _lib, atom as amatom, frames as fr, synth  # noqa: E402

if os.environ.get("AMOFB_LIB"):          # a variant built by tools/build_variants.sh
    _lib._SO = os.path.abspath(os.environ["AMOFB_LIB"])
    import ctypes
    _old = ctypes.CDLL(_lib._SO)            # older variants may lack newer entry points: time what they have
    for _name in list(_lib.SIGNATURES):
        if not hasattr(_old, _name):
            _lib.SIGNATURES.pop(_name)

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 214
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
nbins = int(sys.argv[4]) if len(sys.argv) > 4 else 999        # fewer bins leave more shared memory to a tile's atoms
backend = _lib.get_backend()
ctx = backend.ctx
traj = synth.make_trajectory(name, T)
zs, spec = fr.species_index(traj.numbers)
cut = amatom.cutoff_matrix(amatom.format_cutoff({'Zn-N': 2.5, 'C-N': 1.728, 'C-C': 1.752}), zs)
dev = ctx.device_alloc(traj.positions.nbytes)
ctx.h2d(dev, traj.positions)
ctx.set_profiling(True)
for r in range(reps):
    t0 = time.perf_counter()
    res = backend.pair_counts(spec, len(zs), [(dev.value, traj.cells)], rmax=10.0, nbins=nbins, cn_cutoff=cut)
    dt = time.perf_counter() - t0
    ms, n = ctx.pair_kernel_time(reset=True)
    print("rep %d: %.1f ms wall, pair kernels %.3f ms over %d launches, %d pairs" % (r, dt * 1e3, ms, n, int(res["hist"].sum()) // 2))
