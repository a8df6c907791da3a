"""End-to-end Rdf+CN over a LIST of Atoms objects (the reference's trajectory type) vs the array-backed trajectory:
how much the per-frame Python packing costs."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amof_b200  # noqa: E402
from amof_b200 import synth  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
arr = synth.make_trajectory("c2", T)
lst = [arr[k] for k in range(T)]
sets = {'Zn-N': 2.5, 'C-N': 1.728, 'C-C': 1.752}
for name, traj in (("ArrayTrajectory", arr), ("list[Atoms]", lst)):
    for r in range(3):
        t0 = time.perf_counter()
        rdf, cn = amof_b200.rdf.rdf_and_cn(traj, sets, dr=0.01, rmax=10.0)
        dt = time.perf_counter() - t0
    print("%-16s %7.1f ms for %d frames -> %8.0f frames/s" % (name, dt * 1e3, T, T / dt))
