"""File -> result throughput: a C2 trajectory written as extended XYZ, analysed through amof_b200.stream.XyzStream
(chunks parsed by the native amofb_xyz_parse on all host cores into page-locked buffers while the GPU counts) against the same frames already in memory.
    python tools/profile_stream.py [frames] [threads]"""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amof_b200  # noqa: E402
from amof_b200 import stream, synth  # noqa: E402
from amof_b200.elements import chemical_symbols  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 400
threads = int(sys.argv[2]) if len(sys.argv) > 2 else None
traj = synth.make_trajectory("c2", T)
sym = np.array([chemical_symbols[z] for z in traj.numbers])
path = os.path.join(tempfile.gettempdir(), "c2_%d.xyz" % T)
t0 = time.perf_counter()
with open(path, "w") as fh:
    header = '%d\nLattice="%s" Properties=species:S:1:pos:R:3 pbc="T T T"\n' % (len(sym), " ".join(repr(float(x)) for x in traj.cells[0].ravel()))
    for k in range(T):
        fh.write(header)
        p = traj.positions[k]
        fh.write("\n".join("%s %.17g %.17g %.17g" % (s, x, y, z) for s, (x, y, z) in zip(sym, p.tolist())))
        fh.write("\n")
print("wrote %s: %.1f MB in %.1f s" % (path, os.path.getsize(path) / 1e6, time.perf_counter() - t0))
sets = {'Zn-N': 2.5, 'C-N': 1.728, 'C-C': 1.752}
want_r, want_c = amof_b200.rdf.rdf_and_cn(traj, sets, dr=0.01, rmax=10.0)
for rep in range(2):
    t0 = time.perf_counter()
    s = stream.XyzStream(path, threads=threads)
    t1 = time.perf_counter()
    r, c = amof_b200.rdf.rdf_and_cn(s, sets, dr=0.01, rmax=10.0)
    t2 = time.perf_counter()
    ok = np.array_equal(r.counts, want_r.counts) and np.array_equal(c.counts, want_c.counts)
    print("rep %d: open (frame offsets, headers) %.2f s, rdf_and_cn from the file %.2f s -> %.0f frames/s file to result (%.0f MB/s of text), identical to the in-memory result: %s"
          % (rep, t1 - t0, t2 - t1, T / (t2 - t0), os.path.getsize(path) / 1e6 / (t2 - t0), ok))
t0 = time.perf_counter()
amof_b200.rdf.rdf_and_cn(traj, sets, dr=0.01, rmax=10.0)
print("same frames from memory: %.0f frames/s" % (T / (time.perf_counter() - t0)))
os.remove(path)
