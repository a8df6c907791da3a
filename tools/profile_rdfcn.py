import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amof_b200
from amof_b200 import synth
traj = synth.make_trajectory("c2", 40)
lst = [traj[k] for k in range(40)]
t0 = time.perf_counter()
o = amof_b200.rdf.CoordinationNumber.from_trajectory(lst, {'Zn-N': 2.5}, dr=1e-4)
print("rdf.CoordinationNumber: %.1f ms per frame" % ((time.perf_counter() - t0) / 40 * 1e3), o.data['Zn-N'][:3].tolist())
