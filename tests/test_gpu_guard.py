"""Out-of-bounds writes: compute-sanitizer is refused on this GPU pool (profiles/r02_sanitizer_refused.log), so the library
carries its own check -- AMOFB_GUARD=1 puts canary bytes around every pooled device block and compares them when the block
is returned.  Every analysis runs once under it, in a fresh process (the switch is read at context creation), on shapes that
exercise the tiled and the generic pair kernel, the bond-angle pool and both MSD paths."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import numpy as np
import amof_b200
from amof_b200 import _lib, synth, atom as amatom
backend = _lib.get_backend()
assert backend.ctx.guard_violations() == 0
sets = {'Zn-N': 2.5, 'C-N': 1.728, 'C-C': 1.752}
for name, frames in (("c2", 5), ("c1", 9), ("c3", 2)):
    traj = synth.make_trajectory(name, frames)
    amof_b200.rdf.rdf_and_cn(traj, sets, dr=0.01, rmax=10.0 if name != "c1" else 6.0)
    amof_b200.rdf.Rdf.from_trajectory(traj, dr=1e-3, rmax=5.0)            # histograms beyond shared memory: generic kernel
    amof_b200.cn.CoordinationNumber.from_trajectory(traj, sets)
    amof_b200.bad.Bad.from_trajectory(traj, {'Zn-N': 2.5})
    amof_b200.bad.BadByCn.from_trajectory(traj, {'Zn-N': 2.5})
    amatom.get_neighborlist(traj[0], amatom.format_cutoff(sets))
walk = synth.make_trajectory("c1", 101)
amof_b200.msd.WindowMsd.from_trajectory(walk, delta_time=5, timestep=1, mutate=False)
amof_b200.msd.WindowMsd.from_trajectory(walk, delta_time=5, timestep=1, mutate=False, unwrap=True)
ortho = amof_b200.ArrayTrajectory(walk.numbers, walk.positions, np.diag(np.diag(walk.cells[0])))
amof_b200.msd.DirectMsd.from_trajectory(ortho)          # DirectMsd only takes orthogonal cells
amof_b200.msd.WindowMsd.from_trajectory(ortho, delta_time=5, timestep=1, mutate=False)
n = backend.ctx.guard_violations()           # every analysis has returned its blocks by now
print("guard violations:", n)
assert n == 0
'''


@pytest.mark.gpu
def test_no_out_of_bounds_writes_under_the_guard():
    env = dict(os.environ, AMOFB_GUARD="1", PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-c", SCRIPT], env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "guard violations: 0" in out.stdout and "amofb guard:" not in out.stderr, out.stdout + out.stderr
