import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def zif4():
    """The reference's 272-atom ZIF-4 example frame as an amof_b200.Atoms."""
    from amof_b200 import synth
    from amof_b200.atoms import Atoms
    numbers, pos, cell = synth.zif4_unit()
    return Atoms(numbers=numbers, positions=pos, cell=cell)


@pytest.fixture(scope="session")
def backend():
    """The GPU backend (fails loudly when libamofb.so or the device is missing)."""
    from amof_b200 import _lib
    return _lib.get_backend()


def random_box(seed, n, nspec=3, triclinic=True, size=12.0, scale_pos=1.0):
    """Small random periodic box; positions deliberately spill outside the cell to exercise wrapping."""
    rng = np.random.default_rng(seed)
    cell = np.diag(size * (1.0 + 0.3 * rng.random(3)))
    if triclinic:
        cell[1, 0] = 0.25 * size * (rng.random() - 0.5)
        cell[2, 0] = 0.25 * size * (rng.random() - 0.5)
        cell[2, 1] = 0.25 * size * (rng.random() - 0.5)
    frac = rng.random((n, 3)) * scale_pos - 0.5 * (scale_pos - 1.0)
    pos = frac @ cell
    spec = rng.integers(0, nspec, size=n).astype(np.uint8)
    return pos, cell, spec
