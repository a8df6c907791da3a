"""
amof_b200 -- B200-native (sm_100a) implementation of aMOF's frame-parallel structural analyses.

Same module and class names as the reference package ``amof``::

    amof_b200.rdf.Rdf / rdf.CoordinationNumber     (amof/rdf.py)
    amof_b200.cn.CoordinationNumber                (amof/cn.py)
    amof_b200.bad.Bad / bad.BadByCn                (amof/bad.py)
    amof_b200.msd.WindowMsd / msd.DirectMsd        (amof/msd.py)
    amof_b200.atom, amof_b200.trajectory, amof_b200.files.path

All counting runs in hand-written CUDA kernels behind the C ABI of include/amofb.h (libamofb.so, loaded with
ctypes on first use).  There is no CPU fallback: computing without the library or without a CUDA device raises.
"""
from . import asap_compat, atom, bad, cn, elements, files, msd, rdf, sq, synth, trajectory  # noqa: F401
from .atoms import Atoms, read_extxyz  # noqa: F401
from .frames import ArrayTrajectory  # noqa: F401

__version__ = "0.1.0"
