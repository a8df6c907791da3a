"""The C-ABI library: loads, exports every symbol include/amofb.h declares, refuses to run without a GPU."""
import ctypes
import os
import re

import pytest

from amof_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "amofb.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(amofb_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_in_tree():
    so = _lib.library_path()
    if not os.path.exists(so):
        build.build()
    assert os.path.exists(so) and os.path.dirname(so) == os.path.join(ROOT, "amof_b200")


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 40 and "amofb_pair_push" in names and "amofb_msd_window" in names
    lib = ctypes.CDLL(_lib.library_path())
    for n in names:
        assert hasattr(lib, n), "libamofb.so does not export %s" % n
    assert sorted(_lib.SIGNATURES) == names, "amof_b200/_lib.py SIGNATURES and include/amofb.h disagree"
    typed = _lib.load_library()
    assert typed.amofb_version().startswith(b"amofb")


def test_header_cites_the_reference_interfaces():
    text = open(HEADER).read()
    for cite in ("amof/rdf.py:87-93", "amof/atom.py:72-87", "amof/cn.py:58-74", "amof/bad.py:70-114", "amof/msd.py:186-268",
                 "amof/trajectory.py:285-303"):
        assert cite in text


def test_no_cpu_fallback():
    """Without a CUDA device the context cannot be created and the analysis classes raise (never compute on the CPU)."""
    lib = _lib.load_library()
    h = ctypes.c_void_p()
    rc = lib.amofb_create(0, ctypes.byref(h))
    if rc == 0:
        lib.amofb_destroy(h)
        pytest.skip("a CUDA device is present")
    assert rc < 0 and not h.value
    with pytest.raises(RuntimeError):
        _lib.Context(0)
    assert lib.amofb_sync(None) < 0 and lib.amofb_pair_push(None, 1, None, None) < 0


def test_package_does_not_import_the_oracle():
    import subprocess
    import sys
    code = "import sys; sys.path.insert(0, %r); import amof_b200, amof_b200._lib, amof_b200.synth; " \
           "assert not [m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]" % ROOT
    subprocess.check_call([sys.executable, "-c", code])
    for dirpath, _, files in os.walk(os.path.join(ROOT, "amof_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src, f
