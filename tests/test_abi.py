"""The drop-in boundary on a machine without a GPU: libpsc_b200.so loads, exports every
function include/psc_b200.h declares (and nothing declared is missing from the ctypes
binding the tests and bench.py use), and refuses to work without a device -- there is no
CPU fallback behind the C ABI."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "psc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(psc_b200_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_declares_the_operator_surface():
    names = _declared()
    for must in ("psc_b200_create", "psc_b200_destroy", "psc_b200_push_mprts", "psc_b200_sort",
                 "psc_b200_bnd_particles", "psc_b200_push_E", "psc_b200_push_H", "psc_b200_bnd_fill_ghosts",
                 "psc_b200_bnd_add_ghosts", "psc_b200_marder", "psc_b200_step", "psc_b200_balance",
                 "psc_b200_nccl_init", "psc_b200_moment_1st", "psc_b200_last_error"):
        assert must in names, must
    assert len(names) >= 45


def test_library_exports_every_declared_symbol():
    import psc_b200
    lib = C.CDLL(psc_b200._lib.lib_path())
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, "declared in include/psc_b200.h but not exported: %s" % missing
    lib.psc_b200_version.restype = C.c_char_p
    assert lib.psc_b200_version()


def test_no_cpu_fallback():
    """without a CUDA device a context cannot be created: error code + message, no silent
    host path (skipped where a GPU is present: the gpu-marked tests cover that side)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import psc_b200 as pb
    with pytest.raises(pb.PscB200Error):
        pb.Grid(gdims=(8, 8, 8), length=(8., 8., 8.), np=(1, 1, 1), dt=0.1, kinds=((-1., 1.),), nicell=1)
