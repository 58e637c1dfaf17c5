"""The C-ABI library loads and exports every symbol include/vboc_b200.h declares (no compute calls
without a GPU), and fails loudly when no CUDA device is there."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    h = open(os.path.join(ROOT, "include", "vboc_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(vboc_[a-z0-9_]+)\s*\(", h)))


def test_header_symbols_are_exported():
    from vboc_b200 import _lib
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), n
    assert set(_lib.EXPORTS) == set(names)
    assert b"sm_100a" in L.vboc_version()


def test_default_opts_match_reference_settings():
    """VBOC/triplependulum_class_vboc.py:129-141"""
    from vboc_b200 import engine
    o = engine.default_opts("vboc")
    assert (o.tol_stat, o.qp_tol_stat, o.qp_iter_max, o.max_iter) == (1e-3, 1e-3, 100, 1000)
    assert (o.alpha_reduction, o.alpha_min, o.levenberg_marquardt, o.globalization) == (0.3, 1e-2, 1e-5, 1)
    a = engine.default_opts("al")
    assert a.globalization == 0 and a.qp_iter_max == 50


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from vboc_b200 import engine
    from vboc_b200._lib import VbocError
    with pytest.raises(VbocError):
        engine.BatchSolver(3, "vboc", 4, 100)
    with pytest.raises(VbocError):  # the streaming engine has no host path either
        engine.StreamSolver(3, "vboc", 4, 100)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vboc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle's", "").replace("the oracle", "").lower() or \
                    "import oracle" not in src and "liboracle" not in src, f
