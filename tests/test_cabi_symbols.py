"""CPU-only: the C-ABI library builds for sm_100a, loads, and exports every symbol include/gsb200.h
declares; the product has no CPU fallback and never touches oracle/."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gsb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gsb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import gsb200  # noqa: F401
    from gsb200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/gsb200.h but not exported"
    assert L.gsb_version() == 100
    # every symbol bound by the Python layer is declared in the header
    assert set(_lib._SIGS) <= set(names)


def test_library_is_sm100a_only():
    from gsb200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_and_no_oracle_in_product():
    import torch
    import gsb200  # noqa: F401
    from gsb200 import _lib
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            _lib.context()
    pkg = os.path.join(ROOT, "3dgs-native_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in src and "libgs_oracle" not in src and "gso_" not in src, fn
