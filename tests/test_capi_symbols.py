"""CPU: the C-ABI library loads and exports every symbol include/*.h declares (no compute calls)."""
import ctypes
import glob
import os
import re

from conftest import ROOT


def declared_symbols():
    syms = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        txt = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        syms |= set(re.findall(r"\b(genpc_[a-z0-9_]+)\s*\(", txt))
    return sorted(syms)


def test_header_declares_something():
    assert "genpc_chamfer_forward" in declared_symbols()


def test_library_exports_every_declared_symbol():
    from genpc_b200.csrc import build

    so = build.build()
    L = ctypes.CDLL(so)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, f"declared in include/ but not exported: {missing}"


def test_python_binding_table_matches_header():
    from genpc_b200 import _lib

    assert sorted(_lib._SIGNATURES) == declared_symbols()
    L = _lib.lib()
    assert L.genpc_version().decode().startswith("genpc_b200")
    assert L.genpc_chamfer_workspace_bytes(2, 3, 5) == (2 * 3 + 2 * 5) * 8 + 16


def test_no_cpu_fallback():
    import pytest
    import torch

    from genpc_b200 import _lib
    from genpc_b200.loss_functions import chamfer_3DDist

    with pytest.raises(_lib.GenpcError):
        chamfer_3DDist()(torch.rand(1, 4, 3), torch.rand(1, 5, 3))
