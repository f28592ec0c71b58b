"""CPU-side checks of the C-ABI boundary: the library builds for sm_100a, loads, and exports every symbol
that include/p24.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

from p24 import build as p24_build
from p24 import lib as p24_lib

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared():
    names = []
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        text = open(os.path.join(ROOT, "include", fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"\b(p24_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_library_builds_and_exports_every_declared_symbol():
    path = p24_build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _declared()
    assert "p24_simota_loss_batch" in declared and "p24_loss_finalize" in declared
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/p24.h but not exported"
    # the ctypes prototypes of the host side cover exactly the declared entry points
    assert sorted(p24_lib.exported_names()) == declared


def test_abi_version_error_strings_and_workspace_query():
    lib = p24_lib.load()
    assert lib.p24_abi_version() == p24_lib.ABI_VERSION == 4
    assert b"success" in lib.p24_error_string(0)
    assert b"workspace" in lib.p24_error_string(-2)
    n = lib.p24_workspace_bytes(20, 8400, 50)
    assert n > 0 and n % 256 == 0
    assert lib.p24_workspace_bytes(0, 8400, 50) == 0
    # argument validation happens before any CUDA call
    assert lib.p24_loss_finalize(None, None, None, None, None) == -1


def test_product_path_refuses_cpu_tensors():
    from p24 import synth
    from p24.losses import Loss_Function
    out = synth.make_head_outputs(1, 64, 80, seed=0)
    lab = synth.make_labels(1, 1, 2, 64, 80, seed=0)
    xs, ys, ss = synth.make_grids(64)
    with pytest.raises(p24_lib.P24Error):
        Loss_Function(80).forward((xs, ys, ss, out, []), lab)
