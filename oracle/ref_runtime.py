"""ORACLE side — test / baseline infrastructure, not product code.

Loads the UNMODIFIED reference implementation of the hot path (``models/losses.py`` and ``utils/boxes.py`` of
``/root/reference/yolox_24p``) from a directory tree and runs it on the CPU (or on a GPU for the same-device checks).

Two trees can be loaded:
  * ``/root/reference/yolox_24p``           the build container only (tests/tools/ref_loader.py, make_golden.py);
  * ``oracle/_ref/yolox_24p``               a git-ignored copy of exactly those two files, produced by the committed
                                            recipe ``oracle/make_ref.py`` (``__graft_entry__.build()`` runs it when
                                            /root/reference is present).  It travels to the GPU box with the snapshot,
                                            like the built ``.so``, so that ``bench.py --impl reference`` and the
                                            ``cpu_baseline`` leg time the reference's own code (``kind: "reference"``).

Mechanical accommodations (no reference file is modified): stub modules for imports that are dead on the
hot path and missing in this image (``matplotlib``, ``thop``), and a shim that rewrites the hard-coded
``device='cuda:0'`` of ``losses.py:561,566`` to the device of the run.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_COPY = os.path.join(HERE, "_ref", "yolox_24p")


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def available(root: str = REF_COPY) -> bool:
    return os.path.isfile(os.path.join(root, "models", "losses.py")) and \
        os.path.isfile(os.path.join(root, "utils", "boxes.py"))


_loaded = {}


def load(root: str = REF_COPY):
    """Returns ``(models_module, utils_module)`` of the reference tree at ``root``."""
    root = os.path.abspath(root)
    if root in _loaded:
        return _loaded[root]
    if not available(root):
        raise RuntimeError(f"no reference tree at {root}")
    for name, attrs in (("matplotlib", dict(scale=None)), ("matplotlib.pyplot", dict(axis=None)),
                        ("thop", dict(profile=None)), ("zmq", dict(device=None))):
        try:
            importlib.import_module(name)
        except Exception:
            _stub(name, **attrs)
    if "matplotlib.pyplot" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    # the reference imports top-level ``utils`` / ``models``
    for name in ("utils", "models"):
        mod = sys.modules.get(name)
        if mod is not None and not (getattr(mod, "__file__", "") or "").startswith(root):
            raise RuntimeError(f"a foreign top-level module {name!r} is already imported")
    sys.path.insert(0, root)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import models as ref_models  # type: ignore
            import utils as ref_utils  # type: ignore
    finally:
        sys.path.remove(root)
    _loaded[root] = (ref_models, ref_utils)
    return _loaded[root]


@contextlib.contextmanager
def cuda0_shim(device):
    """Rewrite the hard-coded ``device='cuda:0'`` of ``losses.py:561,566`` to ``device``."""
    import torch

    real_zeros, real_arange = torch.zeros, torch.arange

    def fix(kwargs):
        dev = kwargs.get("device", None)
        if isinstance(dev, str) and dev.startswith("cuda"):
            kwargs["device"] = device
        return kwargs

    def zeros(*a, **k):
        return real_zeros(*a, **fix(k))

    def arange(*a, **k):
        return real_arange(*a, **fix(k))

    torch.zeros, torch.arange = zeros, arange
    try:
        yield
    finally:
        torch.zeros, torch.arange = real_zeros, real_arange
