"""Import the UNMODIFIED reference (``/root/reference/yolox_24p``) in the build container.

Test infrastructure only.  ``/root/reference`` does not exist on the GPU box, so nothing
that runs there (``-m gpu`` tests, ``smoke()``, ``bench.py``) may import this module; it is
used by ``tests/tools/make_golden.py`` (fixture generation) and by the container-only
tests that check the oracle restatement bit-for-bit against the reference.

Two mechanical accommodations, no reference file is modified:

1. three dead imports are missing in this image (``matplotlib``, ``matplotlib.pyplot``,
   ``thop``: ``models/losses.py:1``, ``models/yolo_head_24p.py:5``,
   ``utils/model_utils.py:9``) -> empty stub modules are registered first;
2. ``Loss_Function.pts_in_poly`` hard-codes ``device='cuda:0'`` (``models/losses.py:561,566``)
   -> ``torch.zeros`` / ``torch.arange`` are wrapped while the reference runs so that a
   ``device='cuda:0'`` keyword is rewritten to the device of the run.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

REF_ROOT = "/root/reference/yolox_24p"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "losses.py"))


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


_loaded = None


def load_reference():
    """Returns ``(models_module, utils_module)`` of the untouched reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not present (only available in the build container)")
    try:
        import matplotlib  # noqa: F401
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = _stub("matplotlib", scale=None)
        plt = _stub("matplotlib.pyplot", axis=None)
        mpl.pyplot = plt
    try:
        import thop  # noqa: F401
    except Exception:
        _stub("thop", profile=None)
    # the reference imports top-level ``utils`` / ``models``; make sure ours do not shadow
    for name in ("utils", "models"):
        if name in sys.modules and not getattr(sys.modules[name], "__file__", "").startswith(REF_ROOT):
            raise RuntimeError(f"a foreign top-level module {name!r} is already imported")
    sys.path.insert(0, REF_ROOT)
    try:
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import models as ref_models  # type: ignore
            import utils as ref_utils  # type: ignore
    finally:
        sys.path.remove(REF_ROOT)
    _loaded = (ref_models, ref_utils)
    return _loaded


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
