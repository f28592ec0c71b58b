"""pytest wiring: markers and import paths.

``-m "not gpu"`` runs here (no GPU): oracle vs golden fixtures, host logic, C-ABI export check.
``-m gpu`` runs on a B200: the parity tests proper, through the C-ABI library.
"""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "exploration-of-potential_b200"), os.path.join(ROOT, "tests", "tools"),
          os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "refonly: needs /root/reference (build container only)")
