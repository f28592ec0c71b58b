"""CPU tests of the host-side logic around the C ABI: level-table detection of the anchor grid (ABI v2+), the build
fingerprint, the sharding helpers.  No compute call is made (there is no GPU here and no CPU fallback)."""
import os
import shutil

import pytest
import torch

from p24 import build as p24_build
from p24 import synth
from p24.engine import GridCache
from p24.lib import P24Error


def _cat(lists):
    return [torch.cat([t.reshape(1, -1) for t in lst], 1).reshape(-1).float() for lst in lists]


def test_level_table_of_the_head_grid():
    """yolo_head_24p.py:222-230: three row-major level grids, strides 8 / 16 / 32."""
    for size, want in [(640, [(0, 80, 80), (6400, 40, 40), (8000, 20, 20)]),
                       (1280, [(0, 160, 160), (25600, 80, 80), (32000, 40, 40)]),
                       (320, [(0, 40, 40), (1600, 20, 20), (2000, 10, 10)])]:
        gx, gy, gs = _cat(synth.make_grids(size))
        arr, n = GridCache._levels(gx, gy, gs)
        assert n == 3
        got = [(arr[4 * l], arr[4 * l + 1], arr[4 * l + 2]) for l in range(n)]
        assert got == want


def test_rectangular_and_single_level_grids():
    # one level, 6 columns x 4 rows, stride 16
    ys, xs = torch.meshgrid(torch.arange(4), torch.arange(6), indexing="ij")
    gx, gy = xs.reshape(-1).float(), ys.reshape(-1).float()
    arr, n = GridCache._levels(gx, gy, torch.full((24,), 16.0))
    assert n == 1 and (arr[0], arr[1], arr[2]) == (0, 6, 4)


def test_irregular_grids_are_rejected_loudly():
    gx, gy, gs = _cat(synth.make_grids(320))
    bad = gx.clone()
    bad[5] = 7.0  # not a row-major grid any more
    with pytest.raises(P24Error):
        GridCache._levels(bad, gy, gs)
    perm = torch.randperm(gx.numel(), generator=torch.Generator().manual_seed(0))
    with pytest.raises(P24Error):
        GridCache._levels(gx[perm], gy[perm], gs[perm])
    # five levels: more than the kernels support
    gx5 = torch.zeros(5)
    with pytest.raises(P24Error):
        GridCache._levels(gx5, torch.zeros(5), torch.tensor([8.0, 16.0, 32.0, 64.0, 128.0]))


def test_build_fingerprint_does_not_depend_on_the_checkout_path(tmp_path, monkeypatch):
    """The built library travels with the tree (GPU box snapshot): a moved tree must still look fresh, otherwise every
    rank would rebuild it at import."""
    here = p24_build._fingerprint()
    csrc = tmp_path / "a" / "csrc"
    inc = tmp_path / "b" / "include"
    shutil.copytree(p24_build.CSRC, csrc)
    shutil.copytree(p24_build.INCLUDE, inc)
    monkeypatch.setattr(p24_build, "CSRC", str(csrc))
    monkeypatch.setattr(p24_build, "INCLUDE", str(inc))
    assert p24_build._fingerprint() == here
    with open(csrc / "p24_api.cu", "a") as fh:
        fh.write("\n// changed\n")
    assert p24_build._fingerprint() != here
