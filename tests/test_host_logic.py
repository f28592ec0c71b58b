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


def test_raw_levels_container_and_host_side_head_mirror():
    """p24.engine.RawLevels / p24.head on the CPU: shape validation, the level table handed to the C
    ABI, the cached grids, and the unfused torch decode against the oracle's restatement of yolo_head_24p.py:212-256."""
    import struct
    from oracle import p24_oracle as orc
    from p24 import head as p24_head
    from p24.engine import RawLevels
    reg, obj, cls = synth.make_raw_levels(2, 160, 80, seed=3)
    raw = RawLevels(reg, obj, cls, strides=synth.STRIDES)
    assert (raw.batch, raw.num_classes, raw.num_anchors) == (2, 80, 400 + 100 + 25)
    lv, n = raw.level_table()
    f2i = lambda x: struct.unpack("<i", struct.pack("<f", x))[0]
    assert n == 3 and list(lv) == [0, 20, 20, f2i(8.0), 400, 10, 10, f2i(16.0), 500, 5, 5, f2i(32.0)]
    with pytest.raises(IndexError):
        RawLevels(reg, obj[:2], cls)
    with pytest.raises(IndexError):
        RawLevels([r[:, :25] for r in reg], obj, cls)
    with pytest.raises(IndexError):
        RawLevels(reg, obj, cls).level_table()          # no strides: only the loss entry can use it (grids carry them)
    with pytest.raises(P24Error):
        raw.planes()                                    # CPU tensors: no fallback
    xs, ys, ss, out, _ = p24_head.train_outputs(reg, obj, cls, synth.STRIDES, fused=False)
    ox, oy, os_, want = orc.head_decode_train(reg, obj, cls, list(synth.STRIDES))
    assert torch.equal(out, want)
    assert all(torch.equal(a, b) for a, b in zip(xs + ys + ss, ox + oy + os_))
    mx, my, ms = synth.make_grids(160)
    assert all(torch.equal(a, b) for a, b in zip(xs + ys + ss, mx + my + ms))
    assert p24_head.train_outputs(reg, obj, cls, synth.STRIDES)[0][0] is xs[0]     # grids cached per shape
    assert isinstance(p24_head.train_outputs(reg, obj, cls, synth.STRIDES)[3], RawLevels)
    assert torch.equal(p24_head.infer_outputs(reg, obj, cls, synth.STRIDES, fused=False),
                       orc.head_decode_infer(reg, obj, cls, list(synth.STRIDES)))


def test_label_packing_and_polar_decode_have_no_cpu_fallback():
    import numpy as np
    from p24 import boxes as p24_boxes
    from p24.data import TrainTransform
    with pytest.raises(P24Error):
        TrainTransform(50).pack([np.zeros((2, 51))], [(640, 640)], (640, 640), device="cpu")
    # the coefficients themselves are host constants: spiral (the reference's) and polar (the drawn polygon)
    sx, sy = p24_boxes.spiral_coefficients("cpu")
    px, py = p24_boxes.polar_coefficients("cpu")
    th = torch.arange(24) * torch.tensor(15 * np.pi / 180)
    assert torch.equal(sx, th * torch.cos(th)) and torch.equal(py, torch.sin(th)) and float(px[0]) == 1.0 and float(sy[0]) == 0.0
    rows = torch.zeros(2, 29)
    rows[:, 0], rows[:, 1], rows[:, 2:26] = 10.0, 20.0, 5.0
    poly = p24_boxes.decode_polygons(rows)
    assert poly.shape == (2, 24, 2) and torch.allclose(((poly - torch.tensor([10.0, 20.0])) ** 2).sum(-1).sqrt(), torch.tensor(5.0))
