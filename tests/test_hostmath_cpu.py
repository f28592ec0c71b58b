"""CPU check of the formulas in csrc/p24_math.cuh (the `__host__ __device__` scalar arithmetic the kernels are built
from) against the oracle, through a host build of the header (tests/tools/hostmath.cpp, g++ -ffp-contract=off).
fp32 add / mul / div / sqrt round identically on both sides; the transcendental functions come from glibc here and from
ATen's vectorised kernels in the oracle, hence the 1e-5 relative tolerance of the north star (written below)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from oracle import p24_oracle as orc

RTOL = 1e-5
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def hm(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("hostmath") / "libhostmath.so")
    subprocess.run([gxx, "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++",
                    os.path.join(ROOT, "tests", "tools", "hostmath.cpp"), "-o", out, "-lm"], check=True)
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _rays(seed, n):
    g = torch.Generator().manual_seed(seed)
    rg = torch.rand(n, generator=g) * 150 + 1
    rp = torch.rand(n, generator=g) * 150 + 1
    d = torch.rand(n, generator=g) * 400
    m = torch.arange(n) % 5
    d = torch.where(m == 0, (rg + rp) * (1 + (torch.rand(n, generator=g) - 0.5) * 0.02), d)          # near tangent outside
    d = torch.where(m == 1, (rg - rp).abs() * (1 + (torch.rand(n, generator=g) - 0.5) * 0.02), d)    # near tangent inside
    return rg.float(), rp.float(), d.float()


def test_ray_loss_and_intersection_match_the_oracle(hm):
    rg, rp, d = _rays(0, 100_000)
    n = rg.numel()
    # oracle on pairs whose 24 rays are identical: GT = regular polygon of radius rg, prediction radii rp, distance d
    k = torch.arange(24, dtype=torch.float64) * (np.pi / 12)
    tgt = torch.zeros(n, 50)
    tgt[:, 2::2] = (rg.double()[:, None] * torch.cos(k)).float()
    tgt[:, 3::2] = (rg.double()[:, None] * torch.sin(k)).float()
    pred = torch.zeros(n, 26)
    pred[:, 0] = d
    pred[:, 2:] = rp[:, None]
    want, _ = orc.iou_loss_forward(pred, tgt)
    _, _, r_gt = orc._gt_radii(tgt)
    loss = np.empty((n, 24), dtype=np.float32)
    inter = np.empty((n, 24), dtype=np.float32)
    rgn = np.ascontiguousarray(r_gt.numpy().reshape(-1))
    rpn = np.ascontiguousarray(pred[:, 2:].numpy().reshape(-1))
    dn = np.ascontiguousarray(np.repeat(d.numpy(), 24))
    hm.hm_ray_loss(n * 24, _p(rgn), _p(rpn), _p(dn), _p(loss), _p(inter))
    # rays whose regime test sits within fp noise of the boundary may legitimately take the other branch
    edge = ((r_gt + pred[:, 2:] - d[:, None]).abs() < 1e-3 * (r_gt + pred[:, 2:])) | \
           (((r_gt - pred[:, 2:]).abs() - d[:, None]).abs() < 1e-3 * (r_gt + pred[:, 2:]))
    ok = ~edge.numpy()
    np.testing.assert_allclose(loss[ok], want.numpy()[ok], rtol=RTOL, atol=2e-6)
    winter, _ = orc.circle_inter_matched(tgt[:, 0], tgt[:, 1], r_gt, pred[:, 0], pred[:, 1], pred[:, 2:])
    np.testing.assert_allclose(inter[ok], winter.numpy()[ok], rtol=RTOL, atol=1e-2)


def test_pair_value_matches_bboxes_iou(hm):
    g = torch.Generator().manual_seed(3)
    G, P = 12, 400
    k = torch.arange(24, dtype=torch.float64) * (np.pi / 12)
    rg = torch.rand(G, 24, generator=g) * 100 + 5
    gt = torch.zeros(G, 50)
    gt[:, 0:2] = torch.rand(G, 2, generator=g) * 600 + 20
    gt[:, 2::2] = gt[:, 0:1] + (rg.double() * torch.cos(k)).float()
    gt[:, 3::2] = gt[:, 1:2] + (rg.double() * torch.sin(k)).float()
    pred = torch.zeros(P, 26)
    pred[:, 0:2] = torch.rand(P, 2, generator=g) * 640
    pred[:, 2:] = torch.rand(P, 24, generator=g) * 60 + 2
    want = orc.bboxes_iou(gt, pred)                       # [G, P]
    _, _, r_gt = orc._gt_radii(gt)
    dist = torch.sqrt((gt[:, None, 0] - pred[None, :, 0]) ** 2 + (gt[:, None, 1] - pred[None, :, 1]) ** 2)
    rgn = np.ascontiguousarray(r_gt[:, None, :].expand(G, P, 24).numpy().reshape(-1, 24))
    rpn = np.ascontiguousarray(pred[None, :, 2:].expand(G, P, 24).numpy().reshape(-1, 24))
    dn = np.ascontiguousarray(dist.numpy().reshape(-1))
    out = np.empty(G * P, dtype=np.float32)
    hm.hm_pair_value(G * P, _p(rgn), _p(rpn), _p(dn), _p(out))
    np.testing.assert_allclose(out.reshape(G, P), want.numpy(), rtol=RTOL, atol=1e-6)


def test_angle_sum_and_centre_window(hm):
    g = torch.Generator().manual_seed(5)
    k = torch.arange(24, dtype=torch.float64) * (np.pi / 12)
    r = (torch.rand(24, generator=g) * 60 + 30).double()
    vx = (300 + r * torch.cos(k)).float().numpy()
    vy = (280 + r * torch.sin(k)).float().numpy()
    x = (torch.rand(20000, generator=g) * 400 + 100).float()
    y = (torch.rand(20000, generator=g) * 400 + 80).float()
    out = np.empty(x.numel(), dtype=np.float32)
    hm.hm_angle_sum(x.numel(), _p(vx), _p(vy), _p(x.numpy()), _p(y.numpy()), _p(out))
    # the reference's sum (losses.py:566-588) in torch
    px, py = torch.from_numpy(vx), torch.from_numpy(vy)
    sx, sy = px[None, :] - x[:, None], py[None, :] - y[:, None]
    ex, ey = torch.roll(sx, -1, 1), torch.roll(sy, -1, 1)
    ang = torch.rad2deg(torch.atan2((sx * ey - ex * sy).abs(), sx * ex + sy * ey))
    want = ang[:, 0]
    for j in range(1, 24):
        want = want + ang[:, j]
    np.testing.assert_allclose(out, want.numpy(), rtol=RTOL, atol=1e-3)
    assert (out >= 350).sum() > 1000 and (out < 350).sum() > 1000
    # centre window: pure fp32 add / mul / compare -> bit-exact decisions, also on the boundary
    for stride in (8.0, 16.0, 32.0):
        W = int(640 // stride)
        xs = torch.randint(0, W, (50000,), generator=g).float()
        ys = torch.randint(0, W, (50000,), generator=g).float()
        gcx = ((xs + 0.5) * stride + (torch.randint(-3, 4, (50000,), generator=g).float() * 2.5 * stride / 3)).float()
        gcy = ((ys + 0.5) * stride + (torch.rand(50000, generator=g) - 0.5) * 6 * stride).float()
        res = np.empty(50000, dtype=np.int32)
        hm.hm_in_centre(50000, _p(gcx.numpy()), _p(gcy.numpy()), _p(xs.numpy()), _p(ys.numpy()), C.c_float(stride), _p(res))
        st = torch.tensor(stride)
        xc, yc = xs * st + 0.5 * st, ys * st + 0.5 * st
        rr = 2.5 * st
        d = torch.stack([xc - (gcx - rr), (gcx + rr) - xc, yc - (gcy - rr), (gcy + rr) - yc], 0).min(0).values
        assert np.array_equal(res.astype(bool), (d > 0.0).numpy())


def test_bce_and_cost(hm):
    g = torch.Generator().manual_seed(9)
    x = (torch.randn(50000, generator=g) * 6).float()
    t = torch.rand(50000, generator=g).float()
    out = np.empty(50000, dtype=np.float32)
    hm.hm_bce_logits(50000, _p(x.numpy()), _p(t.numpy()), _p(out))
    want = torch.nn.functional.binary_cross_entropy_with_logits(x, t, reduction="none")
    np.testing.assert_allclose(out, want.numpy(), rtol=RTOL, atol=1e-6)
    n, nc = 3000, 80
    cls = (torch.randn(n, nc, generator=g) * 2 - 3).float()
    obj = (torch.randn(n, generator=g) * 2 - 2).float()
    gcls = torch.randint(0, nc, (n,), generator=g).int()
    val = torch.rand(n, generator=g).float() * 0.98 + 0.01
    valid = (torch.rand(n, generator=g) > 0.3).int()
    cost = np.empty(n, dtype=np.float32)
    hm.hm_cost(n, nc, _p(cls.numpy()), _p(obj.numpy()), _p(gcls.numpy()), _p(val.numpy()), _p(valid.numpy()), _p(cost))
    p = torch.sqrt(torch.sigmoid(cls) * torch.sigmoid(obj)[:, None])                      # losses.py:409-413
    onehot = torch.nn.functional.one_hot(gcls.long(), nc).float()
    cls_cost = torch.nn.functional.binary_cross_entropy(p, onehot, reduction="none").sum(-1)
    want = cls_cost + 3.0 * (-torch.log(val + 1e-8)) + 100000.0 * (~valid.bool())          # losses.py:420-424
    np.testing.assert_allclose(cost, want.numpy(), rtol=RTOL, atol=1e-5)
