"""CPU proof obligations of the decision-preserving bounds, checked against the oracle's ray arithmetic:

  loss(rg, rp, d) <= max(1, 2 - 4 (rg^2 + rp^2) / (rg + rp + d)^2)          (p24_ray_loss_ub: every regime)
  loss(rg, rp, d) <= max(1, 2 - 4 rg^2 / ((rg + d)^2 + rg^2))               (H*: whatever the predicted radius)
  loss(rg, rp, d) <= max(1, 2 - 4 rg^2 / (rg + t)^2)  for t >= rp + d        (H)
  |loss - (2 - 4 (rg^2 + rp^2) / (rg + rp + d)^2)| < 3e-6 * 2  when d >= rg + rp  (closed form of the apart regime)

The kernels use these to skip work; the radii follow the kernels' own precondition (>= 0.25 px: smaller ones disable
the bounds)."""
import numpy as np
import torch

from oracle import p24_oracle as orc


def _ray_loss(rg, rp, d):
    """Oracle loss of rays with GT radius rg, predicted radius rp, centre distance d (all [N])."""
    n = rg.numel()
    k = torch.arange(24, dtype=torch.float64) * (np.pi / 12)
    tgt = torch.zeros(n, 50, dtype=torch.float32)
    tgt[:, 2::2] = (rg.double()[:, None] * torch.cos(k)[None, :]).float()
    tgt[:, 3::2] = (rg.double()[:, None] * torch.sin(k)[None, :]).float()
    pred = torch.zeros(n, 26, dtype=torch.float32)
    pred[:, 0] = d
    pred[:, 2:] = rp[:, None]
    loss, _ = orc.iou_loss_forward(pred, tgt)
    _, _, r_gt = orc._gt_radii(tgt)
    return loss[:, 0].double(), r_gt[:, 0].double()


def _samples(seed, n):
    g = torch.Generator().manual_seed(seed)
    u = lambda lo, hi: torch.exp(torch.rand(n, generator=g) * (np.log(hi) - np.log(lo)) + np.log(lo))
    rg, rp = u(0.25, 400.0), u(0.25, 400.0)
    d = u(1e-3, 2000.0)
    # a third of the samples sit on the regime boundaries (tangent inside / outside) and at d ~ 0
    m = torch.arange(n) % 6
    d = torch.where(m == 0, (rg + rp) * (1 + (torch.rand(n, generator=g) - 0.5) * 1e-3), d)
    d = torch.where(m == 1, (rg - rp).abs() * (1 + (torch.rand(n, generator=g) - 0.5) * 1e-3) + 1e-6, d)
    return rg.float(), rp.float(), d.float()


def test_upper_bounds_hold_in_every_regime():
    worst = {"ub": -1.0, "hstar": -1.0, "h": -1.0}
    for seed in range(4):
        rg32, rp, d = _samples(seed, 200_000)
        loss, rg = _ray_loss(rg32, rp, d)
        rp64, d64 = rp.double(), d.double()
        ub = torch.clamp(2 - 4 * (rg**2 + rp64**2) / (rg + rp64 + d64) ** 2, min=1.0)
        hstar = torch.clamp(2 - 4 * rg**2 / ((rg + d64) ** 2 + rg**2), min=1.0)
        t = rp64 + d64
        h = torch.clamp(2 - 4 * rg**2 / (rg + t) ** 2, min=1.0)
        # the kernels add 2e-5 per VALUE (mean of 24 rays / 2), i.e. 4e-5 per ray loss, to the bounds they compare
        tol = 4e-5
        for name, b in (("ub", ub), ("hstar", hstar), ("h", h)):
            worst[name] = max(worst[name], float((loss - b).max()))
            assert float((loss - b).max()) < tol, (name, seed, float((loss - b).max()))
        assert bool((hstar + 1e-12 >= ub - 1e-9).all())  # H* is ub maximised over the predicted radius
    print("largest loss - bound:", worst)


def test_apart_closed_form():
    for seed in range(3):
        rg32, rp, d = _samples(100 + seed, 200_000)
        loss, rg = _ray_loss(rg32, rp, d)
        apart = d >= (rg32 + rp)  # the kernels use the reference's own fp32 comparison
        closed = 2 - 4 * (rg**2 + rp.double() ** 2) / (rg + rp.double() + d.double()) ** 2
        err = (loss - closed)[apart].abs().max()
        assert float(err) < 6e-6, float(err)  # 3e-6 per value


def test_centre_window_lies_in_the_enumerated_7x7_block():
    """k_pass enumerates, per level, the 7 x 7 block of cells starting at floor(c / stride) - 3 (window_origin) and applies
    the reference's strict test (losses.py:523-542) to those cells only: no passing cell may lie outside the block."""
    g = torch.Generator().manual_seed(7)
    for stride, W in ((8.0, 80), (16.0, 40), (32.0, 20)):
        st = torch.tensor(stride)
        idx = torch.arange(W, dtype=torch.float32)
        xc = (idx * st) + (0.5 * st)  # anchor centres, the reference's arithmetic
        c = torch.cat([torch.rand(20000, generator=g) * 700 - 30,          # anywhere, also outside the image
                       torch.arange(0, 660, 0.5), torch.arange(0, 660, 0.5) + 1e-4, torch.arange(0, 660, 0.5) - 1e-4])
        r = 2.5 * st
        ok = (torch.minimum(xc[None, :] - (c[:, None] - r), (c[:, None] + r) - xc[None, :]) > 0.0)   # [n, W]
        origin = torch.clamp(torch.floor(c / st) - 3.0, -1.0e6, 1.0e6).to(torch.int64)
        cells = torch.arange(W)[None, :].expand_as(ok)
        inside = (cells >= origin[:, None]) & (cells < origin[:, None] + 7)
        assert bool((ok & ~inside).sum() == 0)
        assert int(ok.sum(1).max()) <= 5


def test_nms_cells_keep_overlapping_boxes_adjacent():
    """k_post_nms bins boxes by x0 into cells at least as wide as the widest box: boxes that overlap in x must land in the
    same or in adjacent cells (fp32 arithmetic of the kernel), also under the class offsets of batched NMS."""
    g = torch.Generator().manual_seed(11)
    n = 4000
    x0 = (torch.rand(n, generator=g) * 900 - 200).float()
    w = (torch.rand(n, generator=g) * 580 + 1).float()
    cls = torch.randint(0, 80, (n,), generator=g).float()
    x1 = x0 + w
    step = x1.max() + 1.0
    x0s, x1s = x0 + cls * step, x1 + cls * step
    wmx = (x1s - x0s).max()
    xlo, xhi = x0s.min(), x0s.max()
    span = xhi - xlo
    cellw = torch.maximum(torch.maximum(wmx, span * (1.0 / 2048.0)) * 1.0001, torch.tensor(1e-6))
    ncell = min(2048, int(span / cellw) + 1)
    cell = torch.clamp(((x0s - xlo) / cellw).to(torch.int64), 0, ncell - 1)
    overlap = (x0s[:, None] < x1s[None, :]) & (x0s[None, :] < x1s[:, None])
    far = (cell[:, None] - cell[None, :]).abs() > 1
    assert int((overlap & far).sum()) == 0
    assert ncell >= 16
