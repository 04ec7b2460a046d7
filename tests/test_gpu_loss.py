"""GPU (B200): the fused training objective (csrc/snerf_loss.cu) through snerf_b200.losses -> C ABI, against the
fixtures produced by the reference's own RgbLoss / DepthLoss + torch autograd (tests/golden/loss_*.npz), the float64 oracle,
and torch ops on the GPU at training-batch and full-image sizes."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from test_oracle_golden import LOSS_CASES, check_loss_against_golden

pytestmark = pytest.mark.gpu


def run_ours(g, dev, upstream):
    from snerf_b200.losses import RgbDepthLoss
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    rgb, depth, depth0 = t("rgb").requires_grad_(True), t("depth").requires_grad_(True), t("depth0").requires_grad_(True)
    conf = t("confidence").requires_grad_(True) if bool(g["with_conf"]) else None
    crit = RgbDepthLoss(float(g["depth_lambda"]), float(g["coarse_depth_mult"]), bool(g["disparity"]))
    with torch.enable_grad():
        loss = crit(rgb, t("target"), depth, depth0, t("target_depth"), conf)
        (loss * upstream).backward()
    torch.cuda.synchronize()
    grads = {"rgb": rgb.grad.cpu().numpy(), "depth": depth.grad.cpu().numpy(), "depth0": depth0.grad.cpu().numpy()}
    if conf is not None:
        grads["confidence"] = conf.grad.cpu().numpy()
    return float(loss), float(crit.last["img_loss"]), float(crit.last["depth_loss"]), grads, crit


@pytest.mark.parametrize("name", LOSS_CASES)
def test_loss_grad_matches_reference_fixture(name, cuda_device):
    g = load_golden(name)
    loss, img, dep, grads, crit = run_ours(g, cuda_device, float(g["upstream"]))
    check_loss_against_golden(g, loss, img, dep, grads)
    assert int(crit.last["masked"]) == int(np.sum(g["target_depth"] != 0))


def test_loss_grad_matches_torch_at_full_size(cuda_device):
    """1,440,000 rays (one frame): against the same expression in torch on the GPU, incl. the coarse colour term."""
    from snerf_b200.losses import RgbDepthLoss
    dev = cuda_device
    gen = torch.Generator(device=dev).manual_seed(9)
    N = 1_440_000
    r = lambda *s: torch.rand(*s, device=dev, generator=gen)
    tgt, tdep, conf = r(N, 3), r(N) * 98 + 2, r(N)
    tdep[r(N) < 0.3] = 0.0
    leaves = [r(N, 3), r(N, 3), r(N) * 98 + 2, r(N) * 98 + 2]
    outs = []
    for fused in (True, False):
        rgb, rgb0, d, d0 = [x.clone().requires_grad_(True) for x in leaves]
        with torch.enable_grad():
            if fused:
                loss = RgbDepthLoss(0.1, 0.2, True, rgb0_weight=1.0)(rgb, tgt, d, d0, tdep, conf, rgb_coarse=rgb0)
            else:
                m = tdep != 0
                dl = (conf[m] * ((1 / d[m] - 1 / tdep[m]).abs() + 0.2 * (1 / d0[m] - 1 / tdep[m]).abs())).mean()
                loss = ((rgb - tgt) ** 2).mean() + ((rgb0 - tgt) ** 2).mean() + 0.1 * dl
            loss.backward()
        outs.append((float(loss), [x.grad for x in (rgb, rgb0, d, d0)]))
    (la, ga), (lb, gb) = outs
    assert abs(la - lb) <= 1e-5 * abs(lb)
    for a, b in zip(ga, gb):
        assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max())


def test_loss_errors_and_optional_terms(cuda_device):
    from snerf_b200.losses import RgbDepthLoss
    dev = cuda_device
    crit = RgbDepthLoss()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        crit(torch.zeros(4, 3), torch.zeros(4, 3))
    rgb, tgt = torch.rand(100, 3, device=dev), torch.rand(100, 3, device=dev)
    only_rgb = crit(rgb, tgt)                                          # no depth term: plain RgbLoss
    assert abs(float(only_rgb) - float(((rgb - tgt) ** 2).mean())) < 1e-6
    nolidar = crit(rgb, tgt, torch.ones(100, device=dev), torch.ones(100, device=dev), torch.zeros(100, device=dev))
    assert torch.isnan(nolidar)                                        # mean over zero masked rays, as torch
    with pytest.raises(RuntimeError, match=r"\[N, 3\]"):
        crit(torch.rand(100, 4, device=dev), tgt)
