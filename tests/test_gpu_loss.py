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


# ------------------------------------------------------------------ ProposalLoss (loss_factory.py:54-73)
from test_oracle_golden import PROPOSAL_CASES, check_proposal_against_golden  # noqa: E402


@pytest.mark.parametrize("name", PROPOSAL_CASES)
def test_proposal_loss_grad_matches_reference_fixture(name, cuda_device):
    from snerf_b200.losses import ProposalLoss
    g = load_golden(name)
    t = lambda k: torch.from_numpy(g[k]).to(cuda_device)
    wc = t("weights_c").requires_grad_(True)
    with torch.enable_grad():
        loss = ProposalLoss(float(g["weight"]))(t("s_vals_f"), t("weights_f"), t("s_vals_c"), wc)
        (loss * float(g["upstream"])).backward()
    torch.cuda.synchronize()
    check_proposal_against_golden(g, float(loss), wc.grad.cpu().numpy())


def test_proposal_loss_grad_full_size_properties(cuda_device):
    """One training batch of the shipped config and a full frame's worth of rays: zero when the coarse histogram dominates
    the fine one, matches the same expression in torch on the GPU, finite-difference check of the gradient."""
    from snerf_b200.losses import ProposalLoss
    dev = cuda_device
    gen = torch.Generator(device=dev).manual_seed(4)
    N, Sf, Sc = 65536, 128, 64
    sc = torch.sort(torch.rand(N, Sc + 1, device=dev, generator=gen), dim=-1).values
    sc[:, 0], sc[:, -1] = 0.0, 1.0
    sf = torch.sort(0.02 + 0.95 * torch.rand(N, Sf + 1, device=dev, generator=gen), dim=-1).values
    wf = torch.rand(N, Sf, device=dev, generator=gen) ** 3
    wf = wf / wf.sum(-1, keepdim=True)
    wc = torch.rand(N, Sc, device=dev, generator=gen) ** 3
    wc = (wc / wc.sum(-1, keepdim=True)).requires_grad_(True)
    crit = ProposalLoss(1.0)
    with torch.enable_grad():
        loss = crit(sf, wf, sc, wc)
        loss.backward()
    # the reference's expression, evaluated by torch on the same device (its CUDA cumsum is an fp32 scan: looser)
    wc2 = wc.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        inds = torch.searchsorted(sc, sf, right=True)
        W = torch.cumsum(wc2, dim=1)
        left = torch.gather(W, 1, torch.clamp(inds[:, :-1] - 1, min=0))
        right = torch.gather(W, 1, torch.clamp(inds[:, 1:] - 1, max=Sf - 1))
        ref = (torch.clamp(wf - (right - left), min=0) ** 2 / (wf + 1e-8)).sum(1).mean()
        ref.backward()
    assert abs(float(loss) - float(ref)) <= 1e-4 * abs(float(ref))
    # gradient entries fed by fine intervals with w_f -> 0 are ill-conditioned (g = -2 e / (w_f + 1e-8) amplifies the
    # rounding of the cumulative sum, which torch computes as an fp32 scan on CUDA and in double on the CPU): compare in
    # relative L2 and by the fraction of close entries; the bit-level bar is the CPU-reference fixture above
    dg = (wc.grad - wc2.grad).double()
    assert float(dg.norm() / wc2.grad.double().norm()) <= 5e-2
    assert float((dg.abs() <= 1e-3 * wc2.grad.abs().max()).float().mean()) > 0.98
    # identical edges: the reference's bound of fine interval i is the coarse weight of interval i + 1 (and 0 for the last
    # one); a coarse histogram above the fine one there gives exactly zero loss and gradient
    wf0 = wf[:, :Sc].clone()
    wf0[:, -1] = 0.0
    big = torch.full((N, Sc), 2.0, device=dev, requires_grad=True)
    widths = 0.5 + torch.rand(N, Sc, device=dev, generator=gen)                  # strictly increasing edges (no ties)
    se = torch.cat([torch.zeros(N, 1, device=dev), torch.cumsum(widths, -1)], -1) / (1.5 * Sc)
    with torch.enable_grad():
        z = crit(se, wf0, se, big)
        z.backward()
    assert float(z) == 0.0 and float(big.grad.abs().max()) == 0.0
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        crit(sf.cpu(), wf.cpu(), sc.cpu(), wc.detach().cpu())
