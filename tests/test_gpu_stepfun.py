"""GPU (B200): zip-NeRF's proposal resampling step (csrc/snerf_stepfun.cu) through snerf_b200.stepfun -> C ABI, against
the fixtures produced by the reference's own stepfun.py (tests/golden/stepfun_*.npz), the numpy oracle, and
size-independent properties at render-chunk sizes."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import load_golden, stepfun_cdf
from test_oracle_golden import STEPFUN_CASES, check_stepfun_against_golden

pytestmark = pytest.mark.gpu


def run_ours(g, dev):
    from snerf_b200 import stepfun
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    n, dilate, single = int(g["n"]), bool(g["dilate"]), bool(g["single_jitter"])
    jitter = t(g["jitter"]) if "jitter" in g else None
    t_dil = w_dil = None
    if dilate:
        t_dil, w_dil = stepfun.max_dilate_weights(t(g["sdist"]), t(g["weights"]), float(g["dilation"]), domain=(0., 1.), renormalize=True)
        t_dil, w_dil = t_dil.cpu().numpy(), w_dil.cpu().numpy()
    out, cen = stepfun.resample_intervals(jitter is not None, t(g["sdist"]), t(g["weights"]), n,
                                          dilation=float(g["dilation"]) if dilate else None, domain=(0., 1.),
                                          anneal=float(g["anneal"]), resample_padding=1e-5, single_jitter=single,
                                          _centers=True, _jitter=jitter)
    torch.cuda.synchronize()
    return t_dil, w_dil, cen.cpu().numpy(), out.cpu().numpy()


@pytest.mark.parametrize("name", STEPFUN_CASES)
def test_stepfun_resample_matches_reference_fixture(name, cuda_device):
    g = load_golden(name)
    check_stepfun_against_golden(g, *run_ours(g, cuda_device))


@pytest.mark.parametrize("name", STEPFUN_CASES)
def test_stepfun_stage_entry_points_match(name, cuda_device):
    """The reference-named operators compose to the fused pass: max_dilate_weights -> slices -> logits ->
    sample_intervals (the loop of models.py:174-204 written with our operators) == resample_intervals, bit for bit."""
    from snerf_b200 import stepfun
    g = load_golden(name)
    dev = cuda_device
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    n, dilate, single = int(g["n"]), bool(g["dilate"]), bool(g["single_jitter"])
    jitter = t(g["jitter"]) if "jitter" in g else None
    sd, w = t(g["sdist"]), t(g["weights"])
    if dilate:
        sd, w = stepfun.max_dilate_weights(sd, w, float(g["dilation"]), domain=(0., 1.), renormalize=True)
        sd, w = sd[..., 1:-1], w[..., 1:-1]
    logits = torch.where(sd[..., 1:] > sd[..., :-1], float(g["anneal"]) * torch.log(w + 1e-5), torch.full_like(w, -torch.inf))
    staged = stepfun.sample_intervals(jitter is not None, sd, logits, n, single_jitter=single, domain=(0., 1.), _jitter=jitter)
    _, _, _, fused = run_ours(g, dev)
    d = np.abs(staged.cpu().numpy() - fused)
    assert float(np.mean(d <= 1e-6)) >= 0.98 and float(d.max()) < 1e-2      # torch.log vs logf: an ulp apart at most


def test_stepfun_full_size_properties(cuda_device):
    """65,536 rays x 64 bins -> 64 intervals (one zip-NeRF chunk): sorted, inside the domain, the sampled centres invert
    the CDF (F(centre_i) == u_i), deterministic runs repeat bit for bit, ragged ray counts are prefixes."""
    from snerf_b200 import stepfun
    dev = cuda_device
    gen = torch.Generator(device=dev).manual_seed(3)
    N, S, n = 1 << 16, 64, 64
    sd = torch.sort(torch.rand(N, S + 1, device=dev, generator=gen), dim=-1).values
    sd[:, 0], sd[:, -1] = 0.0, 1.0
    w = torch.rand(N, S, device=dev, generator=gen) ** 6
    w = w / w.sum(-1, keepdim=True)
    out, cen = stepfun.resample_intervals(None, sd, w, n, dilation=0.01, domain=(0., 1.), _centers=True)
    out2 = stepfun.resample_intervals(None, sd, w, n, dilation=0.01, domain=(0., 1.))
    assert torch.equal(out, out2)
    assert out.shape == (N, n + 1) and bool((out[:, 1:] >= out[:, :-1]).all())
    assert float(out.min()) >= 0.0 and float(out.max()) <= 1.0
    # round trip through the dilated step function our own operators return
    td, wd = stepfun.max_dilate_weights(sd, w, 0.01, domain=(0., 1.), renormalize=True)
    td, wd = td[:, 1:-1], wd[:, 1:-1]
    assert bool((td[:, 1:] >= td[:, :-1]).all()) and abs(float(wd.sum(-1).mean()) - 1.0) < 1e-2
    logits = torch.where(td[:, 1:] > td[:, :-1], torch.log(wd + 1e-5), torch.full_like(wd, -torch.inf))
    sub = slice(0, 512)
    F = stepfun_cdf(td[sub].cpu().numpy(), logits[sub].cpu().numpy(), cen[sub].cpu().numpy())
    pad = 1 / (2 * n)
    u = np.linspace(pad, 1 - pad, n)
    assert float(np.max(np.abs(F - u[None]))) < 5e-6
    for m in (1, 3, 5, 130):
        part = stepfun.resample_intervals(None, sd[:m], w[:m], n, dilation=0.01, domain=(0., 1.))
        assert torch.equal(part, out[:m])
    # randomized: stays sorted / in range, jitter moves every ray
    outr = stepfun.resample_intervals(True, sd, w, n, dilation=0.01, domain=(0., 1.), single_jitter=True)
    assert bool((outr[:, 1:] >= outr[:, :-1]).all()) and float(outr.min()) >= 0.0 and float(outr.max()) <= 1.0
    assert float((outr != out).float().mean()) > 0.5


def test_stepfun_errors_and_empty(cuda_device):
    from snerf_b200 import _lib, stepfun
    dev = cuda_device
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        stepfun.resample_intervals(None, torch.zeros(2, 5), torch.ones(2, 4), 8)
    with pytest.raises(ValueError):
        stepfun.sample_intervals(None, torch.zeros(2, 5, device=dev), torch.ones(2, 4, device=dev), 1)
    with pytest.raises(RuntimeError, match="bins"):
        stepfun.resample_intervals(None, torch.zeros(2, 200, device=dev), torch.ones(2, 199, device=dev), 8, dilation=0.01)
    out = stepfun.resample_intervals(None, torch.zeros(0, 5, device=dev), torch.ones(0, 4, device=dev), 8)
    assert out.shape == (0, 9)
    out = stepfun.resample_intervals(None, torch.zeros(2, 3, 5, device=dev).add_(torch.linspace(0, 1, 5, device=dev)),
                                     torch.ones(2, 3, 4, device=dev), 8)
    assert out.shape == (2, 3, 9)                                    # prefix shapes are kept
    o = _lib.StepfunOpts(0, 0, 0, 0.0, 0.0, 1.0, 1.0, 1e-5, 0.0)
    assert _lib.load().snerf_stepfun_resample(C.byref(o), None, None, 4, 4, None, None, 0, 8, None, None, None, None, None) != 0
