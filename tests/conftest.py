import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_params(g):
    """Regenerate the fixture's network weights from their seeds (oracle.make_nerf_params)."""
    from oracle import snerf_oracle as O
    D, W = int(g["D"]), int(g["W"])
    if "trunk_gain" in g:
        kw = dict(trunk_gain=float(g["trunk_gain"]), sigma_bias=float(g["sigma_bias"]))
        return (O.make_nerf_params(int(g["seed_coarse"]), D=D, W=W, **kw),
                O.make_nerf_params(int(g["seed_fine"]), D=D, W=W, **kw))
    return O.make_nerf_params(int(g["seed_coarse"]), D=D, W=W), None


def err_metric(a, b, floor=1e-2):
    """max |a-b| / (|b| + floor*rms(b)): relative error with a floor so near-zero entries of a
    tensor are judged against the tensor's own scale (use floor=0.1 for pre-activation MLP outputs,
    which are sums with cancellation and cross zero)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    m = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), m), "finite/non-finite pattern differs"
    if not m.any():
        return 0.0
    rms = np.sqrt(np.mean(b[m] ** 2)) + 1e-30
    return float(np.max(np.abs(a[m] - b[m]) / (np.abs(b[m]) + floor * rms)))


def bins_of_samples(z_vals, z_samples):
    """The inverse-CDF bin every resampled depth fell into (`below = max(0, inds - 1)` of run_nerf_helpers.py:363-365),
    recovered from the depths themselves: position of z_sample among the coarse mid-points."""
    z = np.asarray(z_vals, np.float32)
    mid = (np.float32(0.5) * (z[:, 1:] + z[:, :-1])).astype(np.float32)
    b = np.stack([np.searchsorted(mid[r], z_samples[r], side="right") for r in range(mid.shape[0])]) - 1
    return np.clip(b, 0, mid.shape[1] - 2)


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.fixture(autouse=True)
def _inference_unless_training(request):
    """Rendering tests mirror the reference's evaluation calls, which run under `torch.no_grad()` (train.py / eval.py);
    the training tests (`train` in their name) keep autograd on: with trainable parameters and grad mode enabled,
    render_rays records its autograd node like any torch module would."""
    import torch
    if "train" in request.node.name or "grad" in request.node.name:
        yield
        return
    with torch.no_grad():
        yield


def stepfun_cdf(t, logits, x):
    """Piecewise-linear CDF of the step function (t [N, T+1], softmax(logits) [N, T]) evaluated at x [N, n], float64.
    Inverse-CDF samples are compared in CDF space: where a bin holds almost no mass the sample position is
    ill-conditioned (an ulp of the CDF moves it across the bin), its CDF value is not."""
    t = np.asarray(t, np.float64); x = np.asarray(x, np.float64)
    l = np.asarray(logits, np.float64)
    w = np.exp(l - l.max(-1, keepdims=True))
    w /= w.sum(-1, keepdims=True)
    cw = np.concatenate([np.zeros((t.shape[0], 1)), np.cumsum(w, -1)], -1)
    return np.stack([np.interp(x[r], t[r], cw[r]) for r in range(t.shape[0])])


VARIANT_ARCH = {"d4": (4, 256), "w128": (6, 128)}


def variant_networks(g, case):
    """(coarse params, fine params, frozen sigma params of the coarse / fine pass) of one grad_variants.npz case."""
    from oracle import snerf_oracle as O, snerf_oracle_grad as OG
    sc, sf, sa = int(g["seed_c"]), int(g["seed_f"]), int(g["seed_alpha"])
    if case == "novd":
        return OG.variant_params(sc, "novd"), OG.variant_params(sf, "novd"), None, None
    if case in VARIANT_ARCH:       # coarse network of its own depth / width, flagship fine network
        Dc, Wc = VARIANT_ARCH[case]
        return (O.make_nerf_params(sc, D=Dc, W=Wc, trunk_gain=1.5, sigma_bias=0.5),
                O.make_nerf_params(sf, trunk_gain=1.5, sigma_bias=0.5), None, None)
    pa = O.make_nerf_params(sa, trunk_gain=1.5, sigma_bias=0.5)
    if case == "rgb":
        return OG.variant_params(sc, "rgb"), OG.variant_params(sf, "rgb"), pa, pa
    return pa, OG.variant_params(sf, "rgb"), None, pa        # nocoarse: alpha_model itself serves the coarse pass
