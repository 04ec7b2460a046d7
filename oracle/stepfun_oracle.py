"""TEST INFRASTRUCTURE ONLY -- numpy restatement of zip-NeRF's proposal resampling step
(s-nerfpp/zipnerf/internal/stepfun.py: weight_to_pdf :64-67, pdf_to_weight :70-72, max_dilate :75-88,
max_dilate_weights :91-105, integrate_weights :108-128, invert_cdf :154-161, sample :175-218, sample_intervals :251-294;
internal/math.py: sorted_interp :88-107; the call site internal/models.py:183-213).  BASELINE configs[3], SURVEY section 8
row f-2(ii).

Only tests/ may import this module; the product path is the CUDA library (csrc/snerf_stepfun.cu).

Parity status: pinned by tests/golden/stepfun_*.npz -- outputs of the reference's own stepfun.py imported from
/root/reference and run on torch-CPU (oracle/make_golden_stepfun.py).  fp32 throughout; the sums (softmax denominator,
cumulative sum, renormalisation) are order-dependent, everything else follows the reference operation by operation.
"""
import numpy as np

F = np.float32
EPS = np.finfo(np.float32).eps


def max_dilate_weights(t, w, dilation, domain=(-np.inf, np.inf), renormalize=False):
    """stepfun.py:75-105.  t [N, S+1] sorted, w [N, S] -> t_dilate [N, 3S+1], w_dilate [N, 3S]."""
    t = np.asarray(t, F); w = np.asarray(w, F)
    p = (w / np.maximum(t[:, 1:] - t[:, :-1], F(EPS))).astype(F)
    t0 = (t[:, :-1] - F(dilation)).astype(F)
    t1 = (t[:, 1:] + F(dilation)).astype(F)
    td = np.sort(np.concatenate([t, t0, t1], axis=-1), axis=-1)
    td = np.clip(td, F(domain[0]), F(domain[1])).astype(F)
    cover = (t0[:, None, :] <= td[:, :, None]) & (t1[:, None, :] > td[:, :, None])
    pd = np.where(cover, p[:, None, :], F(0)).max(axis=-1)[:, :-1]
    wd = (pd * (td[:, 1:] - td[:, :-1]).astype(F)).astype(F)
    if renormalize:
        wd = (wd / np.maximum(wd.sum(axis=-1, keepdims=True, dtype=F), F(EPS))).astype(F)
    return td, wd


def integrate_weights(w):
    """stepfun.py:108-128."""
    cw = np.minimum(np.cumsum(w[:, :-1], axis=-1, dtype=F), F(1))
    z = np.zeros((w.shape[0], 1), F)
    return np.concatenate([z, cw, z + F(1)], axis=-1)


def sorted_interp(x, xp, fp):
    """math.py:88-107 for sorted xp / fp: x [N, n], xp / fp [N, T+1]."""
    N, n = x.shape
    out = np.empty((N, n), F)
    for r in range(N):
        idx = np.searchsorted(xp[r], x[r], side="right") - 1          # last i with x >= xp[i]
        has = idx >= 0
        i0 = np.where(has, idx, 0)
        i1 = np.minimum(idx + 1, xp.shape[1] - 1)
        x0, x1, f0, f1 = xp[r][i0], xp[r][i1], fp[r][i0], fp[r][i1]
        with np.errstate(divide="ignore", invalid="ignore"):
            off = ((x[r] - x0).astype(F) / (x1 - x0).astype(F)).astype(F)
        off = np.clip(np.nan_to_num(off, nan=0.0), 0, 1).astype(F)
        out[r] = (f0 + (off * (f1 - f0).astype(F)).astype(F)).astype(F)
    return out


def softmax(l):
    m = l.max(axis=-1, keepdims=True)
    e = np.exp((l - m).astype(F)).astype(F)
    return (e / e.sum(axis=-1, keepdims=True, dtype=F)).astype(F)


def uniform_samples(n, deterministic_center, jitter=None, single_jitter=False):
    """stepfun.py:199-216: the u of `sample`; jitter = the torch.rand draw [N, 1 or n] (None = rand is None)."""
    import torch                                      # torch.linspace's exact fp32 values (its halfway formula)
    if jitter is None:
        if deterministic_center:
            pad = 1 / (2 * n)
            return torch.linspace(pad, 1. - pad - EPS, n).numpy()[None], 0.0
        return torch.linspace(0, 1. - EPS, n).numpy()[None], 0.0
    u_max = EPS + (1 - EPS) / n
    max_jitter = (1 - u_max) / (n - 1) - EPS
    return torch.linspace(0, 1 - u_max, n).numpy()[None], max_jitter


def sample_intervals(t, w_logits, n, jitter=None, single_jitter=False, domain=(-np.inf, np.inf)):
    """stepfun.py:251-294 (+ sample :175-218 with deterministic_center=True, invert_cdf :154-161)."""
    t = np.asarray(t, F); l = np.asarray(w_logits, F)
    u_base, max_jitter = uniform_samples(n, True, jitter, single_jitter)
    u = np.broadcast_to(u_base, (t.shape[0], n)).astype(F)
    if jitter is not None:
        u = (u + (np.asarray(jitter, F) * F(max_jitter)).astype(F)).astype(F)
    cw = integrate_weights(softmax(l))
    centers = sorted_interp(u, cw, t)
    mid = ((centers[:, 1:] + centers[:, :-1]).astype(F) / F(2)).astype(F)
    first = np.maximum((F(2) * centers[:, :1] - mid[:, :1]).astype(F), F(domain[0]))
    last = np.minimum((F(2) * centers[:, -1:] - mid[:, -1:]).astype(F), F(domain[1]))
    return np.concatenate([first, mid, last], axis=-1).astype(F)


def resample_level(sdist, weights, n, dilate, dilation, domain, anneal=1.0, resample_padding=1e-5, jitter=None,
                   single_jitter=False):
    """One pass of the sampling loop, models.py:176-213: (dilate, drop the end bins,) annealed logits, sample_intervals."""
    sdist = np.asarray(sdist, F); weights = np.asarray(weights, F)
    if dilate:
        sdist, weights = max_dilate_weights(sdist, weights, dilation, domain=domain, renormalize=True)
        sdist, weights = sdist[:, 1:-1], weights[:, 1:-1]
    with np.errstate(divide="ignore"):
        logits = np.where(sdist[:, 1:] > sdist[:, :-1], (F(anneal) * np.log((weights + F(resample_padding)).astype(F))).astype(F),
                          F(-np.inf)).astype(F)
    return sample_intervals(sdist, logits, n, jitter, single_jitter, domain)
