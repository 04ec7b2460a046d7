"""Proposal resampling on libsnerf_b200.so -- the drop-in for the step-function operators zip-NeRF's sampling loop calls
(`s-nerfpp/zipnerf/internal/stepfun.py`: `max_dilate_weights`, `sample_intervals`; call site `internal/models.py:156-213`;
BASELINE configs[3]).

`max_dilate_weights` and `sample_intervals` keep the reference's signatures; `resample_intervals` is the whole pass of the
loop (dilate -> drop end bins -> annealed logits -> sample_intervals) as ONE kernel launch, one warp per ray
(csrc/snerf_stepfun.cu).  Like the reference's loop (`stop_level_grad`, models.py:208-209) the result carries no gradient.
No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_EPS = torch.finfo(torch.float32).eps


def _require(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"snerf_b200.stepfun.{what}: tensors must live on a CUDA sm_100 device (no CPU fallback)")


def _uniform(rand, num_samples, single_jitter, prefix, device, _jitter=None):
    """The `u` of stepfun.sample with deterministic_center=True (stepfun.py:199-216), split into its linspace term
    (evaluated with torch.linspace, as the reference does) and the torch.rand draw the kernel scales and adds."""
    if not rand:
        pad = 1 / (2 * num_samples)
        return torch.linspace(pad, 1. - pad - _EPS, num_samples, device=device), None, 0.0
    u_max = _EPS + (1 - _EPS) / num_samples
    max_jitter = (1 - u_max) / (num_samples - 1) - _EPS
    d = 1 if single_jitter else num_samples
    base = torch.linspace(0, 1 - u_max, num_samples, device=device)
    draw = torch.rand(tuple(prefix) + (d,), device=device) if _jitter is None else _jitter      # _jitter: a recorded draw (tests)
    return base, draw, max_jitter


def _launch(t, w, opts, u_base, jitter, n, want_out, want_centers, want_dilate):
    _require(t, "resample")
    prefix = t.shape[:-1]
    S = w.shape[-1]
    if t.shape[-1] != S + 1 or w.shape[:-1] != prefix:
        raise RuntimeError("stepfun: t must be [..., S+1] and w [..., S]")
    t2 = t.detach().reshape(-1, S + 1).contiguous().float()
    w2 = w.detach().reshape(-1, S).contiguous().float()
    N = t2.shape[0]
    dev = t.device
    out = torch.empty(N, n + 1, device=dev) if want_out else None
    cen = torch.empty(N, n, device=dev) if want_centers else None
    td = torch.empty(N, 3 * S + 1, device=dev) if want_dilate else None
    wd = torch.empty(N, 3 * S, device=dev) if want_dilate else None
    jit = jitter.reshape(N, -1).contiguous().float() if jitter is not None else None
    with torch.cuda.device(dev):
        _lib.check(_lib.load().snerf_stepfun_resample(C.byref(opts), _lib.ptr(t2), _lib.ptr(w2), N, S, _lib.ptr(u_base),
                                                      _lib.ptr(jit), jit.shape[1] if jit is not None else 0, n, _lib.ptr(out),
                                                      _lib.ptr(cen), _lib.ptr(td), _lib.ptr(wd), _lib.stream_ptr(dev)),
                   "snerf_stepfun_resample")
    shp = lambda x, k: None if x is None else x.view(tuple(prefix) + (k,))
    return shp(out, n + 1), shp(cen, n), shp(td, 3 * S + 1), shp(wd, 3 * S)


def max_dilate_weights(t, w, dilation, domain=(-torch.inf, torch.inf), renormalize=False):
    """stepfun.py:91-105: dilate (max-pool) a set of weights -> (t_dilate [..., 3S+1], w_dilate [..., 3S])."""
    opts = _lib.StepfunOpts(1, int(bool(renormalize)), 0, float(dilation), float(domain[0]), float(domain[1]), 1.0, 0.0, 0.0)
    _, _, td, wd = _launch(t, w, opts, None, None, 0, False, False, True)
    return td, wd


def sample_intervals(rand, t, w_logits, num_samples, single_jitter=False, domain=(-torch.inf, torch.inf), _centers=False,
                     _jitter=None):
    """stepfun.py:251-294: sample `num_samples` intervals from the step function (t, softmax(w_logits))."""
    if num_samples <= 1:
        raise ValueError(f'num_samples must be > 1, is {num_samples}.')
    base, jitter, max_jitter = _uniform(rand, num_samples, single_jitter, t.shape[:-1], t.device, _jitter)
    opts = _lib.StepfunOpts(0, 0, 1, 0.0, float(domain[0]), float(domain[1]), 1.0, 0.0, float(max_jitter))
    out, cen, _, _ = _launch(t, w_logits, opts, base, jitter, num_samples, True, _centers, False)
    return (out, cen) if _centers else out


def resample_intervals(rand, sdist, weights, num_samples, dilation=None, domain=(0., 1.), anneal=1., resample_padding=1e-5,
                       single_jitter=True, _centers=False, _jitter=None):
    """One pass of Model.forward's sampling loop (models.py:156-213) in one launch: `dilation` not None ->
    max_dilate_weights(renormalize=True) and the [1:-1] slices; logits = where(dt > 0, anneal * log(w + padding), -inf);
    sample_intervals.  Returns the new sdist [..., num_samples + 1]."""
    if num_samples <= 1:
        raise ValueError(f'num_samples must be > 1, is {num_samples}.')
    base, jitter, max_jitter = _uniform(rand, num_samples, single_jitter, sdist.shape[:-1], sdist.device, _jitter)
    opts = _lib.StepfunOpts(int(dilation is not None), 1, 0, float(dilation or 0.0), float(domain[0]), float(domain[1]),
                            float(anneal), float(resample_padding), float(max_jitter))
    out, cen, _, _ = _launch(sdist, weights, opts, base, jitter, num_samples, True, _centers, False)
    return (out, cen) if _centers else out
