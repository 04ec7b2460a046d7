"""TEST INFRASTRUCTURE ONLY -- numpy (float64) restatement of the training objective of s-nerf/train.py:149-209:
RgbLoss (model/loss_factory.py:5-11), DepthLoss (:26-37) and calc_depth_loss's reduction (model/confidence.py:211-226),
with analytic gradients.  Pinned by tests/golden/loss_*.npz (the reference's classes + torch autograd on CPU,
oracle/make_golden_loss.py).  Only tests/ and smoke() may import this module."""
import numpy as np


def rgb_depth_loss(rgb, target, depth=None, depth0=None, target_depth=None, confidence=None, depth_lambda=0.1,
                   coarse_depth_mult=0.2, disparity=True, rgb0=None, rgb0_weight=0.0, upstream=1.0):
    """Returns (loss, img_loss, depth_loss, grads dict) in float64."""
    rgb, target = np.asarray(rgb, np.float64), np.asarray(target, np.float64)
    N = rgb.shape[0]
    img = np.mean((rgb - target) ** 2)
    grads = {"rgb": upstream * 2 * (rgb - target) / (3 * N)}
    loss = img
    if rgb0 is not None and rgb0_weight:
        r0 = np.asarray(rgb0, np.float64)
        loss = loss + rgb0_weight * np.mean((r0 - target) ** 2)
        grads["rgb0"] = upstream * rgb0_weight * 2 * (r0 - target) / (3 * N)
    dep = 0.0
    if depth is not None:
        d, d0, t = (np.asarray(a, np.float64) for a in (depth, depth0, target_depth))
        m = t != 0
        f = (lambda x: 1.0 / x) if disparity else (lambda x: x)
        fp = (lambda x: -1.0 / x ** 2) if disparity else (lambda x: np.ones_like(x))
        c = np.asarray(confidence, np.float64)[m] if confidence is not None else 1.0
        e, e0 = f(d[m]) - f(t[m]), f(d0[m]) - f(t[m])
        term = np.abs(e) + coarse_depth_mult * np.abs(e0)
        cnt = m.sum()
        dep = np.sum(c * term) / cnt if cnt else np.nan
        loss = loss + depth_lambda * dep
        k = upstream * depth_lambda / cnt if cnt else np.nan
        for key, val in (("depth", k * c * np.sign(e) * fp(d[m])), ("depth0", k * c * coarse_depth_mult * np.sign(e0) * fp(d0[m])),
                         ("confidence", k * term)):
            g = np.zeros(N)
            g[m] = val
            grads[key] = g
    return loss, img, dep, grads


def proposal_loss(s_vals_f, weights_f, s_vals_c, weights_c, weight=1.0):
    """ProposalLoss (model/loss_factory.py:54-73) with the analytic gradient w.r.t. weights_c.  The per-interval terms
    follow the reference's fp32 operations (cumsum accumulated in double and rounded, as torch does on the CPU; fp32 bound /
    clamp / divide: the division by
    w_f + 1e-8 amplifies any rounding of the bound for tiny fine weights, so a float64 restatement would not pin it);
    only the sums over intervals / rays are accumulated in float64.  Returns (loss, grad_weights_c)."""
    f = np.float32
    sf, wf, sc, wc = (np.asarray(a, f) for a in (s_vals_f, weights_f, s_vals_c, weights_c))
    N, Sf = wf.shape
    Sc = wc.shape[1]
    loss, grad = 0.0, np.zeros(wc.shape, np.float64)
    for n in range(N):
        inds = np.searchsorted(sc[n], sf[n], side="right")
        W = np.cumsum(wc[n], dtype=np.float64).astype(f)      # torch's CPU cumsum accumulates float in double
        l = np.clip(np.maximum(inds[:-1] - 1, 0), 0, Sc - 1)
        r = np.clip(np.minimum(inds[1:] - 1, Sf - 1), 0, Sc - 1)
        bound = (W[r] - W[l]).astype(f)
        e = np.maximum((wf[n] - bound).astype(f), f(0))
        den = (wf[n] + f(1e-8)).astype(f)
        loss += np.sum(((e * e).astype(f) / den).astype(f), dtype=np.float64)
        g = (f(-2) * e / den).astype(np.float64)
        for i in range(Sf):
            if e[i] > 0 and r[i] != l[i]:
                lo, hi = min(l[i], r[i]), max(l[i], r[i])
                grad[n, lo + 1:hi + 1] += g[i] if r[i] > l[i] else -g[i]
    return loss / N * weight, grad / N * weight
