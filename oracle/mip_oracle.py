"""TEST INFRASTRUCTURE ONLY -- numpy (fp32) restatement of the sampling / encoding / compositing functions of the mip-NeRF
path the shipped S-NeRF config trains with: s-nerf/model/mip.py -- sample_along_rays :192-212,
resample_along_rays :215-238, cast_rays :80-91 with conical_frustum_to_gaussian :56-70 and lift_gaussian :31-53 (diag),
integrated_pos_enc :94-118 with expected_sin :24-28, volumetric_rendering :121-148 -- and model/math_ops.py --
sorted_piecewise_constant_pdf :19-76, safe_sin / safe_cos :6-16.

GROUNDWORK for the next round: there is no CUDA path for these functions yet, so nothing in the product
or the GPU tests uses this module; it is pinned now (tests/test_oracle_vs_reference_live.py sweeps it against the unmodified
reference on torch-CPU) so that the kernels can be written against a trusted checker.
"""
import numpy as np

F = np.float32
EPS32 = np.finfo(np.float32).eps


def _linspace01(n):
    import torch                                    # torch.linspace's exact fp32 values (halfway formula)
    return torch.linspace(0., 1., n).numpy()


def sample_along_rays_t(near, far, num_samples, lindisp=False, t_rand=None):
    """The t_vals of sample_along_rays (mip.py:192-212); near / far [N, 1]; t_rand [N, S+1] = the torch.rand draw or None."""
    t = _linspace01(num_samples + 1)[None]
    near, far = np.asarray(near, F), np.asarray(far, F)
    if lindisp:
        t = (F(1) / ((F(1) / near * (F(1) - t)).astype(F) + (F(1) / far * t).astype(F))).astype(F)
    else:
        t = ((near * (F(1) - t)).astype(F) + (far * t).astype(F)).astype(F)
    if t_rand is not None:
        mids = (F(0.5) * (t[..., 1:] + t[..., :-1])).astype(F)
        upper = np.concatenate([mids, t[..., -1:]], -1)
        lower = np.concatenate([t[..., :1], mids], -1)
        t = (lower + ((upper - lower).astype(F) * np.asarray(t_rand, F)).astype(F)).astype(F)
    return np.broadcast_to(t, (near.shape[0], num_samples + 1)).astype(F)


def cast_rays(t_vals, origins, directions, radii):
    """cast_rays(..., ray_shape='cone', diag=True) (mip.py:80-91, 56-70 stable branch, 31-45): means [N,S,3], cov_diag [N,S,3]."""
    t = np.asarray(t_vals, F)
    d, o, r = np.asarray(directions, F), np.asarray(origins, F), np.asarray(radii, F)
    t0, t1 = t[..., :-1], t[..., 1:]
    mu = ((t0 + t1) / F(2)).astype(F)
    hw = ((t1 - t0) / F(2)).astype(F)
    den = (F(3) * mu ** 2 + hw ** 2).astype(F)
    t_mean = (mu + ((F(2) * mu * hw ** 2).astype(F) / den).astype(F)).astype(F)
    t_var = ((hw ** 2) / F(3) - F(4 / 15) * ((hw ** 4 * (F(12) * mu ** 2 - hw ** 2)).astype(F) / den ** 2)).astype(F)
    r_var = (r ** 2 * ((mu ** 2) / F(4) + F(5 / 12) * hw ** 2 - F(4 / 15) * (hw ** 4) / den)).astype(F)
    mean = (d[..., None, :] * t_mean[..., None]).astype(F)
    d_mag_sq = np.maximum(F(1e-10), np.sum(d ** 2, axis=-1, keepdims=True, dtype=F))
    d_outer = (d ** 2).astype(F)
    null_outer = (F(1) - d_outer / d_mag_sq).astype(F)
    cov = ((t_var[..., None] * d_outer[..., None, :]).astype(F) + (r_var[..., None] * null_outer[..., None, :]).astype(F)).astype(F)
    return (mean + o[..., None, :]).astype(F), cov


def _safe_trig(x, fn, t=F(100 * np.pi)):
    return fn(np.where(np.abs(x) < t, x, np.mod(x, t))).astype(F)


def integrated_pos_enc(means, cov_diag, min_deg, max_deg):
    """integrated_pos_enc(diag=True) (mip.py:94-118): [.., 3] x2 -> [.., 2 * 3 * (max_deg - min_deg)]."""
    x, c = np.asarray(means, F), np.asarray(cov_diag, F)
    scales = np.array([2 ** i for i in range(min_deg, max_deg)], F)
    shape = x.shape[:-1] + (-1,)
    y = (x[..., None, :] * scales[:, None]).astype(F).reshape(shape)
    y_var = (c[..., None, :] * (scales[:, None] ** 2).astype(F)).astype(F).reshape(shape)
    xx = np.concatenate([y, (y + F(0.5 * np.pi)).astype(F)], -1)
    vv = np.concatenate([y_var, y_var], -1)
    return (np.exp((F(-0.5) * vv).astype(F)).astype(F) * _safe_trig(xx, np.sin)).astype(F)


def sorted_piecewise_constant_pdf(bins, weights, num_samples, u_rand=None):
    """math_ops.py:19-76.  bins [N, S+1] sorted, weights [N, S]; u_rand = the uniform_(to = 1/n - eps) draw [N, n] or None."""
    bins, w = np.asarray(bins, F), np.asarray(weights, F)
    eps = F(1e-5)
    wsum = np.sum(w, axis=-1, keepdims=True, dtype=F)
    padding = np.maximum(F(0), (eps - wsum).astype(F))
    w = (w + (padding / F(w.shape[-1])).astype(F)).astype(F)
    wsum = (wsum + padding).astype(F)
    pdf = (w / wsum).astype(F)
    cdf = np.minimum(F(1), np.cumsum(pdf[..., :-1].astype(np.float64), axis=-1).astype(F))   # torch CPU cumsum: double accumulate
    z = np.zeros(cdf.shape[:-1] + (1,), F)
    cdf = np.concatenate([z, cdf, z + F(1)], -1)
    if u_rand is not None:
        s = 1 / num_samples
        u = ((np.arange(num_samples) * s).astype(F) + np.asarray(u_rand, F)).astype(F)
        u = np.minimum(u, F(1) - F(EPS32))
    else:
        import torch
        u = np.broadcast_to(torch.linspace(0., 1. - EPS32, num_samples).numpy(), cdf.shape[:-1] + (num_samples,)).astype(F)
    out = np.empty(u.shape, F)
    for r in range(u.shape[0]):
        idx = np.searchsorted(cdf[r], u[r], side="right") - 1
        i0 = np.maximum(idx, 0)
        i1 = np.minimum(idx + 1, cdf.shape[1] - 1)
        with np.errstate(divide="ignore", invalid="ignore"):
            t = ((u[r] - cdf[r][i0]).astype(F) / (cdf[r][i1] - cdf[r][i0]).astype(F)).astype(F)
        t = np.clip(np.nan_to_num(t, nan=0.0), 0, 1).astype(F)
        out[r] = (bins[r][i0] + (t * (bins[r][i1] - bins[r][i0]).astype(F)).astype(F)).astype(F)
    return out


def resample_t(t_vals, weights, resample_padding, u_rand=None):
    """The new t_vals of resample_along_rays (mip.py:215-238): blurred max-pool of the weights + padding, then the sorted pdf."""
    w = np.asarray(weights, F)
    pad = np.concatenate([w[..., :1], w, w[..., -1:]], -1)
    wmax = np.maximum(pad[..., :-1], pad[..., 1:])
    blur = (F(0.5) * (wmax[..., :-1] + wmax[..., 1:])).astype(F)
    return sorted_piecewise_constant_pdf(t_vals, (blur + F(resample_padding)).astype(F), np.asarray(t_vals).shape[-1], u_rand)


def volumetric_rendering(rgb, density, t_vals, dirs, white_bkgd=False):
    """mip.py:121-148 -> comp_rgb [N,3], distance [N], acc [N], weights [N,S]."""
    rgb, den, t, d = (np.asarray(a, F) for a in (rgb, density, t_vals, dirs))
    t_mids = (F(0.5) * (t[..., :-1] + t[..., 1:])).astype(F)
    delta = ((t[..., 1:] - t[..., :-1]).astype(F) * np.sqrt(np.sum(d.astype(np.float64) ** 2, -1)).astype(F)[..., None]).astype(F)
    dd = (den[..., 0] * delta).astype(F)
    alpha = (F(1) - np.exp(-dd).astype(F)).astype(F)
    csum = np.cumsum(dd[..., :-1].astype(np.float64), axis=-1).astype(F)
    trans = np.exp(-np.concatenate([np.zeros_like(dd[..., :1]), csum], -1)).astype(F)
    w = (alpha * trans).astype(F)
    comp = np.sum((w[..., None] * rgb).astype(F), axis=-2, dtype=np.float64).astype(F)
    acc = np.sum(w, axis=-1, dtype=np.float64).astype(F)
    dist = np.sum((w * t_mids).astype(F), axis=-1, dtype=np.float64).astype(F)
    dist = np.clip(np.nan_to_num(dist, nan=np.inf), t[:, 0], t[:, -1]).astype(F)
    if white_bkgd:
        comp = (comp + (F(1) - acc[..., None])).astype(F)
    return comp, dist, acc, w
