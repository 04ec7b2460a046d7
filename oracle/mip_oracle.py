"""TEST INFRASTRUCTURE ONLY -- numpy (fp32) restatement of the sampling / encoding / compositing functions of the mip-NeRF
path the shipped S-NeRF config trains with: s-nerf/model/mip.py -- sample_along_rays :192-212,
resample_along_rays :215-238, cast_rays :80-91 with conical_frustum_to_gaussian :56-70 and lift_gaussian :31-53 (diag),
integrated_pos_enc :94-118 with expected_sin :24-28, volumetric_rendering :121-148 -- and model/math_ops.py --
sorted_piecewise_constant_pdf :19-76, safe_sin / safe_cos :6-16.

The second half restates the WARP path, the one the shipped config actually runs (no_warp_sample = 0, fn = 1,
transform_idx = 0, ray_shape = 'cone'; models.py:72-187): warp_sample_along_rays mip.py:268-291, sample2enc :375-390
(transform :393-400, cast_rays, contraction fn2 :364-367, Jacobi_g :339-358), integrated_pos_enc(diag=False) :104-114,
warp_resample_along_rays :294-320, real_volumetric_rendering :151-189, pos_enc :12-21, the proposal / MLP networks
models.py:217-325 and MipNerfModel.forward models.py:72-187 (`mip_forward`).

Pinned by tests/test_oracle_vs_reference_live.py (sweeps against the unmodified reference on torch-CPU) and by
tests/golden/mip_*.npz (outputs of the reference's own MipNerfModel.forward, oracle/make_golden_mip.py).  It is the
checker of the CUDA path in snerf_b200/csrc/snerf_mip.cu (tests/test_gpu_mip.py); the product never imports it.
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


# ======================================================================================================================
# warp path (no_warp_sample = 0): what configs/nuScenes_depth_6cams runs
# ======================================================================================================================
def transform(s, near, far, transform_idx=0):
    """mip.py:7-9, 393-400: 0 = log, 1 = disparity, 2 = linear spacing of s in [0, 1] between near and far."""
    s, near, far = np.asarray(s, F), np.asarray(near, F), np.asarray(far, F)
    if transform_idx == 0:
        return (near * np.exp((s * np.log((far / near).astype(F)).astype(F)).astype(F)).astype(F)).astype(F)
    if transform_idx == 1:
        return (F(1) / (((F(1) - s) / near).astype(F) + (s / far).astype(F)).astype(F)).astype(F)
    return ((near * (F(1) - s)).astype(F) + (far * s).astype(F)).astype(F)


def warp_s_vals(n_rays, num_samples, s_rand=None):
    """s_vals of warp_sample_along_rays (mip.py:268-288): linspace(0, 1, S+1), jittered between mid-points when randomized."""
    s = _linspace01(num_samples + 1)[None]
    if s_rand is not None:
        mids = (F(0.5) * (s[..., 1:] + s[..., :-1])).astype(F)
        upper = np.concatenate([mids, s[..., -1:]], -1)
        lower = np.concatenate([s[..., :1], mids], -1)
        s = (lower + ((upper - lower).astype(F) * np.asarray(s_rand, F)).astype(F)).astype(F)
    return np.broadcast_to(s, (n_rays, num_samples + 1)).astype(F)


def contract(x, radius=3.0):
    """fn2 of warp_fn (mip.py:364-367): mip-360 style contraction outside `radius`, x / radius inside."""
    x = np.asarray(x, F)
    l = (np.sqrt(np.sum(x.astype(np.float64) ** 2, -1)).astype(F) + F(1e-8))[..., None]
    out = ((F(2) - F(radius) / l) * x / l).astype(F)
    return np.where(l > F(radius), out, (x / F(radius)).astype(F)).astype(F)


def jacobi_g(x, radius=3.0):
    """Jacobi_g (mip.py:339-358): J = (-r ln^2 + 2 ln) I + (2 r ln^4 - 2 ln^3) x x^T outside the radius, I / r inside."""
    x = np.asarray(x, F)
    nrm = np.sqrt(np.sum(x.astype(np.float64) ** 2, -1)).astype(F)
    ln = (F(1) / (nrm + F(1e-5))).astype(F)
    L = (x[..., :, None] * x[..., None, :]).astype(F)
    I = np.eye(3, dtype=F)
    P1 = ((-F(radius) * ln ** 2 + F(2) * ln).astype(F))[..., None, None] * I
    P2 = ((F(2) * F(radius) * ln ** 4 - F(2) * ln ** 3).astype(F))[..., None, None] * L
    J1 = (P1 + P2).astype(F)
    l = (nrm + F(1e-5))[..., None, None]
    return np.where(l >= F(radius), J1, (I / F(radius)).astype(F)).astype(F)


def sample2enc(s_vals, origins, directions, radii, near, far, transform_idx=0, radius=3.0):
    """sample2enc (mip.py:375-390): s -> t -> conical frustums -> contracted means + J diag(cov) J^T (full 3x3)."""
    t = transform(s_vals, near, far, transform_idx)
    means, cov = cast_rays(t, origins, directions, radii)
    J = jacobi_g(means, radius)
    f_cov = np.einsum("...ai,...i,...ib->...ab", J.astype(np.float64), cov.astype(np.float64), J.astype(np.float64)).astype(F)
    return contract(means, radius), f_cov


def integrated_pos_enc_full(means, cov, min_deg, max_deg):
    """integrated_pos_enc(diag=False) (mip.py:104-118): basis = [2^i I], y = x @ basis, y_var = diag(basis^T cov basis)."""
    x, c = np.asarray(means, F), np.asarray(cov, F)
    scales = np.array([2.0 ** i for i in range(min_deg, max_deg)], F)
    y = (x[..., None, :] * scales[:, None]).astype(F).reshape(x.shape[:-1] + (-1,))
    diag = np.stack([c[..., 0, 0], c[..., 1, 1], c[..., 2, 2]], -1)
    y_var = (diag[..., None, :] * (scales[:, None] ** 2).astype(F)).astype(F).reshape(x.shape[:-1] + (-1,))
    xx = np.concatenate([y, (y + F(0.5 * np.pi)).astype(F)], -1)
    vv = np.concatenate([y_var, y_var], -1)
    return (np.exp((F(-0.5) * vv).astype(F)).astype(F) * _safe_trig(xx, np.sin)).astype(F)


def pos_enc(x, min_deg, max_deg):
    """pos_enc(append_identity=True) (mip.py:12-21): [x, sin(2^i x) (degree-major), sin(2^i x + pi/2)]."""
    x = np.asarray(x, F)
    scales = np.array([2.0 ** i for i in range(min_deg, max_deg)], F)
    xb = (x[..., None, :] * scales[:, None]).astype(F).reshape(x.shape[:-1] + (-1,))
    four = np.sin(np.concatenate([xb, (xb + F(0.5 * np.pi)).astype(F)], -1)).astype(F)
    return np.concatenate([x, four], -1).astype(F)


def real_volumetric_rendering(rgb, density, s_vals, dirs, near, far, white_bkgd=False, transform_idx=0):
    """mip.py:151-189: like volumetric_rendering with t = T(s); rgb may be None (proposal level) -> comp_rgb None."""
    t = transform(s_vals, near, far, transform_idx)
    den, d = np.asarray(density, F), np.asarray(dirs, F)
    t_mids = (F(0.5) * (t[..., :-1] + t[..., 1:])).astype(F)
    delta = ((t[..., 1:] - t[..., :-1]).astype(F) * np.sqrt(np.sum(d.astype(np.float64) ** 2, -1)).astype(F)[..., None]).astype(F)
    dd = (den[..., 0] * delta).astype(F)
    alpha = (F(1) - np.exp(-dd).astype(F)).astype(F)
    csum = np.cumsum(dd[..., :-1].astype(np.float64), axis=-1).astype(F)
    trans = np.exp(-np.concatenate([np.zeros_like(dd[..., :1]), csum], -1)).astype(F)
    w = (alpha * trans).astype(F)
    comp = None if rgb is None else np.sum((w[..., None] * np.asarray(rgb, F)).astype(F), axis=-2, dtype=np.float64).astype(F)
    acc = np.sum(w, axis=-1, dtype=np.float64).astype(F)
    dist = np.sum((w * t_mids).astype(F), axis=-1, dtype=np.float64).astype(F)
    dist = np.clip(np.nan_to_num(dist, nan=np.inf), t[:, 0], t[:, -1]).astype(F)
    if white_bkgd and comp is not None:
        comp = (comp + (F(1) - acc[..., None])).astype(F)
    return comp, dist, acc, w


def warp_resample_s(s_vals, weights, n_fine, resample_padding=0.01, u_rand=None):
    """New s_vals of warp_resample_along_rays (mip.py:294-313): blurred max-pool + padding, n_fine draws from the sorted pdf."""
    w = np.asarray(weights, F)
    pad = np.concatenate([w[..., :1], w, w[..., -1:]], -1)
    wmax = np.maximum(pad[..., :-1], pad[..., 1:])
    blur = (F(0.5) * (wmax[..., :-1] + wmax[..., 1:])).astype(F)
    return sorted_piecewise_constant_pdf(s_vals, (blur + F(resample_padding)).astype(F), n_fine, u_rand)


def _dense(x, w, b, relu=True):
    y = (x @ w.T + b).astype(F)
    return np.maximum(y, 0) if relu else y


def proposal_forward(P, enc):
    """proposal.forward (models.py:300-325): 4 x DenseBlock + density_layer -> raw_density [M, 1]."""
    x = enc
    i = 0
    while f"proposal.layers.{i}.layers.0.weight" in P:
        x = _dense(x, P[f"proposal.layers.{i}.layers.0.weight"], P[f"proposal.layers.{i}.layers.0.bias"])
        i += 1
    return _dense(x, P["proposal.density_layer.weight"], P["proposal.density_layer.bias"], relu=False)


def mlp_forward(P, enc, cond, skip_layer=4):
    """MLP.forward (models.py:263-297): 8 x DenseBlock, [x, inputs] concatenated AFTER layer 4, density head, bottleneck
    (DenseBlock), condition layers on [bottleneck, cond], rgb head.  enc [M, F], cond [M, C] -> raw_rgb [M,3], raw_density [M,1]."""
    x = enc
    i = 0
    while f"mlp.layers.{i}.layers.0.weight" in P:
        x = _dense(x, P[f"mlp.layers.{i}.layers.0.weight"], P[f"mlp.layers.{i}.layers.0.bias"])
        if i % skip_layer == 0 and i > 0:
            x = np.concatenate([x, enc], -1)
        i += 1
    raw_density = _dense(x, P["mlp.density_layer.weight"], P["mlp.density_layer.bias"], relu=False)
    x = _dense(x, P["mlp.bottleneck_layer.layers.0.weight"], P["mlp.bottleneck_layer.layers.0.bias"])
    x = np.concatenate([x, cond], -1)
    j = 0
    while f"mlp.cond_layers.{j}.layers.0.weight" in P:
        x = _dense(x, P[f"mlp.cond_layers.{j}.layers.0.weight"], P[f"mlp.cond_layers.{j}.layers.0.bias"])
        j += 1
    return _dense(x, P["mlp.rgb_layer.weight"], P["mlp.rgb_layer.bias"], relu=False), raw_density


def make_mip_params(seed, hidden_layer=1024, rgb_layer=3, proposal_hidden=256, max_deg=16, gain=1.0, density_gain=3.0,
                    density_shift=-3.0):
    """Seeded weights with the reference's parameter names and shapes (MipNerfModel of models.py:10-70 built with
    hidden_layer / rgb_layer / max_degree as make_mipnerf passes them).  Xavier-uniform weights like the reference's
    DenseBlock (models.py:207), uniform(-1/sqrt(fan_in), ..) biases like nn.Linear; `density_gain` scales the density heads
    and `density_shift` moves their biases so that, at metric scene scale (1.8 .. 110 m), rays are neither empty nor opaque after a few samples.  Regenerated from the seed wherever needed (never stored: 8.6 M values)."""
    rs = np.random.RandomState(seed)
    feat = 6 * max_deg
    P = {}

    def lin(name, out_f, in_f, g=gain):
        a = g * np.sqrt(6.0 / (in_f + out_f))
        P[name + ".weight"] = rs.uniform(-a, a, (out_f, in_f)).astype(F)
        P[name + ".bias"] = rs.uniform(-1, 1, (out_f,)).astype(F) / F(np.sqrt(in_f))

    for i in range(8):
        in_f = feat if i == 0 else (hidden_layer + feat if i == 5 else hidden_layer)
        lin(f"mlp.layers.{i}.layers.0", hidden_layer, in_f)
    lin("mlp.density_layer", 1, hidden_layer, density_gain)
    lin("mlp.bottleneck_layer.layers.0", hidden_layer, hidden_layer)
    for j in range(rgb_layer):
        lin(f"mlp.cond_layers.{j}.layers.0", 128, hidden_layer + 27 if j == 0 else 128)
    lin("mlp.rgb_layer", 3, 128)
    for i in range(4):
        lin(f"proposal.layers.{i}.layers.0", proposal_hidden, feat if i == 0 else proposal_hidden)
    lin("proposal.density_layer", 1, proposal_hidden, density_gain)
    for k in ("mlp.density_layer.bias", "proposal.density_layer.bias"):
        P[k] = (P[k] + F(density_shift)).astype(F)
    return P


def mip_forward(P, origins, directions, viewdirs, radii, near, far, n_samples=128, n_fine=128, randomized=False, white_bkgd=False,
                s_rand=None, u_rand=None, noise0=None, noise1=None, max_deg=16, deg_view=4, density_bias=-1.0, rgb_padding=0.001,
                resample_padding=0.01, transform_idx=0, radius=3.0):
    """MipNerfModel.forward (models.py:72-187) on the warp path.  Returns the reference's list
    [[None, distance0, acc0, s_vals0, weights0], [rgb, distance, acc, None, s_vals1, weights1]] plus a dict of intermediates.
    `s_rand` / `u_rand` / `noise*` replay the random draws of `randomized` (None = deterministic)."""
    o, d, vd, r = (np.asarray(a, F) for a in (origins, directions, viewdirs, radii))
    near, far = np.asarray(near, F), np.asarray(far, F)
    N = o.shape[0]
    inter = {}
    s0 = warp_s_vals(N, n_samples, s_rand)
    m0, c0 = sample2enc(s0, o, d, r, near, far, transform_idx, radius)
    e0 = integrated_pos_enc_full(m0, c0, 0, max_deg)
    raw_d0 = proposal_forward(P, e0.reshape(-1, e0.shape[-1])).reshape(N, n_samples, 1)
    if noise0 is not None:
        raw_d0 = (raw_d0 + np.asarray(noise0, F)).astype(F)
    den0 = softplus((raw_d0 + F(density_bias)).astype(F))
    _, dist0, acc0, w0 = real_volumetric_rendering(None, den0, s0, d, near, far, white_bkgd, transform_idx)
    s1 = warp_resample_s(s0, w0, n_fine, resample_padding, u_rand)
    m1, c1 = sample2enc(s1, o, d, r, near, far, transform_idx, radius)
    e1 = integrated_pos_enc_full(m1, c1, 0, max_deg)
    S1 = n_fine - 1
    cond = np.repeat(pos_enc(vd, 0, deg_view)[:, None, :], S1, 1).reshape(N * S1, -1)
    raw_rgb, raw_d1 = mlp_forward(P, e1.reshape(-1, e1.shape[-1]), cond)
    raw_rgb, raw_d1 = raw_rgb.reshape(N, S1, 3), raw_d1.reshape(N, S1, 1)
    if noise1 is not None:
        raw_d1 = (raw_d1 + np.asarray(noise1, F)).astype(F)
    rgb = ((F(1) / (F(1) + np.exp(-raw_rgb))).astype(F) * F(1 + 2 * rgb_padding) - F(rgb_padding)).astype(F)
    den1 = softplus((raw_d1 + F(density_bias)).astype(F))
    comp, dist1, acc1, w1 = real_volumetric_rendering(rgb, den1, s1, d, near, far, white_bkgd, transform_idx)
    inter.update(enc0=e0, raw_density0=raw_d0, enc1=e1, raw_rgb=raw_rgb, raw_density1=raw_d1, means0=m0, cov0=c0)
    return [[None, dist0, acc0, s0, w0], [comp, dist1, acc1, None, s1, w1]], inter


def softplus(x):
    """F.softplus (beta 1, threshold 20)."""
    x = np.asarray(x, F)
    return np.where(x > F(20), x, np.log1p(np.exp(np.minimum(x, F(20)))).astype(F)).astype(F)
