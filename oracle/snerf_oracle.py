"""CPU oracle for the S-NeRF `render_rays` hot path (TEST INFRASTRUCTURE ONLY).

This module is a from-scratch numpy restatement of the reference algorithm. It is
the *checker*, never the product: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it.  The product
path (`snerf_b200/`) never routes through it and fails loudly without its CUDA
library.

Parity status: the reference ships no tests or golden vectors for this path
 -- "parity unpinned by the reference".  The oracle is pinned
instead against outputs of the unmodified reference code imported in the build
container (`oracle/make_golden.py` -> `tests/golden/*.npz`,
`oracle/check_against_reference.py`).

Reference lines restated (all relative to /root/reference/s-nerf/model/):
  * stratified depths + jitter ........ render.py:330-354
  * positional encoding ............... run_nerf_helpers.py:22-70
  * MLP ............................... run_nerf_helpers.py:74-126
  * run_network (encode + concat) ..... run_nerf_helpers.py:460-474
  * sigma->alpha composite ............ run_nerf_helpers.py:381-424
  * inverse-CDF resampling ............ run_nerf_helpers.py:336-379
  * hierarchical driver ............... render.py:281-409
  * pinhole ray generator ............. run_nerf_helpers.py:247-258
  * render() ray packing .............. render.py:22-91

Numerics notes (measured against torch 2.11 CPU):
  * all elementwise arithmetic is IEEE fp32 round-to-nearest, no FMA contraction;
  * torch CPU cumsum / cumprod accumulate in fp64 and round every element to fp32;
  * torch.linspace(0,1,N) fp32 uses the symmetric two-sided formula below.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------- #
# small numeric helpers
# --------------------------------------------------------------------------- #
def linspace01(n: int) -> np.ndarray:
    """fp32 `torch.linspace(0., 1., n)` (render.py:330, run_nerf_helpers.py:345).

    torch computes step = (end-start)/(n-1) in fp32 and fills the lower half as
    start + step*i and the upper half as end - step*(n-1-i), each with a single
    rounding (the product is formed exactly).
    """
    if n == 1:
        return np.zeros(1, F32)
    step = F32(1.0) / F32(n - 1)
    i = np.arange(n)
    lo = (np.float64(step) * i).astype(F32)
    hi = (1.0 - np.float64(step) * (n - 1 - i)).astype(F32)
    return np.where(i < n // 2, lo, hi).astype(F32)


def _scan64(x: np.ndarray, op: str) -> np.ndarray:
    """Inclusive scan along the last axis with fp64 accumulation, fp32 outputs."""
    acc = np.cumsum(x.astype(np.float64), -1) if op == "sum" else np.cumprod(x.astype(np.float64), -1)
    return acc.astype(F32)


# --------------------------------------------------------------------------- #
# ray generation / packing (render.py:22-91, run_nerf_helpers.py:247-258)
# --------------------------------------------------------------------------- #
def pinhole_rays(H: int, W: int, focal: float, c2w: np.ndarray, ori_points=None):
    """Pixel-centre pinhole rays; returns (origins[H,W,3], dirs[H,W,3])."""
    c2w = np.asarray(c2w, F32)
    cx, cy = (W * 0.5, H * 0.5) if not ori_points else ori_points
    px = np.arange(W, dtype=F32)[None, :].repeat(H, 0)
    py = np.arange(H, dtype=F32)[:, None].repeat(W, 1)
    f = F32(focal)
    cam = np.stack([((px + F32(0.5)) - F32(cx)) / f,
                    -((py + F32(0.5)) - F32(cy)) / f,
                    -np.ones_like(px)], -1).astype(F32)
    # rays_d[k] = sum_j cam[j] * c2w[k, j]   (fp32 products, fp32 left-to-right sum)
    prod = cam[..., None, :] * c2w[:3, :3]
    d = ((prod[..., 0] + prod[..., 1]) + prod[..., 2]).astype(F32)
    o = np.broadcast_to(c2w[:3, 3], d.shape).astype(F32)
    return o, d


def pack_ray_batch(rays_o, rays_d, near, far, use_viewdirs=True, depths=None):
    """[N, 8|9|11|12] ray batch exactly as render() builds it (render.py:56-79)."""
    o = np.asarray(rays_o, F32).reshape(-1, 3)
    d = np.asarray(rays_d, F32).reshape(-1, 3)
    ones = np.ones_like(d[:, :1])
    cols = [o, d, F32(near) * ones, F32(far) * ones]
    if depths is not None:
        cols.append(np.asarray(depths, F32).reshape(-1, 1))
    if use_viewdirs:
        nrm = np.sqrt(((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]).astype(F32))
        cols.append((d / nrm[:, None]).astype(F32))
    return np.concatenate(cols, -1).astype(F32)


# --------------------------------------------------------------------------- #
# stage functions
# --------------------------------------------------------------------------- #
def stratified_depths(near, far, n, lindisp=False, t_rand=None):
    """Coarse sample depths z[N,n] (render.py:330-352)."""
    near = np.asarray(near, F32).reshape(-1, 1)
    far = np.asarray(far, F32).reshape(-1, 1)
    t = linspace01(n)[None, :]
    one = F32(1.0)
    if not lindisp:
        z = near * (one - t) + far * t
    else:
        z = one / (one / near * (one - t) + one / far * t)
    z = np.ascontiguousarray(np.broadcast_to(z, (near.shape[0], n))).astype(F32)
    if t_rand is not None:
        mid = F32(0.5) * (z[:, 1:] + z[:, :-1])
        hi = np.concatenate([mid, z[:, -1:]], -1)
        lo = np.concatenate([z[:, :1], mid], -1)
        z = (lo + (hi - lo) * np.asarray(t_rand, F32)).astype(F32)
    return z


def posenc(x: np.ndarray, n_freqs: int) -> np.ndarray:
    """[..., 3] -> [..., 3 + 6*n_freqs]: x, then per octave sin(2^k x), cos(2^k x)."""
    x = np.asarray(x, F32)
    parts = [x]
    for k in range(n_freqs):
        xf = x * F32(2.0 ** k)
        parts.append(np.sin(xf, dtype=F32))
        parts.append(np.cos(xf, dtype=F32))
    return np.concatenate(parts, -1).astype(F32)


def _linear(x, w, b):
    return (x @ w.T + b).astype(F32)


def mlp_forward(params: dict, enc_pts: np.ndarray, enc_dirs: np.ndarray | None, alpha_params: dict | None = None) -> np.ndarray:
    """NeRF.forward on already-encoded inputs; returns [M, 4] = (r, g, b, sigma) raw.

    `params` uses the reference's state_dict names (numpy fp32 arrays).
    Skip-concat after trunk layer index 4 is [encoding, hidden] (encoding first).
    """
    D = sum(1 for k in params if k.startswith("pts_linears.") and k.endswith(".weight"))
    skips = params.get("_skips", (4,))
    h = enc_pts
    for i in range(D):
        h = np.maximum(_linear(h, params[f"pts_linears.{i}.weight"], params[f"pts_linears.{i}.bias"]), F32(0))
        if i in skips:
            h = np.concatenate([enc_pts, h], -1)
    if "feature_linear.weight" in params:
        if "alpha_linear.weight" in params:
            sigma = _linear(h, params["alpha_linear.weight"], params["alpha_linear.bias"])
        else:  # NeRF_RGB: sigma from the frozen alpha_model (run_nerf_helpers.py:198-199)
            sigma = mlp_forward(alpha_params, enc_pts, enc_dirs)[:, 3:4]
        feat = _linear(h, params["feature_linear.weight"], params["feature_linear.bias"])
        hv = np.concatenate([feat, enc_dirs], -1)
        hv = np.maximum(_linear(hv, params["views_linears.0.weight"], params["views_linears.0.bias"]), F32(0))
        rgb = _linear(hv, params["rgb_linear.weight"], params["rgb_linear.bias"])
        return np.concatenate([rgb, sigma], -1).astype(F32)
    return _linear(h, params["output_linear.weight"], params["output_linear.bias"])


_BACKEND = {"mlp": "numpy"}


def set_backend(name: str, threads: int | None = None):
    """'numpy' (default; dependency-free checker) or 'torch' (the same restatement with the encode +
    MLP evaluated by torch CPU ops on all host threads -- what the reference's CPU path itself uses;
    this is the variant bench.py times as the CPU baseline)."""
    assert name in ("numpy", "torch")
    _BACKEND["mlp"] = name
    _BACKEND["threads"] = threads


def _query_network_torch(params, pts, viewdirs, multires, multires_views, chunk):
    import torch
    import torch.nn.functional as F
    torch.set_num_threads(_BACKEND.get("threads") or max(1, __import__("os").cpu_count()))
    tp = {k: torch.from_numpy(v) for k, v in params.items() if isinstance(v, np.ndarray)}
    N, S, _ = pts.shape

    def enc(x, L):
        parts = [x]
        for k in range(L):
            xf = x * float(2.0 ** k)
            parts += [torch.sin(xf), torch.cos(xf)]
        return torch.cat(parts, -1)

    with torch.no_grad():
        e = enc(torch.from_numpy(np.ascontiguousarray(pts)).reshape(-1, 3), multires)
        if viewdirs is not None:
            ed = enc(torch.from_numpy(np.ascontiguousarray(viewdirs)), multires_views)
            ed = ed[:, None, :].expand(N, S, ed.shape[-1]).reshape(N * S, -1)
        D = sum(1 for k in tp if k.startswith("pts_linears.") and k.endswith(".weight"))
        outs = []
        for s0 in range(0, e.shape[0], chunk):
            x = e[s0:s0 + chunk]
            h = x
            for i in range(D):
                h = F.relu(F.linear(h, tp[f"pts_linears.{i}.weight"], tp[f"pts_linears.{i}.bias"]))
                if i == 4 and D > 5:
                    h = torch.cat([x, h], -1)
            sigma = F.linear(h, tp["alpha_linear.weight"], tp["alpha_linear.bias"])
            feat = F.linear(h, tp["feature_linear.weight"], tp["feature_linear.bias"])
            hv = F.relu(F.linear(torch.cat([feat, ed[s0:s0 + chunk]], -1), tp["views_linears.0.weight"],
                                 tp["views_linears.0.bias"]))
            rgb = F.linear(hv, tp["rgb_linear.weight"], tp["rgb_linear.bias"])
            outs.append(torch.cat([rgb, sigma], -1))
        return torch.cat(outs, 0).reshape(N, S, 4).numpy()


def query_network(params, pts, viewdirs, multires=10, multires_views=4, chunk=1 << 16, alpha_params=None):
    """run_network: encode points (+ per-ray dirs broadcast over samples), run MLP."""
    if _BACKEND["mlp"] == "torch" and viewdirs is not None and "alpha_linear.weight" in params:
        return _query_network_torch(params, pts, viewdirs, multires, multires_views, chunk)
    N, S, _ = pts.shape
    e = posenc(pts.reshape(-1, 3), multires)
    ed = None
    if viewdirs is not None:
        ed = np.repeat(posenc(viewdirs, multires_views)[:, None, :], S, 1).reshape(N * S, -1)
    outs = []
    for s in range(0, e.shape[0], chunk):
        outs.append(mlp_forward(params, e[s:s + chunk], None if ed is None else ed[s:s + chunk], alpha_params))
    return np.concatenate(outs, 0).reshape(N, S, -1)


def composite(raw, z, rays_d, noise=None, white_bkgd=False):
    """raw2outputs: returns (rgb_map, disp_map, acc_map, weights, depth_map)."""
    raw = np.asarray(raw, F32); z = np.asarray(z, F32); rays_d = np.asarray(rays_d, F32)
    N, S = z.shape
    dist = np.concatenate([z[:, 1:] - z[:, :-1], np.full((N, 1), 1e10, F32)], -1)
    dn = np.sqrt(((rays_d[:, 0] * rays_d[:, 0] + rays_d[:, 1] * rays_d[:, 1])
                  + rays_d[:, 2] * rays_d[:, 2]).astype(F32))
    dist = (dist * dn[:, None]).astype(F32)
    rgb = (F32(1) / (F32(1) + np.exp(-raw[..., :3], dtype=F32))).astype(F32)
    sig = raw[..., 3] if noise is None else (raw[..., 3] + np.asarray(noise, F32))
    alpha = (F32(1) - np.exp(-np.maximum(sig, F32(0)) * dist, dtype=F32)).astype(F32)
    trans_in = np.concatenate([np.ones((N, 1), F32), (F32(1) - alpha) + F32(1e-10)], -1)
    trans = _scan64(trans_in, "prod")[:, :-1]
    w = (alpha * trans).astype(F32)
    rgb_map = (w[..., None] * rgb).sum(-2, dtype=F32)
    depth = (w * z).sum(-1, dtype=F32)
    acc = w.sum(-1, dtype=F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        disp = (F32(1) / np.maximum(F32(1e-10), depth / acc)).astype(F32)
    if white_bkgd:
        rgb_map = rgb_map + (F32(1) - acc[:, None])
    return rgb_map.astype(F32), disp, acc, w, depth


def pdf_cdf(weights):
    """cdf[N, B] with leading zero from interior weights[N, B-1] (run_nerf_helpers.py:338-341)."""
    w = (np.asarray(weights, F32) + F32(1e-5)).astype(F32)
    pdf = (w / w.sum(-1, keepdims=True, dtype=F32)).astype(F32)
    cdf = _scan64(pdf, "sum")
    return np.concatenate([np.zeros_like(cdf[:, :1]), cdf], -1)


def invert_cdf(bins, cdf, u):
    """searchsorted(right=True) + linear interpolation; returns (samples, inds int64)."""
    bins = np.asarray(bins, F32); cdf = np.asarray(cdf, F32); u = np.asarray(u, F32)
    N, B = cdf.shape
    if u.ndim == 1:
        u = np.broadcast_to(u, (N, u.shape[0]))
    inds = (cdf[:, None, :] <= u[:, :, None]).sum(-1).astype(np.int64)
    below = np.maximum(inds - 1, 0)
    above = np.minimum(inds, B - 1)
    cb = np.take_along_axis(cdf, below, 1); ca = np.take_along_axis(cdf, above, 1)
    bb = np.take_along_axis(bins, below, 1); ba = np.take_along_axis(bins, above, 1)
    den = (ca - cb).astype(F32)
    den = np.where(den < F32(1e-5), F32(1), den)
    t = ((u - cb) / den).astype(F32)
    return (bb + t * (ba - bb)).astype(F32), inds


def sample_pdf(bins, weights, n_samples, u=None):
    """Deterministic (u=None -> linspace) or caller-supplied-u importance resampling."""
    cdf = pdf_cdf(weights)
    if u is None:
        u = linspace01(n_samples)
    s, inds = invert_cdf(bins, cdf, u)
    return s, inds, cdf


# --------------------------------------------------------------------------- #
# the hot path
# --------------------------------------------------------------------------- #
def render_rays(ray_batch, net_coarse, net_fine, n_samples, n_importance=0, *,
                multires=10, multires_views=4, lindisp=False, white_bkgd=False,
                t_rand=None, u=None, noise0=None, noise1=None, retraw=False,
                return_intermediates=False, alpha_coarse=None, alpha_fine=None):
    """Full per-ray pipeline (render.py:281-409).  `t_rand`/`u`/`noise*` inject the
    random draws of the perturb / raw_noise_std paths (None = deterministic)."""
    rb = np.asarray(ray_batch, F32)
    o, d = rb[:, 0:3], rb[:, 3:6]
    vd = rb[:, -3:] if rb.shape[1] > 9 else None
    z = stratified_depths(rb[:, 6], rb[:, 7], n_samples, lindisp, t_rand)
    pts = (o[:, None, :] + d[:, None, :] * z[:, :, None]).astype(F32)
    raw = query_network(net_coarse, pts, vd, multires, multires_views, alpha_params=alpha_coarse)
    rgb, disp, acc, w, depth = composite(raw, z, d, noise0, white_bkgd)
    out = {"z_vals_map": z, "weights": w}
    inter = {"raw_coarse": raw}
    if n_importance > 0:
        out.update(rgb0=rgb, disp0=disp, acc0=acc)
        inter["depth0"] = depth
        z_mid = (F32(0.5) * (z[:, 1:] + z[:, :-1])).astype(F32)
        zs, inds, cdf = sample_pdf(z_mid, w[:, 1:-1], n_importance, u)
        z_all = np.sort(np.concatenate([z, zs], -1), -1)
        pts = (o[:, None, :] + d[:, None, :] * z_all[:, :, None]).astype(F32)
        raw = query_network(net_fine if net_fine is not None else net_coarse, pts, vd, multires, multires_views,
                            alpha_params=alpha_fine if net_fine is not None else alpha_coarse)
        rgb, disp, acc, w_f, depth = composite(raw, z_all, d, noise1, white_bkgd)
        mean = zs.mean(-1, dtype=np.float64)
        out["z_std"] = np.sqrt(((zs.astype(np.float64) - mean[:, None]) ** 2).mean(-1)).astype(F32)
        inter.update(cdf=cdf, inds=inds, z_samples=zs, z_all=z_all, raw_fine=raw, weights_fine=w_f)
    out.update(rgb_map=rgb, disp_map=disp, acc_map=acc, depth_map=depth)
    if retraw:
        out["raw"] = raw
    if return_intermediates:
        out["_inter"] = inter
    return out


# --------------------------------------------------------------------------- #
# deterministic synthetic networks (portable across torch versions)
# --------------------------------------------------------------------------- #
def make_nerf_params(seed: int, D=8, W=256, input_ch=63, input_ch_views=27, skips=(4,),
                     trunk_gain=1.0, sigma_bias=0.0):
    """nn.Linear-default-like init (U(-1/sqrt(fan_in), 1/sqrt(fan_in))) from a numpy
    legacy RandomState so every box regenerates bit-identical weights from the seed.
    Names/shapes follow the reference state_dict (run_nerf_helpers.py:75-101)."""
    rs = np.random.RandomState(seed)

    def lin(out_f, in_f, gain=1.0):
        b = 1.0 / np.sqrt(in_f)
        w = rs.uniform(-b, b, size=(out_f, in_f)) * gain
        bias = rs.uniform(-b, b, size=(out_f,))
        return w.astype(F32), bias.astype(F32)

    p = {}
    for i in range(D):
        in_f = input_ch if i == 0 else (W + input_ch if (i - 1) in skips else W)
        p[f"pts_linears.{i}.weight"], p[f"pts_linears.{i}.bias"] = lin(W, in_f, trunk_gain)
    p["views_linears.0.weight"], p["views_linears.0.bias"] = lin(W // 2, W + input_ch_views)
    p["feature_linear.weight"], p["feature_linear.bias"] = lin(W, W)
    p["alpha_linear.weight"], p["alpha_linear.bias"] = lin(1, W)
    p["alpha_linear.bias"] = (p["alpha_linear.bias"] + F32(sigma_bias)).astype(F32)
    p["rgb_linear.weight"], p["rgb_linear.bias"] = lin(3, W // 2)
    return p
