"""Differentiable CPU oracle of the S-NeRF `render_rays` path (TEST INFRASTRUCTURE ONLY).

The forward oracle (`snerf_oracle.py`) is numpy; training needs d(outputs)/d(network parameters),
which the reference obtains from torch autograd over its eager ops.  This module restates the
differentiable part of the path with torch CPU fp32 ops so autograd yields the checker's
gradients; the non-differentiable part (inverse-CDF resampling on detached weights,
render.py:379-381 `z_samples.detach()`) goes through the numpy oracle.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / reference legs may import
this.  Pinned by `tests/golden/grad_cfg3.npz` (gradients of the unmodified reference, produced by
`oracle/make_golden_grad.py`).

Reference lines restated (relative to /root/reference/s-nerf/model/):
  * MLP ............................... run_nerf_helpers.py:103-126
  * sigma->alpha composite ............ run_nerf_helpers.py:381-424
  * hierarchical driver ............... render.py:330-409
"""
from __future__ import annotations

import numpy as np
import torch

from . import snerf_oracle as O

F32 = np.float32


def params_to_torch(params: dict, requires_grad=True) -> dict:
    out = {}
    for k, v in params.items():
        if k.startswith("_"):
            continue
        t = torch.from_numpy(np.array(v, dtype=F32, copy=True))
        t.requires_grad_(requires_grad)
        out[k] = t
    return out


def mlp(P: dict, enc_pts: torch.Tensor, enc_dirs: torch.Tensor, skips=(4,), alpha_P: dict | None = None) -> torch.Tensor:
    """NeRF.forward (run_nerf_helpers.py:103-126): rows [M,63],[M,27] -> [M,4] (r,g,b,sigma).  Without viewdirs heads
    (`output_linear` in P) the first four columns of output_linear (:124).  NeRF_RGB (no `alpha_linear`, :189-206):
    sigma = alpha_model(x)[..., 3] under no_grad, from `alpha_P`."""
    D = sum(1 for k in P if k.startswith("pts_linears.") and k.endswith(".weight"))
    h = enc_pts
    for i in range(D):
        h = torch.relu(h @ P[f"pts_linears.{i}.weight"].t() + P[f"pts_linears.{i}.bias"])
        if i in skips:
            h = torch.cat([enc_pts, h], -1)
    if "output_linear.weight" in P:
        return (h @ P["output_linear.weight"].t() + P["output_linear.bias"])[..., :4]
    if "alpha_linear.weight" not in P:
        with torch.no_grad():
            sigma = mlp(alpha_P, enc_pts, enc_dirs, skips)[..., 3:4]
    else:
        sigma = h @ P["alpha_linear.weight"].t() + P["alpha_linear.bias"]
    feat = h @ P["feature_linear.weight"].t() + P["feature_linear.bias"]
    v = torch.relu(torch.cat([feat, enc_dirs], -1) @ P["views_linears.0.weight"].t() + P["views_linears.0.bias"])
    rgb = v @ P["rgb_linear.weight"].t() + P["rgb_linear.bias"]
    return torch.cat([rgb, sigma], -1)


def query(P: dict, pts: np.ndarray, viewdirs: np.ndarray, multires=10, multires_views=4, alpha_P=None) -> torch.Tensor:
    """run_network (run_nerf_helpers.py:460-474): pts [N,S,3], viewdirs [N,3] -> raw [N,S,4]."""
    N, S, _ = pts.shape
    e = torch.from_numpy(O.posenc(pts.reshape(-1, 3).astype(F32), multires))
    d = torch.from_numpy(O.posenc(np.repeat(viewdirs.astype(F32)[:, None, :], S, 1).reshape(-1, 3), multires_views))
    return mlp(P, e, d, alpha_P=alpha_P).reshape(N, S, 4)


def composite(raw: torch.Tensor, z: torch.Tensor, rays_d: torch.Tensor, noise=None, white_bkgd=False):
    """raw2outputs (run_nerf_helpers.py:381-424) -> rgb_map, disp_map, acc_map, weights, depth_map."""
    dists = z[..., 1:] - z[..., :-1]
    dists = torch.cat([dists, torch.full_like(dists[..., :1], 1e10)], -1)
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)
    rgb = torch.sigmoid(raw[..., :3])
    sig = raw[..., 3] if noise is None else raw[..., 3] + noise
    alpha = 1. - torch.exp(-torch.relu(sig) * dists)
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[..., :1]), 1. - alpha + 1e-10], -1), -1)[..., :-1]
    w = alpha * trans
    rgb_map = torch.sum(w[..., None] * rgb, -2)
    depth = torch.sum(w * z, -1)
    acc = torch.sum(w, -1)
    disp = 1. / torch.max(1e-10 * torch.ones_like(depth), depth / acc)
    if white_bkgd:
        rgb_map = rgb_map + (1. - acc[..., None])
    return rgb_map, disp, acc, w, depth


def render_rays(ray_batch: np.ndarray, Pc: dict, Pf: dict | None, n_samples: int, n_importance: int = 0, *,
                lindisp=False, white_bkgd=False, t_rand=None, u=None, noise0=None, noise1=None,
                z_all=None, alpha_c: dict | None = None, alpha_f: dict | None = None) -> dict:
    """render_rays (render.py:281-409); `Pc`/`Pf` are dicts of torch tensors (leaf parameters); `alpha_c`/`alpha_f` the
    frozen sigma networks of NeRF_RGB passes (render.py:361-371).
    `z_all` (optional [N, Nc+Nf]) replaces the merged depths: resampling is not differentiated and flips bins
    under 1-ulp changes of the cdf (snerf_oracle notes), so gradient checks fix the depths being compared."""
    rb = np.asarray(ray_batch, F32)
    o, d, vd = rb[:, 0:3], rb[:, 3:6], rb[:, -3:]
    z = O.stratified_depths(rb[:, 6], rb[:, 7], n_samples, lindisp, t_rand)
    pts = (o[:, None, :] + d[:, None, :] * z[:, :, None]).astype(F32)
    tz, td = torch.from_numpy(z), torch.from_numpy(d.copy())
    tn = lambda a: None if a is None else torch.from_numpy(np.asarray(a, F32))
    raw = query(Pc, pts, vd, alpha_P=alpha_c)
    rgb, disp, acc, w, depth = composite(raw, tz, td, tn(noise0), white_bkgd)
    out = dict(rgb_map=rgb, disp_map=disp, acc_map=acc, depth_map=depth, weights=w, raw=raw)
    if n_importance > 0:
        out.update(rgb0=rgb, disp0=disp, acc0=acc, depth0=depth, raw_coarse=raw)
        wn = w.detach().numpy()
        z_mid = (F32(0.5) * (z[:, 1:] + z[:, :-1])).astype(F32)
        zs, _, _ = O.sample_pdf(z_mid, wn[:, 1:-1], n_importance, u)
        z_all = np.sort(np.concatenate([z, zs], -1), -1) if z_all is None else np.asarray(z_all, F32)
        pts = (o[:, None, :] + d[:, None, :] * z_all[:, :, None]).astype(F32)
        raw_f = query(Pf if Pf is not None else Pc, pts, vd, alpha_P=alpha_f if Pf is not None else alpha_c)
        rgb, disp, acc, w_f, depth = composite(raw_f, torch.from_numpy(z_all), td, tn(noise1), white_bkgd)
        out.update(rgb_map=rgb, disp_map=disp, acc_map=acc, depth_map=depth, raw=raw_f, weights_fine=w_f)
        out["_z_all"] = z_all
    out["_z_vals"] = z
    return out


GRAD_KEYS = ("rgb_map", "disp_map", "acc_map", "depth_map", "rgb0", "disp0", "acc0", "weights")


def cotangents(shapes: dict, seed: int) -> dict:
    """Seeded upstream gradients for the linear functional L = sum_k <out_k, G_k> the gradient fixtures use."""
    rs = np.random.RandomState(seed)
    return {k: rs.standard_normal(shapes[k]).astype(F32) for k in GRAD_KEYS if k in shapes}


def loss_from(out: dict, G: dict) -> torch.Tensor:
    tot = None
    for k, g in G.items():
        term = (out[k] * torch.from_numpy(g)).sum()
        tot = term if tot is None else tot + term
    return tot


def variant_params(seed: int, kind: str, D=8, W=256, trunk_gain=1.5, sigma_bias=0.5) -> dict:
    """Seeded numpy state_dicts of the other network shapes `render_rays` trains: 'novd' = NeRF(use_viewdirs=False,
    output_ch=5) (trunk + output_linear, run_nerf_helpers.py:99-100); 'rgb' = NeRF_RGB (no alpha_linear, :157-186)."""
    base = O.make_nerf_params(seed, D=D, W=W, trunk_gain=trunk_gain, sigma_bias=sigma_bias)
    if kind == "rgb":
        return {k: v for k, v in base.items() if not k.startswith("alpha_linear")}
    assert kind == "novd"
    rs = np.random.RandomState(seed + 1000)
    b = 1.0 / np.sqrt(W)
    p = {k: v for k, v in base.items() if k.startswith("pts_linears.")}
    p["output_linear.weight"] = rs.uniform(-b, b, size=(5, W)).astype(F32)
    bias = rs.uniform(-b, b, size=(5,)).astype(F32)
    bias[3] += F32(sigma_bias)
    p["output_linear.bias"] = bias
    return p
