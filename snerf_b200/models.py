"""`MipNerfModel`-shaped front door for the fused renderer.

The reference's `train.py` / `eval.py` do not call `render_rays` directly: they build a model with
`make_mipnerf(args, device)`, call `model(rays: Rays, randomized, white_bg, viewc)` (`models.py:72`, `train.py:112-129`)
and render test images with `models.render_image(render_fn, rays, rank, chunk)` (`models.py:328-360`, `eval.py:146`).
This module offers the same three names with the same call / return structure on top of the vanilla hot path
(`snerf_b200.render.render_rays`), so those scripts can be pointed at it:

    model(rays, randomized, white_bg, viewc) -> [[rgb_c, dist_c, acc_c(, s, w)], [rgb_f, dist_f, acc_f, None(, s, w)]]

It is an interface adapter, not a mip-NeRF implementation: the integrated positional encoding, cone casting, scene
contraction and proposal network of the reference's mip path are out of scope; `radii`, `lossmult`
and `app` of `Rays` are accepted and ignored, `distance` is the expected depth of the vanilla compositor.
"""
from __future__ import annotations

import collections

import torch
from torch import nn

from .render import make_query_fn, render_rays
from .run_nerf_helpers import NeRF

# same field order as utils/sample_utils.py:11-13
Rays = collections.namedtuple('Rays', ('origins', 'directions', 'viewdirs', 'radii', 'lossmult', 'near', 'far', 'app'))


def namedtuple_map(fn, tup):
    return type(tup)(*[None if x is None else fn(x) for x in tup])


class FusedNerfModel(nn.Module):
    """Coarse + fine `NeRF` behind `MipNerfModel.forward`'s signature and return structure."""

    def __init__(self, n_samples: int = 64, N_fine: int = 128, use_viewdirs: bool = True, lindisp: bool = False,
                 density_noise: float = 1., proposal_loss: bool = False, netdepth: int = 8, netwidth: int = 256,
                 multires: int = 10, multires_views: int = 4, **mip_only_kwargs):
        super().__init__()
        self.n_samples, self.N_fine = n_samples, N_fine
        self.use_viewdirs, self.lindisp = use_viewdirs, lindisp
        self.density_noise, self.proposal_loss = density_noise, proposal_loss
        self.ignored_kwargs = sorted(mip_only_kwargs)  # ray_shape, fn, radius, max_deg_point, hidden_layer, ...
        self.query_fn, in_ch, in_ch_views = make_query_fn(multires, multires_views, 0, use_viewdirs)
        out_ch = 5 if N_fine > 0 else 4
        self.network_fn = NeRF(D=netdepth, W=netwidth, input_ch=in_ch, input_ch_views=in_ch_views, output_ch=out_ch,
                               skips=[4], use_viewdirs=use_viewdirs)
        self.network_fine = (NeRF(D=netdepth, W=netwidth, input_ch=in_ch, input_ch_views=in_ch_views, output_ch=out_ch,
                                  skips=[4], use_viewdirs=use_viewdirs) if N_fine > 0 else None)

    def forward(self, rays, randomized, white_bg, viewc=None):
        o = rays.origins.reshape(-1, 3).float()
        d = rays.directions.reshape(-1, 3).float()
        cols = [o, d, rays.near.reshape(-1, 1).float(), rays.far.reshape(-1, 1).float()]
        if self.use_viewdirs:
            cols.append(rays.viewdirs.reshape(-1, 3).float())
        batch = torch.cat(cols, -1)
        want = ("depth0",) + (("z_all", "weights_fine") if self.proposal_loss else ())
        out = render_rays(batch, self.network_fn, self.query_fn, self.n_samples, lindisp=self.lindisp,
                          perturb=1. if randomized else 0., N_importance=self.N_fine, network_fine=self.network_fine,
                          white_bkgd=bool(white_bg),
                          raw_noise_std=float(self.density_noise) if randomized else 0., _outputs=want)
        if self.N_fine > 0:
            coarse = [out["rgb0"], out["depth0"], out["acc0"]]
            fine = [out["rgb_map"], out["depth_map"], out["acc_map"], None]  # 4th slot: semantic logits (none here)
            if self.proposal_loss:
                # ProposalLoss wants interval EDGES [N, S+1] with one weight per interval (loss_factory.py:59-73).  The
                # vanilla compositor's weight i belongs to [z_i, z_{i+1}) and its last interval is open (dist 1e10,
                # run_nerf_helpers.py:397): the S depths are the edges of the S-1 closed intervals, the last weight is dropped.
                coarse += [out["z_vals_map"], out["weights"][:, :-1]]
                fine += [out["z_all"], out["weights_fine"][:, :-1]]
            return [coarse, fine]
        single = [out["rgb_map"], out["depth_map"], out["acc_map"]]
        return [single + ([out["z_vals_map"], out["weights"][:, :-1]] if self.proposal_loss else []), single + [None]]


def make_fused_nerf(args, device):
    """`make_mipnerf(args, device)`-shaped factory (models.py:190-198): reads the flags the vanilla path understands."""
    g = lambda name, default: getattr(args, name, default)
    model = FusedNerfModel(n_samples=g("N_samples", 64), N_fine=g("N_fine", 128), use_viewdirs=g("use_viewdirs", True),
                           lindisp=g("lindisp", False), density_noise=g("density_noise", 1.),
                           proposal_loss=g("proposal_loss", False), netdepth=g("netdepth", 8), netwidth=g("netwidth", 256),
                           multires=g("multires", 10), multires_views=g("multires_views", 4))
    return model.to(device)


def render_image(render_fn, rays, rank=0, chunk=8192):
    """Render every pixel of `rays` ([H, W, ...] fields) -> (rgb[H,W,3], distance[H,W], acc[H,W], semantic) like
    models.py:328-360.  The fused kernel keeps nothing per sample in HBM, so `chunk=None` renders the frame in one
    launch (the reference needs 352 chunks of 4096 rays per 1600x900 image)."""
    height, width = rays.origins.shape[:2]
    n = height * width
    flat = namedtuple_map(lambda r: r.reshape(n, -1), rays)
    step = n if chunk is None else chunk
    parts = []
    for i in range(0, n, step):
        parts.append(render_fn(namedtuple_map(lambda r: r[i:i + step], flat))[-1][:3])
    rgb, distance, acc = [torch.cat(p, 0) for p in zip(*parts)]
    return rgb.reshape(height, width, -1), distance.reshape(height, width), acc.reshape(height, width), None
