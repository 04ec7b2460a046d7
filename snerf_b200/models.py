"""`MipNerfModel`-shaped front door for the fused renderer.

The reference's `train.py` / `eval.py` do not call `render_rays` directly: they build a model with
`make_mipnerf(args, device)`, call `model(rays: Rays, randomized, white_bg, viewc)` (`models.py:72`, `train.py:112-129`)
and render test images with `models.render_image(render_fn, rays, rank, chunk)` (`models.py:328-360`, `eval.py:146`).
This module offers the same three names with the same call / return structure on top of the vanilla hot path
(`snerf_b200.render.render_rays`), so those scripts can be pointed at it:

    model(rays, randomized, white_bg, viewc) -> [[rgb_c, dist_c, acc_c(, s, w)], [rgb_f, dist_f, acc_f, None(, s, w)]]

`FusedNerfModel` is an interface adapter over the VANILLA hot path (`radii`, `lossmult`, `app` of `Rays` are accepted
and ignored, `distance` is the expected depth of the vanilla compositor).

`MipNerfModel` / `MLP` / `proposal` / `DenseBlock` / `make_mipnerf` are the reference's own model (models.py:10-325) on the
warp path its shipped config runs (no_warp_sample = 0, fn = 1, ray_shape = 'cone'): same constructor arguments, same
parameter names (a reference checkpoint's `model_param` loads with `load_state_dict`), same `forward` return structure;
every tensor operation of the forward runs in csrc/snerf_mip.cu (sampling + integrated positional encoding, tcgen05
layer GEMMs, compositing + resampling).  Inference only: the reference trains this model through torch autograd, which
the kernels do not record.
"""
from __future__ import annotations

import collections
import ctypes as C
import warnings

import torch
from torch import nn

from . import _lib
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


# ======================================================================================================================
# the reference's mip-NeRF model (s-nerf/model/models.py) on csrc/snerf_mip.cu
# ======================================================================================================================
class DenseBlock(nn.Module):
    """nn.Linear + ReLU with Xavier-uniform weights (models.py:200-215); parameter names `layers.0.weight / bias`."""

    def __init__(self, input_channnels, output_channels: int = 256):
        super().__init__()
        lin = nn.Linear(input_channnels, output_channels)
        torch.nn.init.xavier_uniform_(lin.weight)
        self.layers = nn.Sequential(lin, nn.ReLU(inplace=True))

    @property
    def linear(self):
        return self.layers[0]


class MLP(nn.Module):
    """Parameter container of the reference's `MLP` (models.py:217-297): 8 DenseBlocks with `[x, inputs]` concatenated
    after layer `skip_layer`, density head, bottleneck, condition layers on `[bottleneck, condition]`, rgb head."""

    def __init__(self, n_layers: int = 8, n_units: int = 256, n_layers_condition: int = 1, n_units_condition: int = 128,
                 skip_layer: int = 4, n_rgb_channels: int = 3, n_density_channels: int = 1, feature_dim: int = 96,
                 cond_dim: int = 27, condition=None, semantic=False, semantic_class_num=0):
        super().__init__()
        if semantic:
            raise RuntimeError("snerf_b200.models.MLP: the semantic head is not part of the ported path")
        self.n_layers, self.skip_layer, self.n_units, self.feature_dim = n_layers, skip_layer, n_units, feature_dim
        self.n_units_condition, self.cond_dim = n_units_condition, cond_dim
        self.layers = nn.ModuleList([DenseBlock(feature_dim, n_units)])
        for i in range(n_layers - 1):
            self.layers.append(DenseBlock(feature_dim + n_units if (i % skip_layer == 0 and i > 0) else n_units, n_units))
        self.density_layer = nn.Linear(n_units, n_density_channels)
        self.bottleneck_layer = DenseBlock(n_units, n_units)
        self.cond_layers = nn.Sequential(*[DenseBlock(n_units + cond_dim if i == 0 else n_units_condition, n_units_condition)
                                           for i in range(n_layers_condition)])
        self.rgb_layer = nn.Linear(n_units_condition, n_rgb_channels)
        torch.nn.init.xavier_uniform_(self.density_layer.weight)
        torch.nn.init.xavier_uniform_(self.rgb_layer.weight)


class proposal(nn.Module):
    """Parameter container of the reference's `proposal` network (models.py:300-325): 4 DenseBlocks + density head."""

    def __init__(self, n_units=256, n_layers: int = 4, n_density_channels: int = 1, feature_dim: int = 96):
        super().__init__()
        self.n_layers, self.n_units, self.feature_dim = n_layers, n_units, feature_dim
        self.layers = nn.ModuleList([DenseBlock(feature_dim, n_units)] + [DenseBlock(n_units, n_units) for _ in range(n_layers - 1)])
        self.density_layer = nn.Linear(n_units, n_density_channels)


def _round_up(x, m):
    return (x + m - 1) // m * m


class MipNerfModel(nn.Module):
    """The reference's `MipNerfModel` (models.py:10-187), same constructor and `forward(rays, randomized, white_bg, viewc)`
    -> [[None, distance, acc(, s_vals, weights)], [rgb, distance, acc, None(, s_vals, weights)]].

    Supported configuration = the one the reference can run and ships: `no_warp_sample=0` (with 1 the reference's forward
    uses an undefined `s_vals`, models.py:177), `fn=1` (mip-360 contraction; `viewc` is then unused), view directions,
    no appearance embedding, no semantic head.  `white_bg=True` raises inside the reference (mip.py:188 adds to the
    proposal level's `None` colour); here it whitens the final level only."""

    def __init__(self, n_samples: int = 128, n_levels: int = 2, resample_padding: float = 0.01, stop_level_grad: bool = True,
                 use_viewdirs: bool = True, lindisp: bool = False, ray_shape: str = "cylinder", min_deg_point: int = 0,
                 max_deg_point: int = 16, deg_view: int = 4, density_noise: float = 1., density_bias: float = -1.,
                 rgb_padding: float = 0.001, disable_integration: bool = False, no_warp_sample=True, fn=None, radius=None,
                 real=False, transform_idx=0, rgb_layer=1, hidden_layer=256, encode_appearance=False, N_vocab=100,
                 proposal_hidden_layer=256, proposal_loss=False, N_fine=128, semantic=False, semantic_class_num=0):
        super().__init__()
        problems = []
        if no_warp_sample: problems.append("no_warp_sample must be 0 (the reference's forward is broken otherwise)")
        if fn != 1: problems.append("fn must be 1 (mip-360 contraction)")
        if not use_viewdirs: problems.append("use_viewdirs must be True")
        if encode_appearance or semantic: problems.append("appearance embedding / semantic head are not ported")
        if disable_integration or min_deg_point != 0 or n_levels != 2: problems.append("disable_integration / min_deg_point / n_levels")
        if ray_shape not in ("cone", "cylinder"): problems.append("ray_shape")
        if hidden_layer % 128 or proposal_hidden_layer % 128 or 6 * max_deg_point > 128:
            problems.append("hidden widths must be multiples of 128 and 6 * max_deg_point <= 128")
        if problems:
            raise RuntimeError("snerf_b200.models.MipNerfModel: unsupported configuration: " + "; ".join(problems))
        self.n_levels, self.stop_level_grad, self.deg_view, self.n_samples = n_levels, stop_level_grad, deg_view, n_samples
        self.lindisp, self.ray_shape, self.resample_padding = lindisp, ray_shape, resample_padding
        self.min_deg_point, self.max_deg_point, self.use_viewdirs = min_deg_point, max_deg_point, use_viewdirs
        self.density_noise, self.rgb_padding, self.density_bias = density_noise, rgb_padding, density_bias
        self.no_warp_sample, self.fn, self.radius, self.real, self.transform_idx = no_warp_sample, fn, radius, real, transform_idx
        self.mlp = MLP(feature_dim=max_deg_point * 6, n_layers_condition=rgb_layer, n_units=hidden_layer, cond_dim=3 + 6 * deg_view)
        self.proposal = proposal(n_units=proposal_hidden_layer, feature_dim=max_deg_point * 6)
        self.proposal_loss, self.semantic, self.N_fine = proposal_loss, semantic, N_fine
        self.max_rows = 1 << 21          # rows (ray samples) per internal chunk: 2 M rows x (2 x hidden + 128) bf16
        self._packed = None

    # ---- packed bf16 weights (rebuilt when a parameter changed)
    def invalidate_packed(self):
        self._packed = None

    def _pack(self):
        ps = list(self.parameters())
        stamp = tuple((p.data_ptr(), p._version) for p in ps)
        if self._packed is not None and self._packed["stamp"] == stamp:
            return self._packed
        if ps[0].device.type != "cuda":
            raise RuntimeError("snerf_b200.models.MipNerfModel: parameters must be on a CUDA device (no CPU fallback)")

        def mat(w, segs):
            """[n, sum(real)] fp32 -> bf16 [round_up(n, 128), sum(padded)]: every K segment zero-padded to a multiple of 64"""
            cols, off = [], 0
            for real, padded in segs:
                cols.append(torch.nn.functional.pad(w[:, off:off + real], (0, padded - real)))
                off += real
            out = torch.cat(cols, 1)
            out = torch.nn.functional.pad(out, (0, 0, 0, _round_up(out.shape[0], 128) - out.shape[0]))
            return out.to(torch.bfloat16).contiguous()

        with torch.no_grad():
            F_, H, Hp = self.mlp.feature_dim, self.mlp.n_units, self.proposal.n_units
            def f32(t):        # fp32, contiguous, 16-byte aligned (the GEMM epilogue reads biases / heads as 16-byte vectors)
                t = t.detach().float().contiguous()
                return t.clone() if t.data_ptr() % 16 else t
            P = {"stamp": stamp, "prop": [], "mlp": [], "cond": []}
            for i, blk in enumerate(self.proposal.layers):
                lin = blk.linear
                P["prop"].append((mat(lin.weight.float(), [(F_, 128)] if i == 0 else [(Hp, Hp)]), f32(lin.bias)))
            P["prop_head"] = (f32(self.proposal.density_layer.weight), float(self.proposal.density_layer.bias.detach().float()[0]))
            for i, blk in enumerate(self.mlp.layers):
                lin = blk.linear
                skip_in = lin.weight.shape[1] == H + F_
                segs = [(F_, 128)] if i == 0 else ([(H, H), (F_, 128)] if skip_in else [(H, H)])
                P["mlp"].append((mat(lin.weight.float(), segs), f32(lin.bias), skip_in))
            P["mlp_head"] = (f32(self.mlp.density_layer.weight), float(self.mlp.density_layer.bias.detach().float()[0]))
            bl = self.mlp.bottleneck_layer.linear
            P["bottleneck"] = (mat(bl.weight.float(), [(H, H)]), f32(bl.bias))
            for j, blk in enumerate(self.mlp.cond_layers):
                lin = blk.linear
                if j == 0:
                    P["cond"].append((mat(lin.weight.float()[:, :H], [(H, H)]), None))
                    P["cond0_full"] = (f32(lin.weight), f32(lin.bias))
                else:
                    P["cond"].append((mat(lin.weight.float(), [(lin.weight.shape[1], lin.weight.shape[1])]), f32(lin.bias)))
            P["rgb_head"] = (f32(self.mlp.rgb_layer.weight), [float(v) for v in self.mlp.rgb_layer.bias.detach().float()])
        self._packed = P
        return P

    # ---- one layer on the tensor cores
    @staticmethod
    def _linear(lib, st, a0, k0, w, n, bias, out, m_rows, m_pad, relu=True, a1=None, k1=0, ray_bias=None, rows_per_ray=1,
                head_w=None, head_out=None):
        L = _lib.Linear()
        L.a0, L.lda0, L.k0 = a0.data_ptr(), a0.stride(0), k0
        L.a1, L.lda1, L.k1 = (a1.data_ptr() if a1 is not None else None), (a1.stride(0) if a1 is not None else 0), k1
        L.w, L.n, L.n_pad = w.data_ptr(), n, w.shape[0]
        L.bias = bias.data_ptr() if bias is not None else None
        L.ray_bias, L.rows_per_ray = (ray_bias.data_ptr() if ray_bias is not None else None), rows_per_ray
        L.relu = 1 if relu else 0
        L.out, L.ldo = (out.data_ptr() if out is not None else None), (out.stride(0) if out is not None else 0)
        L.head_w, L.n_heads = (head_w.data_ptr() if head_w is not None else None), (head_w.shape[0] if head_w is not None else 0)
        L.head_out = head_out.data_ptr() if head_out is not None else None
        L.m_rows, L.m_pad = m_rows, m_pad
        _lib.check(lib.snerf_linear_tc(C.byref(L), st), "snerf_linear_tc")

    def _forward_chunk(self, rays9, viewdirs, randomized, white_bg):
        lib = _lib.load()
        dev = rays9.device
        st = _lib.stream_ptr(dev)
        P = self._pack()
        N = rays9.shape[0]
        S0, NF = self.n_samples, self.N_fine
        S1 = NF - 1
        H, Hp = self.mlp.n_units, self.proposal.n_units
        f32 = dict(dtype=torch.float32, device=dev)
        eps = torch.finfo(torch.float32).eps

        def encode(S, s_in, s_rand, m_pad):
            enc = torch.empty((m_pad, 128), dtype=torch.bfloat16, device=dev)
            s_out = None if s_in is not None else torch.empty((N, S + 1), **f32)
            e = _lib.MipEncode()
            e.rays, e.n_rays, e.n_samples, e.rows_per_ray = rays9.data_ptr(), N, S, S
            s_lin = torch.linspace(0., 1., S + 1, device=dev) if s_in is None else None
            e.s_lin = s_lin.data_ptr() if s_lin is not None else None
            e.s_rand = s_rand.data_ptr() if s_rand is not None else None
            e.s_in = s_in.data_ptr() if s_in is not None else None
            e.s_out = s_out.data_ptr() if s_out is not None else None
            e.transform_idx, e.max_deg, e.ray_cone, e.radius = int(self.transform_idx), self.max_deg_point, int(self.ray_shape == "cone"), 3.0
            e.enc, e.enc_f32, e.m_pad = enc.data_ptr(), None, m_pad
            _lib.check(lib.snerf_mip_encode(C.byref(e), st), "snerf_mip_encode")
            return enc, (s_in if s_in is not None else s_out)

        def composite(S, s_vals, dens, rgb, head_b, rgb_b, noise, resample):
            c = _lib.MipComposite()
            c.rays, c.n_rays, c.n_samples, c.rows_per_ray = rays9.data_ptr(), N, S, S
            c.s_vals, c.raw_density = s_vals.data_ptr(), dens.data_ptr()
            c.raw_rgb = rgb.data_ptr() if rgb is not None else None
            c.noise = noise.data_ptr() if noise is not None else None
            c.density_head_bias, c.density_bias, c.rgb_padding = head_b, float(self.density_bias), float(self.rgb_padding)
            for k in range(3):
                c.rgb_head_bias[k] = rgb_b[k] if rgb_b is not None else 0.0
            c.transform_idx, c.white_bkgd = int(self.transform_idx), int(bool(white_bg) and rgb is not None)
            comp = torch.empty((N, 3), **f32) if rgb is not None else None
            dist, acc, wts = torch.empty(N, **f32), torch.empty(N, **f32), torch.empty((N, S), **f32)
            c.comp_rgb = comp.data_ptr() if comp is not None else None
            c.distance, c.acc, c.weights = dist.data_ptr(), acc.data_ptr(), wts.data_ptr()
            keep = []
            s_new = None
            if resample:
                s_new = torch.empty((N, NF), **f32)
                c.n_fine, c.resample_padding, c.s_new = NF, float(self.resample_padding), s_new.data_ptr()
                if randomized:
                    u = torch.empty((N, NF), **f32).uniform_(0, 1.0 / NF - eps)      # math_ops.py:49-52
                    c.u_rand = u.data_ptr()
                else:
                    u = torch.linspace(0., 1. - eps, NF, device=dev)                 # math_ops.py:56
                    c.u_lin = u.data_ptr()
                keep.append(u)
            _lib.check(lib.snerf_mip_composite(C.byref(c), st), "snerf_mip_composite")
            return comp, dist, acc, wts, s_new

        with torch.cuda.device(dev):
            # ---------------- level 0: proposal network on stratified samples
            m0, m0p = N * S0, _round_up(N * S0, 128)
            s_rand = torch.rand((N, S0 + 1), **f32) if randomized else None                  # mip.py:283
            enc0, s0 = encode(S0, None, s_rand, m0p)
            bufs = [torch.empty((m0p, Hp), dtype=torch.bfloat16, device=dev) for _ in range(2)]
            dens0 = torch.zeros(m0p, **f32)
            x, k = enc0, 128
            nl = len(P["prop"])
            for i, (w, b) in enumerate(P["prop"]):
                last = i == nl - 1
                self._linear(lib, st, x, k, w, Hp, b, None if last else bufs[i & 1], m0, m0p,
                             head_w=P["prop_head"][0] if last else None, head_out=dens0 if last else None)
                x, k = bufs[i & 1], Hp
            noise0 = (self.density_noise * torch.randn((N, S0), **f32)) if (randomized and self.density_noise > 0) else None
            _, dist0, acc0, w0, s1 = composite(S0, s0, dens0, None, P["prop_head"][1], None, noise0, True)
            del bufs, enc0, dens0
            # ---------------- level 1: MLP on the resampled intervals
            m1, m1p = N * S1, _round_up(N * S1, 128)
            enc1, _ = encode(S1, s1, None, m1p)
            bufs = [torch.empty((m1p, H), dtype=torch.bfloat16, device=dev) for _ in range(2)]
            dens1, rgb1 = torch.zeros(m1p, **f32), torch.zeros((m1p, 3), **f32)
            x, k = enc1, 128
            nl = len(P["mlp"])
            for i, (w, b, skip_in) in enumerate(P["mlp"]):
                last = i == nl - 1
                out = bufs[i & 1]
                if skip_in:   # [x, inputs] (models.py:274-275): hidden state first, encoding second
                    self._linear(lib, st, x, H, w, H, b, out, m1, m1p, a1=enc1, k1=128,
                                 head_w=P["mlp_head"][0] if last else None, head_out=dens1 if last else None)
                else:
                    self._linear(lib, st, x, k, w, H, b, out, m1, m1p,
                                 head_w=P["mlp_head"][0] if last else None, head_out=dens1 if last else None)
                x, k = out, H
            hb = bufs[nl & 1]
            self._linear(lib, st, x, H, P["bottleneck"][0], H, P["bottleneck"][1], hb, m1, m1p)          # bottleneck DenseBlock
            # first condition layer: bottleneck half on the tensor cores, view-direction half as a per-ray bias
            Wc, bc = P["cond0_full"]
            Cn = Wc.shape[0]
            rbias = torch.empty((N, Cn), **f32)
            vd = viewdirs.float().contiguous()
            _lib.check(lib.snerf_mip_cond_bias(vd.data_ptr(), N, self.deg_view, Wc.data_ptr(), Wc.shape[1], H, bc.data_ptr(), Cn,
                                               rbias.data_ptr(), st), "snerf_mip_cond_bias")
            cb = [x[:, :Cn], x[:, Cn:2 * Cn]]          # the other ping-pong buffer is free now: two [M, 128] views of it
            nc = len(P["cond"])
            src, ksrc = hb, H
            for j, (w, b) in enumerate(P["cond"]):
                last = j == nc - 1
                self._linear(lib, st, src, ksrc, w, Cn, b, cb[j & 1], m1, m1p, ray_bias=rbias if j == 0 else None, rows_per_ray=S1,
                             head_w=P["rgb_head"][0] if last else None, head_out=rgb1 if last else None)
                src, ksrc = cb[j & 1], Cn
            noise1 = (self.density_noise * torch.randn((N, S1), **f32)) if (randomized and self.density_noise > 0) else None
            comp, dist1, acc1, w1, _ = composite(S1, s1, dens1, rgb1, P["mlp_head"][1], P["rgb_head"][1], noise1, False)
        coarse, fine = [None, dist0, acc0], [comp, dist1, acc1, None]
        if self.proposal_loss:
            coarse += [s0, w0]
            fine += [s1, w1]
        return [coarse, fine]

    def forward(self, rays, randomized, white_bg, viewc=None):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            warnings.warn("snerf_b200.models.MipNerfModel: the mip path is inference-only (its kernels record no autograd graph); "
                          "wrap the call in torch.no_grad()", UserWarning, stacklevel=2)
        o = rays.origins.reshape(-1, 3).float()
        if not o.is_cuda:
            raise RuntimeError("snerf_b200.models.MipNerfModel: rays must live on a CUDA sm_100 device (no CPU fallback)")
        n = o.shape[0]
        rays9 = torch.cat([o, rays.directions.reshape(-1, 3).float(), rays.radii.reshape(-1, 1).float(),
                           rays.near.reshape(-1, 1).float(), rays.far.reshape(-1, 1).float()], -1).contiguous()
        vd = rays.viewdirs.reshape(-1, 3)
        if n == 0:      # an empty batch: the reference returns empty tensors of the same shapes
            e = lambda *shape: torch.empty(shape, dtype=torch.float32, device=o.device)
            coarse, fine = [None, e(0), e(0)], [e(0, 3), e(0), e(0), None]
            if self.proposal_loss:
                coarse += [e(0, self.n_samples + 1), e(0, self.n_samples)]
                fine += [e(0, self.N_fine), e(0, self.N_fine - 1)]
            return [coarse, fine]
        per = max(1, self.max_rows // max(self.n_samples, self.N_fine))
        parts = [self._forward_chunk(rays9[i:i + per], vd[i:i + per], bool(randomized), white_bg) for i in range(0, n, per)]
        if len(parts) == 1:
            return parts[0]
        cat = lambda xs: None if xs[0] is None else torch.cat(xs, 0)
        return [[cat([p[l][k] for p in parts]) for k in range(len(parts[0][l]))] for l in range(2)]


def make_mipnerf(args, device):
    """make_mipnerf(args, device) of the reference (models.py:190-198): same argument names."""
    g = lambda name, default: getattr(args, name, default)
    model = MipNerfModel(no_warp_sample=g("no_warp_sample", 1), disable_integration=g("disable_integration", False),
                         ray_shape=g("ray_shape", "cone"), fn=g("fn", 1), max_deg_point=g("max_degree", 16), radius=g("radius", 3.),
                         transform_idx=g("transform_idx", 0), real=g("real", False), rgb_layer=g("rgb_layer", 1),
                         hidden_layer=g("hidden_layer", 256), density_noise=g("density_noise", 1.),
                         encode_appearance=g("encode_appearance", False), n_samples=g("N_samples", 128),
                         proposal_loss=g("proposal_loss", False), N_fine=g("N_fine", 128), semantic=g("semantic", False),
                         semantic_class_num=g("semantic_class_num", 0))
    return model.to(device)
