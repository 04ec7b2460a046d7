"""Host-side mirror of the reference's `model/render.py`: render, batchify_rays, render_rays, create_nerf.

Signatures, dict keys and return structure follow /root/reference/s-nerf/model/render.py
(line numbers cited per function).  `render_rays` is ONE launch of the fused sm_100a kernel
(`snerf_render_rays_fwd`): stratified sampling, encoding, coarse MLP, compositing, inverse-CDF
resampling, sort, fine MLP and the final composite never leave the SM; only the per-ray outputs
are written to HBM.  No CPU / PyTorch fallback exists.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from . import autograd as _autograd
from .run_nerf_helpers import (NeRF, NeRF_RGB, _MODE, _draw_noise, _draw_u, _f32c, _require_cuda, get_embedder, get_rays,
                               ndc_rays, run_network, to8b)

__all__ = ["batchify_rays", "render", "render_rays", "render_path", "create_nerf"]

_LINSPACE_CACHE = {}


def _linspace01(n: int, device) -> torch.Tensor:
    """torch.linspace(0,1,n) computed by torch itself (bit-identical to the reference), cached per device."""
    key = (n, str(device))
    t = _LINSPACE_CACHE.get(key)
    if t is None:
        t = torch.linspace(0., 1., steps=n, dtype=torch.float32).to(device)
        _LINSPACE_CACHE[key] = t
    return t


def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):
    """Render rays in chunks (render.py:8-19).  The fused kernel needs no chunking for memory (it has
    no per-sample HBM state); `chunk` is honoured so results and call pattern stay drop-in."""
    if chunk is None or chunk >= rays_flat.shape[0]:
        return render_rays(rays_flat, **kwargs)
    all_ret = {}
    for i in range(0, rays_flat.shape[0], chunk):
        ret = render_rays(rays_flat[i:i + chunk], **kwargs)
        for k in ret:
            all_ret.setdefault(k, []).append(ret[k])
    return {k: torch.cat(all_ret[k], 0) for k in all_ret}


def render(H, W, focal, chunk=1024 * 32, rays=None, c2w=None, ndc=True,
           near=0., far=1.,
           use_viewdirs=False, c2w_staticcam=None, depths=None, ori_points=None,
           **kwargs):
    """Image- or batch-level entry point (render.py:22-91) -> [rgb_map, disp_map, acc_map, depth_map, extras].

    Builds the `[N, 8|9|11|12]` ray batch (origin, direction, near, far, [depth], [unit view direction]) the kernel
    consumes, renders it (optionally in `chunk`s) and folds every output back to the leading shape of the rays."""
    if (c2w is not None and use_viewdirs and not ndc and c2w_staticcam is None and depths is None
            and not _autograd.wants_grad(kwargs.get("network_fn"), kwargs.get("network_fine"))):
        # get_rays (run_nerf_helpers.py:247-258) + the view-direction normalisation below run in the renderer's prologue:
        # the only ray input of the kernel is the camera (SnerfOpts.camera); nothing per ray is read from HBM
        return _render_camera(H, W, focal, c2w, ori_points, near, far, chunk, **kwargs)
    origins, dirs = get_rays(H, W, focal, c2w, ori_points) if c2w is not None else rays
    _require_cuda(dirs, "render")

    unit_dirs = None
    if use_viewdirs:
        src = dirs
        if c2w_staticcam is not None:  # fixed camera, moving view directions (visualisation of view dependence)
            origins, dirs = get_rays(H, W, focal, c2w_staticcam)
        unit_dirs = (src / src.norm(dim=-1, keepdim=True)).reshape(-1, 3).float()

    lead = list(dirs.shape[:-1])
    if ndc:  # forward-facing scenes
        origins, dirs = ndc_rays(H, W, focal, 1., origins, dirs)
    o_flat, d_flat = origins.reshape(-1, 3).float(), dirs.reshape(-1, 3).float()
    one = torch.ones_like(d_flat[:, :1])
    parts = [o_flat, d_flat, near * one, far * one]
    if depths is not None:
        parts.append(depths.reshape(-1, 1).to(d_flat))
    if unit_dirs is not None:
        parts.append(unit_dirs)
    ray_batch = torch.cat(parts, -1)

    flat = batchify_rays(ray_batch, chunk, **kwargs)
    shaped = {k: v.reshape(lead + list(v.shape[1:])) for k, v in flat.items()}
    main = ('rgb_map', 'disp_map', 'acc_map', 'depth_map')
    return [shaped[k] for k in main] + [{k: v for k, v in shaped.items() if k not in main}]


def _render_camera(H, W, focal, c2w, ori_points, near, far, chunk, **kwargs):
    """render() for a whole pinhole image without a ray batch: consecutive pixel ranges of `chunk` rays (one range when
    chunk is None or covers the image), each ONE launch that builds its rays from the camera."""
    if not ori_points:
        ori_points = [W * 0.5, H * 0.5]
    c2w_t = torch.as_tensor(c2w, dtype=torch.float32)
    dev = c2w_t.device if c2w_t.is_cuda else next(kwargs["network_fn"].parameters()).device
    m = np.ascontiguousarray(c2w_t.detach().cpu().numpy()[:3, :4], dtype=np.float32).reshape(-1)
    n = H * W
    step = n if (chunk is None or chunk >= n) else int(chunk)
    parts = {}
    for first in range(0, n, step):
        cam = _lib.Camera(H, W, float(np.float32(focal)), float(ori_points[0]), float(ori_points[1]), float(near), float(far),
                          (C.c_float * 12)(*m), first)
        ret = render_rays(None, _camera=(cam, min(step, n - first), dev), **kwargs)
        for k, v in ret.items():
            parts.setdefault(k, []).append(v)
    flat = {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in parts.items()}
    shaped = {k: v.reshape([H, W] + list(v.shape[1:])) for k, v in flat.items()}
    main = ('rgb_map', 'disp_map', 'acc_map', 'depth_map')
    return [shaped[k] for k in main] + [{k: v for k, v in shaped.items() if k not in main}]


def render_path(render_poses, hwf, chunk, render_kwargs, gt_imgs=None, savedir=None, render_factor=0,
                ori_points=None, render_masks=None):
    """render.py:94-135: loop over poses, return (rgbs, disps) as numpy stacks."""
    H, W, focal = hwf
    focal = np.array(focal).mean()
    if render_factor != 0:
        H, W, focal = H // render_factor, W // render_factor, focal / render_factor
    rgbs, disps = [], []
    for i, c2w in enumerate(render_poses):
        op = ori_points[i] if ori_points else None
        rgb, disp, acc, depth, extras = render(H, W, focal, ori_points=op, chunk=chunk, c2w=c2w[:3, :4], retraw=True,
                                               **render_kwargs)
        if render_masks:
            rgb[render_masks[i]] = 0
        rgbs.append(rgb.cpu().numpy())
        disps.append(disp.cpu().numpy())
    return np.stack(rgbs, 0), np.stack(disps, 0)


def render_rays(ray_batch,
                network_fn,
                network_query_fn,
                N_samples,
                retraw=False,
                lindisp=False,
                perturb=0.,
                N_importance=0,
                network_fine=None,
                white_bkgd=False,
                raw_noise_std=0.,
                verbose=False,
                pytest=False,
                _extras=False,
                _outputs=(),
                _camera=None):
    """Volumetric rendering of a ray batch (render.py:281-409), one fused kernel launch.

    Returns the reference's dict: rgb_map, disp_map, acc_map, depth_map, z_vals_map (coarse),
    weights (coarse), [raw], and with N_importance > 0: rgb0, disp0, acc0, z_std.
    `_extras=True` additionally returns the stage intermediates (tests); `_outputs` names individual extra
    buffers of SnerfOut to return as well (e.g. "depth0", "z_all", "weights_fine" for the MipNerfModel-shaped adapter).
    """
    if _camera is None:
        _require_cuda(ray_batch, "render_rays")
    # Which module serves each pass (render.py:359-371,387): `main` produces rgb (and sigma unless it is a NeRF_RGB,
    # whose sigma comes from its frozen `alpha_model`, evaluated on the same samples inside the same kernel).
    def split(net):
        return (net, net.alpha_model) if isinstance(net, NeRF_RGB) else (net, None)

    for m in (network_fn, network_fine):
        if m is not None and not isinstance(m, NeRF):
            raise RuntimeError("snerf_b200.render_rays: network_fn / network_fine must be snerf_b200.NeRF / NeRF_RGB "
                               "modules (no PyTorch fallback path)")
    if network_fn is None:
        if not isinstance(network_fine, NeRF_RGB):
            raise RuntimeError("snerf_b200.render_rays: network_fn=None requires network_fine to be a NeRF_RGB")
        coarse = (network_fine.alpha_model, None) if network_fine.alpha_model is not None else split(network_fine)
    else:
        coarse = split(network_fn)
    fine = split(network_fine if network_fine is not None else network_fn)
    for main, alpha in (coarse, fine):
        if isinstance(main, NeRF_RGB) and alpha is None:
            raise RuntimeError("snerf_b200.render_rays: NeRF_RGB without an alpha_model has no density")
    uses_alpha = coarse[1] is not None or fine[1] is not None
    network_fn, network_fine = coarse[0], (fine[0] if (network_fine is not None or fine[0] is not coarse[0]) else None)
    multires = getattr(network_query_fn, "multires", None)
    multires_views = getattr(network_query_fn, "multires_views", None)
    if multires is None:
        raise RuntimeError("snerf_b200.render_rays: network_query_fn must be the closure built by "
                           "snerf_b200.create_nerf / make_query_fn (it carries the embedder sizes)")
    if _camera is not None:       # (camera struct, rays in this launch, device): the kernel builds the rays itself
        cam, N, dev = _camera
        rb, width = None, 11
    else:
        dev = ray_batch.device
        rb = _f32c(ray_batch)
        N, width = rb.shape
    Nc, Nf = int(N_samples), int(N_importance)
    S = Nc + Nf
    d = network_fn.desc()
    mode = _MODE["mode"]
    df = None      # SnerfOpts.desc_fine: the fine network's own architecture (create_nerf: netdepth_fine / netwidth_fine)
    if network_fine is not None:
        df = network_fine.desc()
        differs = [f for f, _ in _lib.NetDesc._fields_ if getattr(d, f) != getattr(df, f)]
        if not differs:
            df = None
        elif any(f not in ("D", "W", "skip") for f in differs):
            raise RuntimeError("snerf_b200.render_rays: coarse and fine networks must agree on input_ch, input_ch_views, "
                               f"use_viewdirs and output_ch (differ in {differs})")
        elif mode != _lib.MODE_FP32 and not ((d.D, d.W, d.skip) == (4, 256, -1) and (df.D, df.W, df.skip) == (8, 256, 4)):
            raise RuntimeError("snerf_b200.render_rays: coarse and fine networks of different depth / width run in fp32 "
                               "mode; the tensor-core modes take 8x256 pairs and the pair coarse 4x256 + fine 8x256 "
                               "(netdepth = 4 / netdepth_fine = 8 of the shipped configs)")
    for a in (coarse[1], fine[1]):
        if a is not None and a.W > max(network_fn.W, network_fine.W if network_fine is not None else 0):
            raise RuntimeError("snerf_b200.render_rays: alpha_model is wider than the networks it serves")
    if uses_alpha and mode != _lib.MODE_FP32:
        raise RuntimeError("snerf_b200.render_rays: NeRF_RGB / alpha_model networks run in fp32 mode only")

    stochastic = perturb > 0.
    t_rand = u_rand = None
    if stochastic:
        if pytest:
            np.random.seed(0)
            t_rand = torch.Tensor(np.random.rand(N, Nc)).to(dev)
        else:
            t_rand = torch.rand((N, Nc), device=dev)
        if Nf > 0:
            u_rand = _draw_u([N], Nf, False, pytest, dev)
    noise0 = _draw_noise([N, Nc], raw_noise_std, pytest, dev)
    noise1 = _draw_noise([N, S], raw_noise_std, pytest, dev) if Nf > 0 else None

    from .run_nerf_helpers import _TRAIN
    tc_training = _TRAIN["precision"] == "bf16"     # set_train_precision: the tensor-core training step
    if _autograd.wants_grad(network_fn, network_fine) and mode != _lib.MODE_FP32 and not tc_training:
        _autograd.warn_inference_only()
    elif _autograd.wants_grad(network_fn, network_fine):
        if rb is None:
            raise RuntimeError("snerf_b200.render_rays: camera mode is inference-only (wrap the call in torch.no_grad())")
        # training: one autograd node around the fused forward (activations saved) and the backward kernels
        if (uses_alpha or not network_fn.use_viewdirs or df is not None) and _TRAIN["precision"] != "fp32":
            raise RuntimeError("snerf_b200.render_rays: NeRF_RGB / alpha_model networks, use_viewdirs=False and coarse / fine "
                               "networks of different depth or width train at set_train_precision('fp32') only")
        f = lambda t: None if t is None else _f32c(t)
        call = _autograd._Call(rb, network_fn, network_fine, multires, multires_views, Nc, Nf, lindisp, white_bkgd,
                               _linspace01(Nc, dev), _linspace01(Nf, dev) if Nf > 0 else None,
                               f(t_rand), f(u_rand), f(noise0), f(noise1), alpha_c=coarse[1], alpha_f=fine[1], desc_fine=df)
        res = _autograd.render_rays_train(call)
        keys = ["rgb_map", "disp_map", "acc_map", "depth_map", "z_vals_map", "weights"]
        if retraw:
            keys.append("raw")
        if Nf > 0:
            keys += ["rgb0", "disp0", "acc0", "z_std"]
        ret = {k: res[k] for k in keys}
        for name in _outputs:
            if name in res:
                ret[name] = res[name]
        if _extras:
            ret["_extras"] = {k: v for k, v in res.items() if k not in ret}
        return ret

    def new(*shape):
        return torch.empty(shape, dtype=torch.float32, device=dev)

    bufs = {"rgb_map": new(N, 3), "disp_map": new(N), "acc_map": new(N), "depth_map": new(N),
            "z_vals_map": new(N, Nc), "weights": new(N, Nc)}
    if Nf > 0:
        bufs.update(rgb0=new(N, 3), disp0=new(N), acc0=new(N), z_std=new(N))
    if retraw:
        bufs["raw"] = new(N, S, 4)
    if _extras:
        bufs["raw_coarse"] = new(N, Nc, 4)
        if Nf > 0:
            bufs.update(depth0=new(N), z_samples=new(N, Nf), z_all=new(N, S), weights_fine=new(N, S))
            bufs.setdefault("raw", new(N, S, 4))

    extra_shapes = {"depth0": (N,), "z_samples": (N, Nf), "z_all": (N, S), "raw_coarse": (N, Nc, 4),
                    "weights_fine": (N, S), "raw": (N, S, 4)}
    for name in _outputs:
        if name not in bufs and (Nf > 0 or name == "raw_coarse"):
            bufs[name] = new(*extra_shapes[name])

    rays = _lib.Rays(rb.data_ptr(), N, width, rb.stride(0)) if rb is not None else _lib.Rays(None, N, 11, 11)
    opts = _lib.Opts()
    if rb is None:
        opts.camera = C.cast(C.pointer(cam), C.c_void_p)
    if df is not None:
        opts.desc_fine = C.cast(C.pointer(df), C.c_void_p)
    opts.n_samples, opts.n_importance = Nc, Nf
    opts.lindisp, opts.white_bkgd, opts.mode = int(bool(lindisp)), int(bool(white_bkgd)), mode
    opts.multires, opts.multires_views = multires, (multires_views if multires_views is not None else 0)
    t_vals = _linspace01(Nc, dev)
    u_vals = _linspace01(Nf, dev) if Nf > 0 else None
    opts.t_vals, opts.u_vals = t_vals.data_ptr(), (u_vals.data_ptr() if u_vals is not None else None)
    for name, t in (("t_rand", t_rand), ("u_rand", u_rand), ("noise0", noise0), ("noise1", noise1)):
        setattr(opts, name, None if t is None else _f32c(t).data_ptr())
    out = _lib.Out()
    for k, t in bufs.items():
        setattr(out, k, t.data_ptr())

    lib = _lib.load()
    img_c = network_fn.packed(mode)
    img_f = network_fine.packed(mode) if network_fine is not None else None
    img_ac = coarse[1].packed(mode) if coarse[1] is not None else None
    img_af = fine[1].packed(mode) if fine[1] is not None else None
    opts.packed_alpha_coarse = img_ac.data_ptr() if img_ac is not None else None
    opts.packed_alpha_fine = img_af.data_ptr() if img_af is not None else None
    with torch.cuda.device(dev):
        _lib.check(lib.snerf_render_rays_fwd(C.byref(rays), C.byref(d), _lib.ptr(img_c), _lib.ptr(img_f),
                                             C.byref(opts), C.byref(out), None, 0, _lib.stream_ptr(dev)),
                   "snerf_render_rays_fwd")

    keys = ["rgb_map", "disp_map", "acc_map", "depth_map", "z_vals_map", "weights"]
    if retraw:
        keys.append("raw")
    if Nf > 0:
        keys += ["rgb0", "disp0", "acc0", "z_std"]
    ret = {k: bufs[k] for k in keys}
    if _extras:
        ret["_extras"] = {k: v for k, v in bufs.items() if k not in ret}
    for name in _outputs:
        if name in bufs:
            ret[name] = bufs[name]
    return ret


class _QueryFn:
    """network_query_fn(inputs, viewdirs, network_fn) closure of create_nerf (render.py:215-218);
    carries the embedder sizes so render_rays can hand them to the fused kernel."""

    def __init__(self, embed_fn, embeddirs_fn, netchunk):
        self.embed_fn, self.embeddirs_fn, self.netchunk = embed_fn, embeddirs_fn, netchunk
        self.multires = embed_fn.multires
        self.multires_views = embeddirs_fn.multires if embeddirs_fn is not None else None

    def __call__(self, inputs, viewdirs, network_fn):
        return run_network(inputs, viewdirs, network_fn, embed_fn=self.embed_fn, embeddirs_fn=self.embeddirs_fn,
                           netchunk=self.netchunk)


def make_query_fn(multires=10, multires_views=4, i_embed=0, use_viewdirs=True, netchunk=1024 * 64):
    embed_fn, input_ch = get_embedder(multires, i_embed)
    embeddirs_fn, input_ch_views = (get_embedder(multires_views, i_embed) if use_viewdirs else (None, 0))
    return _QueryFn(embed_fn, embeddirs_fn, netchunk), input_ch, input_ch_views


def create_nerf(args):
    """Instantiate the coarse / fine NeRF, optimizer and render kwargs (render.py:165-278).
    Returns (render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer, model_confidence)."""
    device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    network_query_fn, input_ch, input_ch_views = make_query_fn(args.multires, args.multires_views, args.i_embed,
                                                               args.use_viewdirs, args.netchunk)
    output_ch = 5 if args.N_importance > 0 else 4
    skips = [4]
    def make(cls, depth, width, **extra):
        return cls(D=depth, W=width, input_ch=input_ch, output_ch=output_ch, skips=skips, input_ch_views=input_ch_views,
                   use_viewdirs=args.use_viewdirs, **extra).to(device)

    model_fine = None
    if getattr(args, "alpha_model_path", None) is None:
        model = make(NeRF, args.netdepth, args.netwidth)
        grad_vars = list(model.parameters())
        if args.N_importance > 0:
            model_fine = make(NeRF, args.netdepth_fine, args.netwidth_fine)
            grad_vars += list(model_fine.parameters())
    else:  # frozen density network + colour networks (render.py:182-208)
        alpha_model = make(NeRF, args.netdepth_fine, args.netwidth_fine)
        ckpt = torch.load(args.alpha_model_path, map_location=device)
        alpha_model.load_state_dict(ckpt['network_fine_state_dict'])
        if not getattr(args, "no_coarse", False):
            model = make(NeRF_RGB, args.netdepth, args.netwidth, alpha_model=alpha_model)
            grad_vars = [p_ for n_, p_ in model.named_parameters() if not n_.startswith("alpha_model.")]
        else:
            model, grad_vars = None, []
        if args.N_importance > 0:
            model_fine = make(NeRF_RGB, args.netdepth_fine, args.netwidth_fine, alpha_model=alpha_model)
            grad_vars += [p_ for n_, p_ in model_fine.named_parameters() if not n_.startswith("alpha_model.")]
    model_confidence = None  # the reference names an undefined DepthConfNet here (render.py:211-213)

    optimizer = torch.optim.Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))
    start = 0
    basedir, expname = args.basedir, args.expname
    if getattr(args, "ft_path", None) is not None and args.ft_path != 'None':
        ckpts = [args.ft_path]
    else:
        d = os.path.join(basedir, expname)
        ckpts = [os.path.join(d, f) for f in sorted(os.listdir(d)) if 'tar' in f] if os.path.isdir(d) else []
    if len(ckpts) > 0 and not args.no_reload:
        ckpt = torch.load(ckpts[-1], map_location=device)
        start = ckpt['global_step']
        optimizer.load_state_dict(ckpt['optimizer_state_dict'])
        model.load_state_dict(ckpt['network_fn_state_dict'])
        if model_fine is not None:
            model_fine.load_state_dict(ckpt['network_fine_state_dict'])

    # same keys as render.py:251-273 (render() consumes use_viewdirs / ndc, render_rays the rest)
    render_kwargs_train = dict(network_query_fn=network_query_fn, perturb=args.perturb, N_importance=args.N_importance,
                               network_fine=model_fine, N_samples=args.N_samples, network_fn=model,
                               use_viewdirs=args.use_viewdirs, white_bkgd=args.white_bkgd,
                               raw_noise_std=args.raw_noise_std)
    forward_facing = args.dataset_type == 'llff' and not args.no_ndc  # NDC only suits LLFF-style data
    render_kwargs_train['ndc'] = forward_facing
    if not forward_facing:
        render_kwargs_train['lindisp'] = args.lindisp
    render_kwargs_test = {**render_kwargs_train, 'perturb': False, 'raw_noise_std': 0.}
    return render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer, model_confidence
