"""Training path of `render_rays`: the fused forward with saved activations + the hand-written backward kernels,
exposed to torch as one autograd node.

The reference trains by running torch autograd over its eager ops (train.py / render.py:281-409).  Here
`render_rays` returns tensors attached to a single `torch.autograd.Function` whose backward calls
`snerf_render_rays_bwd` (csrc/snerf_train.cu): the caller's loss, optimizer and `loss.backward()` stay ordinary
torch code, the gradients of both networks' parameters come from our kernels.  Training arithmetic is fp32
(`set_mode` selects the inference arithmetic only).  The resampled depths are detached as in the reference
(render.py:381); rays carry no gradient.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .run_nerf_helpers import _TRAIN, _f32c

_DIFF_OUT = ("rgb_map", "disp_map", "acc_map", "depth_map", "weights", "rgb0", "disp0", "acc0", "depth0", "raw")
_NODIFF_OUT = ("z_vals_map", "z_std", "z_all", "weights_fine")


_DEBUG_KEEP = None   # set to a list to keep each step's workspace alive after its backward (debugging tools only)


class _Call:
    """Everything one render_rays training call needs to keep between forward and backward."""

    def __init__(self, rb, net_c, net_f, multires, multires_views, Nc, Nf, lindisp, white_bkgd, t_vals, u_vals,
                 t_rand, u_rand, noise0, noise1, alpha_c=None, alpha_f=None, desc_fine=None):
        self.rb, self.net_c, self.net_f = rb, net_c, net_f
        self.alpha_c, self.alpha_f = alpha_c, alpha_f      # frozen sigma networks of NeRF_RGB passes (forward only)
        self.Nc, self.Nf = Nc, Nf
        self.keep = [t_vals, u_vals, t_rand, u_rand, noise0, noise1]
        o = _lib.Opts()
        o.n_samples, o.n_importance = Nc, Nf
        o.lindisp, o.white_bkgd, o.mode = int(bool(lindisp)), int(bool(white_bkgd)), _lib.MODE_FP32
        o.multires, o.multires_views = multires, (multires_views if multires_views is not None else 0)
        o.save_for_backward = 1
        o.t_vals = t_vals.data_ptr()
        o.u_vals = u_vals.data_ptr() if u_vals is not None else None
        for name, t in (("t_rand", t_rand), ("u_rand", u_rand), ("noise0", noise0), ("noise1", noise1)):
            setattr(o, name, None if t is None else t.data_ptr())
        self.opts = o
        self.rays = _lib.Rays(rb.data_ptr(), rb.shape[0], rb.shape[1], rb.stride(0))
        self.desc = net_c.desc()
        self.desc_fine = desc_fine                           # the fine network's own architecture, or None (same)
        if desc_fine is not None:
            o.desc_fine = C.cast(C.pointer(desc_fine), C.c_void_p)
        self.ws = None
        self.names = None


class _RenderRaysTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, call: _Call, *params):
        ctx.set_materialize_grads(False)     # outputs the loss never touches arrive as None in backward, not as zero tensors
        rb, dev = call.rb, call.rb.device
        N, Nc, Nf = rb.shape[0], call.Nc, call.Nf
        S = Nc + Nf
        lib = _lib.load()
        prec = _TRAIN["precision"]            # fixed for the whole step (forward and backward must agree)
        call.prec = prec
        call.tf32 = prec == "tf32"
        # (forward image, opts.mode, backward image) of the precision level
        call.modes = {"fp32": (_lib.MODE_FP32, _lib.MODE_FP32, _lib.PACK_FP32_BWD),
                      "tf32": (_lib.PACK_TF32_FWD, _lib.MODE_TF32, _lib.PACK_TF32_BWD),
                      "bf16": (_lib.MODE_BF16, _lib.MODE_BF16, _lib.PACK_BF16_BWD)}[prec]
        nbytes = lib.snerf_train_workspace_bytes_pair(C.byref(call.desc), C.byref(call.desc_fine) if call.desc_fine is not None
                                                      else None, Nc, Nf, N, call.modes[1])
        if nbytes == 0:
            raise RuntimeError("snerf_train_workspace_bytes: " + _lib.last_error())
        try:
            call.ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        except torch.OutOfMemoryError:
            # (asking the driver for the free size up front costs ~17 ms per call on a busy context: only on failure)
            free, _ = torch.cuda.mem_get_info(dev)
            raise RuntimeError(
                f"snerf_b200.render_rays (training): {N} rays need a {nbytes / 2**30:.1f} GiB activation store "
                f"({nbytes // max(N, 1) / 2**20:.1f} MiB per ray) but only {free / 2**30:.1f} GiB are free.  Use fewer rays per "
                "step, or wrap pure rendering in torch.no_grad() (render() / render_path() for evaluation).") from None

        def new(*shape):
            return torch.empty(shape, dtype=torch.float32, device=dev)

        bufs = {"rgb_map": new(N, 3), "disp_map": new(N), "acc_map": new(N), "depth_map": new(N),
                "weights": new(N, Nc), "z_vals_map": new(N, Nc), "raw": new(N, S, 4)}
        if Nf > 0:
            bufs.update(rgb0=new(N, 3), disp0=new(N), acc0=new(N), depth0=new(N), z_std=new(N), z_all=new(N, S),
                        weights_fine=new(N, S))
        out = _lib.Out()
        for k, t in bufs.items():
            setattr(out, k, t.data_ptr())
        fwd_mode = call.modes[0]
        call.opts.mode = call.modes[1]
        # the optimizer has touched every parameter since the last step: refresh the forward AND the backward images of
        # both networks in one launch (the backward finds them current)
        from .run_nerf_helpers import pack_many
        pack_many([(call.net_c, fwd_mode), (call.net_f, fwd_mode), (call.net_c, call.modes[2]), (call.net_f, call.modes[2])])
        img_c = call.net_c.packed(fwd_mode)
        img_f = call.net_f.packed(fwd_mode) if call.net_f is not None else None
        img_a = [a.packed(_lib.MODE_FP32) if a is not None else None for a in (call.alpha_c, call.alpha_f)]
        call.opts.packed_alpha_coarse = img_a[0].data_ptr() if img_a[0] is not None else None
        call.opts.packed_alpha_fine = img_a[1].data_ptr() if img_a[1] is not None else None
        with torch.cuda.device(dev):
            _lib.check(lib.snerf_render_rays_fwd(C.byref(call.rays), C.byref(call.desc), _lib.ptr(img_c),
                                                 _lib.ptr(img_f), C.byref(call.opts), C.byref(out),
                                                 _lib.ptr(call.ws), nbytes, _lib.stream_ptr(dev)),
                       "snerf_render_rays_fwd(save_for_backward)")
        call.names = [k for k in _DIFF_OUT + _NODIFF_OUT if k in bufs]
        ctx.call = call
        ctx.mark_non_differentiable(*[bufs[k] for k in _NODIFF_OUT if k in bufs])
        return tuple(bufs[k] for k in call.names)

    @staticmethod
    def backward(ctx, *gouts):
        call: _Call = ctx.call
        if call.ws is None:
            raise RuntimeError("snerf_b200: backward through render_rays a second time (workspace already released)")
        dev = call.rb.device
        lib = _lib.load()
        g = _lib.OutGrad()
        keep = []
        for name, t in zip(call.names, gouts):
            if t is None or name not in _lib.GRAD_FIELDS:
                continue
            t = _f32c(t)
            keep.append(t)
            setattr(g, name, t.data_ptr())
        st_c, grads_c, _ = call.net_c.grad_buffers()
        st_f, grads_f = None, []
        if call.net_f is not None:
            st_f, grads_f, _ = call.net_f.grad_buffers()
        pack_mode = call.modes[2]
        bwd_c = call.net_c.packed(pack_mode)
        bwd_f = call.net_f.packed(pack_mode) if call.net_f is not None else None
        call.opts.mode = call.modes[1]
        with torch.cuda.device(dev):
            _lib.check(lib.snerf_render_rays_bwd(C.byref(call.rays), C.byref(call.desc), _lib.ptr(bwd_c),
                                                 _lib.ptr(bwd_f), C.byref(call.opts), C.byref(g), C.byref(st_c),
                                                 C.byref(st_f) if st_f is not None else None, _lib.ptr(call.ws),
                                                 call.ws.numel(), _lib.stream_ptr(dev)), "snerf_render_rays_bwd")
        if _DEBUG_KEEP is not None:   # tools/tc_train_check.py inspects the stores after the backward
            _DEBUG_KEEP.append(call.ws)
        call.ws = None  # the activation store is the big allocation: release it as soon as it has been consumed
        # with a bound FlatGradients buffer the kernels have already accumulated into every p.grad: nothing for autograd to add
        if getattr(call.net_c, "_flat_grad", None) is not None:
            grads_c = [None] * len(grads_c)
        if call.net_f is not None and getattr(call.net_f, "_flat_grad", None) is not None:
            grads_f = [None] * len(grads_f)
        return (None, *grads_c, *grads_f)


def render_rays_train(call: _Call) -> dict:
    """Runs the autograd node; returns {output name: tensor} (differentiable w.r.t. the networks' parameters)."""
    params = [p for _, _, p in call.net_c._slots()]
    if call.net_f is not None:
        params += [p for _, _, p in call.net_f._slots()]
    outs = _RenderRaysTrain.apply(call, *params)
    return dict(zip(call.names, outs))


def warn_inference_only():
    """Trainable parameters + grad mode on, but a tensor-core mode is selected: those kernels are inference-only."""
    import warnings
    warnings.warn("snerf_b200: set_mode('bf16' / 'fp16' / 'fp16x3') selects inference-only kernels; the outputs of this "
                  "render_rays call carry no autograd graph.  To train, call snerf_b200.set_train_precision('bf16') (tensor "
                  "cores) or snerf_b200.set_mode('fp32') (fp32 / tf32 training); or wrap rendering in torch.no_grad().",
                  UserWarning, stacklevel=3)


def wants_grad(*nets) -> bool:
    return torch.is_grad_enabled() and any(p.requires_grad for n in nets if n is not None for p in n.parameters())
