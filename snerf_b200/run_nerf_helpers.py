"""Host-side mirror of the reference's `model/run_nerf_helpers.py` for the render_rays hot path.

Same names, signatures, state_dict keys and error behaviour as the reference
(/root/reference/s-nerf/model/run_nerf_helpers.py; line numbers cited per item), but every
tensor operation runs in libsnerf_b200.so (hand-written sm_100a CUDA) through the C ABI of
include/snerf_b200.h.  PyTorch is used for device memory, streams, RNG and parameters only.
There is no CPU path: calling a compute function with CPU tensors raises.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib

__all__ = ["Embedder", "get_embedder", "NeRF", "NeRF_RGB", "get_rays", "ndc_rays", "sample_pdf", "raw2outputs",
           "batchify", "run_network", "img2mse", "mse2psnr", "to8b", "set_mode", "get_mode"]

# Misc (run_nerf_helpers.py:15-18)
img2mse = lambda x, y: torch.mean((x - y) ** 2)
mse2psnr = lambda x: -10. * torch.log(x) / math.log(10.)
to8b = lambda x: (255 * np.clip(x, 0, 1)).astype(np.uint8)

_MODE = {"mode": _lib.MODE_FP32}
_MODE_NAMES = {"fp32": _lib.MODE_FP32, "bf16": _lib.MODE_BF16, "fp16": _lib.MODE_FP16, "fp16x3": _lib.MODE_FP16X3}
_TRAIN = {"precision": "fp32"}


def set_train_precision(name: str) -> None:
    """Arithmetic of the training step (`render_rays` with autograd on), forward AND backward:

    'fp32'  every GEMM on CUDA-core FFMA, fp32 stores: matches the reference's fp32 autograd to ~1e-6 (the parity level);
    'tf32'  layer-batched tcgen05 GEMMs with tf32 operands over fp32 activation stores, fp32 accumulation (forward too);
    'bf16'  the throughput level: the fused tensor-core renderer as forward (bf16 operands, fp32 accumulation, every
            layer's output kept as bf16), one fused dX-chain kernel and one grouped weight-gradient GEMM as backward
            (bf16 gradients / activations as operands, fp32 accumulation, fp32 parameter gradients).  Needs
            NeRF(8x256, skips=[4], viewdirs) and the sample counts of the tensor-core renderer.
    (An fp16 forward cannot feed this backward: tcgen05 kind::f16 rejects a bf16 x fp16 operand pair -- measured on B200:
    illegal instruction -- and gradients need bf16's exponent range.)"""
    if name not in ("fp32", "tf32", "bf16"):
        raise ValueError("train precision must be 'fp32', 'tf32' or 'bf16'")
    _TRAIN["precision"] = name


def get_train_precision() -> str:
    return _TRAIN["precision"]




def set_mode(mode: str):
    """MLP arithmetic of the fused renderer: 'fp32' (reference-accurate, CUDA cores), 'bf16' (tcgen05 tensor
    cores, bf16 operands, fp32 accumulate), 'fp16' (same kernel, fp16 operands: 8x tighter rounding, needs
    activations / weights inside fp16 range) or 'fp16x3' (fp32-class on the tensor cores: fp16 hi/lo operand
    split, three MMA passes -- the 1e-4 parity bar of 'fp32' at about ten times its rate)."""
    if mode not in _MODE_NAMES:
        raise ValueError(f"mode must be one of {sorted(_MODE_NAMES)}")
    _MODE["mode"] = _MODE_NAMES[mode]


def get_mode() -> str:
    return {v: k for k, v in _MODE_NAMES.items()}[_MODE["mode"]]


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"snerf_b200.{what}: tensors must live on a CUDA (sm_100) device; there is no CPU fallback")


_STAGE_WARNED = set()


def _warn_stage_no_grad(name: str, *tensors_or_modules):
    """The stage entry points (NeRF.forward, run_network, raw2outputs, sample_pdf) are forward kernels: their results carry
    no autograd graph, unlike the reference's eager versions.  Say so once when the caller could be expecting gradients
    (grad mode on and a trainable parameter / a tensor that requires grad among the inputs); training goes through
    render_rays (snerf_b200/autograd.py)."""
    if name in _STAGE_WARNED or not torch.is_grad_enabled():
        return
    for t in tensors_or_modules:
        if t is None:
            continue
        needs = any(p.requires_grad for p in t.parameters()) if isinstance(t, nn.Module) else bool(getattr(t, "requires_grad", False))
        if needs:
            import warnings
            _STAGE_WARNED.add(name)
            warnings.warn(f"snerf_b200.{name}: the result carries no autograd graph (stage entry points are forward-only "
                          "kernels); train through snerf_b200.render_rays, or wrap the call in torch.no_grad()",
                          UserWarning, stacklevel=3)
            return


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


# --------------------------------------------------------------------------------------
# Positional encoding (run_nerf_helpers.py:22-70)
# --------------------------------------------------------------------------------------
class Embedder:
    """cat[x, sin(f0 x), cos(f0 x), sin(f1 x), ...] with the reference's kwargs."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        self.create_embedding_fn()

    def create_embedding_fn(self):
        kw = self.kwargs
        d = kw["input_dims"]
        n = kw["num_freqs"]
        if kw["log_sampling"]:
            bands = 2. ** torch.linspace(0., kw["max_freq_log2"], steps=n)
        else:
            bands = torch.linspace(2. ** 0., 2. ** kw["max_freq_log2"], steps=n)
        self.freq_bands = bands
        self.out_dim = (d if kw["include_input"] else 0) + d * n * len(kw["periodic_fns"])
        # the CUDA encoder implements exactly the reference's default configuration
        self.standard = (kw["include_input"] and d == 3 and kw["log_sampling"] and kw["max_freq_log2"] == n - 1
                         and list(kw["periodic_fns"]) == [torch.sin, torch.cos])

    def embed(self, inputs: torch.Tensor) -> torch.Tensor:
        if not self.standard:
            raise RuntimeError("snerf_b200.Embedder: only the reference's default embedding "
                               "(include_input, log_sampling, [sin, cos]) is implemented")
        _require_cuda(inputs, "Embedder.embed")
        x = _f32c(inputs).reshape(-1, 3)
        out = torch.empty((x.shape[0], self.out_dim), dtype=torch.float32, device=x.device)
        lib = _lib.load()
        with torch.cuda.device(x.device):
            _lib.check(lib.snerf_posenc(_lib.ptr(x), x.shape[0], self.kwargs["num_freqs"], _lib.ptr(out),
                                        _lib.stream_ptr(x.device)), "snerf_posenc")
        return out.reshape(*inputs.shape[:-1], self.out_dim)


class _EmbedFn:
    """Callable returned by get_embedder; carries `multires` so run_network can fuse it."""

    def __init__(self, embedder: Embedder, multires: int):
        self.embedder = embedder
        self.multires = multires

    def __call__(self, x):
        return self.embedder.embed(x)


class _IdentityEmbed(nn.Identity):
    multires = -1


def get_embedder(multires, i=0):
    if i == -1:
        return _IdentityEmbed(), 3
    eo = Embedder(include_input=True, input_dims=3, max_freq_log2=multires - 1, num_freqs=multires,
                  log_sampling=True, periodic_fns=[torch.sin, torch.cos])
    return _EmbedFn(eo, multires), eo.out_dim


# --------------------------------------------------------------------------------------
# Model (run_nerf_helpers.py:74-126)
# --------------------------------------------------------------------------------------
class NeRF(nn.Module):
    """Same constructor, parameter names and forward() contract as the reference `NeRF`.

    Parameters stay ordinary fp32 `nn.Linear` tensors (checkpoints are interchangeable with the
    reference's `network_fn_state_dict`); the kernels read a packed image that is rebuilt
    whenever a parameter changes (`packed(mode)`).
    """

    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, output_ch=4, skips=[4], use_viewdirs=False):
        super().__init__()
        self.D = D
        self.W = W
        self.input_ch = input_ch
        self.input_ch_views = input_ch_views
        self.output_ch = output_ch
        self.skips = skips
        self.use_viewdirs = use_viewdirs

        self.pts_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)] +
            [nn.Linear(W + input_ch, W) if i in self.skips else nn.Linear(W, W) for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W // 2)])
        if use_viewdirs:
            self.feature_linear = nn.Linear(W, W)
            self.alpha_linear = nn.Linear(W, 1)
            self.rgb_linear = nn.Linear(W // 2, 3)
        else:
            self.output_linear = nn.Linear(W, output_ch)
        self._packed = {}

    # ---- C-ABI views of the module
    def desc(self) -> _lib.NetDesc:
        live = [s for s in self.skips if 0 <= s < self.D - 1]
        if len(live) > 1 or any(s == self.D - 1 for s in self.skips):
            raise RuntimeError(f"snerf_b200.NeRF: skips={self.skips} with D={self.D} is not supported "
                               "(at most one skip connection, not after the last trunk layer)")
        return _lib.NetDesc(self.D, self.W, self.input_ch, self.input_ch_views, live[0] if live else -1,
                            1 if self.use_viewdirs else 0, self.output_ch)

    def _param_list(self):
        ps = []
        for l in self.pts_linears:
            ps += [l.weight, l.bias]
        if self.use_viewdirs:
            ps += [self.views_linears[0].weight, self.views_linears[0].bias, self.feature_linear.weight,
                   self.feature_linear.bias, self.rgb_linear.weight, self.rgb_linear.bias]
            if hasattr(self, "alpha_linear"):
                ps += [self.alpha_linear.weight, self.alpha_linear.bias]
        else:
            ps += [self.output_linear.weight, self.output_linear.bias]
        return ps

    def _slots(self):
        """(struct field, trunk index or None, parameter) for every parameter, the member order of SnerfNetF32."""
        out = []
        for i, l in enumerate(self.pts_linears):
            out += [("pts_w", i, l.weight), ("pts_b", i, l.bias)]
        if self.use_viewdirs:
            out += [("views_w", None, self.views_linears[0].weight), ("views_b", None, self.views_linears[0].bias),
                    ("feature_w", None, self.feature_linear.weight), ("feature_b", None, self.feature_linear.bias)]
            if hasattr(self, "alpha_linear"):
                out += [("alpha_w", None, self.alpha_linear.weight), ("alpha_b", None, self.alpha_linear.bias)]
            out += [("rgb_w", None, self.rgb_linear.weight), ("rgb_b", None, self.rgb_linear.bias)]
        else:
            out += [("output_w", None, self.output_linear.weight), ("output_b", None, self.output_linear.bias)]
        return out

    def invalidate_packed(self):
        """Drop the cached weight images.  `packed()` notices in-place updates through each parameter's version counter
        (optimizer steps, `p.copy_()` under no_grad); writes through `p.data` (`p.data.copy_`, `p.data.clamp_`, EMA code)
        do NOT move that counter -- call this after them."""
        self._packed.clear()

    def _bind_flat_grad(self, flat):
        """snerf_b200.parallel.FlatGradients: the backward kernels accumulate straight into `flat` (this network's
        slice of the shared gradient buffer; every `p.grad` is a view of it) and autograd receives no per-parameter
        tensors."""
        self._flat_grad = flat

    def grad_buffers(self):
        """Gradient buffers (one flat allocation) + the SnerfNetGradF32 pointing into it; grads[i] matches `_slots()[i]`.
        Zeroed and fresh per call, unless a FlatGradients buffer is bound: then its (accumulating) views are returned."""
        slots = self._slots()
        dev = slots[0][2].device
        bound = getattr(self, "_flat_grad", None)
        flat = bound if bound is not None else torch.zeros(sum(p.numel() for _, _, p in slots), dtype=torch.float32, device=dev)
        st = _lib.NetGradF32()
        grads, off = [], 0
        for field, idx, p in slots:
            g = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
            grads.append(g)
            if idx is None:
                setattr(st, field, g.data_ptr())
            else:
                getattr(st, field)[idx] = g.data_ptr()
        return st, grads, flat

    def _pack_request(self, mode: int):
        """None when the cached image of `mode` is current, else (desc, NetF32, image tensor, nbytes, stamp, keep-alive list)
        for `snerf_pack_weights` / `snerf_pack_weights_batch`."""
        ps = self._param_list()
        dev = ps[0].device
        if dev.type != "cuda":
            raise RuntimeError("snerf_b200.NeRF: parameters must be on a CUDA device (call .cuda()); no CPU fallback")
        stamp = tuple((p.data_ptr(), p._version) for p in ps)
        hit = self._packed.get((mode, dev.index))
        if hit is not None and hit[0] == stamp:
            return None
        lib = _lib.load()
        d = self.desc()
        nbytes = lib.snerf_packed_bytes(C.byref(d), mode)
        if nbytes == 0:
            raise RuntimeError("snerf_packed_bytes: " + _lib.last_error())
        img = hit[1] if hit is not None and hit[1].numel() == nbytes else torch.empty(nbytes, dtype=torch.uint8, device=dev)
        src = _lib.NetF32()
        keep = []

        def p32(t):
            t = _f32c(t)
            keep.append(t)
            return t.data_ptr()

        for i, l in enumerate(self.pts_linears):
            src.pts_w[i] = p32(l.weight)
            src.pts_b[i] = p32(l.bias)
        if self.use_viewdirs:
            src.views_w, src.views_b = p32(self.views_linears[0].weight), p32(self.views_linears[0].bias)
            src.feature_w, src.feature_b = p32(self.feature_linear.weight), p32(self.feature_linear.bias)
            if hasattr(self, "alpha_linear"):  # NeRF_RGB has none: sigma comes from its frozen alpha_model
                src.alpha_w, src.alpha_b = p32(self.alpha_linear.weight), p32(self.alpha_linear.bias)
            src.rgb_w, src.rgb_b = p32(self.rgb_linear.weight), p32(self.rgb_linear.bias)
        else:
            src.output_w, src.output_b = p32(self.output_linear.weight), p32(self.output_linear.bias)
        return d, src, img, nbytes, stamp, keep

    def packed(self, mode: int) -> torch.Tensor:
        """Device image of the weights for `mode`, refreshed when any parameter was modified."""
        req = self._pack_request(mode)
        dev = self._param_list()[0].device
        if req is None:
            return self._packed[(mode, dev.index)][1]
        d, src, img, nbytes, stamp, keep = req
        with torch.cuda.device(dev):
            _lib.check(_lib.load().snerf_pack_weights(C.byref(d), C.byref(src), _lib.ptr(img), nbytes, mode,
                                                      _lib.stream_ptr(dev)), "snerf_pack_weights")
        self._packed[(mode, dev.index)] = (stamp, img)
        return img

    # ---- tensor-core stage path (set_mode('bf16')): NeRF.forward / network_query_fn on their own as a chain of
    # snerf_linear_tc launches (csrc/snerf_mip.cu: persistent tcgen05 GEMM, bf16 operands, fp32 accumulate)
    def _packed_tc(self):
        """bf16 [n_pad, K_pad] weight matrices (K segments padded to multiples of 64) + fp32 biases / heads, cached on the
        parameters' version counters like `packed()`."""
        ps = self._param_list()
        stamp = tuple((p.data_ptr(), p._version) for p in ps)
        hit = self._packed.get(("tc_stage", ps[0].device.index))
        if hit is not None and hit[0] == stamp:
            return hit[1]
        if not (self.use_viewdirs and hasattr(self, "alpha_linear")):
            raise RuntimeError("snerf_b200.NeRF: the tensor-core stage path needs use_viewdirs=True and an alpha head "
                               "(use set_mode('fp32') for other networks)")
        if self.W % 64 or self.input_ch > 64 or self.input_ch_views > 64:
            raise RuntimeError("snerf_b200.NeRF: the tensor-core stage path needs W % 64 == 0 and encodings of at most 64 channels")
        pad = torch.nn.functional.pad

        def mat(w, segs):
            cols, off = [], 0
            for real, padded in segs:
                cols.append(pad(w[:, off:off + real], (0, padded - real)))
                off += real
            out = torch.cat(cols, 1)
            return pad(out, (0, 0, 0, (-out.shape[0]) % 128)).to(torch.bfloat16).contiguous()

        with torch.no_grad():
            W, ic, icv = self.W, self.input_ch, self.input_ch_views
            def f32(t):        # fp32, contiguous, 16-byte aligned (parameters that are views of a flat buffer need not be)
                t = t.detach().float().contiguous()
                return t.clone() if t.data_ptr() % 16 else t
            P = {"trunk": []}
            for i, l in enumerate(self.pts_linears):
                has_enc = l.weight.shape[1] in (ic, ic + W)
                segs = ([(ic, 64)] if has_enc else []) + ([(W, W)] if l.weight.shape[1] >= W and i > 0 else [])
                P["trunk"].append((mat(l.weight.float(), segs), f32(l.bias), has_enc, i > 0))
            P["feature"] = (mat(self.feature_linear.weight.float(), [(W, W)]), f32(self.feature_linear.bias))
            vl = self.views_linears[0]
            P["views"] = (mat(vl.weight.float(), [(W, W), (icv, 64)]), f32(vl.bias))
            P["alpha_w"] = f32(self.alpha_linear.weight)
            P["rgb_w"] = f32(self.rgb_linear.weight)
            P["head_bias"] = torch.cat([f32(self.rgb_linear.bias), f32(self.alpha_linear.bias)])      # (r, g, b, sigma)
        self._packed[("tc_stage", ps[0].device.index)] = (stamp, P)
        return P

    def _forward_tc(self, enc_pts, enc_dirs, m_rows, m_pad):
        """enc_pts / enc_dirs: bf16 [m_pad, 64] zero-padded operand rows -> fp32 [m_rows, 4]."""
        lib = _lib.load()
        dev = enc_pts.device
        st = _lib.stream_ptr(dev)
        P = self._packed_tc()
        W = self.W
        bufs = [torch.empty((m_pad, W), dtype=torch.bfloat16, device=dev) for _ in range(2)]
        out4 = P["head_bias"].expand(m_pad, 4).contiguous()          # heads accumulate on top of their biases

        def linear(a0, k0, w, n, bias, out, relu=True, a1=None, k1=0, head_w=None, head_col=0):
            L = _lib.Linear()
            L.a0, L.lda0, L.k0 = a0.data_ptr(), a0.stride(0), k0
            L.a1, L.lda1, L.k1 = (a1.data_ptr() if a1 is not None else None), (a1.stride(0) if a1 is not None else 0), k1
            L.w, L.n, L.n_pad, L.bias = w.data_ptr(), n, w.shape[0], bias.data_ptr()
            L.ray_bias, L.rows_per_ray, L.relu = None, 1, 1 if relu else 0
            L.out, L.ldo = (out.data_ptr() if out is not None else None), (out.stride(0) if out is not None else 0)
            if head_w is not None:
                L.head_w, L.n_heads, L.head_ld = head_w.data_ptr(), head_w.shape[0], 4
                L.head_out = out4.data_ptr() + 4 * head_col
            L.m_rows, L.m_pad = m_rows, m_pad
            _lib.check(lib.snerf_linear_tc(C.byref(L), st), "snerf_linear_tc")

        with torch.cuda.device(dev):
            h = None
            n_trunk = len(P["trunk"])
            for i, (w, b, has_enc, has_hidden) in enumerate(P["trunk"]):
                out = bufs[i & 1]
                head = P["alpha_w"] if i == n_trunk - 1 else None     # sigma = alpha_linear(h) on the last trunk output
                if has_enc and has_hidden:      # skip layer: cat[input_pts, h] (run_nerf_helpers.py:109-110)
                    linear(enc_pts, 64, w, W, b, out, a1=h, k1=W, head_w=head, head_col=3)
                elif has_enc:
                    linear(enc_pts, 64, w, W, b, out, head_w=head, head_col=3)
                else:
                    linear(h, W, w, W, b, out, head_w=head, head_col=3)
                h = out
            feat = bufs[n_trunk & 1]
            linear(h, W, P["feature"][0], W, P["feature"][1], feat, relu=False)
            linear(feat, W, P["views"][0], W // 2, P["views"][1], None, a1=enc_dirs, k1=64, head_w=P["rgb_w"], head_col=0)
        return out4[:m_rows]

    def forward(self, x):
        """x[..., input_ch + input_ch_views] (already encoded) -> [..., 4] (rgb, sigma) (reference
        returns output_ch columns without viewdirs; the renderer only ever reads the first four)."""
        _require_cuda(x, "NeRF.forward")
        _warn_stage_no_grad("NeRF.forward", self, x)
        width = self.input_ch + (self.input_ch_views if self.use_viewdirs else 0)
        if x.shape[-1] < width:
            raise RuntimeError(f"NeRF.forward: expected last dim >= {width}, got {x.shape[-1]}")
        x2 = _f32c(x).reshape(-1, x.shape[-1])
        if _MODE["mode"] == _lib.MODE_BF16 and type(self) is NeRF:
            M = x2.shape[0]
            m_pad = (M + 127) // 128 * 128
            lib = _lib.load()
            ep = torch.empty((m_pad, 64), dtype=torch.bfloat16, device=x2.device)
            ed = torch.empty((m_pad, 64), dtype=torch.bfloat16, device=x2.device)
            with torch.cuda.device(x2.device):
                st = _lib.stream_ptr(x2.device)
                _lib.check(lib.snerf_rows_to_bf16(_lib.ptr(x2), M, x2.shape[1], 0, self.input_ch, 1, _lib.ptr(ep), 64, m_pad, st), "snerf_rows_to_bf16")
                _lib.check(lib.snerf_rows_to_bf16(_lib.ptr(x2), M, x2.shape[1], self.input_ch, self.input_ch_views, 1, _lib.ptr(ed), 64,
                                                  m_pad, st), "snerf_rows_to_bf16")
            return self._forward_tc(ep, ed, M, m_pad).reshape(*x.shape[:-1], 4)
        out = torch.empty((x2.shape[0], 4), dtype=torch.float32, device=x2.device)
        lib = _lib.load()
        d = self.desc()
        img = self.packed(_lib.MODE_FP32)
        with torch.cuda.device(x2.device):
            _lib.check(lib.snerf_nerf_forward(C.byref(d), _lib.ptr(img), _lib.MODE_FP32, _lib.ptr(x2), x2.shape[0],
                                              x2.shape[1], _lib.ptr(out), _lib.stream_ptr(x2.device)),
                       "snerf_nerf_forward")
        return out.reshape(*x.shape[:-1], 4)

    def load_weights_from_keras(self, weights):
        """Import the flat weight list of the original (Keras) NeRF release (run_nerf_helpers.py:128-155): kernels are
        stored [in, out] and come in (kernel, bias) pairs ordered trunk layers, feature, views, rgb, alpha."""
        assert self.use_viewdirs, "Not implemented if use_viewdirs=False"
        layers = list(self.pts_linears) + [self.feature_linear, self.views_linears[0], self.rgb_linear, self.alpha_linear]
        for k, layer in enumerate(layers):
            kernel, bias = weights[2 * k], weights[2 * k + 1]
            layer.weight.data = torch.from_numpy(np.transpose(kernel)).to(layer.weight.device)
            layer.bias.data = torch.from_numpy(np.transpose(bias)).to(layer.bias.device)
        self._packed.clear()                 # fresh .data tensors: packed images must be rebuilt


class NeRF_RGB(NeRF):
    """Colour network with a frozen density network (run_nerf_helpers.py:157-212): same trunk / feature / views / rgb
    parameters as `NeRF` but NO alpha_linear; sigma = alpha_model(x)[..., 3] under no_grad.  fp32 mode only."""

    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, output_ch=4, skips=[4], use_viewdirs=False,
                 alpha_model=None):
        super().__init__(D=D, W=W, input_ch=input_ch, input_ch_views=input_ch_views, output_ch=output_ch, skips=skips,
                         use_viewdirs=use_viewdirs)
        if use_viewdirs:
            del self.alpha_linear
        self.alpha_model = alpha_model

    def forward(self, x):
        out = super().forward(x)  # columns 0..2 from this network (column 3 is not produced without an alpha head)
        if self.use_viewdirs:
            with torch.no_grad():
                out[..., 3] = self.alpha_model(x)[..., 3]
        return out


def pack_many(items) -> None:
    """Refresh the weight images of several (network, mode) pairs with ONE library call (`snerf_pack_weights_batch`: the
    tensor-core images of all stale items are written by one kernel launch).  The training step uses it for the forward
    and backward images of both networks, which the optimizer invalidates every iteration."""
    reqs = []
    for net, mode in items:
        if net is None or any(net is n and mode == m for n, m, _ in reqs):
            continue
        r = net._pack_request(mode)
        if r is not None:
            reqs.append((net, mode, r))
    if not reqs:
        return
    n = len(reqs)
    dev = reqs[0][0]._param_list()[0].device
    descs = (C.POINTER(_lib.NetDesc) * n)(*[C.pointer(r[0]) for _, _, r in reqs])
    srcs = (C.POINTER(_lib.NetF32) * n)(*[C.pointer(r[1]) for _, _, r in reqs])
    imgs = (C.c_void_p * n)(*[r[2].data_ptr() for _, _, r in reqs])
    sizes = (C.c_size_t * n)(*[r[3] for _, _, r in reqs])
    modes = (C.c_int32 * n)(*[m for _, m, _ in reqs])
    with torch.cuda.device(dev):
        _lib.check(_lib.load().snerf_pack_weights_batch(n, descs, srcs, imgs, sizes, modes, _lib.stream_ptr(dev)),
                   "snerf_pack_weights_batch")
    for net, mode, (d, src, img, nbytes, stamp, keep) in reqs:
        net._packed[(mode, dev.index)] = (stamp, img)


# --------------------------------------------------------------------------------------
# Ray helpers (run_nerf_helpers.py:247-258, 314-334)
# --------------------------------------------------------------------------------------
def get_rays(H, W, focal, c2w, ori_points=None, device=None):
    """Pinhole rays of an H x W image (pixel centres at +0.5); returns (rays_o, rays_d) [H, W, 3]."""
    if not ori_points:
        ori_points = [W * 0.5, H * 0.5]
    c2w_t = torch.as_tensor(c2w, dtype=torch.float32)
    if device is None:
        device = c2w_t.device if c2w_t.is_cuda else torch.device("cuda", torch.cuda.current_device())
    m = np.ascontiguousarray(c2w_t.detach().cpu().numpy()[:3, :4], dtype=np.float32)
    ro = torch.empty((H, W, 3), dtype=torch.float32, device=device)
    rd = torch.empty((H, W, 3), dtype=torch.float32, device=device)
    lib = _lib.load()
    with torch.cuda.device(device):
        _lib.check(lib.snerf_get_rays(H, W, float(np.float32(focal)), m.ctypes.data_as(C.POINTER(C.c_float)),
                                      float(ori_points[0]), float(ori_points[1]), _lib.ptr(ro), _lib.ptr(rd),
                                      _lib.stream_ptr(device)), "snerf_get_rays")
    return ro, rd


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """NDC reprojection for forward-facing scenes (host-side plumbing above the hot path)."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    sx, sy = -1. / (W / (2. * focal)), -1. / (H / (2. * focal))
    oz = rays_o[..., 2]
    o = torch.stack([sx * rays_o[..., 0] / oz, sy * rays_o[..., 1] / oz, 1. + 2. * near / oz], -1)
    d = torch.stack([sx * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / oz),
                     sy * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / oz),
                     -2. * near / oz], -1)
    return o, d


# --------------------------------------------------------------------------------------
# Hierarchical sampling (run_nerf_helpers.py:336-379)
# --------------------------------------------------------------------------------------
def _draw_u(shape, n, det, pytest, device):
    """The u the reference would draw: linspace (det) or uniform; numpy seed 0 under pytest."""
    if pytest:
        np.random.seed(0)
        u = np.broadcast_to(np.linspace(0., 1., n), shape + [n]) if det else np.random.rand(*(shape + [n]))
        return torch.Tensor(np.ascontiguousarray(u)).to(device)
    if det:
        return torch.linspace(0., 1., steps=n).to(device).expand(shape + [n]).contiguous()
    return torch.rand(shape + [n], device=device)


def sample_pdf(bins, weights, N_samples, det=False, pytest=False, return_inds=False):
    _require_cuda(bins, "sample_pdf")
    _warn_stage_no_grad("sample_pdf", weights, bins)
    lead = list(bins.shape[:-1])
    B = bins.shape[-1]
    b2, w2 = _f32c(bins).reshape(-1, B), _f32c(weights).reshape(-1, B - 1)
    u = _draw_u(lead, N_samples, det, pytest, bins.device).reshape(-1, N_samples).contiguous()
    n = b2.shape[0]
    samples = torch.empty((n, N_samples), dtype=torch.float32, device=bins.device)
    inds = torch.empty((n, N_samples), dtype=torch.int64, device=bins.device) if return_inds else None
    lib = _lib.load()
    with torch.cuda.device(bins.device):
        _lib.check(lib.snerf_sample_pdf_fwd(_lib.ptr(b2), _lib.ptr(w2), None, _lib.ptr(u), 1, n, B, N_samples,
                                            _lib.ptr(samples), _lib.ptr(inds), None, _lib.stream_ptr(bins.device)),
                   "snerf_sample_pdf_fwd")
    samples = samples.reshape(lead + [N_samples])
    return (samples, inds.reshape(lead + [N_samples])) if return_inds else samples


# --------------------------------------------------------------------------------------
# Compositing (run_nerf_helpers.py:381-424)
# --------------------------------------------------------------------------------------
def _draw_noise(shape, std, pytest, device):
    if std <= 0.:
        return None
    if pytest:  # the reference's pytest hook draws UNIFORM noise (run_nerf_helpers.py:406-410)
        np.random.seed(0)
        return torch.Tensor(np.random.rand(*shape) * std).to(device)
    return torch.empty(shape, dtype=torch.float32, device=device).normal_(0.0, float(std))   # one launch (randn * std: two)


def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, white_bkgd=False, pytest=False):
    """-> (rgb_map, disp_map, acc_map, weights, depth_map)"""
    _require_cuda(raw, "raw2outputs")
    _warn_stage_no_grad("raw2outputs", raw, z_vals)
    n, S = z_vals.shape
    raw4 = _f32c(raw[..., :4]) if raw.shape[-1] != 4 else _f32c(raw)
    z, d = _f32c(z_vals), _f32c(rays_d)
    noise = _draw_noise([n, S], raw_noise_std, pytest, raw.device)
    dev = raw.device
    rgb = torch.empty((n, 3), dtype=torch.float32, device=dev)
    disp, acc, depth = (torch.empty((n,), dtype=torch.float32, device=dev) for _ in range(3))
    weights = torch.empty((n, S), dtype=torch.float32, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.snerf_composite_fwd(_lib.ptr(raw4), _lib.ptr(z), _lib.ptr(d), _lib.ptr(noise), n, S,
                                           1 if white_bkgd else 0, _lib.ptr(rgb), _lib.ptr(disp), _lib.ptr(acc),
                                           _lib.ptr(weights), _lib.ptr(depth), _lib.stream_ptr(dev)),
                   "snerf_composite_fwd")
    return rgb, disp, acc, weights, depth


# --------------------------------------------------------------------------------------
# Network query (run_nerf_helpers.py:450-474)
# --------------------------------------------------------------------------------------
def batchify(fn, chunk):
    if chunk is None:
        return fn

    def ret(inputs):
        return torch.cat([fn(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)
    return ret


def run_network(inputs, viewdirs, fn, embed_fn, embeddirs_fn, netchunk=1024 * 64):
    """inputs[N,S,3], viewdirs[N,3] -> raw[N,S,4].  With the standard embedders and a snerf_b200
    `NeRF` this is ONE kernel (encode + MLP, no [M,90] tensor, no netchunk loop); any other `fn`
    gets the reference's generic composition."""
    _require_cuda(inputs, "run_network")
    if isinstance(fn, NeRF):
        _warn_stage_no_grad("run_network", fn, inputs)
    fused = (isinstance(fn, NeRF) and hasattr(embed_fn, "multires")
             and (viewdirs is None or hasattr(embeddirs_fn, "multires")))
    if (fused and inputs.dim() == 3 and _MODE["mode"] == _lib.MODE_BF16 and type(fn) is NeRF and viewdirs is not None
            and fn.use_viewdirs):
        # tensor-core stage path: Embedder kernels -> bf16 operand rows (view directions broadcast over the ray's samples,
        # run_nerf_helpers.py:467-470) -> the snerf_linear_tc chain of NeRF._forward_tc
        N, S, _ = inputs.shape
        M = N * S
        m_pad = (M + 127) // 128 * 128
        ep32 = embed_fn(_f32c(inputs).reshape(-1, 3))
        ed32 = embeddirs_fn(_f32c(viewdirs))
        lib = _lib.load()
        ep = torch.empty((m_pad, 64), dtype=torch.bfloat16, device=inputs.device)
        ed = torch.empty((m_pad, 64), dtype=torch.bfloat16, device=inputs.device)
        with torch.cuda.device(inputs.device):
            st = _lib.stream_ptr(inputs.device)
            _lib.check(lib.snerf_rows_to_bf16(_lib.ptr(ep32), M, ep32.shape[1], 0, ep32.shape[1], 1, _lib.ptr(ep), 64, m_pad, st), "snerf_rows_to_bf16")
            _lib.check(lib.snerf_rows_to_bf16(_lib.ptr(ed32), N, ed32.shape[1], 0, ed32.shape[1], S, _lib.ptr(ed), 64, m_pad, st), "snerf_rows_to_bf16")
        return fn._forward_tc(ep, ed, M, m_pad).reshape(N, S, 4)
    if fused and inputs.dim() == 3:
        N, S, _ = inputs.shape
        pts = _f32c(inputs)
        vd = _f32c(viewdirs) if (viewdirs is not None and fn.use_viewdirs) else None
        raw = torch.empty((N, S, 4), dtype=torch.float32, device=inputs.device)
        lib = _lib.load()
        d = fn.desc()
        img = fn.packed(_lib.MODE_FP32)
        with torch.cuda.device(inputs.device):
            _lib.check(lib.snerf_query_network(C.byref(d), _lib.ptr(img), _lib.MODE_FP32, embed_fn.multires,
                                               embeddirs_fn.multires if vd is not None else 0, _lib.ptr(pts),
                                               _lib.ptr(vd), N, S, _lib.ptr(raw), _lib.stream_ptr(inputs.device)),
                       "snerf_query_network")
        return raw
    inputs_flat = torch.reshape(inputs, [-1, inputs.shape[-1]])
    embedded = embed_fn(inputs_flat)
    if viewdirs is not None:
        input_dirs = viewdirs[:, None].expand(inputs.shape)
        embedded = torch.cat([embedded, embeddirs_fn(torch.reshape(input_dirs, [-1, input_dirs.shape[-1]]))], -1)
    outputs_flat = batchify(fn, netchunk)(embedded)
    return torch.reshape(outputs_flat, list(inputs.shape[:-1]) + [outputs_flat.shape[-1]])
