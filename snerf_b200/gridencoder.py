"""Multi-resolution hash-grid encoder on libsnerf_b200.so -- the drop-in for the reference's torch extension
`s-nerfpp/zipnerf/gridencoder` (grid.py: `_grid_encode`, `grid_encode`, `GridEncoder`; BASELINE configs[3]).

Same module interface (constructor arguments, `embeddings` / `offsets` / `idx` / `grid_sizes` state-dict entries,
`forward(inputs, bound=1, cal_input_grad=False)`, `grad_total_variation(...)`), same autograd contract
(`grid_encode(inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs, gridtype, align_corners,
interpolation) -> [B, L*C]`).  The kernels (csrc/snerf_grid.cu) read and write the [B, L*C] layout directly, so the
permute / contiguous passes of the reference wrapper (grid.py:57,72) do not exist here.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib

_gridtype_to_id = {'hash': 0, 'tiled': 1}
_interp_to_id = {'linear': 0, 'smoothstep': 1}


def _desc(D, Cdim, L, S, H, gridtype, align_corners, interpolation, dtype):
    if dtype not in (torch.float32, torch.float16):
        raise RuntimeError("snerf_b200.gridencoder: embeddings must be float32 or float16")
    return _lib.GridDesc(D, Cdim, L, int(H), int(gridtype), int(bool(align_corners)), int(interpolation),
                         0 if dtype == torch.float32 else 1, float(S))


def _level_table(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size, align_corners):
    """Per-level grid size and first table row, with the values grid.py:123-140 computes: level i has
    ceil(base * scale^i) cells per axis (+1 vertices unless align_corners), stored densely while size^D fits 2^log2_T,
    else hashed into 2^log2_T rows; every level's row count is rounded up to a multiple of 8."""
    cap = 2 ** log2_hashmap_size
    sizes = np.array([int(np.ceil(base_resolution * per_level_scale ** i)) + (0 if align_corners else 1)
                      for i in range(num_levels)], dtype=np.int32)
    rows = [int(np.ceil(min(cap, int(r) ** input_dim) / 8) * 8) for r in sizes]
    starts = np.concatenate([[0], np.cumsum(rows)]).astype(np.int32)
    return sizes, starts


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"snerf_b200.gridencoder.{what}: tensors must live on a CUDA sm_100 device "
                           "(no CPU fallback: the encoder is a CUDA kernel)")


class _grid_encode(Function):
    """grid.py:24-90.  inputs [B, D] float in [0, 1]; embeddings [sO, C]; offsets [L+1] int32 -> [B, L*C]."""

    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0,
                align_corners=False, interpolation=0):
        _require_cuda(inputs, "grid_encode")
        inputs = inputs.contiguous().float()
        B, D = inputs.shape
        L = offsets.shape[0] - 1
        Cdim = embeddings.shape[1]
        S = np.log2(per_level_scale)
        H = base_resolution
        # autocast handling of the reference (grid.py:41-42): half embeddings when C is even
        if torch.is_autocast_enabled() and Cdim % 2 == 0:
            embeddings = embeddings.to(torch.half)
        embeddings = embeddings.contiguous()
        outputs = torch.empty(B, L * Cdim, device=inputs.device, dtype=embeddings.dtype)
        dy_dx = torch.empty(B, L * D * Cdim, device=inputs.device, dtype=embeddings.dtype) if calc_grad_inputs else None
        d = _desc(D, Cdim, L, S, H, gridtype, align_corners, interpolation, embeddings.dtype)
        with torch.cuda.device(inputs.device):
            _lib.check(_lib.load().snerf_grid_encode_fwd(C.byref(d), _lib.ptr(inputs), _lib.ptr(embeddings), _lib.ptr(offsets),
                                                         _lib.ptr(outputs), Cdim, L * Cdim, _lib.ptr(dy_dx), B,
                                                         _lib.stream_ptr(inputs.device)), "snerf_grid_encode_fwd")
        ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
        ctx.dims = [B, D, Cdim, L, S, H, gridtype, interpolation]
        ctx.align_corners = align_corners
        return outputs

    @staticmethod
    def backward(ctx, grad):
        inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        B, D, Cdim, L, S, H, gridtype, interpolation = ctx.dims
        grad = grad.contiguous().to(embeddings.dtype)          # [B, L*C]: consumed in place, no permute
        grad_embeddings = torch.zeros_like(embeddings)
        grad_inputs = torch.zeros_like(inputs, dtype=embeddings.dtype) if dy_dx is not None else None
        d = _desc(D, Cdim, L, S, H, gridtype, ctx.align_corners, interpolation, embeddings.dtype)
        with torch.cuda.device(inputs.device):
            _lib.check(_lib.load().snerf_grid_encode_bwd(C.byref(d), _lib.ptr(grad), Cdim, L * Cdim, _lib.ptr(inputs),
                                                         _lib.ptr(offsets), _lib.ptr(grad_embeddings), _lib.ptr(dy_dx),
                                                         _lib.ptr(grad_inputs), B, _lib.stream_ptr(inputs.device)),
                       "snerf_grid_encode_bwd")
        if dy_dx is not None:
            grad_inputs = grad_inputs.to(inputs.dtype)
        return grad_inputs, grad_embeddings, None, None, None, None, None, None, None


grid_encode = _grid_encode.apply


class _grid_encode_ms(Function):
    """Fused zip-NeRF featurisation (models.py:481-507): means [N, M, 3] in [-bound, bound], stds [N, M] ->
    [N, L*C (+ L)] = mean over the M multisamples of encoder features x erf down-weighting (+ featurized_w columns).
    Differentiable w.r.t. the table only (zip-NeRF detaches the sample positions, models.py:208-209)."""

    @staticmethod
    def forward(ctx, means, stds, embeddings, offsets, grid_sizes, level_gain, bound, per_level_scale, base_resolution,
                gridtype, align_corners, interpolation):
        _require_cuda(means, "grid_encode_multisample")
        if embeddings.dtype != torch.float32 or torch.is_autocast_enabled():
            raise RuntimeError("snerf_b200.gridencoder.grid_encode_multisample: fp32 only (run it outside autocast)")
        means = means.contiguous().float()
        stds = stds.contiguous().float()
        N, M, D = means.shape
        if D != 3 or tuple(stds.shape) != (N, M):
            raise RuntimeError("grid_encode_multisample: means must be [N, M, 3] and stds [N, M]")
        L = offsets.shape[0] - 1
        Cdim = embeddings.shape[1]
        width = L * Cdim + (L if level_gain is not None else 0)
        embeddings = embeddings.contiguous()
        grid_sizes = grid_sizes.to(torch.int32).contiguous()
        # the kernels write / read vector pairs: an odd row width (even level_dim, odd num_levels, featurized_w columns)
        # gets a row stride padded to the next even number; the caller sees the [N, width] view
        pitch = width + (width & 1)
        out = torch.empty(N, pitch, device=means.device, dtype=torch.float32)
        d = _desc(3, Cdim, L, np.log2(per_level_scale), base_resolution, gridtype, align_corners, interpolation, torch.float32)
        with torch.cuda.device(means.device):
            _lib.check(_lib.load().snerf_grid_encode_ms_fwd(C.byref(d), _lib.ptr(means), _lib.ptr(stds), float(bound),
                                                            _lib.ptr(embeddings), _lib.ptr(offsets), _lib.ptr(grid_sizes),
                                                            _lib.ptr(level_gain), _lib.ptr(out), pitch, N, M,
                                                            _lib.stream_ptr(means.device)), "snerf_grid_encode_ms_fwd")
        ctx.save_for_backward(means, stds, embeddings, offsets, grid_sizes)
        ctx.cfg = (d, float(bound), N, M)
        return out[:, :width] if pitch != width else out

    @staticmethod
    def backward(ctx, grad):
        means, stds, embeddings, offsets, grid_sizes = ctx.saved_tensors
        d, bound, N, M = ctx.cfg
        grad = grad.float()
        if grad.shape[1] & 1:                    # same even row stride as the forward
            grad = torch.nn.functional.pad(grad, (0, 1))
        grad = grad.contiguous()
        grad_embeddings = torch.zeros_like(embeddings)
        with torch.cuda.device(means.device):
            _lib.check(_lib.load().snerf_grid_encode_ms_bwd(C.byref(d), _lib.ptr(grad), grad.shape[1], _lib.ptr(means),
                                                            _lib.ptr(stds), bound, _lib.ptr(offsets), _lib.ptr(grid_sizes),
                                                            _lib.ptr(grad_embeddings), N, M, _lib.stream_ptr(means.device)),
                       "snerf_grid_encode_ms_bwd")
        return (None, None, grad_embeddings) + (None,) * 9


grid_encode_multisample = _grid_encode_ms.apply


class GridEncoder(nn.Module):
    """grid.py:96-200 -- same constructor, buffers and parameter, so checkpoints are interchangeable."""

    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, gridtype='hash', align_corners=False,
                 interpolation='linear', init_std=1e-4):
        super().__init__()
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = _gridtype_to_id[gridtype]
        self.interpolation = interpolation
        self.interp_id = _interp_to_id[interpolation]
        self.align_corners = align_corners
        self.init_std = init_std

        sizes, starts = _level_table(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size, align_corners)
        total = int(starts[-1])
        self.max_params = 2 ** log2_hashmap_size
        self.register_buffer('offsets', torch.from_numpy(starts))
        self.register_buffer('idx', torch.repeat_interleave(torch.arange(num_levels), torch.from_numpy(np.diff(starts).astype(np.int64))))
        self.register_buffer('grid_sizes', torch.from_numpy(sizes))
        self.n_params = self.offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(total, level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        std = self.init_std
        self.embeddings.data.uniform_(-std, std)

    def __repr__(self):
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> {int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))} "
                f"per_level_scale={self.per_level_scale:.4f} params={tuple(self.embeddings.shape)} gridtype={self.gridtype} "
                f"align_corners={self.align_corners} interpolation={self.interpolation}")

    def forward(self, inputs, bound=1, cal_input_grad=False):
        """inputs [..., input_dim] in [-bound, bound] -> [..., num_levels * level_dim] (grid.py:155-176)."""
        inputs = (inputs + bound) / (2 * bound)
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        outputs = grid_encode(inputs, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                              inputs.requires_grad, self.gridtype_id, self.align_corners, self.interp_id)
        return outputs.view(prefix_shape + [self.output_dim])

    @torch.no_grad()
    def level_gain(self):
        """sqrt(init_std^2 + per-level mean of |embedding|^2) [L] -- the `scale_featurization` factor of
        models.py:496-503 (torch_scatter.segment_coo in the reference), one reduction kernel over the table."""
        _require_cuda(self.embeddings, "level_gain")
        emb = self.embeddings.detach().float().contiguous()
        L = self.offsets.shape[0] - 1
        scratch = torch.zeros(L, device=emb.device, dtype=torch.float64)
        gain = torch.empty(L, device=emb.device, dtype=torch.float32)
        d = _desc(3, emb.shape[1], L, np.log2(self.per_level_scale), self.base_resolution, self.gridtype_id,
                  self.align_corners, self.interp_id, torch.float32)
        with torch.cuda.device(emb.device):
            _lib.check(_lib.load().snerf_grid_level_gain(C.byref(d), _lib.ptr(emb), _lib.ptr(self.offsets), float(self.init_std),
                                                         _lib.ptr(scratch), _lib.ptr(gain), _lib.stream_ptr(emb.device)),
                       "snerf_grid_level_gain")
        return gain

    def encode_multisample(self, means, stds, bound=1, scale_featurization=True):
        """The feature vector zip-NeRF's MLP.predict_density feeds its density layer (models.py:481-507), in one
        kernel: means [..., M, 3] in [-bound, bound], stds [..., M] -> [..., L*C (+ L)]."""
        if self.input_dim != 3:
            raise RuntimeError("encode_multisample: input_dim must be 3")
        if means.requires_grad or stds.requires_grad:
            raise RuntimeError("encode_multisample is differentiable w.r.t. the table only: detach means / stds "
                               "(zip-NeRF does, models.py:208-209) or use forward() for pose refinement")
        prefix = list(means.shape[:-2])
        M = means.shape[-2]
        gain = self.level_gain() if scale_featurization else None
        out = grid_encode_multisample(means.reshape(-1, M, 3), stds.reshape(-1, M), self.embeddings, self.offsets,
                                      self.grid_sizes, gain, bound, self.per_level_scale, self.base_resolution,
                                      self.gridtype_id, self.align_corners, self.interp_id)
        return out.view(prefix + [out.shape[-1]])

    @torch.autocast("cuda", enabled=False)
    def grad_total_variation(self, weight=1e-7, inputs=None, bound=1, B=1000000):
        """Adds the total-variation gradient to `embeddings.grad` (grid.py:178-200); call between backward() and step()."""
        D = self.input_dim
        if inputs is None:
            inputs = torch.rand(B, self.input_dim, device=self.embeddings.device)
        else:
            inputs = ((inputs + bound) / (2 * bound)).view(-1, self.input_dim)
            B = inputs.shape[0]
        if self.embeddings.grad is None:
            raise ValueError('grad is None, should be called after loss.backward() and before optimizer.step()!')
        _require_cuda(self.embeddings, "grad_total_variation")
        inputs = inputs.contiguous().float()
        d = _desc(D, self.embeddings.shape[1], self.offsets.shape[0] - 1, np.log2(self.per_level_scale),
                  self.base_resolution, self.gridtype_id, self.align_corners, 0, self.embeddings.dtype)
        with torch.cuda.device(inputs.device):
            _lib.check(_lib.load().snerf_grid_grad_tv(C.byref(d), _lib.ptr(inputs), _lib.ptr(self.embeddings.data),
                                                      _lib.ptr(self.embeddings.grad), _lib.ptr(self.offsets), float(weight), B,
                                                      _lib.stream_ptr(inputs.device)), "snerf_grid_grad_tv")
