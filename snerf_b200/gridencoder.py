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

        resolutions, offsets, offset = [], [], 0
        self.max_params = 2 ** log2_hashmap_size
        for i in range(num_levels):
            resolution = int(np.ceil(base_resolution * per_level_scale ** i))
            resolution = resolution if align_corners else resolution + 1
            params_in_level = min(self.max_params, resolution ** input_dim)
            params_in_level = int(np.ceil(params_in_level / 8) * 8)
            resolutions.append(resolution)
            offsets.append(offset)
            offset += params_in_level
        offsets.append(offset)
        offsets = torch.from_numpy(np.array(offsets, dtype=np.int32))
        self.register_buffer('offsets', offsets)
        idx = torch.empty(offset, dtype=torch.long)
        for i in range(self.num_levels):
            idx[offsets[i]:offsets[i + 1]] = i
        self.register_buffer('idx', idx)
        self.register_buffer('grid_sizes', torch.from_numpy(np.array(resolutions, dtype=np.int32)))
        self.n_params = offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(offset, level_dim))
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
