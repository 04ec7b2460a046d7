"""The training objective of `s-nerf/train.py` on libsnerf_b200.so: the rgb and
depth / confidence losses consumed straight from the renderer's outputs in ONE reduction kernel (+ one backward kernel),
instead of ~80 torch launches per step.

    loss = RgbLoss(rgb_fine, target)                                       (model/loss_factory.py:5-11, train.py:149)
         + depth_lambda * calc_depth_loss(...)                             (model/confidence.py:211-226, train.py:205-209)
    calc_depth_loss = mean over rays with target_depth != 0 of
                      confidence * (|f(depth) - f(t)| + coarse_depth_mult * |f(depth_coarse) - f(t)|)     (DepthLoss :26-37)
    f(x) = 1/x if disparity_depth else x

`RgbDepthLoss` takes the same hyper-parameters the reference reads from `args` (depth_lambda, coarse_depth_mult,
disparity_depth); `rgb0_weight` adds the vanilla-NeRF coarse colour term (0 = the reference's objective).  Differentiable
w.r.t. rgb, rgb0, depth, depth0 and confidence.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib


def _c(t):
    return None if t is None else t.detach().contiguous().float()


class _fused_loss(Function):
    @staticmethod
    def forward(ctx, rgb, rgb0, target, depth, depth0, target_depth, confidence, depth_lambda, c_weight, rgb0_weight, disparity):
        if not rgb.is_cuda:
            raise RuntimeError("snerf_b200.losses: tensors must live on a CUDA sm_100 device (no CPU fallback)")
        dev = rgb.device
        t = [_c(x) for x in (rgb, rgb0, target, depth, depth0, target_depth, confidence)]
        N = t[0].shape[0]
        if t[0].shape != (N, 3) or t[2].shape != (N, 3) or any(x is not None and x.numel() != N for x in t[3:]):
            raise RuntimeError("snerf_b200.losses: rgb / target must be [N, 3], depths / confidence [N]")
        opts = _lib.LossOpts(float(depth_lambda), float(c_weight), float(rgb0_weight if rgb0 is not None else 0.0), int(bool(disparity)))
        scratch = torch.zeros(5, dtype=torch.float64, device=dev)
        stats = torch.empty(5, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().snerf_loss_fwd(C.byref(opts), *[_lib.ptr(x) for x in t], N, _lib.ptr(scratch), _lib.ptr(stats),
                                                  _lib.stream_ptr(dev)), "snerf_loss_fwd")
        ctx.opts, ctx.t, ctx.stats = opts, t, stats
        ctx.mark_non_differentiable(stats)
        return stats[0], stats

    @staticmethod
    def backward(ctx, g, _g_stats):
        t, N = ctx.t, ctx.t[0].shape[0]
        dev = t[0].device
        need = ctx.needs_input_grad          # rgb, rgb0, target, depth, depth0, target_depth, confidence, ...
        new = lambda like, want: torch.empty_like(like) if (want and like is not None) else None
        g_rgb, g_rgb0 = new(t[0], need[0]), new(t[1], need[1])
        g_d, g_d0, g_c = new(t[3], need[3]), new(t[4], need[4]), new(t[6], need[6])
        g = g.detach().contiguous().float().reshape(1)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().snerf_loss_bwd(C.byref(ctx.opts), *[_lib.ptr(x) for x in t], N, _lib.ptr(ctx.stats), _lib.ptr(g),
                                                  _lib.ptr(g_rgb), _lib.ptr(g_rgb0), _lib.ptr(g_d), _lib.ptr(g_d0), _lib.ptr(g_c),
                                                  _lib.stream_ptr(dev)), "snerf_loss_bwd")
        return g_rgb, g_rgb0, None, g_d, g_d0, None, g_c, None, None, None, None


class RgbDepthLoss(nn.Module):
    """`RgbLoss` + `DepthLoss` / `calc_depth_loss` of the reference as one fused op.

    forward(rgb, target, depth=None, depth_coarse=None, target_depth=None, confidence=None, rgb_coarse=None) -> scalar loss;
    `.last` holds {img_loss, depth_loss, masked, img_loss_coarse} of the latest call (device scalars, for logging / PSNR)."""

    def __init__(self, depth_lambda=0.1, coarse_depth_mult=0.2, disparity_depth=True, rgb0_weight=0.0):
        super().__init__()
        self.depth_lambda, self.c_weight, self.disparity, self.rgb0_weight = depth_lambda, coarse_depth_mult, disparity_depth, rgb0_weight
        self.last = {}

    def forward(self, rgb, target, depth=None, depth_coarse=None, target_depth=None, confidence=None, rgb_coarse=None):
        loss, stats = _fused_loss.apply(rgb, rgb_coarse if self.rgb0_weight else None, target, depth, depth_coarse, target_depth,
                                        confidence, self.depth_lambda, self.c_weight, self.rgb0_weight, self.disparity)
        self.last = {"img_loss": stats[1], "depth_loss": stats[2], "masked": stats[3], "img_loss_coarse": stats[4]}
        return loss


class _proposal_loss(Function):
    @staticmethod
    def forward(ctx, s_vals_f, weights_f, s_vals_c, weights_c, weight):
        if not weights_c.is_cuda:
            raise RuntimeError("snerf_b200.losses: tensors must live on a CUDA sm_100 device (no CPU fallback)")
        dev = weights_c.device
        sf, wf, sc, wc = [_c(x) for x in (s_vals_f, weights_f, s_vals_c, weights_c)]
        N, Sf, Sc = wf.shape[0], wf.shape[1], wc.shape[1]
        if sf.shape != (N, Sf + 1) or sc.shape != (N, Sc + 1) or wc.shape[0] != N:
            raise RuntimeError("ProposalLoss: s_vals must be [N, S+1] and weights [N, S]")
        scratch = torch.zeros(2, dtype=torch.float64, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        grad = torch.empty_like(wc) if ctx.needs_input_grad[3] else None
        with torch.cuda.device(dev):
            _lib.check(_lib.load().snerf_proposal_loss(_lib.ptr(sf), _lib.ptr(wf), _lib.ptr(sc), _lib.ptr(wc), N, Sf, Sc, float(weight),
                                                       _lib.ptr(scratch), _lib.ptr(loss), _lib.ptr(grad), _lib.stream_ptr(dev)),
                       "snerf_proposal_loss")
        ctx.grad_wc = grad
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        return None, None, None, (ctx.grad_wc * g if ctx.grad_wc is not None else None), None


class ProposalLoss(nn.Module):
    """`ProposalLoss` of the reference (model/loss_factory.py:54-73), same call signature; value and the gradient
    w.r.t. `weights_c` (the fine histogram is detached there) come from one kernel launch."""

    def __init__(self, proposal_lambda=1.0):
        super().__init__()
        self.weight = proposal_lambda

    def forward(self, s_vals_f, weights_f, s_vals_c, weights_c):
        return _proposal_loss.apply(s_vals_f, weights_f, s_vals_c, weights_c, self.weight)
