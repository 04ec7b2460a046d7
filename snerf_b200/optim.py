"""Optimizer + training-step plumbing for the fused training path.

The reference trains with `torch.optim.Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))` (model/render.py:222,
train.py:100) and an exponential learning-rate decay written into `optimizer.param_groups` every iteration
(train.py).  `FlatAdam` is that optimizer for networks whose parameters AND gradients live in one flat buffer each
(`snerf_b200.parallel.FlatGradients`): the update is one kernel (`snerf_adam_step`) instead of torch's multi-tensor loop
over 48 small tensors, and step count / learning rate are device scalars, so a whole training step -- forward, loss,
backward, gradient all-reduce, update -- can be captured in ONE CUDA graph (`GraphedTrainStep`).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .parallel import FlatGradients

__all__ = ["FlatAdam", "GraphedTrainStep"]


class FlatAdam:
    """Adam over the parameters of `nets` (snerf_b200.NeRF modules), same update rule and defaults as torch.optim.Adam.

    The parameters are re-pointed at views of one flat fp32 buffer (values preserved; `state_dict()` of the modules is
    unaffected), the gradients at views of `FlatGradients`.  `param_groups[0]['lr']` may be rewritten between steps like
    the reference's decay loop does; it is copied to the device scalar the kernel reads when it changes."""

    def __init__(self, nets, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grads: FlatGradients | None = None):
        self.nets = [n for n in nets if n is not None]
        slots = [p for n in self.nets for _, _, p in n._slots()]
        dev = slots[0].device
        if dev.type != "cuda":
            raise RuntimeError("snerf_b200.FlatAdam: parameters must live on a CUDA device (no CPU fallback)")
        self.flat = torch.empty(sum(p.numel() for p in slots), dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in slots:
                view = self.flat[off:off + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view
                off += p.numel()
        for n in self.nets:
            n.invalidate_packed()
        self.grads = grads if grads is not None else FlatGradients(self.nets)
        if self.grads.flat.numel() != self.flat.numel():
            raise RuntimeError("FlatAdam: the FlatGradients buffer covers a different parameter set")
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self._lr_host = float(lr)
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        self.param_groups = [{"params": slots, "lr": float(lr), "betas": betas, "eps": eps, "weight_decay": weight_decay}]

    def zero_grad(self, set_to_none: bool = False):
        self.grads.zero()

    def sync_lr(self):
        """Copy a rewritten `param_groups[0]['lr']` to the device scalar (call outside a captured graph)."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_host:
            self.lr_dev.fill_(lr)
            self._lr_host = lr

    @torch.no_grad()
    def step(self, _sync_lr: bool = True):
        if _sync_lr:
            self.sync_lr()
        dev = self.flat.device
        with torch.cuda.device(dev):
            _lib.check(_lib.load().snerf_adam_step(_lib.ptr(self.flat), _lib.ptr(self.grads.flat), _lib.ptr(self.exp_avg),
                                                   _lib.ptr(self.exp_avg_sq), self.flat.numel(), _lib.ptr(self.lr_dev),
                                                   float(self.betas[0]), float(self.betas[1]), float(self.eps),
                                                   float(self.weight_decay), _lib.ptr(self.step_count), _lib.stream_ptr(dev)),
                       "snerf_adam_step")
        for n in self.nets:          # the kernel wrote the parameters behind autograd's back: packed images are stale
            n.invalidate_packed()

    def state_dict(self):
        return {"step": self.step_count.clone(), "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "lr": float(self.param_groups[0]["lr"])}

    def load_state_dict(self, sd):
        self.step_count.copy_(sd["step"]); self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.param_groups[0]["lr"] = float(sd["lr"])
        self.sync_lr()


class GraphedTrainStep:
    """One training iteration of the reference (train.py:110-221: render the batch, loss, backward, optimizer step; here
    plus the gradient all-reduce that replaces DataParallel) captured in ONE CUDA graph and replayed per step.

        step = GraphedTrainStep(batch_example, fn, optimizer)      # fn(batch) -> loss (scalar tensor)
        loss = step(batch)                                          # copies `batch` into the static input, replays

    `fn` runs under autograd on a static copy of the batch; everything it launches (our kernels through the C ABI, torch's
    RNG, the fused loss) goes to the capture stream.  `optimizer` is a FlatAdam (its gradient buffer is zeroed, all-reduced
    in place over the default process group and applied inside the graph)."""

    def __init__(self, batch: torch.Tensor, fn, optimizer: FlatAdam, warmup: int = 3, average: bool = True):
        self.opt, self.fn, self.average = optimizer, fn, average
        self.static_in = batch.detach().clone()
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=batch.device)
        side.wait_stream(torch.cuda.current_stream(batch.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):                 # allocator, packed images, NCCL communicators: all warm before capture
                self._body()
        torch.cuda.current_stream(batch.device).wait_stream(side)
        torch.cuda.synchronize(batch.device)
        with torch.cuda.graph(self.graph):
            self.static_loss = self._body()

    def _body(self):
        self.opt.grads.zero()
        with torch.enable_grad():
            loss = self.fn(self.static_in)
            loss.backward()
        self.opt.grads.all_reduce(average=self.average)
        self.opt.step(_sync_lr=False)
        return loss.detach()

    def __call__(self, batch: torch.Tensor) -> torch.Tensor:
        self.opt.sync_lr()
        self.static_in.copy_(batch, non_blocking=True)
        self.graph.replay()
        for n in self.opt.nets:      # the replayed Adam kernel rewrote the parameters behind autograd's back: a later eager
            n.invalidate_packed()    # render (evaluation between steps) must repack instead of trusting the version counters
        return self.static_loss
