"""Proposal-resampling micro-benchmark (csrc/snerf_stepfun.cu) at zip-NeRF shapes: one pass of the sampling loop
(models.py:156-213) for a chunk of rays -- dilate 64 bins -> 190 bins, annealed logits, sample 64 (or 32) intervals.

    python tools/stepfun_bench.py [--rays N] [--steps K]

Algorithmic HBM bytes per ray: (S+1 + S) * 4 in, 4 jitter, (n+1) * 4 out = 780 B at S = n = 64; the kernel is
latency / issue bound (merge ranks, max-pool, scan, binary searches in shared memory), not HBM bound.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def run(dev, rays=1 << 16, S=64, n=64, steps=20, warmup=3):
    import torch
    from snerf_b200 import stepfun
    gen = torch.Generator(device=dev).manual_seed(3)
    sd = torch.sort(torch.rand(rays, S + 1, device=dev, generator=gen), dim=-1).values
    sd[:, 0], sd[:, -1] = 0.0, 1.0
    w = torch.rand(rays, S, device=dev, generator=gen) ** 6
    w = w / w.sum(-1, keepdim=True)
    res = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, kw in (("dilate_resample_rand", dict(rand=True, dilation=0.0025 + 0.5 / 64)), ("resample_det", dict(rand=None, dilation=None))):
        fn = lambda: stepfun.resample_intervals(kw["rand"], sd, w, n, dilation=kw["dilation"], domain=(0., 1.), single_jitter=True)
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name + "_ms"] = e0.elapsed_time(e1) / steps
    ms = res["dilate_resample_rand_ms"]
    bytes_per_ray = (S + 1 + S + 1 + n + 1) * 4
    return {"workload": f"{rays} rays, {S} bins -> max_dilate_weights -> {3 * S - 2} bins -> {n} intervals (models.py:156-213), single jitter",
            **res, "rays_per_s": rays / (ms * 1e-3), "hbm_bytes_per_ray": bytes_per_ray,
            "achieved_gbs": rays * bytes_per_ray / (ms * 1e-3) / 1e9, "launches_per_pass": "1 kernel + torch.linspace + torch.rand",
            "dtype": "f32", "steps": steps}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=1 << 16)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    import torch
    print(json.dumps(run(torch.device("cuda", 0), a.rays, steps=a.steps)))
