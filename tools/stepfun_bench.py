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
    # the unmodified reference's composition of the same pass (stepfun.py max_dilate_weights + sample_intervals at the call
    # site internal/models.py:183-213) on the same GPU, from the byte-compiled modules under oracle/_ref (bench leg only)
    try:
        from oracle import build_ref_python
        ref = build_ref_python.load_zip_stepfun()
    except Exception:
        ref = None
    if ref is not None:
        dil = 0.0025 + 0.5 / 64

        def ref_pass(randomized):
            t, ww = ref.max_dilate_weights(sd, w, dil, domain=(0., 1.), renormalize=True)
            t, ww = t[..., 1:-1], ww[..., 1:-1]
            logits = torch.where(t[..., 1:] > t[..., :-1], torch.log(ww + 1e-5), torch.full_like(t[..., :-1], -torch.inf))
            return ref.sample_intervals(randomized, t, logits, n, single_jitter=True, domain=(0., 1.))

        with torch.no_grad():
            ours = stepfun.resample_intervals(None, sd, w, n, dilation=dil, domain=(0., 1.), single_jitter=True)
            theirs = ref_pass(False)
            # (flat stretches of the CDF -- empty bins under peaky weights -- make the inverse ill-conditioned: the reference's
            #  own GPU cumsum / sort differ from its CPU run there too; the fixtures pin the kernel at 1e-6 on CPU-made goldens)
            diff = (ours - theirs).abs()
            res["max_abs_diff_vs_reference_det"] = float(diff.max())
            res["mean_abs_diff_vs_reference_det"] = float(diff.mean())
            res["frac_diff_gt_1e-5"] = float((diff > 1e-5).float().mean())
            for _ in range(2):
                ref_pass(True)
            torch.cuda.synchronize()
            k = max(2, steps // 4)
            e0.record()
            for _ in range(k):
                ref_pass(True)
            e1.record()
            torch.cuda.synchronize()
        res["reference_composition_ms"] = e0.elapsed_time(e1) / k
        res["speedup_vs_reference_composition"] = res["reference_composition_ms"] / res["dilate_resample_rand_ms"]
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
