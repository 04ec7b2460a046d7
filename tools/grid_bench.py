"""Hash-grid encoder micro-benchmark at BASELINE configs[3] shapes (zip-NeRF main grid: 10 levels x 4 features,
base 16 -> 8192, 2^21-entry tables = 150 MB fp32; one render chunk = 16384 rays x 32 samples x 6 multisamples =
3,145,728 points).  Used by bench.py ("grid" object of the JSON line) and stand-alone:

    python tools/grid_bench.py [--impl ours|reference] [--points N] [--steps K]

`--impl reference` times the reference's own CUDA kernels (oracle/_ref/_gridencoder_ref.so, gridencoder.cu compiled
unmodified) driven as its wrapper drives them (grid.py:24-90: [L,B,C] outputs + permute/reshape copy, grad permute +
contiguous copy, zero-filled grad table) -- the GPU baseline this kernel family is meant to beat.

Algorithmic bytes per point (fp32, D=3, C=4, L=10): forward = 12 (input) + L*8 corners*16 B gathers (1280, served by
L2 / HBM) + L*C*4 (160, output) = 1452 B; backward = 12 + 160 (grad) + 1280 B of vector reductions.  The compulsory
HBM part is 172 B/point + the 150 MB table once per pass.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = dict(input_dim=3, num_levels=10, level_dim=4, base_resolution=16, desired_resolution=8192, log2_hashmap_size=21)
POINTS = 16384 * 32 * 6


def layout():
    """offsets / per-level scale exactly as GridEncoder.__init__ computes them (no oracle import: product-side helper)."""
    from snerf_b200.gridencoder import GridEncoder
    enc = GridEncoder(**CFG)
    return enc.offsets.clone(), float(enc.per_level_scale)


def gather_bytes(B, L=10, D=3, C=4, esz=4):
    return B * (D * 4 + L * (1 << D) * C * esz + L * C * esz)


def run(impl, dev, points=POINTS, steps=10, warmup=3):
    import torch
    offsets, pls = layout()
    off = offsets.to(dev)
    L, Cd, D = CFG["num_levels"], CFG["level_dim"], 3
    S, Hres = float(np.log2(pls)), CFG["base_resolution"]
    gen = torch.Generator(device=dev).manual_seed(7)
    emb = (torch.rand(int(offsets[-1]), Cd, device=dev, generator=gen) * 2 - 1) * 1e-1
    # zip-NeRF feeds contracted coordinates: most samples near the scene centre, multisamples clustered per ray sample
    centre = torch.rand(points // 6, 1, 3, device=dev, generator=gen)
    x = (centre + 2e-3 * torch.randn(points // 6, 6, 3, device=dev, generator=gen)).clamp(0, 1).reshape(-1, 3).contiguous()
    B = x.shape[0]
    dy = torch.randn(B, L * Cd, device=dev, generator=gen)

    if impl == "ours":
        from snerf_b200.gridencoder import grid_encode

        def fwd():
            return grid_encode(x, emb, off, pls, Hres)

        def fwd_bwd():
            e = emb.detach().requires_grad_(True)
            with torch.enable_grad():
                y = grid_encode(x, e, off, pls, Hres)
            y.backward(dy)
            return e.grad
    else:
        from oracle import build_ref_gridencoder as R
        ref = R.load()
        if ref is None:
            return {"unavailable": "oracle/_ref/_gridencoder_ref.so not built"}

        def fwd():
            out = torch.empty(L, B, Cd, device=dev)
            ref.grid_encode_forward(x, emb, off, out, B, D, Cd, L, S, Hres, None, 0, False, 0)
            return out.permute(1, 0, 2).reshape(B, L * Cd)           # grid.py:57 (materialises the copy)

        def fwd_bwd():
            y = fwd()
            g = dy.view(B, L, Cd).permute(1, 0, 2).contiguous()       # grid.py:72
            ge = torch.zeros_like(emb)                                # grid.py:77
            ref.grid_encode_backward(g, x, emb, off, ge, B, D, Cd, L, S, Hres, None, None, 0, False, 0)
            return ge

    # ---- the stage zip-NeRF actually runs per level (models.py:481-507): encode the 6 multisamples of every sample,
    # erf down-weighting, mean, scale_featurization columns -- fused into one kernel here
    Ns, Mm = B // 6, 6
    means = (x * 2 - 1).view(Ns, Mm, 3).contiguous()
    stds = torch.exp(torch.rand(Ns, Mm, device=dev, generator=gen) * 7 - 9)
    dyf = torch.randn(Ns, L * Cd + L, device=dev, generator=gen)
    if impl == "ours":
        from snerf_b200.gridencoder import GridEncoder
        enc = GridEncoder(**CFG).to(dev)
        enc.embeddings.data.copy_(emb)

        def ms_fwd():
            return enc.encode_multisample(means, stds)

        def ms_fwd_bwd():
            enc.embeddings.grad = None
            with torch.enable_grad():
                enc.encode_multisample(means, stds).backward(dyf)
            return enc.embeddings.grad
    else:
        from oracle.make_golden_grid import run_reference_multisample
        gsz = torch.from_numpy(np.array([int(np.ceil(Hres * pls ** i)) + 1 for i in range(L)], np.int32)).to(dev)

        def ms_fwd():
            return run_reference_multisample(ref, means, stds, emb, off, gsz, S, Hres)[0]

        def ms_fwd_bwd():
            with torch.enable_grad():
                return run_reference_multisample(ref, means, stds, emb, off, gsz, S, Hres, 1e-4, dyf)[1]

    res = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for name, fn in (("fwd", fwd), ("fwd_bwd", fwd_bwd), ("ms_fwd", ms_fwd), ("ms_fwd_bwd", ms_fwd_bwd)):
            for _ in range(warmup):
                fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / steps
    ms_f, ms_fb = res["fwd"], res["fwd_bwd"]
    gb_f = gather_bytes(B) / 1e9
    return {"impl": impl, "workload": f"configs[3]: zip-NeRF main grid (L=10, C=4, 16->8192, T=2^21, {emb.numel() * 4 / 1e6:.0f} MB fp32), "
                                      f"{B} points (16384 rays x 32 samples x 6 multisamples), clustered contracted coordinates",
            "points": B, "fwd_ms": ms_f, "fwd_points_per_s": B / (ms_f * 1e-3), "fwd_gather_gbs": gb_f / (ms_f * 1e-3),
            "fwd_bwd_ms": ms_fb, "fwd_bwd_points_per_s": B / (ms_fb * 1e-3), "bwd_ms": ms_fb - ms_f,
            "multisample_fwd_ms": res["ms_fwd"], "multisample_fwd_bwd_ms": res["ms_fwd_bwd"],
            "multisample_samples_per_s": Ns / (res["ms_fwd"] * 1e-3),
            "multisample_note": f"{Ns} samples x {Mm} multisamples -> [{Ns}, {L * Cd + L}] density features (models.py:481-507)",
            "algorithmic_bytes_per_point_fwd": gather_bytes(1), "compulsory_hbm_bytes_per_point_fwd": 12 + L * Cd * 4,
            "dtype": "f32", "steps": steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "both"])
    ap.add_argument("--points", type=int, default=POINTS)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    for impl in (["ours", "reference"] if a.impl == "both" else [a.impl]):
        print(json.dumps(run(impl, dev, a.points, a.steps, a.warmup)))


if __name__ == "__main__":
    main()
