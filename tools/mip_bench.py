"""Throughput of the mip-NeRF path at the shipped configuration (hidden 1024, rgb_layer 3,
128 + 128 samples, eval: randomized = False): rays/s and achieved tensor TFLOP/s.  Called by bench.py (reported, not the
headline);  python tools/mip_bench.py [rays]  prints the JSON object on its own."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def flop_per_ray(hidden=1024, rgb_layer=3, n_samples=128, n_fine=128, feat=96, prop=256, cond=27):
    """MLP FLOPs per ray (2 x MAC, unpadded shapes of models.py:217-325)."""
    p = feat * prop + 3 * prop * prop + prop
    m = feat * hidden + 4 * hidden * hidden + (hidden + feat) * hidden + 2 * hidden * hidden + hidden       # trunk + density
    m += hidden * hidden + (hidden + cond) * 128 + (rgb_layer - 1) * 128 * 128 + 128 * 3                    # bottleneck, condition, rgb
    return 2 * (n_samples * p + (n_fine - 1) * m)


def run(dev, n_rays=8192, steps=5, seed=0):
    from snerf_b200.models import MipNerfModel, Rays
    from tools import synth
    model = MipNerfModel(no_warp_sample=0, ray_shape="cone", fn=1, max_deg_point=16, radius=3.0, transform_idx=0, real=True,
                         rgb_layer=3, hidden_layer=1024, density_noise=0.0, n_samples=128, proposal_loss=False, N_fine=128).to(dev)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("density_layer.bias"):
                p.fill_(-3.0)
    c2w = synth.camera(0)
    o, d = synth.pinhole_rays(900, 1600, 1266.4, c2w, [816.3, 491.5])
    idx = np.random.RandomState(seed).choice(900 * 1600, n_rays, replace=False)
    o, d = torch.from_numpy(o.reshape(-1, 3)[idx]).to(dev), torch.from_numpy(d.reshape(-1, 3)[idx]).to(dev)
    vd = d / d.norm(dim=-1, keepdim=True)
    one = torch.ones(n_rays, 1, device=dev)
    rays = Rays(o, d, vd, one * (2.0 / np.sqrt(12.0) / 1266.4), one, one * 1.8, one * 110.0, None)
    with torch.no_grad():
        for _ in range(2):
            out = model(rays, False, False, None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = model(rays, False, False, None)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    f = flop_per_ray()
    return {"workload": f"mip-NeRF path of configs/nuScenes_depth_6cams (proposal 4x256 + MLP 8x1024 + bottleneck + 3 condition layers, "
                        f"128 + 128 samples, eval), {n_rays} rays per call (the reference's eval chunk)",
            "rays": n_rays, "ms_per_call": ms, "value": n_rays / (ms * 1e-3), "unit": "rays/s", "flop_per_ray": f,
            "tflops": n_rays / (ms * 1e-3) * f / 1e12, "launches_per_call": 4 + 4 + 8 + 1 + 3 + 1 + 2,
            "dtype": "bf16 operands and activations / f32 accumulate (tcgen05), fp32 sampling / encoding / compositing",
            "outputs_finite": bool(torch.isfinite(out[1][0]).all().item())}


if __name__ == "__main__":
    print(json.dumps(run(torch.device("cuda", 0), int(sys.argv[1]) if len(sys.argv) > 1 else 8192)))
