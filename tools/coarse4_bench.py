"""Throughput of the tensor-core renderer on the shipped configs' network pair (coarse NeRF 4x256, fine NeRF 8x256;
create_nerf with netdepth = 4 / netdepth_fine = 8) next to the 8x256 pair, one 1600x900 frame per step.

    python tools/coarse4_bench.py [steps]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import snerf_b200                                              # noqa: E402
from snerf_b200 import NeRF, make_query_fn, render_rays        # noqa: E402
from tools import synth                                        # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    dev = torch.device("cuda", 0)
    kw = dict(W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    fine = NeRF(D=8, **kw)
    fine.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in synth.nerf_params(21, trunk_gain=1.5, sigma_bias=1.0).items()})
    c8 = NeRF(D=8, **kw)
    c8.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in synth.nerf_params(20, trunk_gain=1.5, sigma_bias=1.0).items()})
    c4 = NeRF(D=4, **kw)
    c4.load_state_dict({k: v for k, v in c8.state_dict().items() if not k.startswith("pts_linears.") or int(k.split(".")[1]) < 4},
                       strict=False)
    with torch.no_grad():          # layer 3 of the 4-layer trunk has no skip input: plain [256, 256]
        c4.pts_linears[3].weight.copy_(c8.pts_linears[3].weight)
    nets = {"8x256 + 8x256": c8.to(dev).requires_grad_(False), "4x256 + 8x256": c4.to(dev).requires_grad_(False)}
    fine = fine.to(dev).requires_grad_(False)
    q, _, _ = make_query_fn()
    n = 1600 * 900
    rs = np.random.RandomState(0)
    d = rs.standard_normal((n, 3)).astype(np.float32); d[:, 2] = -1
    rb = torch.from_numpy(synth.ray_batch(rs.standard_normal((n, 3)).astype(np.float32) * 0.1, d, 1.8, 110.0)).to(dev)
    res = {}
    for mode in ("bf16", "fp16x3"):
        snerf_b200.set_mode(mode)
        for name, net in nets.items():
            with torch.no_grad():
                for _ in range(3):
                    render_rays(rb, net, q, 64, N_importance=128, network_fine=fine)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    render_rays(rb, net, q, 64, N_importance=128, network_fine=fine)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            res[f"{mode} {name}"] = {"ms_per_frame": ms, "rays_per_s": n / ms * 1e3}
    snerf_b200.set_mode("fp32")
    print(json.dumps(res))


if __name__ == "__main__":
    main()
