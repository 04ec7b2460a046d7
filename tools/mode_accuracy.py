"""Error of each MLP-arithmetic mode against the fp64 oracle on seeded rays (GPU box; test infrastructure).

  python tools/mode_accuracy.py [n_rays]  ->  one JSON line per mode: rgb / depth / weights error vs the oracle
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import snerf_b200                                           # noqa: E402
from oracle import snerf_oracle as O                        # noqa: E402
from snerf_b200 import make_query_fn, render_rays           # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    dev = torch.device("cuda:0")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_parity import make_net as mk
    pc = O.make_nerf_params(20, trunk_gain=1.5, sigma_bias=1.0)
    pf = O.make_nerf_params(21, trunk_gain=1.5, sigma_bias=1.0)
    rs = np.random.RandomState(3)
    d = rs.standard_normal((n, 3)).astype(np.float32); d[:, 2] = -1.0
    rb = O.pack_ray_batch(rs.standard_normal((n, 3)).astype(np.float32) * 0.1, d, 1.8, 110.0)
    O.set_backend("torch", threads=16)
    ref = O.render_rays(rb, pc, pf, 64, 128)
    nc, nf = mk(pc, 8, 256, dev), mk(pf, 8, 256, dev)
    q, _, _ = make_query_fn()
    t = torch.from_numpy(rb).to(dev)
    for mode in ("fp32", "fp16x3", "fp16", "bf16"):
        snerf_b200.set_mode(mode)
        out = render_rays(t, nc, q, 64, N_importance=128, network_fine=nf, retraw=True)
        row = {"mode": mode, "rays": n}
        for k in ("rgb_map", "depth_map", "acc_map", "rgb0"):
            a, b = out[k].cpu().numpy().astype(np.float64), np.asarray(ref[k], np.float64)
            row[k] = {"l1": float(np.mean(np.abs(a - b))), "max": float(np.max(np.abs(a - b))),
                      "rel_l1": float(np.mean(np.abs(a - b)) / (np.mean(np.abs(b)) + 1e-30))}
        print(json.dumps(row))
    snerf_b200.set_mode("fp32")


if __name__ == "__main__":
    main()
