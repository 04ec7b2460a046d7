"""A few config-3 training steps (512 rays, 64c+128f) for profiling:  python tools/train_steps.py [steps] [rays] [fp32|tf32|bf16]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                               # noqa: E402
import snerf_b200                                          # noqa: E402
from snerf_b200 import make_query_fn, render_rays          # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    n = int(sys.argv[2]) if len(sys.argv) > 2 else bench.TRAIN_RAYS
    dev = torch.device("cuda", 0)
    (net_c, net_f), _ = bench.make_networks(dev)
    q, _, _ = make_query_fn()
    from snerf_b200.optim import FlatAdam
    opt = FlatAdam([net_c, net_f], lr=5e-4)
    if len(sys.argv) > 3:
        snerf_b200.set_train_precision(sys.argv[3])
    rs = np.random.RandomState(0)
    c2w, O = bench.camera_rays_numpy(0)
    o, d = O.pinhole_rays(bench.H, bench.W, bench.FOCAL, c2w, [bench.CX, bench.CY])
    idx = rs.choice(bench.H * bench.W, n, replace=False)
    rb = torch.from_numpy(O.ray_batch(o.reshape(-1, 3)[idx], d.reshape(-1, 3)[idx], bench.NEAR, bench.FAR)).to(dev)
    tgt = torch.rand(n, 3, device=dev)
    dep = 1.0 / (torch.rand(n, device=dev) * 98 + 2)        # LiDAR target as disparity
    conf = torch.rand(n, device=dev)
    snerf_b200.set_mode("fp32")
    from snerf_b200.losses import RgbDepthLoss
    crit = RgbDepthLoss(bench.DEPTH_LAMBDA, bench.COARSE_DEPTH_MULT, disparity_depth=False, rgb0_weight=1.0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    import time
    for it in range(steps):
        torch.cuda.synchronize()
        t_host = time.perf_counter()
        ev[0].record()
        out = render_rays(rb, net_c, q, bench.NC, N_importance=bench.NF, network_fine=net_f, perturb=1.0, raw_noise_std=1.0)
        loss = crit(out["rgb_map"], tgt, out["disp_map"], out["disp0"], dep, conf, rgb_coarse=out["rgb0"])
        ev[1].record()
        opt.zero_grad()
        loss.backward()
        ev[2].record()
        opt.step()
        ev[3].record()
        t_host = (time.perf_counter() - t_host) * 1e3
        torch.cuda.synchronize()
        print(f"step {it}: fwd+loss {ev[0].elapsed_time(ev[1]):.2f} ms, bwd {ev[1].elapsed_time(ev[2]):.2f} ms, "
              f"adam {ev[2].elapsed_time(ev[3]):.2f} ms, total {ev[0].elapsed_time(ev[3]):.2f} ms, host enqueue {t_host:.2f} ms, loss {float(loss):.4f}")


if __name__ == "__main__":
    main()
