"""Per-kernel device times of the config-3 training step (torch.profiler / CUPTI; GPU box):
    python tools/train_profile.py [precision] [steps]
Runs bench.py's training step (fused forward + fused loss + backward kernels + in-place all-reduce + fused Adam) and
prints every kernel's launches per step and mean duration.  Diagnostic only: its numbers are not bench values."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                               # noqa: E402
import snerf_b200                                          # noqa: E402
from snerf_b200 import make_query_fn, render_rays          # noqa: E402
from snerf_b200.losses import RgbDepthLoss                 # noqa: E402
from snerf_b200.optim import FlatAdam                      # noqa: E402


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    n = bench.TRAIN_RAYS
    dev = torch.device("cuda", 0)
    (net_c, net_f), _ = bench.make_networks(dev)
    q, _, _ = make_query_fn()
    opt = FlatAdam([net_c, net_f], lr=5e-4)
    grads = opt.grads
    rs = np.random.RandomState(0)
    c2w, O = bench.camera_rays_numpy(0)
    o, d = O.pinhole_rays(bench.H, bench.W, bench.FOCAL, c2w, [bench.CX, bench.CY])
    idx = rs.choice(bench.H * bench.W, n, replace=False)
    rb = torch.from_numpy(O.ray_batch(o.reshape(-1, 3)[idx], d.reshape(-1, 3)[idx], bench.NEAR, bench.FAR)).to(dev)
    tgt = torch.rand(n, 3, device=dev)
    dep = 1.0 / (torch.rand(n, device=dev) * 98 + 2)
    conf = torch.rand(n, device=dev)
    snerf_b200.set_mode("fp32")
    snerf_b200.set_train_precision(prec)
    crit = RgbDepthLoss(bench.DEPTH_LAMBDA, bench.COARSE_DEPTH_MULT, disparity_depth=False, rgb0_weight=1.0)

    def step():
        out = render_rays(rb, net_c, q, bench.NC, N_importance=bench.NF, network_fine=net_f, perturb=1.0, raw_noise_std=1.0)
        loss = crit(out["rgb_map"], tgt, out["disp_map"], out["disp0"], dep, conf, rgb_coarse=out["rgb0"])
        grads.zero()
        loss.backward()
        grads.all_reduce()
        opt.step()
        return loss

    def infer():   # the same rays through the inference instantiation of the forward kernel (no stores), for comparison
        snerf_b200.set_mode("bf16")
        with torch.no_grad():
            render_rays(rb, net_c, q, bench.NC, N_importance=bench.NF, network_fine=net_f, perturb=1.0, raw_noise_std=1.0)
        snerf_b200.set_mode("fp32")

    for _ in range(5):
        step()
    infer()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record(); torch.cuda.synchronize()
    print(f"[{prec}] {e0.elapsed_time(e1) / steps:.3f} ms per step (events, no profiler)")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            step()
        for _ in range(steps):
            infer()
        torch.cuda.synchronize()
    rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
    rows.sort(key=lambda r: -r[2])
    tot = sum(r[2] for r in rows)
    print(f"kernel time per step {tot / steps / 1e3:.3f} ms over {sum(r[1] for r in rows) / steps:.1f} launches per step")
    for k, c, t in rows:
        print(f"  {t / steps:9.1f} us/step  {c / steps:5.1f} x  {t / c:8.1f} us   {k[:110]}")


if __name__ == "__main__":
    main()
