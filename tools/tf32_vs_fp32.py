"""fp32 vs tf32 training precision on the same rays: output and gradient differences (GPU; debug / profile aid)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import snerf_b200                                          # noqa: E402
from oracle import snerf_oracle as O                       # noqa: E402
from snerf_b200 import make_query_fn, render_rays          # noqa: E402
from test_gpu_parity import make_net                       # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dev = torch.device("cuda", 0)
    pc = O.make_nerf_params(80, trunk_gain=1.5, sigma_bias=0.5)
    pf = O.make_nerf_params(81, trunk_gain=1.5, sigma_bias=0.5)
    rs = np.random.RandomState(11)
    d = rs.standard_normal((n, 3)).astype(np.float32); d[:, 2] = -1.0
    rb = torch.from_numpy(O.pack_ray_batch(rs.standard_normal((n, 3)).astype(np.float32) * 0.1, d, 1.8, 110.0)).to(dev)
    target = torch.from_numpy(rs.rand(n, 3).astype(np.float32)).to(dev)
    q, _, _ = make_query_fn()
    res = {}
    for prec in ("fp32", "tf32"):
        nc, nf = make_net(pc, 8, 256, dev, train=True), make_net(pf, 8, 256, dev, train=True)
        snerf_b200.set_train_precision(prec)
        out = render_rays(rb, nc, q, 64, N_importance=128, network_fine=nf, retraw=True)
        loss = ((out["rgb_map"] - target) ** 2).mean() + ((out["rgb0"] - target) ** 2).mean() + 0.01 * out["depth_map"].mean()
        loss.backward()
        torch.cuda.synchronize()
        g = {**{"c." + k: p.grad.clone() for k, p in nc.named_parameters()},
             **{"f." + k: p.grad.clone() for k, p in nf.named_parameters()}}
        res[prec] = ({k: v.detach().clone() for k, v in out.items()}, g, float(loss))
    snerf_b200.set_train_precision("fp32")
    print("loss", res["fp32"][2], res["tf32"][2])
    for k in ("rgb_map", "rgb0", "depth_map", "acc_map", "weights", "raw"):
        a, b = res["fp32"][0][k], res["tf32"][0][k]
        print(f"out {k:10s} max abs {float((a - b).abs().max()):.3e}  rel L2 {float((a - b).norm() / a.norm()):.3e}")
    for k, a in res["fp32"][1].items():
        b = res["tf32"][1][k]
        print(f"grad {k:28s} rel L2 {float((a - b).norm() / (a.norm() + 1e-30)):.3e}  cos {float((a * b).sum() / (a.norm() * b.norm() + 1e-30)):.6f}")


if __name__ == "__main__":
    main()
