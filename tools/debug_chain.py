"""Compare the gradient stores (dZ of every layer) written by the fp32 and the tf32 backward (debug aid)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                               # noqa: E402
import snerf_b200                                          # noqa: E402
from snerf_b200 import make_query_fn, render_rays          # noqa: E402


def layout(n, Nc=64, Nf=128, D=8, W=256):
    total_ch = 96 + D * W + W + W // 2
    TC, TF = (Nc + 63) // 64, (Nc + Nf + 63) // 64
    Rc, Rf = n * TC * 64, n * TF * 64
    off = 0
    out = {}
    for name, cnt in (("save_c", total_ch * Rc), ("save_f", total_ch * Rf), ("dz_c", (total_ch - 96) * Rc),
                      ("dz_f", (total_ch - 96) * Rf), ("draw_c", 4 * Rc), ("draw_f", 4 * Rf)):
        out[name] = (off, cnt)
        off += (cnt + 31) // 32 * 32
    return out, Rc, Rf, total_ch


def main():
    n = 16
    dev = torch.device("cuda", 0)
    (net_c, net_f), _ = bench.make_networks(dev)
    q, _, _ = make_query_fn()
    rs = np.random.RandomState(0)
    c2w, O = bench.camera_rays_numpy(0)
    o, d = O.pinhole_rays(bench.H, bench.W, bench.FOCAL, c2w, [bench.CX, bench.CY])
    idx = rs.choice(bench.H * bench.W, n, replace=False)
    rb = torch.from_numpy(O.pack_ray_batch(o.reshape(-1, 3)[idx], d.reshape(-1, 3)[idx], bench.NEAR, bench.FAR)).to(dev)
    tgt = torch.rand(n, 3, device=dev)
    L, Rc, Rf, total_ch = layout(n)
    stores = {}
    for prec in ("fp32", "tf32"):
        snerf_b200.set_train_precision(prec)
        net_c.zero_grad(); net_f.zero_grad()
        out = render_rays(rb, net_c, q, 64, N_importance=128, network_fine=net_f)
        ws = out["rgb_map"].grad_fn.call.ws
        loss = ((out["rgb_map"] - tgt) ** 2).mean() + ((out["rgb0"] - tgt) ** 2).mean()
        loss.backward()
        torch.cuda.synchronize()
        f = ws.view(torch.float32)
        stores[prec] = {k: f[o_:o_ + c].clone() for k, (o_, c) in L.items()}
    snerf_b200.set_train_precision("fp32")
    names = [f"h{i}" for i in range(8)] + ["feature", "views"]
    for pas, R in (("dz_c", Rc), ("dz_f", Rf)):
        a = stores["fp32"][pas].view(-1, R)
        b = stores["tf32"][pas].view(-1, R)
        for i, nm in enumerate(names):
            lo, hi = i * 256, min((i + 1) * 256, a.shape[0])
            x, y = a[lo:hi], b[lo:hi]
            print(pas, nm, "norm fp32 %.3e tf32 %.3e rel %.3e" % (float(x.norm()), float(y.norm()),
                                                                 float((x - y).norm() / (x.norm() + 1e-30))))


if __name__ == "__main__":
    main()
