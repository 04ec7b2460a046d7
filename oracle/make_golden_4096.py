"""Generate tests/golden/cfg2_*_4096.npz: the 4096-ray slice of BASELINE configs[1] SURVEY.md section 7-1 asks for, rendered by
the UNMODIFIED reference (s-nerf/model/render.py:281-409, torch-CPU fp32), for weight sets (A) default init and (B) "peaky".
Only the reference's output dict and the inverse-CDF bin indices are stored (outputs 560 B/ray + 128 uint8 indices per ray);
stage intermediates stay with the 128-ray fixtures of oracle/make_golden.py.

    python oracle/make_golden_4096.py          (build container: /root/reference must be mounted)
TEST INFRASTRUCTURE ONLY."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import make_golden as MG, ref_import, snerf_oracle as O    # noqa: E402


def main():
    ref_render, ref_helpers = ref_import.load()
    o, d, idx = MG.nuscenes_like_rays(ref_helpers, 4096, seed=12)
    rb = O.pack_ray_batch(o, d, 1.8, 110.0)
    for name, seeds, gain, sb in (("cfg2_default_4096", (10, 11), 1.0, 0.0), ("cfg2_peaky_4096", (20, 21), 1.5, 1.0)):
        pc = O.make_nerf_params(seeds[0], trunk_gain=gain, sigma_bias=sb)
        pf = O.make_nerf_params(seeds[1], trunk_gain=gain, sigma_bias=sb)
        ret, net_c, net_f, qfn = MG.run_reference(ref_render, ref_helpers, rb, pc, pf, 64, 128, 8, 256)
        # the reference's own searchsorted indices (run_nerf_helpers.py:336-379) recomputed from its coarse outputs
        z = torch.from_numpy(ret["z_vals_map"]); w = torch.from_numpy(ret["weights"])
        z_mid = .5 * (z[..., 1:] + z[..., :-1])
        wts = w[..., 1:-1] + 1e-5
        pdf = wts / torch.sum(wts, -1, keepdim=True)
        cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
        u = torch.linspace(0., 1., steps=128).expand(list(cdf.shape[:-1]) + [128]).contiguous()
        inds = torch.searchsorted(cdf, u, right=True).numpy()
        assert inds.max() <= 63
        keep = ("rgb_map", "disp_map", "acc_map", "depth_map", "weights", "z_vals_map", "rgb0", "disp0", "acc0", "z_std")
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), ray_batch=rb, seed_coarse=seeds[0], seed_fine=seeds[1],
                            trunk_gain=gain, sigma_bias=sb, D=8, W=256, Nc=64, Nf=128, inds=inds.astype(np.uint8),
                            torch_version=torch.__version__, **{"out_" + k: ret[k] for k in keep})
        print(name, "rgb mean", float(ret["rgb_map"].mean()), "acc mean", float(ret["acc_map"].mean()))


if __name__ == "__main__":
    main()
