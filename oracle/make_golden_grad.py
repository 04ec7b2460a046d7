"""Generate `tests/golden/grad_cfg3.npz`: parameter gradients of the UNMODIFIED reference.

    python oracle/make_golden_grad.py

TEST INFRASTRUCTURE ONLY (needs /root/reference; never runs on the GPU box).  The reference's training
step differentiates `render_rays` with torch autograd (train path of BASELINE config 3: perturb=1,
raw_noise_std=1).  This script runs the reference's own `render_rays` (render.py:281-409) with
`pytest=True` (its deterministic-draw hooks), contracts every differentiable output with a seeded
cotangent (`snerf_oracle_grad.cotangents`) and stores d(sum_k <out_k, G_k>)/d(parameter) for both
networks: small tensors in full, the 256-wide matrices as every 16th output row.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_import, snerf_oracle as O, snerf_oracle_grad as OG  # noqa: E402
from oracle.make_golden import nuscenes_like_rays, reference_intermediates, t  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
F32 = np.float32
ROW_STRIDE = 16


def thin(name, g):
    """What the fixture keeps of one gradient tensor."""
    if g.ndim == 2 and g.shape[0] >= 128:
        return g[::ROW_STRIDE].copy()
    return g.copy()


def main():
    ref_render, ref_helpers = ref_import.load()
    n_rays, Nc, Nf = 32, 64, 128
    o, d, idx = nuscenes_like_rays(ref_helpers, n_rays, seed=3)
    rb = O.pack_ray_batch(o, d, 1.8, 110.0)
    seeds, gain, sb = (50, 51), 1.5, 0.5
    pc = O.make_nerf_params(seeds[0], trunk_gain=gain, sigma_bias=sb)
    pf = O.make_nerf_params(seeds[1], trunk_gain=gain, sigma_bias=sb)
    net_c = ref_import.build_reference_net(ref_helpers, pc).train()
    net_f = ref_import.build_reference_net(ref_helpers, pf).train()
    qfn = ref_import.reference_query_fn(ref_helpers)
    ret = ref_render.render_rays(t(rb), net_c, qfn, Nc, retraw=True, N_importance=Nf, network_fine=net_f,
                                 perturb=1.0, raw_noise_std=1.0, pytest=True)
    G = OG.cotangents({k: tuple(v.shape) for k, v in ret.items()}, seed=77)
    loss = OG.loss_from(ret, G)
    loss.backward()
    store = dict(ray_batch=rb, pixel_index=idx, seed_coarse=seeds[0], seed_fine=seeds[1], trunk_gain=gain,
                 sigma_bias=sb, Nc=Nc, Nf=Nf, cot_seed=77, row_stride=ROW_STRIDE, loss=float(loss),
                 torch_version=torch.__version__)
    np.random.seed(0); store["t_rand"] = np.random.rand(n_rays, Nc).astype(F32)
    np.random.seed(0); store["noise0"] = (np.random.rand(n_rays, Nc) * 1.0).astype(F32)
    np.random.seed(0); store["noise1"] = (np.random.rand(n_rays, Nc + Nf) * 1.0).astype(F32)
    np.random.seed(0); store["u"] = np.random.rand(n_rays, Nf).astype(F32)
    for tag, net in (("c", net_c), ("f", net_f)):
        for name, p in net.named_parameters():
            store[f"g{tag}_{name}"] = thin(name, p.grad.numpy())
    inter = reference_intermediates(ref_helpers, rb, {k: v.detach().numpy() for k, v in ret.items()}, net_c, qfn, Nf,
                                    1.0, 1.0, False, True)
    store["mid_z_all"] = inter["z_all"]
    for k in ("rgb_map", "depth_map", "rgb0", "acc_map"):
        store["out_" + k] = ret[k].detach().numpy()
    np.savez_compressed(os.path.join(OUT, "grad_cfg3.npz"), **store)
    print("wrote grad_cfg3.npz, loss", float(loss), "size", os.path.getsize(os.path.join(OUT, "grad_cfg3.npz")))


if __name__ == "__main__":
    main()
