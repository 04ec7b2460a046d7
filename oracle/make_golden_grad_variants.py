"""Generate `tests/golden/grad_variants.npz`: parameter gradients of the UNMODIFIED reference for the network shapes
beside the flagship NeRF(use_viewdirs=True) that `render_rays` trains.

    python oracle/make_golden_grad_variants.py

TEST INFRASTRUCTURE ONLY (needs /root/reference; never runs on the GPU box).  Cases:
  novd      NeRF(use_viewdirs=False, output_ch=5) coarse + fine      (run_nerf_helpers.py:99-100,124)
  rgb       NeRF_RGB(alpha_model) coarse + fine                      (run_nerf_helpers.py:157-206, render.py:182-208)
  nocoarse  network_fn=None, network_fine=NeRF_RGB(alpha_model)      (render.py:361-371: the coarse pass runs alpha_model)
  d4        coarse NeRF 4x256 (no live skip), fine NeRF 8x256        (create_nerf with the shipped configs' netdepth = 4
                                                                      against netdepth_fine = 8, render.py:176-201)
  w128      coarse NeRF 6x128, fine NeRF 8x256                       (netwidth != netwidth_fine)
As in make_golden_grad.py: reference `render_rays` with pytest=True draws, every differentiable output contracted with
a seeded cotangent, gradients of the 256-wide matrices kept as every 16th row.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_import, snerf_oracle as O, snerf_oracle_grad as OG  # noqa: E402
from oracle.make_golden import nuscenes_like_rays, t  # noqa: E402
from oracle.make_golden_grad import ROW_STRIDE, thin  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
F32 = np.float32
N_RAYS, NC, NF = 16, 64, 64


def sd(params, prefix=""):
    return {prefix + k: torch.from_numpy(v.copy()) for k, v in params.items()}


def main():
    ref_render, H = ref_import.load()
    o, d, _ = nuscenes_like_rays(H, N_RAYS, seed=5)
    rb = O.pack_ray_batch(o, d, 1.8, 110.0)
    qfn = ref_import.reference_query_fn(H)
    kw = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4])
    store = dict(ray_batch=rb, Nc=NC, Nf=NF, cot_seed=78, row_stride=ROW_STRIDE, torch_version=torch.__version__,
                 seed_c=80, seed_f=81, seed_alpha=82)
    np.random.seed(0); store["t_rand"] = np.random.rand(N_RAYS, NC).astype(F32)
    np.random.seed(0); store["noise0"] = np.random.rand(N_RAYS, NC).astype(F32)
    np.random.seed(0); store["noise1"] = np.random.rand(N_RAYS, NC + NF).astype(F32)
    np.random.seed(0); store["u"] = np.random.rand(N_RAYS, NF).astype(F32)

    def merged_depths(ret):
        """The fine pass's depths, by the reference's own lines render.py:376-384 on its own coarse outputs."""
        z = ret["z_vals_map"].detach()
        mid = .5 * (z[..., 1:] + z[..., :-1])
        zs = H.sample_pdf(mid, ret["weights"].detach()[..., 1:-1], NF, det=False, pytest=True)
        return torch.sort(torch.cat([z, zs.detach()], -1), -1)[0].numpy()

    def run(case, net_c, net_f, named):
        ret = ref_render.render_rays(t(rb), net_c, qfn, NC, retraw=True, N_importance=NF, network_fine=net_f,
                                     perturb=1.0, raw_noise_std=1.0, pytest=True)
        G = OG.cotangents({k: tuple(v.shape) for k, v in ret.items()}, seed=78)
        loss = OG.loss_from(ret, G)
        loss.backward()
        store[f"{case}_loss"] = float(loss)
        store[f"{case}_mid_z_all"] = merged_depths(ret)
        for tag, net in named:
            for name, p in net.named_parameters():
                if name.startswith("alpha_model."):
                    assert p.grad is None, "the frozen sigma network receives no gradient (no_grad, :198)"
                    continue
                if p.grad is None:          # views_linears of a use_viewdirs=False network: built but never used (:92)
                    assert name.startswith("views_linears.")
                    continue
                store[f"{case}_g{tag}_{name}"] = thin(name, p.grad.numpy())
        for k in ("rgb_map", "rgb0", "acc_map", "depth_map"):
            store[f"{case}_out_{k}"] = ret[k].detach().numpy()
        print(case, "loss", float(loss))

    # ---- novd
    nets = []
    for seed in (80, 81):
        n = H.NeRF(use_viewdirs=False, **kw)
        n.load_state_dict(sd(OG.variant_params(seed, "novd")), strict=False)   # views_linears exists but is unused (:92)
        nets.append(n.train())
    run("novd", nets[0], nets[1], (("c", nets[0]), ("f", nets[1])))

    # ---- coarse and fine networks of different architectures
    for case, (Dc, Wc) in (("d4", (4, 256)), ("w128", (6, 128))):
        pc = O.make_nerf_params(80, D=Dc, W=Wc, trunk_gain=1.5, sigma_bias=0.5)
        pf = O.make_nerf_params(81, trunk_gain=1.5, sigma_bias=0.5)
        nc = ref_import.build_reference_net(H, pc, D=Dc, W=Wc).train()
        nf = ref_import.build_reference_net(H, pf).train()
        run(case, nc, nf, (("c", nc), ("f", nf)))

    # ---- rgb / nocoarse
    p_alpha = O.make_nerf_params(82, trunk_gain=1.5, sigma_bias=0.5)

    def rgb_net(seed):
        alpha = ref_import.build_reference_net(H, p_alpha)
        n = H.NeRF_RGB(use_viewdirs=True, alpha_model=alpha, **kw)
        n.load_state_dict({**sd(OG.variant_params(seed, "rgb")), **sd(p_alpha, "alpha_model.")})
        return n.train()

    nc, nf = rgb_net(80), rgb_net(81)
    run("rgb", nc, nf, (("c", nc), ("f", nf)))
    nf = rgb_net(81)
    for p in nf.alpha_model.parameters():
        p.grad = None
    # network_fn=None: the coarse pass differentiates alpha_model itself (no no_grad there, render.py:361-363)
    ret = ref_render.render_rays(t(rb), None, qfn, NC, retraw=True, N_importance=NF, network_fine=nf, perturb=1.0,
                                 raw_noise_std=1.0, pytest=True)
    G = OG.cotangents({k: tuple(v.shape) for k, v in ret.items()}, seed=78)
    loss = OG.loss_from(ret, G)
    loss.backward()
    store["nocoarse_loss"] = float(loss)
    store["nocoarse_mid_z_all"] = merged_depths(ret)
    for name, p in nf.named_parameters():
        key = "nocoarse_gc_" + name[len("alpha_model."):] if name.startswith("alpha_model.") else "nocoarse_gf_" + name
        store[key] = thin(name, p.grad.numpy())
    for k in ("rgb_map", "rgb0", "acc_map", "depth_map"):
        store[f"nocoarse_out_{k}"] = ret[k].detach().numpy()
    print("nocoarse loss", float(loss))
    path = os.path.join(OUT, "grad_variants.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
