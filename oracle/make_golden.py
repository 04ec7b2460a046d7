"""Generate `tests/golden/*.npz` by running the UNMODIFIED reference in this container.

    python oracle/make_golden.py            # writes tests/golden/

TEST INFRASTRUCTURE ONLY (needs /root/reference; never runs on the GPU box).
The reference has no golden vectors of its own, so these fixtures --
inputs, every `render_rays` output key and the stage intermediates, produced by the
reference's own functions (render.py:281-409, run_nerf_helpers.py:336-424) under
torch CPU fp32 -- are what pins both the numpy oracle and the CUDA path.
Network weights are not stored: they are regenerated from a numpy RandomState seed
(`snerf_oracle.make_nerf_params`), which is bit-stable across boxes.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_import, snerf_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
F32 = np.float32


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def nuscenes_like_rays(ref_helpers, n_rays, seed):
    """CAM_FRONT-like camera of BASELINE config 2 via the reference's own get_rays."""
    H, W, f = 900, 1600, 1266.4
    c2w = torch.eye(4)[:3, :4]
    o, d = ref_helpers.get_rays(H, W, f, c2w, ori_points=[816.3, 491.5])
    rs = np.random.RandomState(seed)
    idx = rs.choice(H * W, n_rays, replace=False)
    return o.reshape(-1, 3)[idx].numpy().astype(F32), d.reshape(-1, 3)[idx].numpy().astype(F32), idx


def run_reference(ref_render, ref_helpers, rb, pc, pf, Nc, Nf, D, W, **kw):
    net_c = ref_import.build_reference_net(ref_helpers, pc, D=D, W=W)
    net_f = ref_import.build_reference_net(ref_helpers, pf, D=D, W=W) if pf is not None else None
    qfn = ref_import.reference_query_fn(ref_helpers)
    with torch.no_grad():
        ret = ref_render.render_rays(t(rb), net_c, qfn, Nc, retraw=True, N_importance=Nf,
                                     network_fine=net_f, **kw)
    return {k: v.numpy() for k, v in ret.items()}, net_c, net_f, qfn


def reference_intermediates(ref_helpers, rb, ret, net_c, qfn, Nf, perturb, raw_noise_std, white_bkgd, pytest):
    """Re-derive the stage values exactly the way render_rays does (render.py:354-391)."""
    o, d, vd = t(rb[:, 0:3]), t(rb[:, 3:6]), t(rb[:, -3:])
    z = t(ret["z_vals_map"])
    with torch.no_grad():
        pts = o[:, None, :] + d[:, None, :] * z[:, :, None]
        raw_c = qfn(pts, vd, net_c)
        _, _, _, w, depth0 = ref_helpers.raw2outputs(raw_c, z, d, raw_noise_std, white_bkgd, pytest=pytest)
        z_mid = .5 * (z[..., 1:] + z[..., :-1])
        zs = ref_helpers.sample_pdf(z_mid, w[..., 1:-1], Nf, det=(perturb == 0.), pytest=pytest)
        z_all, _ = torch.sort(torch.cat([z, zs], -1), -1)
        # cdf / inds, same formulas as run_nerf_helpers.py:338-363
        ww = w[..., 1:-1] + 1e-5
        pdf = ww / torch.sum(ww, -1, keepdim=True)
        cdf = torch.cumsum(pdf, -1)
        cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
        if pytest and perturb != 0.:
            np.random.seed(0)
            u = torch.Tensor(np.random.rand(*(list(cdf.shape[:-1]) + [Nf])))
        else:
            u = torch.linspace(0., 1., steps=Nf).expand(list(cdf.shape[:-1]) + [Nf])
        inds = torch.searchsorted(cdf, u.contiguous(), right=True)
    assert np.array_equal(w.numpy(), ret["weights"])
    return dict(raw_coarse=raw_c.numpy(), depth0=depth0.numpy(), cdf=cdf.numpy(), u=u.numpy().astype(F32),
                inds=inds.numpy().astype(np.int64), z_samples=zs.numpy(), z_all=z_all.numpy())


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_render, ref_helpers = ref_import.load()
    torch.manual_seed(0)
    meta = dict(torch_version=torch.__version__, numpy_version=np.__version__)

    # ---- config 1: 1000 rays, 32 coarse, 4x64 MLP, no fine pass (BASELINE.json configs[0])
    rs = np.random.RandomState(1)
    o = rs.standard_normal((1000, 3)).astype(F32)
    d = rs.standard_normal((1000, 3)).astype(F32)
    rb = O.pack_ray_batch(o, d, 2.0, 6.0)
    pc = O.make_nerf_params(100, D=4, W=64)
    ret, *_ = run_reference(ref_render, ref_helpers, rb, pc, None, 32, 0, 4, 64)
    np.savez_compressed(os.path.join(OUT, "cfg1_plumbing.npz"), ray_batch=rb, seed_coarse=100, D=4, W=64,
                        Nc=32, Nf=0, **{"out_" + k: v for k, v in ret.items()}, **meta)

    # ---- config 2 slices: nuScenes-like camera, 8x256, 64c + 128f
    o, d, idx = nuscenes_like_rays(ref_helpers, 128, seed=2)
    rb = O.pack_ray_batch(o, d, 1.8, 110.0)
    variants = {
        "cfg2_default": dict(seeds=(10, 11), gain=1.0, sb=0.0, kw={}),
        "cfg2_peaky": dict(seeds=(20, 21), gain=1.5, sb=1.0, kw={}),
        "cfg2_lindisp_white": dict(seeds=(30, 31), gain=1.5, sb=0.5, kw=dict(lindisp=True, white_bkgd=True)),
        "cfg2_stochastic": dict(seeds=(40, 41), gain=1.5, sb=0.5,
                                kw=dict(perturb=1.0, raw_noise_std=1.0, pytest=True)),
    }
    for name, v in variants.items():
        pc = O.make_nerf_params(v["seeds"][0], trunk_gain=v["gain"], sigma_bias=v["sb"])
        pf = O.make_nerf_params(v["seeds"][1], trunk_gain=v["gain"], sigma_bias=v["sb"])
        kw = v["kw"]
        ret, net_c, net_f, qfn = run_reference(ref_render, ref_helpers, rb, pc, pf, 64, 128, 8, 256, **kw)
        inter = reference_intermediates(ref_helpers, rb, ret, net_c, qfn, 128, kw.get("perturb", 0.),
                                        kw.get("raw_noise_std", 0.), kw.get("white_bkgd", False),
                                        kw.get("pytest", False))
        extra = {}
        if kw.get("pytest"):
            # the draws the reference makes under pytest=True (render.py:346-350,
            # run_nerf_helpers.py:350-359,406-410): np.random.seed(0) before each
            np.random.seed(0); extra["t_rand"] = np.random.rand(rb.shape[0], 64).astype(F32)
            np.random.seed(0); extra["noise0"] = (np.random.rand(rb.shape[0], 64) * kw["raw_noise_std"]).astype(F32)
            np.random.seed(0); extra["noise1"] = (np.random.rand(rb.shape[0], 192) * kw["raw_noise_std"]).astype(F32)
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"), ray_batch=rb, pixel_index=idx,
            seed_coarse=v["seeds"][0], seed_fine=v["seeds"][1], trunk_gain=v["gain"], sigma_bias=v["sb"],
            D=8, W=256, Nc=64, Nf=128, lindisp=bool(kw.get("lindisp", False)),
            white_bkgd=bool(kw.get("white_bkgd", False)), perturb=float(kw.get("perturb", 0.)),
            raw_noise_std=float(kw.get("raw_noise_std", 0.)),
            **{"out_" + k: val for k, val in ret.items()}, **{"mid_" + k: val for k, val in inter.items()},
            **extra, **meta)
        print(name, {k: tuple(val.shape) for k, val in ret.items()})

    # ---- NeRF_RGB + frozen alpha_model, network_fn=None (render.py:361-371, run_nerf_helpers.py:157-212)
    rb64 = rb[:64]
    p_alpha = O.make_nerf_params(60, trunk_gain=1.5, sigma_bias=1.0)
    p_rgb = {k: v for k, v in O.make_nerf_params(61, trunk_gain=1.5).items() if not k.startswith("alpha_linear")}
    alpha_net = ref_import.build_reference_net(ref_helpers, p_alpha)
    rgb_net = ref_helpers.NeRF_RGB(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4],
                                   use_viewdirs=True, alpha_model=alpha_net)
    rgb_net.load_state_dict({**{k: torch.from_numpy(v.copy()) for k, v in p_rgb.items()},
                             **{"alpha_model." + k: torch.from_numpy(v.copy()) for k, v in p_alpha.items()}})
    qfn = ref_import.reference_query_fn(ref_helpers)
    with torch.no_grad():
        ret = ref_render.render_rays(t(rb64), None, qfn, 64, retraw=True, N_importance=128, network_fine=rgb_net.eval())
    np.savez_compressed(os.path.join(OUT, "cfg2_rgb_alpha.npz"), ray_batch=rb64, seed_alpha=60, seed_rgb=61,
                        trunk_gain=1.5, sigma_bias=1.0, D=8, W=256, Nc=64, Nf=128,
                        **{"out_" + k: v.numpy() for k, v in ret.items()}, **meta)

    # ---- stage fixture: sample_pdf edge cases straight from the reference function
    rs = np.random.RandomState(7)
    B = 63
    bins = np.sort(rs.uniform(2, 100, size=(16, B)).astype(F32), -1)
    w = rs.uniform(0, 1, size=(16, B - 1)).astype(F32)
    w[0] = 0.0                      # all-zero weights -> uniform pdf
    w[1] = 0.0; w[1, 17] = 1.0      # single spike -> many denom < 1e-5
    w[2, :31] = 0.0                 # leading empty bins
    w[3, 31:] = 0.0                 # trailing empty bins
    w[4] = 1e-7                     # tiny but equal
    w[5] = rs.uniform(0, 1e-6, B - 1)
    with torch.no_grad():
        s_det = ref_helpers.sample_pdf(t(bins), t(w), 128, det=True).numpy()
        np.random.seed(0)
        s_rnd = ref_helpers.sample_pdf(t(bins), t(w), 128, det=False, pytest=True).numpy()
    np.random.seed(0)
    u_rnd = np.random.rand(16, 128).astype(F32)
    np.savez_compressed(os.path.join(OUT, "stage_sample_pdf.npz"), bins=bins, weights=w,
                        samples_det=s_det, samples_rand=s_rnd, u_rand=u_rnd, **meta)

    # ---- stage fixture: raw2outputs edge cases (zero density -> acc==0 -> NaN disp, saturated sigma)
    raw = rs.standard_normal((16, 64, 4)).astype(F32) * 2
    raw[0, :, 3] = -1.0             # relu -> 0 everywhere: acc = 0, disp = nan (0/0)
    raw[1, :, 3] = 50.0             # opaque at first sample
    raw[2, :32, 3] = -5.0
    z = np.sort(rs.uniform(1.8, 110, size=(16, 64)).astype(F32), -1)
    dd = rs.standard_normal((16, 3)).astype(F32)
    with torch.no_grad():
        outs = ref_helpers.raw2outputs(t(raw), t(z), t(dd), 0, False)
        outs_w = ref_helpers.raw2outputs(t(raw), t(z), t(dd), 0, True)
    names = ["rgb_map", "disp_map", "acc_map", "weights", "depth_map"]
    np.savez_compressed(os.path.join(OUT, "stage_raw2outputs.npz"), raw=raw, z=z, rays_d=dd,
                        **{n: o_.numpy() for n, o_ in zip(names, outs)},
                        **{n + "_white": o_.numpy() for n, o_ in zip(names, outs_w)}, **meta)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
