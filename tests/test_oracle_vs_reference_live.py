"""CPU, build container only: the oracles against the UNMODIFIED reference imported live from /root/reference, on seeds the
committed fixtures do not contain (skipped where the reference is not mounted, e.g. on the GPU box).  The fixtures pin the
oracles at a few points; this sweeps them."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from conftest import stepfun_cdf

REF_S = "/root/reference/s-nerf"
REF_Z = "/root/reference/s-nerfpp/zipnerf"
pytestmark = pytest.mark.skipif(not (os.path.isdir(REF_S) and os.path.isdir(REF_Z)), reason="reference not mounted")


@pytest.fixture(scope="module")
def ref_stepfun():
    # s-nerf's `model/` and zipnerf's `internal/` are both plain top-level packages: import zipnerf's by path
    if REF_Z not in sys.path:
        sys.path.insert(0, REF_Z)
    from internal import stepfun
    return stepfun


@pytest.fixture(scope="module")
def ref_losses():
    from oracle import ref_import
    ref_import.load()
    import importlib
    stub = types.ModuleType("model.loss")          # loss_factory.py:1 needs only this name from model/loss.py (SmoothLoss)
    stub.edge_aware_loss_v2 = None
    sys.modules.setdefault("model.loss", stub)
    return importlib.import_module("model.loss_factory")


@pytest.mark.parametrize("seed", range(6))
def test_stepfun_oracle_sweep(seed, ref_stepfun):
    from oracle import stepfun_oracle as SO
    from oracle.make_golden_stepfun import make_rays
    rs = np.random.RandomState(2000 + seed)
    N, S, n = 12, int(rs.choice([24, 48, 64, 100])), int(rs.choice([16, 32, 64, 128]))
    s, w = make_rays(3000 + seed, N, S, peaky=bool(seed % 2))
    dilation, anneal, single = float(rs.uniform(0.001, 0.02)), float(rs.choice([1.0, 0.6])), bool(seed % 3)
    sd, wt = torch.from_numpy(s), torch.from_numpy(w)
    td, wd = ref_stepfun.max_dilate_weights(sd, wt, dilation, domain=(0., 1.), renormalize=True)
    o_td, o_wd = SO.max_dilate_weights(s, w, dilation, (0., 1.), True)
    assert np.array_equal(o_td, td.numpy())
    assert float(np.max(np.abs(o_wd - wd.numpy()))) <= 1e-6 * float(wd.max())
    t2, w2 = td[..., 1:-1], wd[..., 1:-1]
    logits = torch.where(t2[..., 1:] > t2[..., :-1], anneal * torch.log(w2 + 1e-5), torch.full_like(w2, -torch.inf))
    torch.manual_seed(seed)
    jitter = torch.rand((N, 1 if single else n)).numpy()
    torch.manual_seed(seed)
    centers = ref_stepfun.sample(True, t2, logits, n, single, deterministic_center=True).numpy()
    torch.manual_seed(seed)
    out = ref_stepfun.sample_intervals(True, t2, logits, n, single_jitter=single, domain=(0., 1.)).numpy()
    o_out = SO.resample_level(s, w, n, True, dilation, (0., 1.), anneal, 1e-5, jitter, single)
    u_base, mj = SO.uniform_samples(n, True, jitter, single)
    u = (np.broadcast_to(u_base, (N, n)).astype(np.float32) + (jitter * np.float32(mj)).astype(np.float32)).astype(np.float32)
    o_cen = SO.sorted_interp(u, SO.integrate_weights(SO.softmax(logits.numpy())), t2.numpy())
    F_o, F_r = stepfun_cdf(t2.numpy(), logits.numpy(), o_cen), stepfun_cdf(t2.numpy(), logits.numpy(), centers)
    assert float(np.max(np.abs(F_o - F_r))) <= 1e-5
    assert float(np.mean(np.abs(o_out - out) <= 1e-5)) >= 0.98


@pytest.mark.parametrize("seed", range(4))
def test_loss_oracles_grad_sweep(seed, ref_losses):
    from oracle import loss_oracle as LO
    from oracle.make_golden_loss import make_histograms, make_inputs, reference_loss
    rs = np.random.RandomState(4000 + seed)
    N, disparity, with_conf = int(rs.choice([64, 257])), bool(seed % 2), bool(seed // 2)
    lam, cw = float(rs.uniform(0.05, 1.0)), float(rs.uniform(0.1, 0.6))
    arrs = make_inputs(5000 + seed, N, float(rs.uniform(0, 0.8)))
    rgb, tgt, depth, depth0, tdepth, conf = [torch.from_numpy(a) for a in arrs]
    for t in (rgb, depth, depth0, conf):
        t.requires_grad_(True)
    loss, img, dep = reference_loss(ref_losses, rgb, tgt, depth, depth0, tdepth, conf if with_conf else None, disparity, lam, cw)
    loss.backward()
    o_loss, o_img, o_dep, g = LO.rgb_depth_loss(arrs[0], arrs[1], arrs[2], arrs[3], arrs[4], arrs[5] if with_conf else None, lam, cw, disparity)
    assert abs(o_loss - float(loss)) <= 2e-5 * abs(float(loss)) and abs(o_dep - float(dep)) <= 2e-5 * abs(float(dep))
    for key, t in (("rgb", rgb), ("depth", depth), ("depth0", depth0)) + ((("confidence", conf),) if with_conf else ()):
        ref = t.grad.numpy()
        assert float(np.max(np.abs(g[key] - ref))) <= 2e-5 * float(np.max(np.abs(ref))), key
    # ProposalLoss
    Sf, Sc = (64, 64) if seed % 2 else (96, 48)
    sf, wf, sc, wc = make_histograms(6000 + seed, 24, Sf, Sc)
    t = [torch.from_numpy(a) for a in (sf, wf, sc, wc)]
    t[3].requires_grad_(True)
    pl = ref_losses.ProposalLoss(types.SimpleNamespace(proposal_lambda=lam))(*t)
    pl.backward()
    o_pl, o_g = LO.proposal_loss(sf, wf, sc, wc, lam)
    assert abs(o_pl - float(pl)) <= 1e-5 * abs(float(pl))
    assert float(np.max(np.abs(o_g - t[3].grad.numpy()))) <= 2e-5 * float(np.max(np.abs(t[3].grad.numpy())))


@pytest.mark.parametrize("seed,D,W,Nc,Nf", [(11, 4, 64, 32, 0), (12, 8, 128, 48, 64), (13, 8, 256, 64, 128)])
def test_render_oracle_sweep(seed, D, W, Nc, Nf):
    """oracle/snerf_oracle.render_rays against the reference's own render_rays (model/render.py:281-409) on networks /
    rays / sample counts the fixtures do not contain (deterministic eval path: perturb = 0, no noise)."""
    from conftest import err_metric
    from oracle import make_golden, ref_import
    from oracle import snerf_oracle as O
    ref_render, ref_helpers = ref_import.load()
    pc = O.make_nerf_params(seed, D=D, W=W, trunk_gain=1.5, sigma_bias=1.0)
    pf = O.make_nerf_params(seed + 100, D=D, W=W, trunk_gain=1.5, sigma_bias=1.0) if Nf else None
    o, d, _ = make_golden.nuscenes_like_rays(ref_helpers, 40, seed)
    rb = O.pack_ray_batch(o, d, 1.8, 110.0)
    ref, *_ = make_golden.run_reference(ref_render, ref_helpers, rb, pc, pf, Nc, Nf, D, W)
    out = O.render_rays(rb, pc, pf, Nc, Nf, retraw=True)
    assert np.array_equal(out["z_vals_map"], ref["z_vals_map"])
    for k in ("rgb_map", "disp_map", "acc_map", "depth_map", "weights") + (("rgb0", "disp0", "acc0", "z_std") if Nf else ()):
        assert err_metric(out[k], ref[k]) < 1e-4, k


@pytest.mark.parametrize("seed", range(4))
def test_mip_oracle_sweep(seed):
    """oracle/mip_oracle.py (groundwork for SURVEY section 8 row f-2(i)) against the reference's model/mip.py and
    model/math_ops.py on torch-CPU: stratified t_vals, cone casting, integrated positional encoding, the blurred
    sorted-pdf resampler (deterministic and randomized, the draw replayed from the seed) and volumetric_rendering."""
    from oracle import mip_oracle as MO, ref_import
    ref_import.load()
    import importlib
    mip, mops = importlib.import_module("model.mip"), importlib.import_module("model.math_ops")
    rs = np.random.RandomState(7000 + seed)
    N, S = 20, int(rs.choice([32, 64, 128]))
    o = rs.standard_normal((N, 3)).astype(np.float32)
    d = rs.standard_normal((N, 3)).astype(np.float32)
    radii = rs.uniform(5e-4, 2e-3, (N, 1)).astype(np.float32)
    near, far = np.full((N, 1), 1.8, np.float32), np.full((N, 1), 110.0, np.float32)
    lindisp, randomized = bool(seed % 2), bool(seed // 2)
    T = torch.from_numpy
    torch.manual_seed(seed)
    t_rand = torch.rand(N, S + 1).numpy() if randomized else None
    torch.manual_seed(seed)
    t_ref, (m_ref, c_ref) = mip.sample_along_rays(T(o), T(d), T(radii), S, T(near), T(far), randomized, lindisp, "cone")
    t_or = MO.sample_along_rays_t(near, far, S, lindisp, t_rand)
    assert float(np.max(np.abs(t_or - t_ref.numpy()) / np.abs(t_ref.numpy()))) <= 2e-7
    m_or, c_or = MO.cast_rays(t_ref.numpy(), o, d, radii)
    assert float(np.max(np.abs(m_or - m_ref.numpy()))) <= 1e-5 * float(np.abs(m_ref.numpy()).max())
    assert float(np.max(np.abs(c_or - c_ref.numpy()))) <= 1e-4 * float(np.abs(c_ref.numpy()).max())
    enc_ref = mip.integrated_pos_enc((m_ref, c_ref), 0, 12, device="cpu").numpy()
    enc_or = MO.integrated_pos_enc(m_ref.numpy(), c_ref.numpy(), 0, 12)
    # sin of arguments up to 2^11 * |x|: an ulp of the argument is 1e-4 of a period at the top octave
    assert float(np.max(np.abs(enc_or - enc_ref))) <= 2e-3 and float(np.mean(np.abs(enc_or - enc_ref))) <= 2e-5
    # resampling
    w = (rs.rand(N, S).astype(np.float32) ** 4)
    w[3] = 0.0
    u_rand = None
    if randomized:
        torch.manual_seed(100 + seed)
        u_rand = torch.empty(N, S + 1).uniform_(to=1 / (S + 1) - torch.finfo(torch.float32).eps).numpy()
        torch.manual_seed(100 + seed)
    new_ref, _ = mip.resample_along_rays(T(o), T(d), T(radii), t_ref, T(w), randomized, "cone", True, 0.01)
    new_or = MO.resample_t(t_ref.numpy(), w, 0.01, u_rand)
    dnew = np.abs(new_or - new_ref.numpy()) / np.abs(new_ref.numpy())
    assert float(np.mean(dnew <= 1e-5)) >= 0.98 and float(dnew.max()) < 1e-2
    # compositing
    rgb, dens = rs.rand(N, S, 3).astype(np.float32), (rs.rand(N, S, 1).astype(np.float32) ** 3) * 0.5
    ref = mip.volumetric_rendering(T(rgb), T(dens), t_ref, T(d), bool(seed % 2))
    got = MO.volumetric_rendering(rgb, dens, t_ref.numpy(), d, bool(seed % 2))
    for a, b in zip(got, ref):
        b = b.numpy()
        assert float(np.max(np.abs(a - b))) <= 1e-5 * max(float(np.abs(b).max()), 1e-6)


@pytest.mark.parametrize("seed", range(3))
def test_mip_warp_oracle_sweep(seed):
    """The warp-path half of oracle/mip_oracle.py (what configs/nuScenes_depth_6cams runs) against the reference's
    model/mip.py on torch-CPU, function by function: s_vals, sample2enc (log transform, cone cast, contraction,
    Jacobian), integrated_pos_enc(diag=False), pos_enc, warp resampling, real_volumetric_rendering."""
    from oracle import mip_oracle as MO, ref_import
    ref_import.load()
    import importlib
    mip = importlib.import_module("model.mip")
    rs = np.random.RandomState(8000 + seed)
    N, S, NF = 16, int(rs.choice([32, 64, 128])), int(rs.choice([33, 64, 128]))
    o = (rs.standard_normal((N, 3)) * 2).astype(np.float32)
    d = rs.standard_normal((N, 3)).astype(np.float32)
    radii = rs.uniform(5e-4, 2e-3, (N, 1)).astype(np.float32)
    near, far = np.full((N, 1), 1.8, np.float32), np.full((N, 1), 110.0, np.float32)
    randomized, tidx = bool(seed % 2), seed % 3
    T = torch.from_numpy
    torch.manual_seed(seed)
    s_rand = torch.rand(N, S + 1).numpy() if randomized else None
    torch.manual_seed(seed)
    s_ref, (m_ref, c_ref) = mip.warp_sample_along_rays(T(o), T(d), T(radii), S, T(near), T(far), randomized, False, "cone",
                                                       viewc=torch.zeros(3), fn_idx=1, radius=3., transform_idx=tidx)
    s_or = MO.warp_s_vals(N, S, s_rand)
    assert np.array_equal(s_or, s_ref.numpy())
    m_or, c_or = MO.sample2enc(s_or, o, d, radii, near, far, tidx)
    assert float(np.max(np.abs(m_or - m_ref.numpy()))) <= 2e-5 * float(np.abs(m_ref.numpy()).max())
    assert float(np.max(np.abs(c_or - c_ref.numpy()))) <= 2e-4 * float(np.abs(c_ref.numpy()).max())
    enc_ref = mip.integrated_pos_enc((m_ref, c_ref), 0, 16, diag=False, device="cpu").numpy()
    enc_or = MO.integrated_pos_enc_full(m_ref.numpy(), c_ref.numpy(), 0, 16)
    assert float(np.max(np.abs(enc_or - enc_ref))) <= 2e-3 and float(np.mean(np.abs(enc_or - enc_ref))) <= 2e-5
    vd = d / np.linalg.norm(d, axis=-1, keepdims=True)
    assert float(np.max(np.abs(MO.pos_enc(vd, 0, 4) - mip.pos_enc(T(vd), 0, 4).numpy()))) <= 1e-6
    # compositing + resampling
    rgb, dens = rs.rand(N, S, 3).astype(np.float32), (rs.rand(N, S, 1).astype(np.float32) ** 3) * 0.5
    ref = mip.real_volumetric_rendering(T(rgb), T(dens), s_ref, T(d), None, False, T(near), T(far), transform_idx=tidx)
    got = MO.real_volumetric_rendering(rgb, dens, s_or, d, near, far, False, tidx)
    for a, b in zip(got, ref[:4]):
        b = b.numpy()
        assert float(np.max(np.abs(a - b))) <= 2e-5 * max(float(np.abs(b).max()), 1e-6)
    w = ref[3].numpy()
    u_rand = None
    if randomized:
        torch.manual_seed(200 + seed)
        u_rand = torch.empty(N, NF).uniform_(to=1 / NF - torch.finfo(torch.float32).eps).numpy()
        torch.manual_seed(200 + seed)
    new_ref, _ = mip.warp_resample_along_rays(T(o), T(d), T(radii), s_ref, T(w), randomized, NF, "cone", True, 0.01,
                                              viewc=torch.zeros(3), near=T(near), far=T(far), fn_idx=1, radius=3., transform_idx=tidx)
    new_or = MO.warp_resample_s(s_or, w, NF, 0.01, u_rand)
    dnew = np.abs(new_or - new_ref.numpy())
    assert new_or.shape == (N, NF) and float(np.mean(dnew <= 1e-6)) >= 0.98 and float(dnew.max()) < 1e-3


def test_mip_forward_oracle_vs_reference_model():
    """oracle.mip_oracle.mip_forward against the unmodified MipNerfModel.forward (models.py:72-187) on a small network the
    fixtures do not contain (hidden 256, rgb_layer 3, 64 + 96 samples, randomized draws replayed)."""
    from oracle import make_golden_mip as G, mip_oracle as MO
    models = G.load_reference_models()
    import collections
    P = MO.make_mip_params(77, 256, 3)
    model = G.build(models, P, 256, 3, 64, 96)
    o, d, vd, radii, near, far = G.rays_like_nuscenes(24, 5)
    Rays = collections.namedtuple('Rays', ('origins', 'directions', 'viewdirs', 'radii', 'lossmult', 'near', 'far', 'app'))
    T = torch.from_numpy
    torch.manual_seed(9)
    s_rand = torch.rand(24, 65).numpy()
    u_rand = torch.empty(24, 96).uniform_(to=1 / 96 - torch.finfo(torch.float32).eps).numpy()
    torch.manual_seed(9)
    with torch.no_grad():
        ref = model(Rays(T(o), T(d), T(vd), T(radii), torch.ones(24, 1), T(near), T(far), None), True, False, torch.zeros(3))
    got, _ = MO.mip_forward(P, o, d, vd, radii, near, far, 64, 96, True, False, s_rand, u_rand)
    assert np.array_equal(got[0][3], ref[0][3].numpy())
    assert float(np.max(np.abs(got[0][4] - ref[0][4].numpy()))) < 2e-5
    assert float(np.max(np.abs(got[1][0] - ref[1][0].numpy()))) < 2e-4
    assert float(np.max(np.abs(got[1][2] - ref[1][2].numpy()))) < 2e-4
    assert float(np.max(np.abs(got[1][1] - ref[1][1].numpy()) / ref[1][1].numpy())) < 1e-3
    assert float(np.mean(np.abs(got[1][4] - ref[1][4].numpy()) < 1e-5)) > 0.97


@pytest.mark.parametrize("case,seed", [("novd", 31), ("rgb", 32), ("nocoarse", 33), ("d4", 34), ("w128", 35)])
def test_grad_oracle_network_variants_live(case, seed):
    """The differentiable oracle against the reference's own autograd for the network shapes beside the flagship pair
    (use_viewdirs=False; NeRF_RGB + frozen alpha_model; network_fn=None; coarse 4x256 / 6x128 under fine 8x256), on seeds,
    rays and sample counts the committed fixture (grad_variants.npz) does not contain."""
    from oracle import make_golden, ref_import
    from oracle import snerf_oracle as O, snerf_oracle_grad as OG
    ref_render, H = ref_import.load()
    n, Nc, Nf = 6, 32, 32
    o, d, _ = make_golden.nuscenes_like_rays(H, n, seed)
    rb = O.pack_ray_batch(o, d, 1.8, 110.0)
    qfn = ref_import.reference_query_fn(H)
    kw = dict(input_ch=63, input_ch_views=27, output_ch=5, skips=[4])
    sd = lambda p, pre="": {pre + k: torch.from_numpy(v.copy()) for k, v in p.items()}
    pa = O.make_nerf_params(seed + 2, trunk_gain=1.5, sigma_bias=0.5)
    ac = af = None
    if case == "novd":
        pc, pf = OG.variant_params(seed, "novd"), OG.variant_params(seed + 1, "novd")
        nets = []
        for p in (pc, pf):
            m = H.NeRF(D=8, W=256, use_viewdirs=False, **kw)
            m.load_state_dict(sd(p), strict=False)
            nets.append(m.train())
        net_c, net_f = nets
    elif case in ("d4", "w128"):
        Dc, Wc = (4, 256) if case == "d4" else (6, 128)
        pc = O.make_nerf_params(seed, D=Dc, W=Wc, trunk_gain=1.5, sigma_bias=0.5)
        pf = O.make_nerf_params(seed + 1, trunk_gain=1.5, sigma_bias=0.5)
        net_c = ref_import.build_reference_net(H, pc, D=Dc, W=Wc).train()
        net_f = ref_import.build_reference_net(H, pf).train()
    else:
        def rgb_net(s):
            m = H.NeRF_RGB(D=8, W=256, use_viewdirs=True, alpha_model=ref_import.build_reference_net(H, pa), **kw)
            m.load_state_dict({**sd(OG.variant_params(s, "rgb")), **sd(pa, "alpha_model.")})
            return m.train()
        pf, af = OG.variant_params(seed + 1, "rgb"), pa
        net_f = rgb_net(seed + 1)
        if case == "rgb":
            pc, ac, net_c = OG.variant_params(seed, "rgb"), pa, rgb_net(seed)
        else:                                  # network_fn=None: alpha_model itself runs (and is differentiated in) the coarse pass
            pc, net_c = pa, None
    ret = ref_render.render_rays(torch.from_numpy(rb), net_c, qfn, Nc, retraw=True, N_importance=Nf, network_fine=net_f,
                                 perturb=1.0, raw_noise_std=1.0, pytest=True)
    G = OG.cotangents({k: tuple(v.shape) for k, v in ret.items()}, seed=seed)
    OG.loss_from(ret, G).backward()
    # the fine pass's depths by the reference's own lines (render.py:376-384) on its own coarse outputs
    z = ret["z_vals_map"].detach()
    zs = H.sample_pdf(.5 * (z[..., 1:] + z[..., :-1]), ret["weights"].detach()[..., 1:-1], Nf, det=False, pytest=True)
    z_all = torch.sort(torch.cat([z, zs.detach()], -1), -1)[0].numpy()
    draws = {}
    for name, shape in (("t_rand", (n, Nc)), ("noise0", (n, Nc)), ("noise1", (n, Nc + Nf)), ("u", (n, Nf))):
        np.random.seed(0)
        draws[name] = np.random.rand(*shape).astype(np.float32)
    Pc, Pf = OG.params_to_torch(pc), OG.params_to_torch(pf)
    T = lambda p: None if p is None else OG.params_to_torch(p, requires_grad=False)
    out = OG.render_rays(rb, Pc, Pf, Nc, Nf, z_all=z_all, alpha_c=T(ac), alpha_f=T(af), **draws)
    OG.loss_from(out, G).backward()
    ref_c = dict(net_f.alpha_model.named_parameters()) if case == "nocoarse" else dict(net_c.named_parameters())
    ref_f = {k: v for k, v in net_f.named_parameters() if not k.startswith("alpha_model.")}
    checked = 0
    for P, R in ((Pc, ref_c), (Pf, ref_f)):
        for name, p in P.items():
            r = R[name].grad.numpy()
            scale = float(np.max(np.abs(r))) + 1e-30
            assert float(np.max(np.abs(p.grad.numpy() - r))) < 1e-4 * scale, (case, name)
            checked += 1
    assert checked >= 36
