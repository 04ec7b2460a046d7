"""GPU parity tests of the mip-NeRF path (csrc/snerf_mip.cu, snerf_b200.models.MipNerfModel)
against oracle/mip_oracle.py (pinned live against the reference and by tests/golden/mip_*.npz = outputs of the
reference's own MipNerfModel.forward, s-nerf/model/models.py:72-187) -- stage by stage through the C ABI, then end to end.

Bars: fp32 stages (sampling, cone cast, contraction, Jacobian, IPE; compositing, resampling) 1e-5-class; the layer
GEMM exact to fp32 summation order given its bf16 operands; the whole model (bf16 operands and activations, fp32
accumulation) rgb L1 < 2e-3 / max 1.5e-2, weights 2e-2 of max, distance 2 % -- the level of the vanilla bf16 mode."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import mip_oracle as MO

pytestmark = pytest.mark.gpu


def _rays(n, seed):
    rs = np.random.RandomState(seed)
    o = (rs.standard_normal((n, 3)) * np.array([2.0, 0.3, 2.0])).astype(np.float32)
    d = rs.standard_normal((n, 3)).astype(np.float32)
    d[:, 2] -= 1.5
    radii = rs.uniform(5e-4, 2e-3, (n, 1)).astype(np.float32)
    near, far = np.full((n, 1), 1.8, np.float32), np.full((n, 1), 110.0, np.float32)
    return o, d, radii, near, far


@pytest.mark.parametrize("S,randomized,tidx,cone", [(128, False, 0, 1), (64, True, 0, 1), (33, True, 1, 1), (128, False, 2, 0)])
def test_mip_encode_vs_oracle(cuda_device, S, randomized, tidx, cone):
    """snerf_mip_encode: s_vals bit-exact; integrated positional encoding vs the oracle (high octaves take sin of arguments
    up to 2^15 |x|, where an fp32 ulp of the argument is already 1e-3 of a period: mean error is the meaningful bar)."""
    from snerf_b200 import _lib
    dev = cuda_device
    n = 50
    o, d, radii, near, far = _rays(n, S)
    rays9 = torch.from_numpy(np.concatenate([o, d, radii, near, far], 1)).to(dev)
    s_rand = torch.rand(n, S + 1, device=dev) if randomized else None
    m_pad = (n * S + 127) // 128 * 128
    enc = torch.full((m_pad, 128), 7.0, dtype=torch.bfloat16, device=dev)
    enc32 = torch.empty((n * S, 96), device=dev)
    s_out = torch.empty((n, S + 1), device=dev)
    s_lin = torch.linspace(0., 1., S + 1, device=dev)
    e = _lib.MipEncode()
    e.rays, e.n_rays, e.n_samples, e.rows_per_ray = rays9.data_ptr(), n, S, S
    e.s_lin, e.s_rand, e.s_in, e.s_out = s_lin.data_ptr(), (s_rand.data_ptr() if randomized else None), None, s_out.data_ptr()
    e.transform_idx, e.max_deg, e.ray_cone, e.radius = tidx, 16, cone, 3.0
    e.enc, e.enc_f32, e.m_pad = enc.data_ptr(), enc32.data_ptr(), m_pad
    _lib.check(_lib.load().snerf_mip_encode(C.byref(e), _lib.stream_ptr(dev)), "snerf_mip_encode")
    torch.cuda.synchronize()
    s_ref = MO.warp_s_vals(n, S, s_rand.cpu().numpy() if randomized else None)
    assert np.array_equal(s_out.cpu().numpy(), s_ref)
    if cone:
        m, c = MO.sample2enc(s_ref, o, d, radii, near, far, tidx)
    else:
        t = MO.transform(s_ref, near, far, tidx)
        t0, t1 = t[..., :-1], t[..., 1:]
        tm, rv, tv = (t0 + t1) / 2, radii ** 2 / 4, (t1 - t0) ** 2 / 12
        dm = np.maximum(1e-10, np.sum(d ** 2, -1, keepdims=True))
        means = d[:, None, :] * tm[..., None] + o[:, None, :]
        cov = tv[..., None] * (d ** 2)[:, None, :] + rv[..., None] * (1 - d ** 2 / dm)[:, None, :]
        J = MO.jacobi_g(means.astype(np.float32))
        c = np.einsum("...ai,...i,...ib->...ab", J, cov, J).astype(np.float32)
        m = MO.contract(means.astype(np.float32))
    ref = MO.integrated_pos_enc_full(m, c, 0, 16).reshape(n * S, 96)
    got = enc32.cpu().numpy()
    low = np.r_[0:24, 48:72]                   # octaves 0..7: arguments below 2^7 |x|
    assert float(np.max(np.abs(got[:, low] - ref[:, low]))) < 2e-4
    assert float(np.mean(np.abs(got - ref))) < 5e-5 and float(np.max(np.abs(got - ref))) < 5e-2
    # the bf16 operand rows: rounding of the fp32 values, zero padding of columns 96..127 and of the rows past n * S
    eb = enc.float().cpu().numpy()
    assert float(np.max(np.abs(eb[:n * S, :96] - got))) <= 4e-3 + 1e-7
    assert not eb[:n * S, 96:].any() and not eb[n * S:].any()


@pytest.mark.parametrize("M,k0,k1,N,relu,heads,ray_bias", [
    (1000, 128, 0, 256, True, 0, False),       # proposal layer 0
    (777, 1024, 128, 1024, True, 1, False),    # the [x, inputs] skip layer + density head
    (640, 1024, 0, 128, True, 0, True),        # first condition layer with the per-ray bias
    (513, 128, 0, 128, False, 3, False),       # last condition layer + rgb head (no ReLU variant)
])
def test_linear_tc_vs_torch(cuda_device, M, k0, k1, N, relu, heads, ray_bias):
    """snerf_linear_tc against torch.matmul on the same bf16 operands with fp32 accumulation (summation order only)."""
    from snerf_b200 import _lib
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(M)
    m_pad = (M + 127) // 128 * 128
    a0 = (torch.randn(m_pad, k0 + 64, device=dev, generator=g)).to(torch.bfloat16)      # row pitch > k0: strided operand
    a1 = torch.randn(m_pad, k1, device=dev, generator=g).to(torch.bfloat16) if k1 else None
    n_pad = (N + 127) // 128 * 128
    w = (torch.randn(n_pad, k0 + k1, device=dev, generator=g) / np.sqrt(k0 + k1)).to(torch.bfloat16)
    bias = torch.randn(N, device=dev, generator=g)
    rpr = 16
    rb = torch.randn((M + rpr - 1) // rpr, N, device=dev, generator=g) if ray_bias else None
    hw = torch.randn(heads, N, device=dev, generator=g) if heads else None
    out = torch.zeros(m_pad, N + 32, dtype=torch.bfloat16, device=dev)
    hout = torch.zeros(m_pad, max(heads, 1), device=dev)
    L = _lib.Linear()
    L.a0, L.lda0, L.k0 = a0.data_ptr(), a0.stride(0), k0
    L.a1, L.lda1, L.k1 = (a1.data_ptr() if k1 else None), (a1.stride(0) if k1 else 0), k1
    L.w, L.n, L.n_pad, L.bias = w.data_ptr(), N, n_pad, bias.data_ptr()
    L.ray_bias, L.rows_per_ray, L.relu = (rb.data_ptr() if ray_bias else None), rpr, int(relu)
    L.out, L.ldo = out.data_ptr(), out.stride(0)
    L.head_w, L.n_heads, L.head_out = (hw.data_ptr() if heads else None), heads, (hout.data_ptr() if heads else None)
    L.m_rows, L.m_pad = M, m_pad
    _lib.check(_lib.load().snerf_linear_tc(C.byref(L), _lib.stream_ptr(dev)), "snerf_linear_tc")
    torch.cuda.synchronize()
    A = a0[:M, :k0].float() if not k1 else torch.cat([a0[:M, :k0].float(), a1[:M].float()], 1)
    ref = A @ w[:N].float().T + bias
    if ray_bias:
        ref = ref + rb[torch.arange(M, device=dev) // rpr]
    if relu:
        ref = torch.relu(ref)
    got = out[:M, :N].float()
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) < 6e-3 * scale          # bf16 rounding of the stored result
    assert not out[M:].any() and not out[:, N:].any()              # rows past M and columns past N untouched
    if heads:
        href = ref @ hw.T
        assert float((hout[:M, :heads] - href).abs().max()) < 2e-4 * float(href.abs().max()) + 1e-5
        assert not hout[M:].any()


@pytest.mark.parametrize("S,NF,randomized,with_rgb", [(128, 128, False, False), (64, 97, True, False), (127, 0, False, True)])
def test_mip_composite_resample_vs_oracle(cuda_device, S, NF, randomized, with_rgb):
    """snerf_mip_composite: softplus / sigmoid heads + real_volumetric_rendering 2e-5; resampled s_vals within 1e-5 for
    > 98 % (an ulp of the cdf moves a sample across a nearly empty bin) and sorted."""
    from snerf_b200 import _lib
    dev = cuda_device
    n = 37
    rs = np.random.RandomState(S + NF)
    o, d, radii, near, far = _rays(n, 3)
    rays9 = torch.from_numpy(np.concatenate([o, d, radii, near, far], 1)).to(dev)
    s_vals = MO.warp_s_vals(n, S, rs.rand(n, S + 1).astype(np.float32))
    raw_d = (rs.standard_normal((n, S)) * 2 - 2).astype(np.float32)
    raw_d[5] = -40.0                                                     # an empty ray: resampling falls back to the padding
    raw_rgb = rs.standard_normal((n, S, 3)).astype(np.float32) if with_rgb else None
    hb, rgbb = 0.25, [0.1, -0.2, 0.3]
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    c = _lib.MipComposite()
    c.rays, c.n_rays, c.n_samples, c.rows_per_ray = rays9.data_ptr(), n, S, S
    sv, rd = T(s_vals), T(raw_d.reshape(-1))
    rr = T(raw_rgb.reshape(-1, 3)) if with_rgb else None
    c.s_vals, c.raw_density, c.raw_rgb, c.noise = sv.data_ptr(), rd.data_ptr(), (rr.data_ptr() if with_rgb else None), None
    c.density_head_bias, c.density_bias, c.rgb_padding = hb, -1.0, 0.001
    for k in range(3):
        c.rgb_head_bias[k] = rgbb[k]
    c.transform_idx, c.white_bkgd = 0, 0
    comp = torch.empty(n, 3, device=dev)
    dist, acc, wts = torch.empty(n, device=dev), torch.empty(n, device=dev), torch.empty(n, S, device=dev)
    c.comp_rgb = comp.data_ptr() if with_rgb else None
    c.distance, c.acc, c.weights = dist.data_ptr(), acc.data_ptr(), wts.data_ptr()
    eps = np.finfo(np.float32).eps
    u_rand = None
    if NF:
        s_new = torch.empty(n, NF, device=dev)
        c.n_fine, c.resample_padding, c.s_new = NF, 0.01, s_new.data_ptr()
        if randomized:
            u_rand = (rs.rand(n, NF) * (1.0 / NF - eps)).astype(np.float32)
            ur = T(u_rand)
            c.u_rand = ur.data_ptr()
        else:
            ul = torch.linspace(0., 1. - eps, NF, device=dev)
            c.u_lin = ul.data_ptr()
    _lib.check(_lib.load().snerf_mip_composite(C.byref(c), _lib.stream_ptr(dev)), "snerf_mip_composite")
    torch.cuda.synchronize()
    den = MO.softplus((raw_d + np.float32(hb) - 1.0).astype(np.float32))[..., None]
    rgb = None
    if with_rgb:
        r = raw_rgb + np.array(rgbb, np.float32)
        rgb = ((1 / (1 + np.exp(-r))) * 1.002 - 0.001).astype(np.float32)
    rc, rdist, racc, rw = MO.real_volumetric_rendering(rgb, den, s_vals, d, near, far)
    assert float(np.max(np.abs(wts.cpu().numpy() - rw))) < 2e-5
    assert float(np.max(np.abs(acc.cpu().numpy() - racc))) < 2e-5
    assert float(np.max(np.abs(dist.cpu().numpy() - rdist) / rdist)) < 2e-5
    if with_rgb:
        assert float(np.max(np.abs(comp.cpu().numpy() - rc))) < 2e-5
    if NF:
        ref = MO.warp_resample_s(s_vals, rw, NF, 0.01, u_rand)
        got = s_new.cpu().numpy()
        assert np.all(np.diff(got, axis=-1) >= 0) and got.min() >= 0 and got.max() <= 1
        dd = np.abs(got - ref)
        assert float(np.mean(dd <= 1e-5)) >= 0.98 and float(dd.max()) < 2e-3, (float(np.mean(dd <= 1e-5)), float(dd.max()))


def _model_from_golden(g, dev):
    from snerf_b200.models import MipNerfModel
    m = MipNerfModel(no_warp_sample=0, ray_shape="cone", fn=1, max_deg_point=16, radius=3.0, transform_idx=0, real=True,
                     rgb_layer=int(g["rgb_layer"]), hidden_layer=int(g["hidden"]), density_noise=0.0, n_samples=int(g["n_samples"]),
                     proposal_loss=True, N_fine=int(g["n_fine"]))
    P = MO.make_mip_params(int(g["seed"]), int(g["hidden"]), int(g["rgb_layer"]))
    # the reference saves `model_param` from a DataParallel wrapper ('module.' prefix, eval.py:72-74): load it the same way
    m = m.to(dev)
    wrapped = torch.nn.DataParallel(m, device_ids=[dev.index or 0])
    wrapped.load_state_dict({"module." + k: torch.from_numpy(v) for k, v in P.items()})
    return m.eval(), P


@pytest.mark.parametrize("name", ["mip_shipped_det", "mip_small"])
def test_mip_model_vs_reference_golden(cuda_device, name):
    """MipNerfModel.forward (shipped configuration: hidden 1024, rgb_layer 3, 128 + 128 samples) against the outputs of the
    reference's own forward (tests/golden/mip_*.npz).  The proposal level decides where the fine samples go, so the fine
    level is ALSO judged at the kernel's own s_vals against the oracle network (removes the resampling discontinuity)."""
    from snerf_b200.models import Rays
    dev = cuda_device
    g = load_golden(name)
    model, P = _model_from_golden(g, dev)
    T = lambda k: torch.from_numpy(g[k]).to(dev)
    rays = Rays(T("origins"), T("directions"), T("viewdirs"), T("radii"), torch.ones(len(g["near"]), 1, device=dev), T("near"), T("far"), None)
    with torch.no_grad():
        (none0, dist0, acc0, s0, w0), (rgb, dist1, acc1, none1, s1, w1) = model(rays, False, False, torch.zeros(3, device=dev))
    assert none0 is None and none1 is None
    N = g["rgb"].shape[0]
    assert np.array_equal(s0.cpu().numpy(), g["s_vals0"])
    assert float(np.max(np.abs(w0.cpu().numpy() - g["weights0"]))) < 2e-2 * float(g["weights0"].max())
    assert float(np.max(np.abs(acc0.cpu().numpy() - g["acc0"]))) < 1e-2
    assert float(np.mean(np.abs(rgb.cpu().numpy() - g["rgb"]))) < 2e-3
    assert float(np.max(np.abs(rgb.cpu().numpy() - g["rgb"]))) < 1.5e-2
    assert float(np.max(np.abs(acc1.cpu().numpy() - g["acc1"]))) < 1e-2
    assert float(np.median(np.abs(dist1.cpu().numpy() - g["dist1"]) / g["dist1"])) < 5e-3
    assert float(np.mean(np.abs(s1.cpu().numpy() - g["s_vals1"]) < 2e-3)) > 0.95
    # fine level at the kernel's own s_vals
    s1n = s1.cpu().numpy()
    m1, c1 = MO.sample2enc(s1n, g["origins"], g["directions"], g["radii"], g["near"], g["far"])
    e1 = MO.integrated_pos_enc_full(m1, c1, 0, 16)
    S1 = s1n.shape[1] - 1
    cond = np.repeat(MO.pos_enc(g["viewdirs"], 0, 4)[:, None, :], S1, 1).reshape(N * S1, -1)
    raw_rgb, raw_d = MO.mlp_forward(P, e1.reshape(N * S1, -1), cond)
    col = ((1 / (1 + np.exp(-raw_rgb.reshape(N, S1, 3)))) * 1.002 - 0.001).astype(np.float32)
    den = MO.softplus((raw_d.reshape(N, S1, 1) - 1.0).astype(np.float32))
    rc, rdist, racc, rw = MO.real_volumetric_rendering(col, den, s1n, g["directions"], g["near"], g["far"])
    assert float(np.max(np.abs(rgb.cpu().numpy() - rc))) < 8e-3
    assert float(np.max(np.abs(w1.cpu().numpy() - rw))) < 2e-2 * float(rw.max())
    assert float(np.max(np.abs(dist1.cpu().numpy() - rdist) / rdist)) < 2e-2


def test_mip_model_randomized_chunks_and_errors(cuda_device):
    """randomized=True draws (s jitter, u) keep outputs finite and sorted; chunked evaluation equals one pass; unsupported
    configurations are refused loudly; CPU tensors raise (no fallback)."""
    from snerf_b200.models import MipNerfModel, Rays, make_mipnerf
    dev = cuda_device
    g = load_golden("mip_small")
    model, _ = _model_from_golden(g, dev)
    T = lambda k: torch.from_numpy(g[k]).to(dev)
    rays = Rays(T("origins"), T("directions"), T("viewdirs"), T("radii"), None, T("near"), T("far"), None)
    with torch.no_grad():
        full = model(rays, False, False, None)
        model.max_rows = 64 * 20                      # 20 rays per internal chunk
        parts = model(rays, False, False, None)
        model.max_rows = 1 << 21
        rnd = model(rays, True, False, None)
    for a, b in zip(full[1], parts[1]):
        assert (a is None and b is None) or torch.equal(a, b)
    s1 = rnd[1][4]
    assert torch.isfinite(rnd[1][0]).all() and bool((s1[:, 1:] >= s1[:, :-1]).all())
    assert float((rnd[1][0] - full[1][0]).abs().max()) > 1e-4     # the jitter does something
    with pytest.raises(RuntimeError, match="unsupported configuration"):
        MipNerfModel(no_warp_sample=1, fn=1, ray_shape="cone")
    with pytest.raises(RuntimeError, match="CUDA"):
        model(Rays(*[None if x is None else x.cpu() for x in rays]), False, False, None)
    import argparse
    ns = argparse.Namespace(no_warp_sample=0, fn=1, ray_shape="cone", hidden_layer=256, rgb_layer=1, N_samples=64, N_fine=64,
                            proposal_loss=True, density_noise=0.0)
    assert isinstance(make_mipnerf(ns, dev), MipNerfModel)


def test_mip_model_empty_and_single_ray(cuda_device):
    """Edge sizes: zero rays (every output is an empty tensor of the right shape) and one ray (a 128-row tile with 127 / 128
    valid rows) equal the corresponding slice of a larger batch."""
    from snerf_b200.models import Rays
    dev = cuda_device
    g = load_golden("mip_small")
    model, _ = _model_from_golden(g, dev)
    T = lambda k, sl: torch.from_numpy(g[k][sl]).to(dev)
    mk = lambda sl: Rays(T("origins", sl), T("directions", sl), T("viewdirs", sl), T("radii", sl), None, T("near", sl), T("far", sl), None)
    with torch.no_grad():
        full = model(mk(slice(0, 8)), False, False, None)
        one = model(mk(slice(3, 4)), False, False, None)
        none = model(mk(slice(0, 0)), False, False, None)
    assert torch.equal(one[1][0], full[1][0][3:4]) and torch.equal(one[1][4], full[1][4][3:4]) and torch.equal(one[0][4], full[0][4][3:4])
    assert none[1][0].shape == (0, 3) and none[1][1].shape == (0,) and none[0][3].shape == (0, 65) and none[1][5].shape == (0, 63)
