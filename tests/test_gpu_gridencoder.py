"""GPU (B200): the hash-grid encoder (csrc/snerf_grid.cu) through snerf_b200.gridencoder -> C ABI, against
  * the reference's OWN kernels run live (oracle/_ref/_gridencoder_ref.so = gridencoder.cu compiled unmodified),
  * the committed golden fixtures those kernels produced (tests/golden/grid_*.npz, oracle/make_golden_grid.py),
  * the numpy oracle (oracle/gridencoder_oracle.py).
Bars: fp32 forward / dy_dx / grad_inputs bit-exact (same arithmetic, no atomics); table gradients and the TV increment
are sums of atomics in undefined order -> 1e-5 of the tensor's max; fp16 within fp16 rounding of the reference."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden
from oracle import build_ref_gridencoder as R
from oracle import gridencoder_oracle as G
from oracle.make_golden_grid import CASES, make_embeddings, make_inputs, run_reference

pytestmark = pytest.mark.gpu
NAMES = list(CASES)
GT = {"hash": 0, "tiled": 1}
IT = {"linear": 0, "smoothstep": 1}


@pytest.fixture(scope="module")
def ref_mod():
    m = R.load()
    if m is None:
        pytest.skip("oracle/_ref/_gridencoder_ref.so not built (needs /root/reference at build time)")
    return m


def case(name):
    """Seeded inputs of a fixture case (identical to what oracle/make_golden_grid.py fed the reference)."""
    i = NAMES.index(name)
    cfg, gridtype, interp, B = CASES[name]
    offsets, sizes, pls = G.level_layout(**cfg)
    D, Cd, H = cfg["input_dim"], cfg["level_dim"], cfg["base_resolution"]
    L = len(offsets) - 1
    emb = make_embeddings(100 + i, int(offsets[-1]), Cd)
    x = make_inputs(200 + i, B, D)
    grad = np.random.RandomState(300 + i).standard_normal((B, L * Cd)).astype(np.float32)
    return dict(cfg=cfg, gridtype=GT[gridtype], interp=IT[interp], B=B, D=D, C=Cd, H=H, L=L, pls=pls, S=float(np.log2(pls)),
                offsets=offsets, emb=emb, x=x, grad=grad, align=bool(cfg.get("align_corners", False)))


def ours(c, dev, dtype=torch.float32, want_input_grad=True, tv_weight=None):
    """forward + backward (+ TV) through the public autograd function; returns numpy arrays."""
    from snerf_b200.gridencoder import grid_encode
    x = torch.from_numpy(c["x"]).to(dev).requires_grad_(want_input_grad)
    emb = torch.from_numpy(c["emb"]).to(dev).to(dtype).requires_grad_(True)
    off = torch.from_numpy(c["offsets"]).to(dev)
    with torch.enable_grad():
        out = grid_encode(x, emb, off, c["pls"], c["H"], want_input_grad, c["gridtype"], c["align"], c["interp"])
        out.backward(torch.from_numpy(c["grad"]).to(dev).to(dtype))
    torch.cuda.synchronize()
    return (out.detach().float().cpu().numpy(), emb.grad.float().cpu().numpy(),
            x.grad.cpu().numpy() if want_input_grad else None)


def ours_dy_dx(c, dev):
    """dy_dx is internal to the autograd node: fetch it through the C ABI."""
    from snerf_b200 import _lib
    x = torch.from_numpy(c["x"]).to(dev)
    emb = torch.from_numpy(c["emb"]).to(dev)
    off = torch.from_numpy(c["offsets"]).to(dev)
    out = torch.empty(c["B"], c["L"] * c["C"], device=dev)
    dy = torch.empty(c["B"], c["L"] * c["D"] * c["C"], device=dev)
    d = _lib.GridDesc(c["D"], c["C"], c["L"], c["H"], c["gridtype"], int(c["align"]), c["interp"], 0, c["S"])
    _lib.check(_lib.load().snerf_grid_encode_fwd(C.byref(d), _lib.ptr(x), _lib.ptr(emb), _lib.ptr(off), _lib.ptr(out),
                                                 c["C"], c["L"] * c["C"], _lib.ptr(dy), c["B"], _lib.stream_ptr(dev)), "fwd")
    torch.cuda.synchronize()
    return out.cpu().numpy(), dy.cpu().numpy()


def reference_live(ref_mod, c, dev, dtype=torch.float32, tv_weight=1e-2):
    t = lambda a: torch.from_numpy(a).to(dev)
    out, dy, ge, gi, tv = run_reference(ref_mod, t(c["x"]), t(c["emb"]).to(dtype), t(c["offsets"]), c["S"], c["H"], c["gridtype"],
                                        c["align"], c["interp"], t(c["grad"]).to(dtype), tv_weight)
    f = lambda a: a.float().cpu().numpy()
    return f(out), f(dy), f(ge), f(gi), f(tv)


def close_to_max(a, b, tol):
    return float(np.max(np.abs(a.astype(np.float64) - b))) <= tol * (float(np.max(np.abs(b))) + 1e-30)


# ------------------------------------------------------------------ vs the reference's own kernels, live
@pytest.mark.parametrize("name", NAMES)
def test_grid_fwd_bwd_grad_matches_reference_kernels(name, cuda_device, ref_mod):
    c = case(name)
    out, ge, gi = ours(c, cuda_device)
    _, dy = ours_dy_dx(c, cuda_device)
    r_out, r_dy, r_ge, r_gi, _ = reference_live(ref_mod, c, cuda_device)
    assert out.shape == (c["B"], c["L"] * c["C"])
    assert np.array_equal(out, r_out), float(np.max(np.abs(out - r_out)))            # bit-exact
    assert np.array_equal(dy, r_dy), float(np.max(np.abs(dy - r_dy)))                # bit-exact
    assert np.array_equal(gi, r_gi), float(np.max(np.abs(gi - r_gi)))                # bit-exact (sequential sums)
    assert np.array_equal(ge != 0, r_ge != 0)                                        # same touched cells
    assert close_to_max(ge, r_ge, 1e-5)                                              # atomics: order-dependent rounding


@pytest.mark.parametrize("name", NAMES)
def test_grid_tv_matches_reference_kernels(name, cuda_device, ref_mod):
    from snerf_b200.gridencoder import GridEncoder
    c = case(name)
    cfg = dict(c["cfg"])
    enc = GridEncoder(gridtype="hash" if c["gridtype"] == 0 else "tiled", **cfg).to(cuda_device)
    assert np.array_equal(enc.offsets.cpu().numpy(), c["offsets"])
    enc.embeddings.data.copy_(torch.from_numpy(c["emb"]))
    enc.embeddings.grad = torch.zeros_like(enc.embeddings)
    enc.grad_total_variation(weight=1e-2, inputs=torch.from_numpy(c["x"]).to(cuda_device) * 2 - 1, bound=1)
    torch.cuda.synchronize()
    tv = enc.embeddings.grad.cpu().numpy()
    # the module maps [-1,1] -> [0,1] ((x*2-1+1)/2 is exact for these inputs except the deliberately odd ones): compare on
    # the reference fed with the SAME mapped inputs
    c2 = dict(c, x=((torch.from_numpy(c["x"]) * 2 - 1 + 1) / 2).numpy())
    *_, r_tv = reference_live(ref_mod, c2, cuda_device)
    assert np.array_equal(tv != 0, r_tv != 0)
    assert close_to_max(tv, r_tv, 1e-5)


@pytest.mark.parametrize("name", ["grid_zip_main", "grid_small_hash_smooth", "grid_2d_tiled_align"])
def test_grid_fp16_grad_matches_reference_kernels(name, cuda_device, ref_mod):
    c = case(name)
    out, ge, gi = ours(c, cuda_device, dtype=torch.float16)
    r_out, _, r_ge, r_gi, _ = reference_live(ref_mod, c, cuda_device, dtype=torch.float16)
    assert close_to_max(out, r_out, 2e-3)
    assert close_to_max(gi, r_gi, 4e-3)
    assert close_to_max(ge, r_ge, 1e-2)     # fp16 atomics: order-dependent at fp16 precision
    frac_exact = float(np.mean(out == r_out))
    assert frac_exact > 0.98, frac_exact    # same rounding sequence: (almost) every element identical


# ------------------------------------------------------------------ vs the committed fixtures (reference outputs)
@pytest.mark.parametrize("name", NAMES)
def test_grid_grad_matches_golden(name, cuda_device):
    g = load_golden(name)
    c = case(name)
    assert np.array_equal(c["x"], g["inputs"]) and int(g["seed_emb"]) == 100 + NAMES.index(name)
    out, ge, gi = ours(c, cuda_device)
    _, dy = ours_dy_dx(c, cuda_device)
    assert np.array_equal(out, g["out"])
    assert np.array_equal(dy, g["dy_dx"])
    assert np.array_equal(gi, g["grad_inputs"])
    rows = g["ge_rows"]
    assert np.array_equal(np.nonzero(np.any(ge != 0, axis=1))[0], rows)
    assert close_to_max(ge[rows], g["ge_vals"], 1e-5)


# ------------------------------------------------------------------ vs the numpy oracle
@pytest.mark.parametrize("name", NAMES)
def test_grid_grad_matches_oracle(name, cuda_device):
    c = case(name)
    out, ge, gi = ours(c, cuda_device)
    o_out, o_dy = G.grid_encode_forward(c["x"], c["emb"], c["offsets"], c["S"], c["H"], c["gridtype"], c["align"], c["interp"],
                                        calc_dy_dx=True)
    L, B, Cd = o_out.shape
    # dyadic per-level scales: the oracle reproduces the kernel's rounding; otherwise the device exp2f (2-ulp approximate)
    # can differ from numpy's by an ulp in `scale` (tests/test_oracle_golden.py::test_grid_oracle_matches_reference_kernels)
    dyadic = float(c["S"]).is_integer()
    assert close_to_max(out, o_out.transpose(1, 0, 2).reshape(B, L * Cd), 1e-6 if dyadic else 2e-5)
    g_lbc = c["grad"].reshape(B, L, Cd).transpose(1, 0, 2)
    o_ge, o_gi = G.grid_encode_backward(g_lbc, c["x"], c["emb"].shape, c["offsets"], c["S"], c["H"], c["gridtype"], c["align"],
                                        c["interp"], dy_dx=o_dy)
    assert close_to_max(ge, o_ge, 1e-5 if dyadic else 4e-5)
    assert close_to_max(gi, o_gi, 1e-5 if dyadic else 4e-5)


# ------------------------------------------------------------------ module interface (grid.py:96-200)
def test_grid_module_grad_interface(cuda_device):
    from snerf_b200.gridencoder import GridEncoder
    torch.manual_seed(0)
    enc = GridEncoder(input_dim=3, num_levels=10, level_dim=4, base_resolution=16, desired_resolution=8192,
                      log2_hashmap_size=21).to(cuda_device)
    assert set(enc.state_dict()) == {"embeddings", "offsets", "idx", "grid_sizes"}
    assert enc.output_dim == 40 and abs(enc.per_level_scale - 2.0) < 1e-12
    assert tuple(enc.embeddings.shape) == (int(enc.offsets[-1]), 4)
    assert float(enc.embeddings.abs().max()) <= 1e-4          # init_std uniform init (grid.py:143-145)
    enc.embeddings.data.uniform_(-1, 1)
    x = (torch.rand(7, 33, 3, device=cuda_device) * 2 - 1).requires_grad_(True)       # arbitrary prefix shape, [-1,1]
    with torch.enable_grad():
        y = enc(x, bound=1)
        assert y.shape == (7, 33, 40) and y.dtype == torch.float32
        (y * y).sum().backward()
    assert x.grad is not None and x.grad.shape == x.shape and float(x.grad.abs().max()) > 0
    assert enc.embeddings.grad.shape == enc.embeddings.shape
    orc = G.GridEncoderOracle(enc.embeddings.detach().cpu().numpy(), input_dim=3, num_levels=10, level_dim=4,
                              base_resolution=16, desired_resolution=8192, log2_hashmap_size=21)
    assert close_to_max(y.detach().cpu().numpy(), orc(x.detach().cpu().numpy()), 1e-6)
    # no input gradient requested -> no dy_dx, inputs.grad stays None
    x2 = x.detach()
    with torch.enable_grad():
        enc(x2).sum().backward()
    assert x2.grad is None
    # autocast: half embeddings when level_dim is even (grid.py:41-42)
    with torch.autocast("cuda", dtype=torch.float16):
        assert enc(x2).dtype == torch.float16


def test_grid_errors_and_ragged_sizes(cuda_device):
    from snerf_b200 import _lib
    from snerf_b200.gridencoder import GridEncoder, grid_encode
    enc = GridEncoder(input_dim=3, num_levels=4, level_dim=2, base_resolution=4, log2_hashmap_size=10)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc(torch.zeros(5, 3))
    enc = enc.to(cuda_device)
    enc.embeddings.data.uniform_(-1, 1)
    assert enc(torch.zeros(0, 3, device=cuda_device)).shape == (0, 8)          # empty batch
    full = enc(torch.linspace(-1, 1, 3 * 1025, device=cuda_device).view(1025, 3))
    for n in (1, 255, 257, 1025):                                                # ragged tails = prefix of the full batch
        part = enc(torch.linspace(-1, 1, 3 * 1025, device=cuda_device).view(1025, 3)[:n])
        assert torch.equal(part, full[:n])
    bad = torch.zeros(int(enc.offsets[-1]), 3, device=cuda_device)                # level_dim = 3 is not supported (as in the reference)
    with pytest.raises(RuntimeError, match="level_dim"):
        grid_encode(torch.zeros(4, 3, device=cuda_device), bad, enc.offsets, 2.0, 4)
    d = _lib.GridDesc(3, 2, 4, 4, 0, 0, 0, 0, 1.0)
    assert _lib.load().snerf_grid_encode_fwd(C.byref(d), None, None, None, None, 2, 8, None, 4, None) != 0


# ------------------------------------------------------------------ full size (BASELINE configs[3] shapes): properties
def test_grid_full_size_properties(cuda_device, ref_mod):
    """2^20 points through the zip-NeRF main grid (10 levels x 4 features, 2^21-entry tables, 150 MB): bit-exact against
    the reference kernels at full size, linear in the table, [L,B,C]-strided output == permuted [B,L*C] output, and
    out-of-range points give zeros."""
    from snerf_b200 import _lib
    from snerf_b200.gridencoder import grid_encode
    dev = cuda_device
    cfg = CASES["grid_zip_main"][0]
    offsets, _, pls = G.level_layout(**cfg)
    L, Cd, B = 10, 4, 1 << 20
    gen = torch.Generator(device=dev).manual_seed(5)
    e1 = torch.rand(int(offsets[-1]), Cd, device=dev, generator=gen) * 2 - 1
    e2 = torch.rand(int(offsets[-1]), Cd, device=dev, generator=gen) * 2 - 1
    x = torch.rand(B, 3, device=dev, generator=gen)
    x[::1000, 1] = 1.5                                             # out of range
    off = torch.from_numpy(offsets).to(dev)
    y1 = grid_encode(x, e1, off, pls, 16)
    y2 = grid_encode(x, e2, off, pls, 16)
    y12 = grid_encode(x, e1 + 0.5 * e2, off, pls, 16)
    assert float((y12 - (y1 + 0.5 * y2)).abs().max()) < 2e-6       # linearity in the table
    assert float(y1[::1000].abs().max()) == 0.0 and float(y1[1::1000].abs().max()) > 0
    # reference kernels, same inputs
    r_out = torch.empty(L, B, Cd, device=dev)
    ref_mod.grid_encode_forward(x, e1, off, r_out, B, 3, Cd, L, float(np.log2(pls)), 16, None, 0, False, 0)
    torch.cuda.synchronize()
    assert torch.equal(y1, r_out.permute(1, 0, 2).reshape(B, L * Cd))
    # the reference's [L, B, C] layout through the strided C ABI
    out_lbc = torch.empty(L, B, Cd, device=dev)
    d = _lib.GridDesc(3, Cd, L, 16, 0, 0, 0, 0, float(np.log2(pls)))
    _lib.check(_lib.load().snerf_grid_encode_fwd(C.byref(d), _lib.ptr(x), _lib.ptr(e1), _lib.ptr(off), _lib.ptr(out_lbc),
                                                 B * Cd, Cd, None, B, _lib.stream_ptr(dev)), "fwd [L,B,C]")
    torch.cuda.synchronize()
    assert torch.equal(out_lbc, r_out)
    # gradient w.r.t. the table: adjoint identity <dY, enc(E)> == <grad_E, E> (the encoder is linear in E)
    e = e1.clone().requires_grad_(True)
    dy = torch.randn(B, L * Cd, device=dev, generator=gen)
    with torch.enable_grad():
        grid_encode(x, e, off, pls, 16).backward(dy)
    lhs = float((dy.double() * y1.double()).sum())
    rhs = float((e.grad.double() * e1.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), 1.0), (lhs, rhs)


# ------------------------------------------------------------------ fused multisample featurisation (models.py:481-507)
from oracle.make_golden_grid import MS_CASES, make_multisamples, run_reference_multisample  # noqa: E402

MS_NAMES = list(MS_CASES)


def ms_case(name):
    i = MS_NAMES.index(name)
    cfg, N, M = MS_CASES[name]
    offsets, sizes, pls = G.level_layout(**cfg)
    Cd = cfg["level_dim"]
    L = len(offsets) - 1
    means, stds = make_multisamples(500 + i, N, M)
    return dict(cfg=cfg, N=N, M=M, C=Cd, L=L, H=cfg["base_resolution"], S=float(np.log2(pls)), offsets=offsets, sizes=sizes,
                emb=make_embeddings(400 + i, int(offsets[-1]), Cd), means=means, stds=stds,
                grad=np.random.RandomState(600 + i).standard_normal((N, L * (Cd + 1))).astype(np.float32))


def ours_ms(c, dev, scale_featurization=True):
    from snerf_b200.gridencoder import GridEncoder
    enc = GridEncoder(**c["cfg"]).to(dev)
    enc.embeddings.data.copy_(torch.from_numpy(c["emb"]))
    with torch.enable_grad():
        out = enc.encode_multisample(torch.from_numpy(c["means"]).to(dev), torch.from_numpy(c["stds"]).to(dev), bound=1,
                                     scale_featurization=scale_featurization)
        out.backward(torch.from_numpy(c["grad"][:, :out.shape[1]]).to(dev))
    torch.cuda.synchronize()
    return out.detach().cpu().numpy(), enc.embeddings.grad.cpu().numpy(), enc


@pytest.mark.parametrize("name", MS_NAMES)
def test_grid_multisample_grad_matches_reference_pipeline(name, cuda_device, ref_mod):
    """ONE kernel vs the reference's composition (its encoder kernels + permute + erf / product / mean / cat in torch +
    autograd + its backward kernel), live on the same inputs."""
    c = ms_case(name)
    out, ge, enc = ours_ms(c, cuda_device)
    t = lambda a: torch.from_numpy(a).to(cuda_device)
    with torch.enable_grad():
        r_out, r_ge = run_reference_multisample(ref_mod, t(c["means"]), t(c["stds"]), t(c["emb"]), t(c["offsets"]), t(c["sizes"]),
                                                c["S"], c["H"], 1e-4, t(c["grad"]))
    r_out, r_ge = r_out.cpu().numpy(), r_ge.cpu().numpy()
    LC = c["L"] * c["C"]
    assert out.shape == (c["N"], LC + c["L"])
    assert close_to_max(out[:, :LC], r_out[:, :LC], 1e-6)            # mean over M: summation order differs from torch's
    assert close_to_max(out[:, LC:], r_out[:, LC:], 1e-5)            # carries a 2M-term fp32 mean (torch) vs double partial sums (ours)
    assert np.array_equal(ge != 0, r_ge != 0)
    assert close_to_max(ge, r_ge, 1e-5)
    # the per-level gain alone (segment mean of |embedding|^2)
    gain = enc.level_gain().cpu().numpy()
    o = G.GridEncoderOracle(c["emb"], **c["cfg"])
    assert np.allclose(gain, o.level_gain(1e-4), rtol=1e-6, atol=0)
    # without scale_featurization: exactly the first L*C columns
    out2, ge2, _ = ours_ms(c, cuda_device, scale_featurization=False)
    assert out2.shape == (c["N"], LC) and np.array_equal(out2, out[:, :LC])


@pytest.mark.parametrize("name", MS_NAMES)
def test_grid_multisample_grad_matches_golden_and_oracle(name, cuda_device):
    c = ms_case(name)
    g = load_golden(name)
    assert np.array_equal(c["means"], g["means"]) and np.array_equal(c["stds"], g["stds"])
    out, ge, _ = ours_ms(c, cuda_device)
    LC = c["L"] * c["C"]
    assert close_to_max(out[:, :LC], g["out"][:, :LC], 1e-6)
    assert close_to_max(out[:, LC:], g["out"][:, LC:], 5e-5)         # fixture's level mean was an fp32 atomic reduction
    rows = g["ge_rows"]
    assert np.array_equal(np.nonzero(np.any(ge != 0, axis=1))[0], rows)
    assert close_to_max(ge[rows], g["ge_vals"], 1e-5)
    o = G.GridEncoderOracle(c["emb"], **c["cfg"])
    assert close_to_max(out, o.encode_multisample(c["means"], c["stds"]), 2e-6)
    assert close_to_max(ge, o.encode_multisample_backward(c["grad"], c["means"], c["stds"]), 1e-5)


def test_grid_multisample_full_size_grad_properties(cuda_device):
    """One zip-NeRF render chunk worth of samples (2^19 samples x 6 multisamples, main grid): equals the unfused
    composition of our own encoder + torch ops, is linear in the table, and satisfies the adjoint identity."""
    from snerf_b200.gridencoder import GridEncoder
    dev = cuda_device
    enc = GridEncoder(**CASES["grid_zip_main"][0]).to(dev)
    gen = torch.Generator(device=dev).manual_seed(11)
    enc.embeddings.data.copy_(torch.rand(enc.embeddings.shape, device=dev, generator=gen) * 2 - 1)
    N, M, L, Cd = 1 << 19, 6, 10, 4
    centre = torch.rand(N, 1, 3, device=dev, generator=gen) * 2 - 1
    stds = torch.exp(torch.rand(N, M, device=dev, generator=gen) * 9 - 10)
    means = centre + stds[..., None] * torch.randn(N, M, 3, device=dev, generator=gen)
    with torch.no_grad():
        fused = enc.encode_multisample(means, stds)
        feats = enc(means, bound=1).unflatten(-1, (L, -1))
        w = torch.erf(1 / torch.sqrt(8 * stds[..., None] ** 2 * enc.grid_sizes ** 2))
        unfused = (feats * w[..., None]).mean(dim=-3).flatten(-2, -1)
        assert fused.shape == (N, L * Cd + L)
        assert float((fused[:, :L * Cd] - unfused).abs().max()) <= 1e-6 * float(unfused.abs().max())
        fw = (2 * w.mean(dim=-2) - 1) * enc.level_gain()
        assert float((fused[:, L * Cd:] - fw).abs().max()) <= 2e-6 * float(fw.abs().max())
    dy = torch.randn(N, L * Cd + L, device=dev, generator=gen)
    with torch.enable_grad():
        enc.encode_multisample(means, stds).backward(dy)
    lhs = float((dy[:, :L * Cd].double() * fused[:, :L * Cd].double()).sum())
    rhs = float((enc.embeddings.grad.double() * enc.embeddings.detach().double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), 1.0), (lhs, rhs)
    with pytest.raises(RuntimeError, match="table only"):
        enc.encode_multisample(means.requires_grad_(True), stds)
