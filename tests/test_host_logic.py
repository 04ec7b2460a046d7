"""CPU: host-side mirror of the reference interface + the C ABI surface (no compute)."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from snerf_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.load()


def test_abi_exports_every_declared_symbol(lib):
    from snerf_b200 import _lib
    header = open(os.path.join(ROOT, "include", "snerf_b200.h")).read()
    declared = set(re.findall(r"\b(snerf_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.snerf_version() == int(re.search(r"#define\s+SNERF_ABI_VERSION\s+(\d+)", header).group(1))


def test_struct_layouts_match_header(lib):
    from snerf_b200 import _lib
    assert ctypes.sizeof(_lib.NetDesc) == 28
    assert ctypes.sizeof(_lib.NetF32) == 8 * (2 * 16 + 10)
    assert ctypes.sizeof(_lib.Rays) == 24
    assert ctypes.sizeof(_lib.Opts) == 32 + 10 * 8     # ... + camera pointer (ABI 10) + desc_fine (ABI 11)
    assert ctypes.sizeof(_lib.Camera) == 4 * 2 + 4 * 5 + 4 * 12 + 4 + 8   # (+4 padding before first_pixel)
    assert ctypes.sizeof(_lib.Linear) == 152 and ctypes.sizeof(_lib.MipEncode) == 96 and ctypes.sizeof(_lib.MipComposite) == 160
    assert ctypes.sizeof(_lib.Out) == 16 * 8
    assert ctypes.sizeof(_lib.GridDesc) == 36          # 8 x int32 + float
    assert ctypes.sizeof(_lib.StepfunOpts) == 36       # 3 x int32 + 6 x float


def test_packed_sizes_and_unsupported_configs(lib):
    from snerf_b200 import _lib
    d = _lib.NetDesc(8, 256, 63, 27, 4, 1, 5)
    n32 = lib.snerf_packed_bytes(ctypes.byref(d), _lib.MODE_FP32)
    n16 = lib.snerf_packed_bytes(ctypes.byref(d), _lib.MODE_BF16)
    # fp32: 2 KiB header + padded K-major weights; bf16: 72 chunks of 16 KiB + packets + dir weights
    assert n32 > 4 * 593408 and n32 < 4 * 700000
    assert n16 == 1024 + 72 * 16384 + 10 * 4160 + 128 * 32 * 4     # packets: 2112 B head + 2048 B bias tile (snerf_packed.h)
    small = _lib.NetDesc(4, 64, 63, 27, -1, 1, 4)
    assert lib.snerf_packed_bytes(ctypes.byref(small), _lib.MODE_FP32) > 0
    assert lib.snerf_packed_bytes(ctypes.byref(small), _lib.MODE_BF16) == 0
    assert b"tensor-core modes support" in lib.snerf_last_error()
    # fp16x3: every chunk twice (hi part, lo part)
    assert lib.snerf_packed_bytes(ctypes.byref(d), _lib.MODE_FP16X3) == 1024 + 144 * 16384 + 10 * 4160 + 128 * 32 * 4
    assert lib.snerf_packed_bytes(ctypes.byref(small), _lib.MODE_FP16X3) == 0
    bad = _lib.NetDesc(8, 100, 63, 27, 4, 1, 4)
    assert lib.snerf_packed_bytes(ctypes.byref(bad), _lib.MODE_FP32) == 0


def test_compute_entry_points_fail_loudly_without_gpu(lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.snerf_device_check(0) != 0
    from snerf_b200 import NeRF, raw2outputs, render_rays, make_query_fn
    net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, use_viewdirs=True)
    q, _, _ = make_query_fn()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        render_rays(torch.zeros(4, 11), net, q, 64, N_importance=128, network_fine=net)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        raw2outputs(torch.zeros(2, 8, 4), torch.zeros(2, 8), torch.zeros(2, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(3, 90))


def test_nerf_state_dict_matches_reference_names():
    from snerf_b200 import NeRF
    net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    sd = net.state_dict()
    shapes = {k: tuple(v.shape) for k, v in sd.items()}
    assert shapes["pts_linears.0.weight"] == (256, 63)
    assert shapes["pts_linears.5.weight"] == (256, 319)
    assert shapes["pts_linears.4.weight"] == (256, 256)
    assert shapes["views_linears.0.weight"] == (128, 283)
    assert shapes["feature_linear.weight"] == (256, 256)
    assert shapes["alpha_linear.weight"] == (1, 256)
    assert shapes["rgb_linear.weight"] == (3, 128)
    assert sum(v.numel() for v in sd.values()) == 595844
    d = net.desc()
    assert (d.D, d.W, d.skip, d.use_viewdirs) == (8, 256, 4, 1)
    # config 1: skips=[4] never reached with D=4
    small = NeRF(D=4, W=64, input_ch=63, input_ch_views=27, output_ch=4, skips=[4], use_viewdirs=True)
    assert small.desc().skip == -1
    no_vd = NeRF(D=8, W=256, input_ch=63, input_ch_views=0, output_ch=5, skips=[4], use_viewdirs=False)
    assert "output_linear.weight" in no_vd.state_dict() and tuple(no_vd.output_linear.weight.shape) == (5, 256)


def test_signatures_mirror_reference():
    from snerf_b200 import render as R, run_nerf_helpers as H
    sig = inspect.signature(R.render_rays)
    assert list(sig.parameters)[:13] == ["ray_batch", "network_fn", "network_query_fn", "N_samples", "retraw", "lindisp",
                                         "perturb", "N_importance", "network_fine", "white_bkgd", "raw_noise_std",
                                         "verbose", "pytest"]
    assert list(inspect.signature(R.render).parameters)[:13] == [
        "H", "W", "focal", "chunk", "rays", "c2w", "ndc", "near", "far", "use_viewdirs", "c2w_staticcam", "depths",
        "ori_points"]
    assert list(inspect.signature(H.sample_pdf).parameters)[:5] == ["bins", "weights", "N_samples", "det", "pytest"]
    assert list(inspect.signature(H.raw2outputs).parameters) == ["raw", "z_vals", "rays_d", "raw_noise_std",
                                                                 "white_bkgd", "pytest"]
    assert list(inspect.signature(H.NeRF.__init__).parameters)[1:] == ["D", "W", "input_ch", "input_ch_views",
                                                                       "output_ch", "skips", "use_viewdirs"]
    assert list(inspect.signature(H.run_network).parameters) == ["inputs", "viewdirs", "fn", "embed_fn",
                                                                 "embeddirs_fn", "netchunk"]


def test_create_nerf_kwargs(tmp_path):
    from types import SimpleNamespace
    from snerf_b200 import create_nerf
    args = SimpleNamespace(multires=10, multires_views=4, i_embed=0, use_viewdirs=True, N_importance=128, N_samples=64,
                           netdepth=8, netwidth=256, netdepth_fine=8, netwidth_fine=256, netchunk=65536,
                           alpha_model_path=None, weighted_loss=False, lrate=5e-4, basedir=str(tmp_path),
                           expname="exp", ft_path=None, no_reload=False, perturb=1.0, white_bkgd=False,
                           raw_noise_std=1.0, dataset_type="nuscenes", no_ndc=True, lindisp=False)
    kw_train, kw_test, start, grad_vars, opt, conf = create_nerf(args)
    assert set(kw_train) == {"network_query_fn", "perturb", "N_importance", "network_fine", "N_samples", "network_fn",
                             "use_viewdirs", "white_bkgd", "raw_noise_std", "ndc", "lindisp"}
    assert kw_test["perturb"] is False and kw_test["raw_noise_std"] == 0.
    assert start == 0 and conf is None and len(grad_vars) == 2 * 24
    assert kw_train["network_query_fn"].multires == 10 and kw_train["network_query_fn"].multires_views == 4
    # checkpoint round trip with the reference's key names (render.py:229-247)
    os.makedirs(tmp_path / "exp")
    torch.save({"global_step": 7, "optimizer_state_dict": opt.state_dict(),
                "network_fn_state_dict": kw_train["network_fn"].state_dict(),
                "network_fine_state_dict": kw_train["network_fine"].state_dict()}, tmp_path / "exp" / "000007.tar")
    _, _, start2, *_ = create_nerf(args)
    assert start2 == 7


def test_embedder_out_dims():
    from snerf_b200 import get_embedder
    fn, dim = get_embedder(10, 0)
    assert dim == 63 and fn.multires == 10
    fn, dim = get_embedder(4, 0)
    assert dim == 27
    fn, dim = get_embedder(10, -1)
    assert dim == 3 and fn.multires == -1


def test_keras_weight_import_matches_reference():
    """NeRF.load_weights_from_keras against the reference's own method (run_nerf_helpers.py:128-155) on the same flat
    Keras-style weight list (needs /root/reference: runs in the build container, skipped on the GPU box)."""
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference not mounted")
    _, ref_h = ref_import.load()
    from snerf_b200 import NeRF
    rs = np.random.RandomState(5)
    dims = [(63, 256)] + [(256, 256)] * 4 + [(319, 256)] + [(256, 256)] * 2 + [(256, 256), (283, 128), (128, 3), (256, 1)]
    weights = []
    for i, o in dims:                       # Keras stores kernels [in, out]
        weights += [rs.standard_normal((i, o)).astype(np.float32), rs.standard_normal(o).astype(np.float32)]
    kw = dict(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    ours, ref = NeRF(**kw), ref_h.NeRF(**kw)
    ours.load_weights_from_keras(weights)
    ref.load_weights_from_keras(weights)
    sd_o, sd_r = ours.state_dict(), ref.state_dict()
    assert set(sd_o) == set(sd_r)
    for k in sd_r:
        assert torch.equal(sd_o[k], sd_r[k]), k
    with pytest.raises(AssertionError):
        NeRF(D=8, W=256, input_ch=63, input_ch_views=0, output_ch=5, use_viewdirs=False).load_weights_from_keras(weights)


def test_product_path_never_imports_the_oracle():
    """oracle/ is test infrastructure: the package must not import it, bench.py only inside its CPU-baseline /
    reference-arm legs (and the tools that time the reference's kernels)."""
    import ast
    pkg = os.path.join(ROOT, "snerf_b200")
    for fn in sorted(os.listdir(pkg)):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            for node in ast.walk(ast.parse(src)):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    names = [node.module or ""]
                assert not any(n == "oracle" or n.startswith("oracle.") for n in names), (fn, names)
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    allowed = {"pick_cpu_threads", "cpu_port_rays_per_s", "cpu_train_rays_per_s", "run_reference_arm", "reference_cpu"}
    for fn in [n for n in tree.body if isinstance(n, ast.FunctionDef)]:
        imports = [n for n in ast.walk(fn) if isinstance(n, ast.ImportFrom) and (n.module or "").startswith("oracle")]
        if fn.name == "main":            # one import, inside the `world == 1 and not args.no_cpu_baseline` leg
            assert len(imports) == 1, [ast.dump(i) for i in imports]
        elif imports:
            assert fn.name in allowed, fn.name


def test_synthetic_workload_helpers_match_the_oracle_data_helpers():
    """tools/synth.py (what the product arms of bench.py use) generates the same weights / rays as the oracle's helpers
    (what the fixtures and the CPU arm use), so both arms of the bench run on identical inputs."""
    from oracle import snerf_oracle as O
    from tools import synth
    a, b = synth.nerf_params(20, trunk_gain=1.5, sigma_bias=1.0), O.make_nerf_params(20, trunk_gain=1.5, sigma_bias=1.0)
    assert set(a) == set(b) and all(np.array_equal(a[k], b[k]) for k in b)
    c2w = synth.camera(2)
    o1, d1 = synth.pinhole_rays(90, 160, 126.64, c2w, [81.63, 49.15])
    o2, d2 = O.pinhole_rays(90, 160, 126.64, c2w, [81.63, 49.15])
    assert np.array_equal(o1, o2) and np.array_equal(d1, d2)
    assert np.array_equal(synth.ray_batch(o1, d1, 1.8, 110.0), O.pack_ray_batch(o2, d2, 1.8, 110.0))


def test_new_entry_points_validate_arguments_before_touching_the_device(lib):
    """Bad arguments are reported through the status code + snerf_last_error() without a GPU (no exceptions, no exit)."""
    from snerf_b200 import _lib
    C_ = ctypes
    err = lambda: _lib.last_error()
    d = _lib.GridDesc(3, 4, 10, 16, 0, 0, 0, 0, 1.0)
    assert lib.snerf_grid_encode_ms_fwd(C_.byref(d), None, None, 1.0, None, None, None, None, None, 50, 4, 6, None) != 0
    assert "bad argument" in err()
    assert lib.snerf_grid_encode_ms_fwd(C_.byref(d), None, None, 1.0, None, None, None, None, None, 50, 0, 6, None) == 0   # empty batch
    bad = _lib.GridDesc(3, 3, 10, 16, 0, 0, 0, 0, 1.0)
    assert lib.snerf_grid_encode_ms_fwd(C_.byref(bad), None, None, 1.0, None, None, None, None, None, 50, 4, 6, None) != 0
    assert "level_dim" in err()
    o = _lib.StepfunOpts(1, 1, 1, 0.01, 0.0, 1.0, 1.0, 1e-5, 0.0)
    buf = (C_.c_float * 64)()
    p = C_.cast(buf, C_.c_void_p)
    assert lib.snerf_stepfun_resample(C_.byref(o), p, p, 2, 4, p, None, 0, 8, p, None, None, None, None) != 0
    assert "logits" in err()                                           # dilation works on weights
    o = _lib.StepfunOpts(0, 0, 0, 0.0, 0.0, 1.0, 1.0, 1e-5, 0.0)
    assert lib.snerf_stepfun_resample(C_.byref(o), p, p, 2, 4, p, p, 3, 8, p, None, None, None, None) != 0
    assert "jitter" in err()
    assert lib.snerf_stepfun_resample(C_.byref(o), p, p, 0, 4, None, None, 0, 8, None, None, None, None, None) == 0       # no rays
    lo = _lib.LossOpts(0.1, 0.2, 0.0, 1)
    assert lib.snerf_loss_fwd(C_.byref(lo), p, None, p, p, None, None, None, 4, p, p, None) != 0
    assert "depth0" in err()
    assert lib.snerf_loss_fwd(C_.byref(lo), p, None, p, None, None, None, p, 4, p, p, None) != 0
    assert "confidence" in err()
    assert lib.snerf_proposal_loss(None, p, p, p, 4, 8, 8, 1.0, p, p, None, None) != 0
    assert "bad argument" in err()


def test_stepfun_uniform_draw_matches_reference_sample():
    """snerf_b200.stepfun._uniform (host side of sample_intervals: the linspace term + the torch.rand draw the kernel
    scales and adds) against the `u` the reference's stepfun.sample builds (stepfun.py:199-216), same seed, on the CPU."""
    from snerf_b200 import stepfun
    eps = torch.finfo(torch.float32).eps
    for n, single in ((64, True), (32, False), (128, True)):
        # deterministic_center=True, rand=None
        base, jitter, mj = stepfun._uniform(None, n, single, (5,), torch.device("cpu"))
        pad = 1 / (2 * n)
        assert jitter is None and mj == 0.0
        assert torch.equal(base, torch.linspace(pad, 1. - pad - eps, n))
        # randomized
        torch.manual_seed(7)
        base, jitter, mj = stepfun._uniform(True, n, single, (5,), torch.device("cpu"))
        torch.manual_seed(7)
        u_max = eps + (1 - eps) / n
        max_jitter = (1 - u_max) / (n - 1) - eps
        d = 1 if single else n
        u_ref = torch.linspace(0, 1 - u_max, n) + torch.rand((5,) + (d,)) * max_jitter        # stepfun.py:211-216
        assert mj == max_jitter and jitter.shape == (5, d)
        assert torch.equal(base + jitter * mj, u_ref)
        assert float(u_ref.max()) < 1.0


def test_committed_bench_lines_follow_the_contract():
    """The JSON lines bench.py printed on the B200 (profiles/r1c_bench_*.json) carry every key of the driver's contract."""
    import json
    load = lambda name: json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().split("\n")[-1])
    ours, ref = load("r1c_bench_n1.json"), load("r1c_bench_reference_arm.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in ours, k
    assert ours["n_gpus"] == 1 and ours["warmup"] >= 3 and ours["scaling"] == "weak" and ours["data"] == "synthetic"
    assert ours["vs_baseline"] is None and "workload" in ours["config"] and "model" not in ours["config"]
    assert set(ours["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and ours["e2e"]["h2d_bytes_per_step"] > 0
    assert 0 < ours["e2e"]["value"] <= ours["value"] * 1.01 and ours["e2e"]["value"] != ours["value"]
    rl = ours["roofline"]
    assert rl["bound"] == "tensor" and rl["unit"] == "TFLOP/s" and abs(rl["frac"] - rl["achieved"] / rl["peak"]) < 1e-9 and rl["traffic"]
    cb = ours["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"]
    assert set(ours["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(ours["clocks"]["reasons"])
    assert ours["gpu_launches"] == ours["steps"]                       # one fused launch per step
    assert ref["impl"] == "reference" and ref["metric"] == ours["metric"] and ref["unit"] == ours["unit"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["value"] == ref["value"] and ref["cpu_baseline"]["value"] == ref["value"]
    two = load("r1c_bench_n2.json")
    assert two["n_gpus"] == 2 and two["value"] > 1.8 * ours["value"] * 0.95 and two["frame6"]["scaling"] == "strong"


def test_new_operators_refuse_cpu_tensors():
    """No CPU / PyTorch fallback anywhere: the zip-NeRF operators and the fused losses raise on CPU tensors."""
    from snerf_b200 import stepfun
    from snerf_b200.gridencoder import GridEncoder
    from snerf_b200.losses import ProposalLoss, RgbDepthLoss
    enc = GridEncoder(input_dim=3, num_levels=4, level_dim=2, base_resolution=4, log2_hashmap_size=10)
    assert set(enc.state_dict()) == {"embeddings", "offsets", "idx", "grid_sizes"} and enc.output_dim == 8
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc(torch.zeros(5, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc.encode_multisample(torch.zeros(5, 6, 3), torch.ones(5, 6), scale_featurization=False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        stepfun.resample_intervals(None, torch.linspace(0, 1, 9)[None], torch.ones(1, 8), 8, dilation=0.01)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        stepfun.max_dilate_weights(torch.linspace(0, 1, 9)[None], torch.ones(1, 8), 0.01)
    with pytest.raises(ValueError):
        stepfun.sample_intervals(None, torch.linspace(0, 1, 9)[None], torch.zeros(1, 8), 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        RgbDepthLoss()(torch.zeros(4, 3), torch.zeros(4, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ProposalLoss()(torch.linspace(0, 1, 9)[None], torch.ones(1, 8), torch.linspace(0, 1, 9)[None], torch.ones(1, 8))


def test_stage_entry_points_warn_once_when_gradients_are_expected():
    """NeRF.forward / run_network / raw2outputs / sample_pdf are forward-only kernels (the reference's are differentiable
    eager code): with grad mode on and trainable inputs they say so once; not under no_grad, not for frozen modules."""
    import warnings
    from snerf_b200 import run_nerf_helpers as H
    H._STAGE_WARNED.discard("NeRF.forward")
    net = H.NeRF(D=2, W=64, input_ch=63, input_ch_views=27, use_viewdirs=True)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        with torch.no_grad():
            H._warn_stage_no_grad("NeRF.forward", net)
        H._warn_stage_no_grad("NeRF.forward", net.requires_grad_(False), torch.zeros(1))
        assert len(w) == 0
        H._warn_stage_no_grad("NeRF.forward", net.requires_grad_(True))
        H._warn_stage_no_grad("NeRF.forward", net)
        assert len(w) == 1 and "no autograd graph" in str(w[0].message)


@pytest.mark.parametrize("rays,n_cta", [(512, 148), (4096, 148), (5, 148), (512, 1), (33, 7)])
def test_weight_gradient_cuts_cover_and_balance(rays, n_cta):
    """Host-side partition of the weight-gradient launch (csrc/snerf_train_tc.cu: dw_compute_cuts; cost model fitted by
    tools/dw_balance.py): the per-SM ranges are contiguous, ordered, cover every 64-row block of all 28 problems exactly
    once, and no SM's modelled cost exceeds the reported makespan, which stays close to the perfectly divisible optimum."""
    import ctypes as C
    from snerf_b200 import _lib
    lib = _lib.load()
    pairs = (rays + 1) // 2
    rows_c, rows_f = pairs * 2 * 64, pairs * 2 * 192
    cut = (C.c_int64 * (n_cta + 1))()
    first = (C.c_int64 * 29)()
    units = (C.c_double * 28)()
    span = C.c_double()
    n = lib.snerf_debug_dw_cuts(rows_c, rows_f, n_cta, cut, first, units, C.byref(span))
    assert n == 28, _lib.last_error()
    cut, first, units = list(cut), list(first), list(units)
    total = first[28]
    assert total == 14 * (rows_c // 64) + 14 * (rows_f // 64)
    assert cut[0] == 0 and cut[-1] == total and all(a <= b for a, b in zip(cut, cut[1:]))
    assert all(a < b for a, b in zip(first, first[1:]))
    # streamed units per SM: what the equal-bytes cut of the first version equalised -- now allowed to differ by problem type
    per_sm = []
    for c in range(n_cta):
        u = 0.0
        for p in range(28):
            lo, hi = max(cut[c], first[p]), min(cut[c + 1], first[p + 1])
            if hi > lo:
                u += (hi - lo) * units[p]
        per_sm.append(u)
    assert abs(sum(per_sm) - sum(units[p] * (first[p + 1] - first[p]) for p in range(28))) < 1e-6
    busy = [u for u in per_sm if u > 0]
    if n_cta == 148 and rays >= 512:
        assert len(busy) == 148 and max(busy) < 1.35 * (sum(busy) / len(busy))
    assert span.value > 0
