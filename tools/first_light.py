"""Bring-up diagnostics for a GPU box: runs each piece in isolation, prints a summary and stores
the raw arrays under gpurun_out/ so failures can be analysed offline.  Not part of the product."""
import os, sys, json, traceback
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "gpurun_out"); os.makedirs(OUT, exist_ok=True)
from conftest import load_golden, golden_params, err_metric
from oracle import snerf_oracle as O
import snerf_b200
from snerf_b200 import _lib
dev = torch.device("cuda", 0)
summary = {"gpu": torch.cuda.get_device_name(0)}

def stage(name):
    def deco(fn):
        try:
            summary[name] = fn()
        except Exception as e:
            summary[name] = "EXC: " + repr(e)
            traceback.print_exc()
            try: torch.cuda.synchronize()
            except Exception as e2: summary[name] += " | sync: " + repr(e2)
        print(name, "->", summary[name], flush=True)
    return deco

def umma(variant):
    rs = np.random.RandomState(0)
    a = rs.standard_normal((128, 64)).astype(np.float32); b = rs.standard_normal((128, 64)).astype(np.float32)
    ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    td = torch.full((128, 128), -777.0, dtype=torch.float32, device=dev)
    _lib.check(_lib.load().snerf_selftest_umma(_lib.ptr(ta), _lib.ptr(tb), _lib.ptr(td), variant, _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    a16 = ta.to(torch.bfloat16).float().cpu().numpy(); b16 = tb.to(torch.bfloat16).float().cpu().numpy()
    ref = a16.astype(np.float64) @ b16.astype(np.float64).T
    d = td.cpu().numpy()
    np.savez(os.path.join(OUT, f"selftest_umma_{variant}.npz"), a=a16, b=b16, d=d, ref=ref)
    return {"max_abs_err": float(np.max(np.abs(d - ref))), "untouched": int((d == -777.0).sum())}

stage("umma_selftest_ss")(lambda: umma(0))
stage("umma_selftest_ts")(lambda: umma(1))

def fused(name, mode):
    from test_gpu_parity import run_fused
    g = load_golden(name)
    out, ex = run_fused(g, dev, mode)
    res = {}
    for k in out:
        if "out_" + k in g: res[k] = err_metric_safe(out[k], g["out_" + k])
    for k in ex:
        if "mid_" + k in g: res["mid_" + k] = err_metric_safe(ex[k], g["mid_" + k])
    res["rgb_l1"] = float(np.nanmean(np.abs(out["rgb_map"] - g["out_rgb_map"])))
    np.savez(os.path.join(OUT, f"fused_{mode}_{name}.npz"), **out, **{"ex_" + k: v for k, v in ex.items()})
    return res

def err_metric_safe(a, b):
    try: return float(f"{err_metric(a, b):.3e}")
    except AssertionError as e: return "nonfinite-mismatch"

for nm in ["cfg1_plumbing", "cfg2_peaky"]:
    stage(f"fused_fp32_{nm}")(lambda nm=nm: fused(nm, "fp32"))
for nm in ["cfg2_peaky", "cfg2_stochastic"]:
    stage(f"fused_bf16_{nm}")(lambda nm=nm: fused(nm, "bf16"))

@stage("timing")
def _():
    from test_gpu_parity import make_net
    from snerf_b200 import make_query_fn, render_rays
    pc = O.make_nerf_params(1, trunk_gain=1.5, sigma_bias=0.5); pf = O.make_nerf_params(2, trunk_gain=1.5, sigma_bias=0.5)
    nc, nf = make_net(pc, 8, 256, dev), make_net(pf, 8, 256, dev)
    q, _, _ = make_query_fn()
    res = {}
    for mode, n in (("fp32", 4096), ("bf16", 65536), ("bf16", 1 << 20)):
        rs = np.random.RandomState(0)
        o = np.zeros((n, 3), np.float32); d = rs.standard_normal((n, 3)).astype(np.float32); d[:, 2] = -1
        rb = torch.from_numpy(O.pack_ray_batch(o, d, 1.8, 110.0)).to(dev)
        snerf_b200.set_mode(mode)
        for _ in range(2): render_rays(rb, nc, q, 64, N_importance=128, network_fine=nf)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); render_rays(rb, nc, q, 64, N_importance=128, network_fine=nf); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        res[f"{mode}_{n}"] = {"ms": ms, "rays_per_s": n / ms * 1e3, "tflops": n / ms * 1e3 * 303.83e6 / 1e12}
    snerf_b200.set_mode("fp32")
    return res

json.dump(summary, open(os.path.join(OUT, "first_light.json"), "w"), indent=1)
print(json.dumps(summary, indent=1))
