"""Per-CTA timing of the weight-gradient kernel (csrc/snerf_train_tc.cu: dw_tc_kernel) for its host-side cost model:

    SNERF_DW_TIMING=1 [SNERF_DW_FLUSH_COST=x] python tools/dw_balance.py

Runs a few bf16 training steps on 512 rays and prints, per CTA of the last launch: total cycles, cycles in accumulator
flushes, number of flushes, 8 KiB units streamed -- and the least-squares fit  cycles ~ a * units + b * flushes + c."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import snerf_b200                                              # noqa: E402
from snerf_b200 import NeRF, _lib, make_query_fn, render_rays  # noqa: E402
from tools import synth                                        # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    nets = []
    for seed in (20, 21):
        m = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
        m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in synth.nerf_params(seed, trunk_gain=1.5, sigma_bias=1.0).items()})
        nets.append(m.to(dev))
    q, _, _ = make_query_fn()
    rs = np.random.RandomState(0)
    n = 512
    d = rs.standard_normal((n, 3)).astype(np.float32); d[:, 2] = -1
    rb = torch.from_numpy(synth.ray_batch(rs.standard_normal((n, 3)).astype(np.float32) * 0.1, d, 1.8, 110.0)).to(dev)
    snerf_b200.set_train_precision("bf16")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(6):
        for m in nets:
            m.zero_grad(set_to_none=True)
        out = render_rays(rb, nets[0], q, 64, N_importance=128, network_fine=nets[1], perturb=1.0, raw_noise_std=1.0)
        loss = out["rgb_map"].sum() + out["rgb0"].sum() + 0.01 * out["depth_map"].sum()
        torch.cuda.synchronize()
        e0.record()
        loss.backward()
        e1.record()
        torch.cuda.synchronize()
    snerf_b200.set_train_precision("fp32")
    print(f"backward (composite + chain + dW): {e0.elapsed_time(e1) * 1e3:.1f} us")
    n_cta = torch.cuda.get_device_properties(dev).multi_processor_count
    buf = (C.c_int64 * (4 * n_cta))()
    _lib.check(_lib.load().snerf_debug_dw_timing(buf, n_cta), "snerf_debug_dw_timing")
    t = np.array(buf, dtype=np.int64).reshape(n_cta, 4).astype(np.float64)
    if t[:, 0].max() == 0:
        print("no timing recorded: set SNERF_DW_TIMING=1")
        return
    print(f"total cycles: min {t[:, 0].min():.0f} mean {t[:, 0].mean():.0f} max {t[:, 0].max():.0f}; flush cycles mean {t[:, 1].mean():.0f} "
          f"max {t[:, 1].max():.0f}; flushes mean {t[:, 2].mean():.2f}; units mean {t[:, 3].mean():.0f} (min {t[:, 3].min():.0f}, max {t[:, 3].max():.0f})")
    A = np.stack([t[:, 3], t[:, 2], np.ones(n_cta)], 1)
    coef, *_ = np.linalg.lstsq(A, t[:, 0], rcond=None)
    print(f"fit: cycles = {coef[0]:.1f} * units + {coef[1]:.0f} * flushes + {coef[2]:.0f};  flush / unit = {coef[1] / coef[0]:.1f} units; "
          f"measured flush cycles per flush {t[:, 1].sum() / max(t[:, 2].sum(), 1):.0f}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.save(os.path.join(ROOT, "gpurun_out", "dw_timing_%s.npy" % os.environ.get("SNERF_DW_FLUSH_COST", "default")), t)
    for c in range(0, n_cta, 12):
        print(c, t[c].astype(np.int64).tolist())


if __name__ == "__main__":
    main()
