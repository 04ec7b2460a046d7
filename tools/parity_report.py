"""Per-fixture, per-mode parity table against the committed golden vectors of the unmodified reference
(GPU box; test infrastructure).  The metric is tests/conftest.err_metric: max |a-b| / (|b| + 1% rms(b)).

  python tools/parity_report.py  ->  one JSON line per (fixture, mode); markdown table on stderr
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import err_metric, golden_params, load_golden   # noqa: E402
from oracle import snerf_oracle as O                           # noqa: E402
from test_gpu_parity import CFG2, run_fused                    # noqa: E402

KEYS = ["rgb_map", "disp_map", "acc_map", "depth_map", "weights", "rgb0", "disp0", "acc0", "z_std"]


def main():
    dev = torch.device("cuda:0")
    rows = []
    with torch.no_grad():
        for name in CFG2:
            g = load_golden(name)
            pc, pf = golden_params(g)
            rb = g["ray_batch"]
            for mode in ("fp32", "fp16x3", "fp16", "bf16"):
                out, ex = run_fused(g, dev, mode)
                row = {"fixture": name, "mode": mode,
                       "z_vals_bit_exact": bool(np.array_equal(out["z_vals_map"], g["out_z_vals_map"])),
                       "rgb_l1": float(np.mean(np.abs(out["rgb_map"] - g["out_rgb_map"])))}
                for k in KEYS:
                    row[k] = err_metric(out[k], g["out_" + k])
                row["raw_coarse"] = err_metric(ex["raw_coarse"], g["mid_raw_coarse"], floor=0.1)
                # fraction of resampled depths that landed in another bin than the reference's
                a, b = ex["z_samples"], g["mid_z_samples"]
                row["z_samples_moved"] = float(np.mean(np.abs(a - b) > 1e-4 * np.abs(b) + 1e-5))
                # fine pass judged at the kernel's own depths (removes the bin-flip discontinuity)
                pts = rb[:, None, 0:3] + rb[:, None, 3:6] * ex["z_all"][:, :, None]
                raw_ref = O.query_network(pf, pts.astype(np.float32), rb[:, -3:])
                noise1 = g["noise1"] if "noise1" in g else None
                rgb, disp, acc, w, depth = O.composite(raw_ref, ex["z_all"], rb[:, 3:6], noise1, bool(g["white_bkgd"]))
                row["fine_at_own_depths"] = {"raw": err_metric(out["raw"], raw_ref, floor=0.1),
                                             "rgb_map": err_metric(out["rgb_map"], rgb),
                                             "depth_map": err_metric(out["depth_map"], depth),
                                             "acc_map": err_metric(out["acc_map"], acc),
                                             "weights_fine": err_metric(ex["weights_fine"], w, floor=0.1)}
                rows.append(row)
                print(json.dumps(row), flush=True)
    hdr = ["fixture", "mode", "rgb_l1"] + KEYS + ["fine rgb@own z", "fine depth@own z", "fine w@own z", "z_samples moved"]
    print("| " + " | ".join(hdr) + " |", file=sys.stderr)
    print("|" + "---|" * len(hdr), file=sys.stderr)
    for r in rows:
        f = r["fine_at_own_depths"]
        cells = [r["fixture"], r["mode"], f"{r['rgb_l1']:.1e}"] + [f"{r[k]:.1e}" for k in KEYS] + \
                [f"{f['rgb_map']:.1e}", f"{f['depth_map']:.1e}", f"{f['weights_fine']:.1e}", f"{r['z_samples_moved']:.3f}"]
        print("| " + " | ".join(cells) + " |", file=sys.stderr)


if __name__ == "__main__":
    main()
