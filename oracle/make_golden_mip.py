"""Generate tests/golden/mip_*.npz from the UNMODIFIED reference: MipNerfModel.forward (s-nerf/model/models.py:72-187) on
the warp path the shipped config runs (configs/nuScenes_depth_6cams: no_warp_sample = 0, fn = 1, transform_idx = 0,
ray_shape = 'cone', hidden_layer = 1024, rgb_layer = 3, N_samples = N_fine = 128, density_noise = 0), torch-CPU fp32.

    python oracle/make_golden_mip.py          (build container: /root/reference must be mounted)

The network weights are NOT stored (8.6 M values): both this script and the tests regenerate them from the seed with
oracle.mip_oracle.make_mip_params.  Stored: rays, the random draws replayed from the torch seed, every output of forward.
TEST INFRASTRUCTURE ONLY."""
import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mip_oracle as MO, ref_import    # noqa: E402


def load_reference_models():
    ref_import.load()
    if "turtle" not in sys.modules:                 # models.py:2 `from turtle import forward` (unused; needs tkinter)
        t = types.ModuleType("turtle"); t.forward = None; sys.modules["turtle"] = t
    return importlib.import_module("model.models")


def rays_like_nuscenes(n, seed):
    rs = np.random.RandomState(seed)
    H, W, focal, cx, cy = 900, 1600, 1266.4, 816.3, 491.5
    i, j = rs.uniform(0, W, n), rs.uniform(0, H, n)
    dirs = np.stack([(i + 0.5 - cx) / focal, -(j + 0.5 - cy) / focal, -np.ones(n)], -1)
    yaw = rs.uniform(0, 2 * np.pi)
    R = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
    d = (dirs @ R.T).astype(np.float32)
    o = (rs.standard_normal((n, 3)) * np.array([2.0, 0.3, 2.0])).astype(np.float32)
    vd = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(np.float32)
    radii = np.full((n, 1), 2.0 / np.sqrt(12.0) / focal, np.float32) * rs.uniform(0.8, 1.2, (n, 1)).astype(np.float32)
    return o, d, vd, radii, np.full((n, 1), 1.8, np.float32), np.full((n, 1), 110.0, np.float32)


def build(models, P, hidden, rgb_layer, n_samples, n_fine, density_noise=0.0):
    m = models.MipNerfModel(no_warp_sample=0, ray_shape="cone", fn=1, max_deg_point=16, radius=3.0, transform_idx=0, real=True,
                            rgb_layer=rgb_layer, hidden_layer=hidden, density_noise=density_noise, n_samples=n_samples,
                            proposal_loss=True, N_fine=n_fine)
    missing = m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in P.items()}, strict=True)
    return m.eval()


def run_case(models, name, seed, n_rays, hidden, rgb_layer, n_samples, n_fine, randomized, white):
    P = MO.make_mip_params(seed, hidden, rgb_layer)
    model = build(models, P, hidden, rgb_layer, n_samples, n_fine)
    o, d, vd, radii, near, far = rays_like_nuscenes(n_rays, seed + 1)
    T = torch.from_numpy
    Rays = models.utils.Rays if hasattr(models.utils, "Rays") else None
    import collections
    if Rays is None:
        Rays = collections.namedtuple('Rays', ('origins', 'directions', 'viewdirs', 'radii', 'lossmult', 'near', 'far', 'app'))
    rays = Rays(T(o), T(d), T(vd), T(radii), torch.ones(n_rays, 1), T(near), T(far), None)
    s_rand = u_rand = None
    if randomized:   # replay the two draws forward() makes: torch.rand [N, S+1] (mip.py:283), uniform_ [N, n_fine] (math_ops.py:51)
        torch.manual_seed(seed)
        s_rand = torch.rand(n_rays, n_samples + 1).numpy()
        u_rand = torch.empty(n_rays, n_fine).uniform_(to=1 / n_fine - torch.finfo(torch.float32).eps).numpy()
        torch.manual_seed(seed)
    with torch.no_grad():
        ret = model(rays, randomized, white, torch.zeros(3))
    out = {"seed": seed, "hidden": hidden, "rgb_layer": rgb_layer, "n_samples": n_samples, "n_fine": n_fine,
           "randomized": int(randomized), "white_bkgd": int(white), "origins": o, "directions": d, "viewdirs": vd, "radii": radii,
           "near": near, "far": far, "torch_version": torch.__version__,
           "dist0": ret[0][1].numpy(), "acc0": ret[0][2].numpy(), "s_vals0": ret[0][3].numpy(), "weights0": ret[0][4].numpy(),
           "rgb": ret[1][0].numpy(), "dist1": ret[1][1].numpy(), "acc1": ret[1][2].numpy(), "s_vals1": ret[1][4].numpy(),
           "weights1": ret[1][5].numpy()}
    assert ret[0][0] is None and ret[1][3] is None
    if randomized:
        out.update(s_rand=s_rand, u_rand=u_rand)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
    print(name, "rgb mean", float(out["rgb"].mean()), "acc1 mean", float(out["acc1"].mean()), "acc0 mean", float(out["acc0"].mean()))


def main():
    models = load_reference_models()
    run_case(models, "mip_shipped_det", 300, 48, 1024, 3, 128, 128, False, False)
    run_case(models, "mip_shipped_rand", 310, 48, 1024, 3, 128, 128, True, False)
    run_case(models, "mip_small", 320, 64, 256, 1, 64, 64, False, False)   # (white_bg=True raises inside the reference: mip.py:188 adds to a None comp_rgb at level 0)


if __name__ == "__main__":
    main()
