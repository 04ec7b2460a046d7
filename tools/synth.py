"""Synthetic workload of the benchmarks (no reference data travels to the GPU box): seeded network weights with the
reference's state_dict names / shapes, pinhole camera rays and the [N, 11] ray batch `render()` builds.  Plain numpy,
bit-reproducible from the seeds on every box.  Workload generation only -- not a compute path and not the oracle."""
import numpy as np

F32 = np.float32


def nerf_params(seed, D=8, W=256, input_ch=63, input_ch_views=27, skips=(4,), trunk_gain=1.0, sigma_bias=0.0):
    """nn.Linear-default-like weights (U(+-1/sqrt(fan_in))) for `NeRF` (run_nerf_helpers.py:75-101 names); trunk_gain /
    sigma_bias make the density field non-degenerate (weight set B)."""
    rs = np.random.RandomState(seed)

    def lin(n_out, n_in, gain=1.0):
        bound = 1.0 / np.sqrt(n_in)
        w = rs.uniform(-bound, bound, size=(n_out, n_in)) * gain
        b = rs.uniform(-bound, bound, size=(n_out,))
        return w.astype(F32), b.astype(F32)

    p = {}
    for i in range(D):
        n_in = input_ch if i == 0 else (W + input_ch if (i - 1) in skips else W)
        p[f"pts_linears.{i}.weight"], p[f"pts_linears.{i}.bias"] = lin(W, n_in, trunk_gain)
    p["views_linears.0.weight"], p["views_linears.0.bias"] = lin(W // 2, W + input_ch_views)
    p["feature_linear.weight"], p["feature_linear.bias"] = lin(W, W)
    p["alpha_linear.weight"], p["alpha_linear.bias"] = lin(1, W)
    p["alpha_linear.bias"] = (p["alpha_linear.bias"] + F32(sigma_bias)).astype(F32)
    p["rgb_linear.weight"], p["rgb_linear.bias"] = lin(3, W // 2)
    return p


def camera(cam_index):
    """Synthetic camera `cam_index` of a 6-camera rig: c2w [3,4], yawed by 60 degrees per camera."""
    a = np.deg2rad(60.0 * cam_index)
    return np.array([[np.cos(a), 0, np.sin(a), 0.5 * cam_index], [0, 1, 0, 0.1], [-np.sin(a), 0, np.cos(a), 1.5]], F32)


def pinhole_rays(H, W, focal, c2w, centre):
    """Pixel-centre pinhole rays (origins [H,W,3], directions [H,W,3]), the convention of get_rays
    (run_nerf_helpers.py:247-258): dirs = ((i + .5 - cx) / f, -(j + .5 - cy) / f, -1) rotated by c2w[:3,:3]."""
    c2w = np.asarray(c2w, F32)
    i = np.arange(W, dtype=F32)[None, :].repeat(H, 0)
    j = np.arange(H, dtype=F32)[:, None].repeat(W, 1)
    f = F32(focal)
    cam = np.stack([((i + F32(0.5)) - F32(centre[0])) / f, -((j + F32(0.5)) - F32(centre[1])) / f, -np.ones_like(i)], -1).astype(F32)
    prod = cam[..., None, :] * c2w[:3, :3]
    d = ((prod[..., 0] + prod[..., 1]) + prod[..., 2]).astype(F32)
    return np.broadcast_to(c2w[:3, 3], d.shape).astype(F32), d


def ray_batch(rays_o, rays_d, near, far):
    """[N, 11] = o, d, near, far, unit viewdirs (render.py:56-79)."""
    o = np.asarray(rays_o, F32).reshape(-1, 3)
    d = np.asarray(rays_d, F32).reshape(-1, 3)
    ones = np.ones_like(d[:, :1])
    nrm = np.sqrt(((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]).astype(F32))
    return np.concatenate([o, d, F32(near) * ones, F32(far) * ones, (d / nrm[:, None]).astype(F32)], -1).astype(F32)


pack_ray_batch = ray_batch          # name used by the profiling tools
