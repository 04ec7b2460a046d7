"""Small invocations of every kernel added in round 2, for `compute-sanitizer --tool memcheck` (GPU box):
    compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_r2.py
Covers: tensor-core training step (odd ray count: padding ray, dummy tiles), one-kernel Adam, the mip-NeRF model (both layer
GEMM variants, ragged row counts), the tensor-core NeRF.forward chain and the camera prologue."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import snerf_b200                                              # noqa: E402
from snerf_b200 import NeRF, make_query_fn, render_rays        # noqa: E402
from snerf_b200.models import MipNerfModel, Rays               # noqa: E402
from snerf_b200.optim import FlatAdam                          # noqa: E402
from snerf_b200.render import render                           # noqa: E402
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
    d = rs.standard_normal((5, 3)).astype(np.float32); d[:, 2] = -1
    rb = torch.from_numpy(synth.ray_batch(np.zeros((5, 3), np.float32), d, 1.8, 110.0)).to(dev)
    # 1. tensor-core training step on 5 rays (3 pairs: one padding ray), twice, with the flat Adam
    opt = FlatAdam(nets, lr=5e-4)
    snerf_b200.set_train_precision("bf16")
    for _ in range(2):
        opt.zero_grad()
        out = render_rays(rb, nets[0], q, 64, N_importance=128, network_fine=nets[1], perturb=1.0, raw_noise_std=1.0)
        (out["rgb_map"].sum() + out["rgb0"].sum() + 0.01 * out["depth_map"].sum()).backward()
        opt.step()
    snerf_b200.set_train_precision("fp32")
    torch.cuda.synchronize()
    assert torch.isfinite(opt.grads.flat).all()
    opt.grads.release()
    # 2. mip-NeRF model: 3 rays (381 fine rows: ragged last tile), both GEMM variants
    for pair in ("1", "0"):
        os.environ["SNERF_LIN_2SM"] = pair      # (read once per process by the library: the second value only documents intent)
        mm = MipNerfModel(no_warp_sample=0, ray_shape="cone", fn=1, rgb_layer=3, hidden_layer=256, density_noise=1.0, n_samples=128,
                          proposal_loss=True, N_fine=128).to(dev)
        o3 = torch.randn(3, 3, device=dev)
        d3 = torch.randn(3, 3, device=dev)
        one = torch.ones(3, 1, device=dev)
        with torch.no_grad():
            ret = mm(Rays(o3, d3, d3 / d3.norm(dim=-1, keepdim=True), one * 1e-3, None, one * 1.8, one * 110.0, None), True, False, None)
        assert torch.isfinite(ret[1][0]).all()
    # 3. tensor-core stage entry points + camera prologue
    for p in nets[0].parameters():
        p.requires_grad_(False)
    for p in nets[1].parameters():
        p.requires_grad_(False)
    snerf_b200.set_mode("bf16")
    with torch.no_grad():
        y = nets[0](torch.randn(100, 90, device=dev))
        c2w = np.array([[1, 0, 0, 0.1], [0, 1, 0, 0.2], [0, 0, 1, 0.3]], np.float32)
        img = render(5, 7, 6.0, chunk=None, c2w=c2w, ndc=False, near=1.8, far=110., use_viewdirs=True, network_fn=nets[0],
                     network_query_fn=q, N_samples=64, N_importance=128, network_fine=nets[1], perturb=0., raw_noise_std=0.)
    snerf_b200.set_mode("fp32")
    torch.cuda.synchronize()
    assert torch.isfinite(y).all() and torch.isfinite(img[0]).all()
    # 4. fp32 training level on the other network shapes: output_linear head, NeRF_RGB + frozen alpha_model (also as
    #    network_fn=None), coarse 2x64 under fine 4x128 (SnerfOpts.desc_fine); 3 rays, ragged 40 + 24 samples
    from snerf_b200.run_nerf_helpers import NeRF_RGB
    kw = dict(input_ch=63, input_ch_views=27, output_ch=5, skips=[4])
    rb3 = rb[:3].contiguous()

    def step(net_c, net_f):
        for n in (net_c, net_f):
            if n is not None:
                n.requires_grad_(True)
        out = render_rays(rb3, net_c, q, 40, N_importance=24, network_fine=net_f, perturb=1.0, raw_noise_std=1.0)
        (out["rgb_map"].sum() + out["rgb0"].sum() + 0.01 * out["depth_map"].sum()).backward()
        torch.cuda.synchronize()
        assert torch.isfinite(out["rgb_map"]).all()

    step(NeRF(D=3, W=64, use_viewdirs=False, **kw).to(dev), NeRF(D=3, W=64, use_viewdirs=False, **kw).to(dev))
    step(NeRF(D=2, W=64, use_viewdirs=True, **kw).to(dev), NeRF(D=4, W=128, use_viewdirs=True, **kw).to(dev))
    alpha = NeRF(D=3, W=128, use_viewdirs=True, **kw).to(dev)
    step(NeRF_RGB(D=3, W=128, use_viewdirs=True, alpha_model=alpha, **kw).to(dev),
         NeRF_RGB(D=3, W=128, use_viewdirs=True, alpha_model=alpha, **kw).to(dev))
    step(None, NeRF_RGB(D=3, W=128, use_viewdirs=True, alpha_model=alpha, **kw).to(dev))
    # 5. the fused renderer (bf16 / fp16); with SNERF_B200_PAIR=1 in the environment this is the cta_group::2 variant
    for mode in ("bf16", "fp16"):
        snerf_b200.set_mode(mode)
        with torch.no_grad():
            o5 = render_rays(rb, nets[0], q, 64, N_importance=128, network_fine=nets[1], retraw=True)
        torch.cuda.synchronize()
        assert torch.isfinite(o5["rgb_map"]).all()
    # 6. the shipped configs' network pair in the tensor-core modes: coarse 4x256 (steps 3..6 skipped) + fine 8x256
    c4 = NeRF(D=4, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True).to(dev).requires_grad_(False)
    for mode in ("bf16", "fp16x3"):
        snerf_b200.set_mode(mode)
        with torch.no_grad():
            o6 = render_rays(rb, c4, q, 64, N_importance=128, network_fine=nets[1])
            o7 = render_rays(rb, c4, q, 128, N_importance=0, network_fine=None)
        torch.cuda.synchronize()
        assert torch.isfinite(o6["rgb_map"]).all() and torch.isfinite(o7["rgb_map"]).all()
    snerf_b200.set_mode("fp32")
    print("sanitize_r2: all invocations finished")


if __name__ == "__main__":
    main()
