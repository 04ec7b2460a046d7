"""TEST INFRASTRUCTURE ONLY -- golden vectors for the fused training objective, produced by the reference's OWN loss
classes (s-nerf/model/loss_factory.py: RgbLoss, DepthLoss) and its reduction (model/confidence.py:211-226,
calc_depth_loss: mask, optional confidence, mean) imported from /root/reference on torch-CPU, with torch autograd for
the gradients.  model/loss.py (needed by loss_factory only for SmoothLoss) has absent dependencies: stubbed.

    python oracle/make_golden_loss.py          # writes tests/golden/loss_*.npz (needs /root/reference)

calc_depth_loss itself drags the whole confidence model (`confidence_depends`): its reduction is the three lines
`depth_loss = depth_loss_fn(pred[mask], pred_c[mask], tgt[mask]); depth_loss *= confidence; return depth_loss.mean()`,
restated below around the reference's DepthLoss.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: (N, disparity, with_confidence, depth_lambda, coarse_depth_mult, zero fraction)
    "loss_disparity_conf": (512, True, True, 0.1, 0.2, 0.3),
    "loss_metric_noconf": (333, False, False, 1.0, 0.5, 0.0),
    "loss_disparity_sparse": (1024, True, True, 0.04, 0.2, 0.9),
}


def make_inputs(seed, N, zero_frac):
    rs = np.random.RandomState(seed)
    rgb, tgt = rs.rand(N, 3).astype(np.float32), rs.rand(N, 3).astype(np.float32)
    depth = rs.uniform(2, 100, N).astype(np.float32)
    depth0 = (depth * rs.uniform(0.7, 1.3, N)).astype(np.float32)
    tdepth = (depth * rs.uniform(0.8, 1.25, N)).astype(np.float32)
    tdepth[rs.rand(N) < zero_frac] = 0.0
    tdepth[0] = depth[0]                       # an exact hit: |x| has zero gradient there
    conf = rs.rand(N).astype(np.float32)
    return rgb, tgt, depth, depth0, tdepth, conf


def reference_loss(lf, rgb, tgt, depth, depth0, tdepth, conf, disparity, depth_lambda, c_weight):
    args = types.SimpleNamespace(coarse_depth_mult=c_weight, disparity_depth=disparity)
    rgb_loss_fn, depth_loss_fn = lf.RgbLoss(args), lf.DepthLoss(args)
    img_loss = rgb_loss_fn(rgb, tgt)                                   # train.py:149
    mask = tdepth != 0                                                 # confidence.py:213
    depth_loss = depth_loss_fn(depth[mask], depth0[mask], tdepth[mask])
    if conf is not None:
        depth_loss = depth_loss * conf[mask]                           # confidence.py:221-222
    depth_loss = depth_loss.mean()
    return img_loss + depth_loss * depth_lambda, img_loss, depth_loss  # train.py:209


def main():
    from oracle import ref_import
    ref_import.load()
    import importlib
    # loss_factory.py:1 imports edge_aware_loss_v2 from model/loss.py, whose own imports (imageio, pyquaternion, ...) are
    # absent here; SmoothLoss is not part of this fixture: give loss_factory a stub for that one name
    stub = types.ModuleType("model.loss")
    stub.edge_aware_loss_v2 = None
    sys.modules["model.loss"] = stub
    lf = importlib.import_module("model.loss_factory")
    for i, (name, (N, disparity, with_conf, lam, cw, zf)) in enumerate(CASES.items()):
        arrs = make_inputs(1100 + i, N, zf)
        rgb, tgt, depth, depth0, tdepth, conf = [torch.from_numpy(a) for a in arrs]
        leaves = [rgb, depth, depth0] + ([conf] if with_conf else [])
        for t in leaves:
            t.requires_grad_(True)
        loss, img, dep = reference_loss(lf, rgb, tgt, depth, depth0, tdepth, conf if with_conf else None, disparity, lam, cw)
        (loss * 1.7).backward()                                        # a non-unit upstream gradient
        rec = dict(rgb=arrs[0], target=arrs[1], depth=arrs[2], depth0=arrs[3], target_depth=arrs[4], confidence=arrs[5],
                   disparity=disparity, with_conf=with_conf, depth_lambda=np.float32(lam), coarse_depth_mult=np.float32(cw),
                   upstream=np.float32(1.7), loss=loss.detach().numpy(), img_loss=img.detach().numpy(), depth_loss=dep.detach().numpy(),
                   g_rgb=rgb.grad.numpy(), g_depth=depth.grad.numpy(), g_depth0=depth0.grad.numpy(),
                   torch_version=torch.__version__)
        if with_conf:
            rec["g_conf"] = conf.grad.numpy()
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **rec)
        print(name, float(loss), float(img), float(dep))
    main_proposal(lf)


PROPOSAL_CASES = {
    # name: (N, n_fine, n_coarse, weight)
    "proploss_64_64": (48, 64, 64, 1.0),
    "proploss_128_64": (32, 128, 64, 0.5),       # more fine than coarse intervals (the reference's fine-count clamp is inert)
}


def make_histograms(seed, N, Sf, Sc):
    """Coarse edges over [0, 1]; fine edges strictly inside them (where the reference's gathers are in range), fine weights
    partly above the coarse envelope so that the loss is active."""
    rs = np.random.RandomState(seed)
    sc = np.sort(rs.rand(N, Sc + 1).astype(np.float32), axis=-1)
    sc[:, 0], sc[:, -1] = 0.0, 1.0
    sf = np.sort((0.02 + 0.95 * rs.rand(N, Sf + 1)).astype(np.float32), axis=-1)
    wc = rs.rand(N, Sc).astype(np.float32) ** 3
    wc /= wc.sum(-1, keepdims=True)
    wf = rs.rand(N, Sf).astype(np.float32) ** 3
    wf /= wf.sum(-1, keepdims=True)
    sf[0, 3] = sc[0, 7]                          # a fine edge exactly on a coarse edge (right=True)
    sf[0] = np.sort(sf[0])
    return sf, wf, sc, wc


def main_proposal(lf):
    for i, (name, (N, Sf, Sc, weight)) in enumerate(PROPOSAL_CASES.items()):
        sf, wf, sc, wc = make_histograms(1300 + i, N, Sf, Sc)
        args = types.SimpleNamespace(proposal_lambda=weight)
        t = [torch.from_numpy(a) for a in (sf, wf, sc, wc)]
        t[3].requires_grad_(True)
        loss = lf.ProposalLoss(args)(*t)
        (loss * 1.3).backward()
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), s_vals_f=sf, weights_f=wf, s_vals_c=sc,
                            weights_c=wc, weight=np.float32(weight), upstream=np.float32(1.3), loss=loss.detach().numpy(),
                            g_weights_c=t[3].grad.numpy(), torch_version=torch.__version__)
        print(name, float(loss.detach()))


if __name__ == "__main__":
    main()
