"""TEST INFRASTRUCTURE ONLY -- golden vectors for the proposal resampling step, produced by the reference's OWN
s-nerfpp/zipnerf/internal/stepfun.py imported from /root/reference and run on torch-CPU (it needs only torch + numpy).

    python oracle/make_golden_stepfun.py          # writes tests/golden/stepfun_*.npz (needs /root/reference)

Each fixture is one pass of the sampling loop of Model.forward (internal/models.py:156-213): optional
max_dilate_weights + [1:-1] slices, annealed logits, sample_intervals -- inputs, the jitter draw, and every intermediate.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/s-nerfpp/zipnerf"


def make_rays(seed, N, S, peaky):
    """Sorted normalised distances in [0, 1] and weights that sum to <= 1, like one level's output."""
    rs = np.random.RandomState(seed)
    s = np.sort(rs.rand(N, S + 1).astype(np.float32), axis=-1)
    s[:, 0], s[:, -1] = 0.0, 1.0
    s[1, 5] = s[1, 4]                                   # a zero-width bin (-inf logit)
    s[2, 10:14] = s[2, 10]                              # several
    w = rs.rand(N, S).astype(np.float32) ** (8 if peaky else 1)
    w[3] = 0.0                                          # an empty ray (all weight -> padding)
    w[4, :] = 0.0; w[4, 17] = 1.0                       # one spike
    w = w / np.maximum(w.sum(-1, keepdims=True), 1.0) * rs.uniform(0.3, 1.0, (N, 1)).astype(np.float32)
    return s, w.astype(np.float32)


CASES = {
    # name: (N, S, n_out, dilate, dilation, anneal, randomized, single_jitter, peaky)
    "stepfun_level0_det": (24, 1, 64, False, 0.0, 1.0, False, True, False),        # first level: sdist=[0,1], weights=1
    "stepfun_level1_det": (24, 64, 64, True, 0.0025 + 0.5 / 64, 1.0, False, True, True),
    "stepfun_level2_rand_single": (24, 64, 32, True, 0.0025 + 0.5 / 4096, 0.7, True, True, True),
    "stepfun_level1_rand_indep": (16, 48, 96, True, 0.01, 1.0, True, False, False),
    "stepfun_nodilate_rand": (16, 128, 128, False, 0.0, 1.0, True, True, True),
}


def main():
    sys.path.insert(0, REF)
    from internal import stepfun
    out_dir = os.path.join(ROOT, "tests", "golden")
    for i, (name, (N, S, n, dilate, dilation, anneal, randomized, single, peaky)) in enumerate(CASES.items()):
        if S == 1:
            s = np.tile(np.array([[0.0, 1.0]], np.float32), (N, 1)); w = np.ones((N, 1), np.float32)
        else:
            s, w = make_rays(700 + i, N, S, peaky)
        sdist, weights = torch.from_numpy(s), torch.from_numpy(w)
        domain = (0.0, 1.0)
        rec = dict(sdist=s, weights=w)
        if dilate:
            sdist, weights = stepfun.max_dilate_weights(sdist, weights, dilation, domain=domain, renormalize=True)
            rec.update(t_dilate=sdist.numpy().copy(), w_dilate=weights.numpy().copy())
            sdist, weights = sdist[..., 1:-1], weights[..., 1:-1]
        logits = torch.where(sdist[..., 1:] > sdist[..., :-1], anneal * torch.log(weights + 1e-5),
                             torch.full_like(sdist[..., :-1], -torch.inf))
        jitter = None
        if randomized:
            # `sample` draws torch.rand(t.shape[:-1] + (d,)) (stepfun.py:215): replay the same draw
            torch.manual_seed(900 + i)
            jitter = torch.rand((N, 1 if single else n)).numpy()
            torch.manual_seed(900 + i)
        out = stepfun.sample_intervals(randomized, sdist, logits, n, single_jitter=single, domain=domain)
        if randomized:
            torch.manual_seed(900 + i)
        centers = stepfun.sample(randomized, sdist, logits, n, single, deterministic_center=True)     # what sample_intervals drew
        rec.update(centers=centers.numpy(), t_sampled=sdist.numpy().copy(), logits=logits.numpy(), out=out.numpy(), n=n, dilate=dilate, dilation=np.float32(dilation), anneal=np.float32(anneal),
                   single_jitter=single, torch_version=torch.__version__)
        if jitter is not None:
            rec["jitter"] = jitter
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **rec)
        print(name, out.shape, float(out.min()), float(out.max()))


if __name__ == "__main__":
    main()
