"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's multi-resolution hash-grid encoder
(s-nerfpp/zipnerf/gridencoder/src/gridencoder.cu, wrapper grid.py).  BASELINE configs[3] / SURVEY section 8 row f-2(ii).

Only tests/, __graft_entry__.smoke() and bench tools may import this module; the product path is the CUDA library.

Parity status: pinned by tests/golden/grid_*.npz -- outputs of the reference's OWN kernels (gridencoder.cu compiled
unmodified for sm_100 by oracle/build_ref_gridencoder.py and run on a B200 through oracle/make_golden_grid.py).

Arithmetic notes (what the restatement has to mimic to agree with nvcc's code for the reference):
  * `pos = x * scale + 0.5f` and `acc += w * g` are contracted to FMAs by nvcc (-fmad=true is the default): restated
    as a float64 product-sum rounded once to float32;
  * `scale = exp2f(level * S) * H - 1` uses the device exp2f; for the shipped configs per_level_scale is exactly 2
    (desired 8192 / base 16 over 10 levels; prop grids 512 and 2048 over 6 / 8 levels) so every scale is an integer;
  * the index arithmetic is uint32 with wrap-around, including the early exit of the stride loop (gridencoder.cu:66-84).
"""
import numpy as np

PRIMES = np.array([1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737], dtype=np.uint64)
M32 = np.uint64(0xFFFFFFFF)


def level_layout(input_dim=3, num_levels=16, level_dim=2, per_level_scale=2.0, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, align_corners=False):
    """offsets / resolutions / per_level_scale exactly as GridEncoder.__init__ computes them (grid.py:96-140)."""
    if desired_resolution is not None:
        per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    max_params = 2 ** log2_hashmap_size
    offsets, resolutions, offset = [], [], 0
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        resolution = resolution if align_corners else resolution + 1
        params = min(max_params, resolution ** input_dim)
        params = int(np.ceil(params / 8) * 8)
        resolutions.append(resolution)
        offsets.append(offset)
        offset += params
    offsets.append(offset)
    return np.array(offsets, np.int32), np.array(resolutions, np.int32), float(per_level_scale)


def level_scale(level, S, H):
    """`exp2f(level * S) * H - 1.0f` (gridencoder.cu:137)."""
    e = np.float32(np.exp2(np.float64(np.float32(np.float32(level) * np.float32(S)))))
    return np.float32(np.float64(e) * np.float64(H) - 1.0)


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def grid_index(pos_grid, hashmap_size, resolution, gridtype, align_corners):
    """get_grid_index (gridencoder.cu:66-84) without the `* C + ch`: pos_grid [B, D] uint64 holding uint32 values."""
    B, D = pos_grid.shape
    stride = np.uint64(1)
    index = np.zeros(B, np.uint64)
    d = 0
    while d < D and stride <= np.uint64(hashmap_size):
        index = (index + pos_grid[:, d] * stride) & M32
        stride = (stride * np.uint64(resolution if align_corners else resolution + 1)) & M32
        d += 1
    if gridtype == 0 and stride > np.uint64(hashmap_size):
        index = np.zeros(B, np.uint64)
        for i in range(D):
            index ^= (pos_grid[:, i] * PRIMES[i]) & M32
    return (index % np.uint64(hashmap_size)).astype(np.int64)


def _locate(x, level, S, H, align_corners, interp):
    scale = level_scale(level, S, H)
    resolution = int(np.ceil(scale)) + 1
    pos = _fma(x, np.full_like(x, scale), np.full_like(x, 0.0 if align_corners else 0.5))
    pg = np.floor(pos)
    pos = (pos - pg).astype(np.float32)
    pos_grid = pg.astype(np.int64).astype(np.uint64) & M32
    deriv = np.ones_like(pos)
    if interp == 1:
        deriv = (np.float32(6) * pos * (np.float32(1) - pos)).astype(np.float32)
        pos = (pos * pos * _fma(np.full_like(pos, -2.0), pos, np.full_like(pos, 3.0))).astype(np.float32)
    return scale, resolution, pos, pos_grid, deriv


def grid_encode_forward(inputs, embeddings, offsets, S, H, gridtype=0, align_corners=False, interp=0, calc_dy_dx=False):
    """kernel_grid (gridencoder.cu:87-245).  inputs [B, D] in [0, 1]; embeddings [sO, C]; returns outputs [L, B, C]
    (and dy_dx [B, L, D, C] when requested)."""
    x = np.ascontiguousarray(inputs, np.float32)
    emb = np.asarray(embeddings, np.float32)
    B, D = x.shape
    L, C = len(offsets) - 1, emb.shape[1]
    out = np.zeros((L, B, C), np.float32)
    dy_dx = np.zeros((B, L, D, C), np.float32) if calc_dy_dx else None
    ok = ~np.any((x < 0) | (x > 1), axis=1)
    xi = x[ok]
    for level in range(L):
        grid = emb[offsets[level]:offsets[level + 1]]
        hsize = int(offsets[level + 1] - offsets[level])
        scale, res, pos, pg, deriv = _locate(xi, level, S, H, align_corners, interp)
        acc = np.zeros((xi.shape[0], C), np.float32)
        for idx in range(1 << D):
            w = np.ones(xi.shape[0], np.float32)
            loc = pg.copy()
            for d in range(D):
                if idx & (1 << d):
                    w = (w * pos[:, d]).astype(np.float32)
                    loc[:, d] = (pg[:, d] + np.uint64(1)) & M32
                else:
                    w = (w * (np.float32(1) - pos[:, d])).astype(np.float32)
            g = grid[grid_index(loc, hsize, res, gridtype, align_corners)]
            acc = _fma(np.repeat(w[:, None], C, 1), g, acc)
        out[level, ok] = acc
        if calc_dy_dx:
            for gd in range(D):
                accg = np.zeros((xi.shape[0], C), np.float32)
                for idx in range(1 << (D - 1)):
                    w = np.full(xi.shape[0], scale, np.float32)
                    loc = pg.copy()
                    for nd in range(D - 1):
                        d = nd + 1 if nd >= gd else nd
                        if idx & (1 << nd):
                            w = (w * pos[:, d]).astype(np.float32)
                            loc[:, d] = (pg[:, d] + np.uint64(1)) & M32
                        else:
                            w = (w * (np.float32(1) - pos[:, d])).astype(np.float32)
                    loc[:, gd] = pg[:, gd]
                    left = grid[grid_index(loc, hsize, res, gridtype, align_corners)]
                    loc[:, gd] = (pg[:, gd] + np.uint64(1)) & M32
                    right = grid[grid_index(loc, hsize, res, gridtype, align_corners)]
                    # `acc += w * (r - l) * deriv`: nvcc contracts the last product into the sum (one rounding)
                    term = (w[:, None] * (right - left).astype(np.float32)).astype(np.float32)
                    accg = _fma(term, np.repeat(deriv[:, gd:gd + 1], C, 1), accg)
                tmp = dy_dx[:, level, gd]
                tmp[ok] = accg
                dy_dx[:, level, gd] = tmp
    return (out, dy_dx) if calc_dy_dx else out


def grid_encode_backward(grad, inputs, embeddings_shape, offsets, S, H, gridtype=0, align_corners=False, interp=0,
                         dy_dx=None):
    """kernel_grid_backward + kernel_input_backward (gridencoder.cu:248-369).  grad [L, B, C] -> grad_embeddings
    [sO, C] (float64 accumulation: the reference's atomics have no defined order) and grad_inputs [B, D] or None."""
    x = np.ascontiguousarray(inputs, np.float32)
    g = np.asarray(grad, np.float32)
    B, D = x.shape
    L, C = len(offsets) - 1, embeddings_shape[1]
    ge = np.zeros(embeddings_shape, np.float64)
    ok = ~np.any((x < 0) | (x > 1), axis=1)
    xi = x[ok]
    for level in range(L):
        hsize = int(offsets[level + 1] - offsets[level])
        scale, res, pos, pg, _ = _locate(xi, level, S, H, align_corners, interp)
        gl = g[level][ok]
        for idx in range(1 << D):
            w = np.ones(xi.shape[0], np.float32)
            loc = pg.copy()
            for d in range(D):
                if idx & (1 << d):
                    w = (w * pos[:, d]).astype(np.float32)
                    loc[:, d] = (pg[:, d] + np.uint64(1)) & M32
                else:
                    w = (w * (np.float32(1) - pos[:, d])).astype(np.float32)
            index = grid_index(loc, hsize, res, gridtype, align_corners) + int(offsets[level])
            np.add.at(ge, index, (w[:, None] * gl).astype(np.float32).astype(np.float64))
    gi = None
    if dy_dx is not None:
        # result += grad[l, b, ch] * dy_dx[b, l, d, ch], l outer, ch inner, fp32 (gridencoder.cu:358-364)
        gi = np.zeros((B, D), np.float32)
        for level in range(L):
            for ch in range(C):
                gi = _fma(np.repeat(g[level, :, ch:ch + 1], D, 1), dy_dx[:, level, :, ch], gi)
    return ge.astype(np.float32), gi


def grad_total_variation(inputs, embeddings, offsets, weight, S, H, gridtype=0, align_corners=False):
    """kernel_grad_tv (gridencoder.cu:506-610): returns the increment added to embeddings.grad (float64 accumulate)."""
    x = np.ascontiguousarray(inputs, np.float32)
    emb = np.asarray(embeddings, np.float32)
    B, D = x.shape
    L, C = len(offsets) - 1, emb.shape[1]
    out = np.zeros(emb.shape, np.float64)
    ok = ~np.any((x < 0) | (x > 1), axis=1)
    xi = x[ok]
    wgt = np.float32(np.float32(weight) / np.float32(2 * D))
    for level in range(L):
        grid = emb[offsets[level]:offsets[level + 1]]
        hsize = int(offsets[level + 1] - offsets[level])
        scale, res, _, pg, _ = _locate(xi, level, S, H, align_corners, 0)
        index = grid_index(pg, hsize, res, gridtype, align_corners)
        results = np.zeros((xi.shape[0], C), np.float32)
        idelta = np.zeros((xi.shape[0], C), np.float32)
        for d in range(D):
            cur = pg[:, d].astype(np.int64)
            for side, valid in ((1, cur < res), (-1, cur > 0)):
                loc = pg.copy()
                loc[:, d] = (cur + side).astype(np.uint64) & M32
                other = grid[grid_index(loc, hsize, res, gridtype, align_corners)]
                gv = (grid[index] - other).astype(np.float32)
                gv = np.where(valid[:, None], gv, np.float32(0))
                results = (results + gv).astype(np.float32)
                idelta = _fma(gv, gv, idelta)
        val = (wgt * results).astype(np.float32) * (np.float32(1) / np.sqrt(idelta + np.float32(1e-9))).astype(np.float32)
        np.add.at(out, index + int(offsets[level]), val.astype(np.float64))
    return out.astype(np.float32)


class GridEncoderOracle:
    """GridEncoder.forward (grid.py:155-176): map [-bound, bound] -> [0, 1], encode, return [..., L * C]."""

    def __init__(self, embeddings, **cfg):
        self.cfg = dict(input_dim=3, num_levels=16, level_dim=2, per_level_scale=2.0, base_resolution=16,
                        log2_hashmap_size=19, desired_resolution=None, align_corners=False)
        self.gridtype = {"hash": 0, "tiled": 1}[cfg.pop("gridtype", "hash")]
        self.interp = {"linear": 0, "smoothstep": 1}[cfg.pop("interpolation", "linear")]
        self.cfg.update(cfg)
        self.offsets, self.grid_sizes, self.per_level_scale = level_layout(**self.cfg)
        self.S = np.log2(self.per_level_scale)
        self.embeddings = np.asarray(embeddings, np.float32)

    def __call__(self, inputs, bound=1):
        x = ((np.asarray(inputs, np.float32) + np.float32(bound)) / np.float32(2 * bound)).astype(np.float32)
        prefix = x.shape[:-1]
        out = grid_encode_forward(x.reshape(-1, x.shape[-1]), self.embeddings, self.offsets, self.S,
                                  self.cfg["base_resolution"], self.gridtype, self.cfg["align_corners"], self.interp)
        L, B, C = out.shape
        return out.transpose(1, 0, 2).reshape(*prefix, L * C)

    # ---- zip-NeRF multisample featurisation (s-nerfpp/zipnerf/internal/models.py:481-507, MLP.predict_density)
    def multisample_weights(self, stds):
        """`erf(1 / sqrt(8 stds^2 grid_sizes^2))` (models.py:493) with torch's fp32 operation order; [..., M] -> [..., M, L]."""
        from scipy.special import erf
        s = np.asarray(stds, np.float32)[..., None]
        g2 = (self.grid_sizes.astype(np.int32) ** 2).astype(np.float32)
        a = ((np.float32(8) * (s * s).astype(np.float32)).astype(np.float32) * g2).astype(np.float32)
        r = (np.float32(1) / np.sqrt(a).astype(np.float32)).astype(np.float32)
        return erf(r.astype(np.float64)).astype(np.float32)

    def level_gain(self, init_std=1e-4):
        """`(init_std^2 + segment mean of |embedding|^2).sqrt()` per level (models.py:496-503)."""
        sq = (self.embeddings.astype(np.float64) ** 2).sum(-1)
        mean = np.array([sq[self.offsets[l]:self.offsets[l + 1]].mean() for l in range(len(self.offsets) - 1)])
        return np.sqrt(np.float32(np.float32(init_std) ** 2) + mean.astype(np.float32)).astype(np.float32)

    def encode_multisample(self, means, stds, bound=1, init_std=1e-4, scale_featurization=True):
        """features = mean_m(encoder(means)[..., m, l, c] * w[..., m, l]); optionally cat (2 mean_m(w) - 1) * level_gain.
        means [..., M, 3] in [-bound, bound], stds [..., M] -> [..., L*C (+ L)]."""
        means = np.asarray(means, np.float32)
        L = len(self.offsets) - 1
        feats = self(means, bound).reshape(*means.shape[:-1], L, -1)                    # [..., M, L, C]
        w = self.multisample_weights(stds)                                              # [..., M, L]
        prod = (feats * w[..., None]).astype(np.float32)
        out = prod.astype(np.float64).mean(axis=-3).astype(np.float32).reshape(*means.shape[:-2], -1)
        if scale_featurization:
            fw = ((np.float32(2) * w.astype(np.float64).mean(axis=-2).astype(np.float32) - np.float32(1))
                  * self.level_gain(init_std)).astype(np.float32)
            out = np.concatenate([out, fw], axis=-1)
        return out

    def encode_multisample_backward(self, grad, means, stds, bound=1):
        """d loss / d embeddings for encode_multisample: grad [N, >= L*C] (featurized_w columns carry no table gradient)."""
        means = np.asarray(means, np.float32)
        N, M, _ = means.shape
        L = len(self.offsets) - 1
        C = self.embeddings.shape[1]
        w = self.multisample_weights(stds)                                              # [N, M, L]
        g = np.asarray(grad, np.float32)[:, :L * C].reshape(N, 1, L, C)
        gf = ((g / np.float32(M)).astype(np.float32) * w[..., None]).astype(np.float32)  # [N, M, L, C]
        x = ((means + np.float32(bound)) / np.float32(2 * bound)).astype(np.float32).reshape(-1, 3)
        ge, _ = grid_encode_backward(gf.reshape(N * M, L, C).transpose(1, 0, 2), x, self.embeddings.shape, self.offsets, self.S,
                                     self.cfg["base_resolution"], self.gridtype, self.cfg["align_corners"], self.interp)
        return ge
