"""snerf_b200 -- B200 (sm_100a) volumetric renderer behind S-NeRF's `render_rays` interface.

Mirrors `s-nerf/model/render.py` and `s-nerf/model/run_nerf_helpers.py` of the reference
(same callables, dict keys and state_dict names) on top of libsnerf_b200.so (C ABI:
include/snerf_b200.h).  Import style matches the reference:

    from snerf_b200 import render, run_nerf_helpers
    from snerf_b200.render import render, render_rays, create_nerf

Next to the renderer: `losses` (RgbLoss + DepthLoss / calc_depth_loss, ProposalLoss of s-nerf/model/loss_factory.py as
fused kernels), `gridencoder` (zip-NeRF's GridEncoder + fused multisample featurisation) and `stepfun` (zip-NeRF's
proposal resampling: max_dilate_weights, sample_intervals, resample_intervals); `optim` (one-kernel Adam over flat
parameter / gradient buffers, whole training step as one CUDA graph) and `parallel` (ray sharding, flat gradient all-reduce).
"""
from . import _lib, gridencoder, losses, models, optim, parallel, render, run_nerf_helpers, stepfun  # noqa: F401
from .render import batchify_rays, create_nerf, make_query_fn, render_path, render_rays  # noqa: F401
from .run_nerf_helpers import (NeRF, get_embedder, get_mode, get_rays, raw2outputs, run_network,  # noqa: F401
                               sample_pdf, set_mode, set_train_precision, get_train_precision)

__version__ = "0.1.0"
