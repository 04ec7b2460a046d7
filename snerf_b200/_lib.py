"""ctypes binding of libsnerf_b200.so (C ABI declared in include/snerf_b200.h).

The library is built in-tree (`snerf_b200/libsnerf_b200.so`) by `build()` -> `make -C csrc`
(nvcc, sm_100a).  There is no CPU fallback: if the library is missing, `load()` raises, and on a
box without an sm_100 GPU every compute entry point returns an error that `check()` raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsnerf_b200.so")
CSRC = os.path.join(_HERE, "csrc")

MODE_FP32 = 0
MODE_BF16 = 1
MODE_FP16 = 2
MODE_TF32 = 3          # snerf_render_rays_bwd only: weight-gradient GEMMs on tcgen05 (tf32 operands)
MODE_FP16X3 = 4        # fp32-class tensor-core arithmetic: fp16 hi/lo operand split, three MMA passes
PACK_FP32_BWD = 16  # snerf_pack_weights mode of the training backward image
PACK_TF32_BWD = 17  # same, weights rounded to tf32 (for the tensor-core backward)
PACK_TF32_FWD = 18  # forward image with tf32-rounded weights (tensor-core training forward)
PACK_BF16_BWD = 19  # backward image of the 16-bit tensor-core training step (fused dX chain)
MAX_TRUNK = 16

_f32p = C.POINTER(C.c_float)


class NetDesc(C.Structure):
    _fields_ = [("D", C.c_int32), ("W", C.c_int32), ("input_ch", C.c_int32), ("input_ch_views", C.c_int32),
                ("skip", C.c_int32), ("use_viewdirs", C.c_int32), ("output_ch", C.c_int32)]


class NetF32(C.Structure):
    _fields_ = [("pts_w", C.c_void_p * MAX_TRUNK), ("pts_b", C.c_void_p * MAX_TRUNK),
                ("views_w", C.c_void_p), ("views_b", C.c_void_p),
                ("feature_w", C.c_void_p), ("feature_b", C.c_void_p),
                ("alpha_w", C.c_void_p), ("alpha_b", C.c_void_p),
                ("rgb_w", C.c_void_p), ("rgb_b", C.c_void_p),
                ("output_w", C.c_void_p), ("output_b", C.c_void_p)]


class NetGradF32(C.Structure):
    _fields_ = NetF32._fields_


class Rays(C.Structure):
    _fields_ = [("ray_batch", C.c_void_p), ("n_rays", C.c_int64), ("width", C.c_int32), ("row_stride", C.c_int32)]


class MipEncode(C.Structure):
    _fields_ = [("rays", C.c_void_p), ("n_rays", C.c_int64), ("n_samples", C.c_int32), ("rows_per_ray", C.c_int32),
                ("s_lin", C.c_void_p), ("s_rand", C.c_void_p), ("s_in", C.c_void_p), ("s_out", C.c_void_p),
                ("transform_idx", C.c_int32), ("max_deg", C.c_int32), ("ray_cone", C.c_int32), ("radius", C.c_float),
                ("enc", C.c_void_p), ("enc_f32", C.c_void_p), ("m_pad", C.c_int64)]


class Linear(C.Structure):
    _fields_ = [("a0", C.c_void_p), ("lda0", C.c_int64), ("k0", C.c_int32),
                ("a1", C.c_void_p), ("lda1", C.c_int64), ("k1", C.c_int32),
                ("w", C.c_void_p), ("n", C.c_int32), ("n_pad", C.c_int32), ("bias", C.c_void_p),
                ("ray_bias", C.c_void_p), ("rows_per_ray", C.c_int32), ("relu", C.c_int32),
                ("out", C.c_void_p), ("ldo", C.c_int64), ("head_w", C.c_void_p), ("n_heads", C.c_int32),
                ("head_out", C.c_void_p), ("head_ld", C.c_int32), ("m_rows", C.c_int64), ("m_pad", C.c_int64)]


class MipComposite(C.Structure):
    _fields_ = [("rays", C.c_void_p), ("n_rays", C.c_int64), ("n_samples", C.c_int32), ("rows_per_ray", C.c_int32),
                ("s_vals", C.c_void_p), ("raw_density", C.c_void_p), ("raw_rgb", C.c_void_p), ("noise", C.c_void_p),
                ("density_head_bias", C.c_float), ("density_bias", C.c_float), ("rgb_padding", C.c_float),
                ("rgb_head_bias", C.c_float * 3), ("transform_idx", C.c_int32), ("white_bkgd", C.c_int32),
                ("comp_rgb", C.c_void_p), ("distance", C.c_void_p), ("acc", C.c_void_p), ("weights", C.c_void_p),
                ("n_fine", C.c_int32), ("u_lin", C.c_void_p), ("u_rand", C.c_void_p), ("resample_padding", C.c_float),
                ("s_new", C.c_void_p)]


class LossOpts(C.Structure):
    _fields_ = [("depth_lambda", C.c_float), ("coarse_depth_mult", C.c_float), ("rgb0_weight", C.c_float), ("disparity", C.c_int32)]


class StepfunOpts(C.Structure):
    _fields_ = [("dilate", C.c_int32), ("renormalize", C.c_int32), ("weights_are_logits", C.c_int32), ("dilation", C.c_float),
                ("domain_lo", C.c_float), ("domain_hi", C.c_float), ("anneal", C.c_float), ("resample_padding", C.c_float),
                ("max_jitter", C.c_float)]


class GridDesc(C.Structure):
    _fields_ = [("D", C.c_int32), ("C", C.c_int32), ("L", C.c_int32), ("H", C.c_int32), ("gridtype", C.c_int32),
                ("align_corners", C.c_int32), ("interp", C.c_int32), ("dtype", C.c_int32), ("S", C.c_float)]


class Opts(C.Structure):
    _fields_ = [("n_samples", C.c_int32), ("n_importance", C.c_int32), ("lindisp", C.c_int32),
                ("white_bkgd", C.c_int32), ("mode", C.c_int32), ("multires", C.c_int32),
                ("multires_views", C.c_int32), ("save_for_backward", C.c_int32),
                ("t_vals", C.c_void_p), ("u_vals", C.c_void_p), ("t_rand", C.c_void_p), ("u_rand", C.c_void_p),
                ("noise0", C.c_void_p), ("noise1", C.c_void_p),
                ("packed_alpha_coarse", C.c_void_p), ("packed_alpha_fine", C.c_void_p), ("camera", C.c_void_p),
                ("desc_fine", C.c_void_p)]


class Camera(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("focal", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("near", C.c_float), ("far", C.c_float), ("c2w", C.c_float * 12), ("first_pixel", C.c_int64)]


OUT_FIELDS = ["rgb_map", "disp_map", "acc_map", "depth_map", "z_vals_map", "weights", "rgb0", "disp0", "acc0",
              "z_std", "raw", "depth0", "z_samples", "z_all", "raw_coarse", "weights_fine"]


class Out(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in OUT_FIELDS]


GRAD_FIELDS = ["rgb_map", "disp_map", "acc_map", "depth_map", "weights", "rgb0", "disp0", "acc0", "depth0", "raw"]


class OutGrad(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in GRAD_FIELDS]


# name -> (restype, argtypes); every symbol include/snerf_b200.h declares
SYMBOLS = {
    "snerf_version": (C.c_int, []),
    "snerf_last_error": (C.c_char_p, []),
    "snerf_device_check": (C.c_int, [C.c_int]),
    "snerf_packed_bytes": (C.c_size_t, [C.POINTER(NetDesc), C.c_int]),
    "snerf_pack_weights": (C.c_int, [C.POINTER(NetDesc), C.POINTER(NetF32), C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "snerf_debug_dw_timing": (C.c_int, [C.POINTER(C.c_int64), C.c_int32]),
    "snerf_debug_dw_cuts": (C.c_int, [C.c_int64, C.c_int64, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                      C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "snerf_pack_weights_batch": (C.c_int, [C.c_int32, C.POINTER(C.POINTER(NetDesc)), C.POINTER(C.POINTER(NetF32)),
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_int32), C.c_void_p]),
    "snerf_query_workspace": (C.c_size_t, [C.POINTER(NetDesc), C.POINTER(Opts), C.c_int64]),
    "snerf_render_rays_fwd": (C.c_int, [C.POINTER(Rays), C.POINTER(NetDesc), C.c_void_p, C.c_void_p, C.POINTER(Opts),
                                        C.POINTER(Out), C.c_void_p, C.c_size_t, C.c_void_p]),
    "snerf_train_workspace_bytes": (C.c_size_t, [C.POINTER(NetDesc), C.c_int32, C.c_int32, C.c_int64]),
    "snerf_train_workspace_bytes_mode": (C.c_size_t, [C.POINTER(NetDesc), C.c_int32, C.c_int32, C.c_int64, C.c_int32]),
    "snerf_train_workspace_bytes_pair": (C.c_size_t, [C.POINTER(NetDesc), C.POINTER(NetDesc), C.c_int32, C.c_int32,
                                                      C.c_int64, C.c_int32]),
    "snerf_render_rays_bwd": (C.c_int, [C.POINTER(Rays), C.POINTER(NetDesc), C.c_void_p, C.c_void_p, C.POINTER(Opts),
                                        C.POINTER(OutGrad), C.POINTER(NetGradF32), C.POINTER(NetGradF32), C.c_void_p,
                                        C.c_size_t, C.c_void_p]),
    "snerf_query_network": (C.c_int, [C.POINTER(NetDesc), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "snerf_nerf_forward": (C.c_int, [C.POINTER(NetDesc), C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int32,
                                     C.c_void_p, C.c_void_p]),
    "snerf_posenc": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "snerf_composite_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "snerf_sample_pdf_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32,
                                       C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "snerf_get_rays": (C.c_int, [C.c_int32, C.c_int32, C.c_float, _f32p, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "snerf_grid_encode_fwd": (C.c_int, [C.POINTER(GridDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "snerf_grid_encode_bwd": (C.c_int, [C.POINTER(GridDesc), C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "snerf_grid_grad_tv": (C.c_int, [C.POINTER(GridDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                     C.c_int64, C.c_void_p]),
    "snerf_grid_encode_ms_fwd": (C.c_int, [C.POINTER(GridDesc), C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p]),
    "snerf_grid_encode_ms_bwd": (C.c_int, [C.POINTER(GridDesc), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_float,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "snerf_grid_level_gain": (C.c_int, [C.POINTER(GridDesc), C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                        C.c_void_p]),
    "snerf_loss_fwd": (C.c_int, [C.POINTER(LossOpts)] + [C.c_void_p] * 7 + [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "snerf_loss_bwd": (C.c_int, [C.POINTER(LossOpts)] + [C.c_void_p] * 7 + [C.c_int64] + [C.c_void_p] * 8),
    "snerf_mip_encode": (C.c_int, [C.POINTER(MipEncode), C.c_void_p]),
    "snerf_linear_tc": (C.c_int, [C.POINTER(Linear), C.c_void_p]),
    "snerf_rows_to_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                     C.c_int64, C.c_void_p]),
    "snerf_mip_cond_bias": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                      C.c_void_p, C.c_void_p]),
    "snerf_mip_composite": (C.c_int, [C.POINTER(MipComposite), C.c_void_p]),
    "snerf_adam_step": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p,
                                  C.c_void_p]),
    "snerf_proposal_loss": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_int32, C.c_int32, C.c_float] + [C.c_void_p] * 4),
    "snerf_stepfun_resample": (C.c_int, [C.POINTER(StepfunOpts), C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                         C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
    "snerf_selftest_umma": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
}

_lib = None
_lock = threading.Lock()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libsnerf_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j8"]
    if force:
        subprocess.run(["make", "-C", CSRC, "clean"], check=True, capture_output=not verbose)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libsnerf_b200.so failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)
    return LIB_PATH


def load():
    """The loaded library (raises if it has not been built -- no silent fallback)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(snerf_b200 has no CPU / PyTorch fallback)")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def last_error() -> str:
    return load().snerf_last_error().decode("utf-8", "replace")


def check(status: int, what: str = "libsnerf_b200"):
    if status != 0:
        raise RuntimeError(f"{what} failed (status {status}): {last_error()}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
