/*
 * snerf_b200.h -- C ABI of libsnerf_b200.so: the B200 (sm_100a) volumetric renderer
 * for S-NeRF's `render_rays` hot path.
 *
 * The reference has no FFI for this path: its boundary is Python callables
 * (SURVEY.md section 8b).  Each entry point below names the reference callable it
 * replaces (paths relative to /root/reference/s-nerf/model/); the Python mirror in
 * snerf_b200/ keeps the reference signatures and calls these through ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless noted;
 *     all float tensors are fp32, row-major, contiguous in their last dimension;
 *   - the library never allocates or frees device memory: buffers (and the packed
 *     weight images) are owned by the caller;
 *   - every launch goes to the `stream` argument (a cudaStream_t passed as void*);
 *   - return value 0 = success, negative = SnerfStatus; snerf_last_error() gives the
 *     message of the last failure on the calling thread.  Asynchronous CUDA faults
 *     surface at the caller's next synchronisation, as with any CUDA library;
 *   - there is no CPU fallback: on a box without an sm_100 GPU every compute entry
 *     point returns SNERF_ERR_ARCH / SNERF_ERR_CUDA.
 */
#ifndef SNERF_B200_H_
#define SNERF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNERF_ABI_VERSION 11
#define SNERF_MAX_TRUNK_LAYERS 16

typedef enum SnerfStatus {
  SNERF_OK = 0,
  SNERF_ERR_BAD_ARG = -1,      /* null pointer, bad shape, misaligned buffer        */
  SNERF_ERR_UNSUPPORTED = -2,  /* configuration outside what the selected mode runs */
  SNERF_ERR_ARCH = -3,         /* device is not sm_100                              */
  SNERF_ERR_CUDA = -4,         /* a CUDA runtime call failed (message has details)  */
  SNERF_ERR_WORKSPACE = -5     /* workspace too small                               */
} SnerfStatus;

/* Arithmetic mode of the MLP (everything outside the MLP is fp32 in both modes). */
typedef enum SnerfMode {
  SNERF_MODE_FP32 = 0, /* FFMA on CUDA cores, fp32 activations: reference-accurate   */
  SNERF_MODE_BF16 = 1, /* tcgen05 tensor cores, bf16 operands, fp32 accumulate (TMEM) */
  SNERF_MODE_FP16 = 2, /* same kernel with fp16 operands: 10-bit mantissa (8x tighter than bf16) at the same rate;
                          operands must stay inside fp16 range (|x| < 65504), true for NeRF-style MLPs */
  SNERF_MODE_TF32 = 3, /* training only (save_for_backward forward + snerf_render_rays_bwd): every MLP GEMM of the step
                          (forward layers, dX chain, weight gradients) on tcgen05 with tf32 operands fetched by TMA from
                          fp32 activation stores, fp32 accumulation; sampling / compositing as SNERF_MODE_FP32 */
  SNERF_MODE_FP16X3 = 4 /* fp32-class arithmetic on the tensor cores: the SNERF_MODE_FP16 kernel with every operand split into
                          fp16 hi + fp16 lo parts and three tcgen05.mma passes per k-block (hi*hi + lo*hi + hi*lo, fp32
                          accumulation; dropped term 2^-22 relative).  Meets the 1e-4 parity bar of SNERF_MODE_FP32 at
                          ~10x its rate; same fp16 range requirement as SNERF_MODE_FP16 */
} SnerfMode;
/* Extra value of the `mode` argument of snerf_packed_bytes / snerf_pack_weights: the image the training backward
 * kernel streams (un-transposed fp32 weight blocks + its step table). */
#define SNERF_PACK_FP32_BWD 16
/* Same image with the streamed weights rounded to the nearest tf32 value: pass this one to snerf_render_rays_bwd when
 * opts->mode is SNERF_MODE_TF32 (the tensor core truncates fp32 operands; pre-rounded operands avoid that bias). */
#define SNERF_PACK_TF32_BWD 17
/* Backward image of the tensor-core training step (opts->mode SNERF_MODE_BF16 with save_for_backward):
 * the transposed 16-bit weight chunks the fused dX-chain kernel streams + alpha_linear / rgb_linear in fp32.  The forward of
 * that step uses the ordinary SNERF_MODE_BF16 image. */
#define SNERF_PACK_BF16_BWD 19
/* The forward (fp32-layout) image with tf32-rounded weights, for snerf_render_rays_fwd with opts->mode = SNERF_MODE_TF32. */
#define SNERF_PACK_TF32_FWD 18

/* Architecture of one `NeRF` module (run_nerf_helpers.py:75-101). */
typedef struct SnerfNetDesc {
  int32_t D;               /* trunk depth (pts_linears), <= SNERF_MAX_TRUNK_LAYERS   */
  int32_t W;               /* trunk width: 64, 128 or 256                            */
  int32_t input_ch;        /* encoded point width, 3 + 6*multires (<= 63)            */
  int32_t input_ch_views;  /* encoded direction width, 3 + 6*multires_views (<= 27)  */
  int32_t skip;            /* trunk layer index after which [enc, h] is concatenated, -1 = none */
  int32_t use_viewdirs;    /* 1: alpha/feature/views/rgb heads; 0: output_linear     */
  int32_t output_ch;       /* width of output_linear when use_viewdirs == 0 (4 or 5) */
} SnerfNetDesc;

/* fp32 parameters of one `NeRF` module exactly as its state_dict holds them
 * (weight[out,in] row-major, bias[out]).  Unused heads are NULL; alpha_w == NULL with use_viewdirs packs the
 * NeRF_RGB variant (no alpha head: sigma comes from a separate frozen network, see SnerfOpts). */
typedef struct SnerfNetF32 {
  const float* pts_w[SNERF_MAX_TRUNK_LAYERS];
  const float* pts_b[SNERF_MAX_TRUNK_LAYERS];
  const float* views_w;   const float* views_b;    /* views_linears.0  [W/2, W+input_ch_views] */
  const float* feature_w; const float* feature_b;  /* feature_linear   [W, W]                  */
  const float* alpha_w;   const float* alpha_b;    /* alpha_linear     [1, W]                  */
  const float* rgb_w;     const float* rgb_b;      /* rgb_linear       [3, W/2]                */
  const float* output_w;  const float* output_b;   /* output_linear    [output_ch, W]          */
} SnerfNetF32;

/* A batch of rays in the reference's `ray_batch` layout (render.py:71-79,324-328):
 * columns 0:3 origin, 3:6 direction (un-normalised), 6 near, 7 far, [8 depth],
 * last three = unit view direction iff width > 9. */
typedef struct SnerfRays {
  const float* ray_batch;  /* [n_rays, width] with row pitch `row_stride` floats */
  int64_t n_rays;
  int32_t width;           /* 8, 9, 11 or 12 */
  int32_t row_stride;
} SnerfRays;

/* Options of render_rays (render.py:281-293) plus the positional-encoding sizes that
 * the reference bakes into network_query_fn (render.py:168-173,215-218). */
struct SnerfCamera;
typedef struct SnerfOpts {
  int32_t n_samples;      /* N_samples (coarse), 2..256                            */
  int32_t n_importance;   /* N_importance (fine); n_samples+n_importance <= 256    */
  int32_t lindisp;
  int32_t white_bkgd;
  int32_t mode;           /* SnerfMode */
  int32_t multires;       /* 10  (-1 = identity embedder, i_embed=-1)              */
  int32_t multires_views; /* 4                                                     */
  int32_t save_for_backward; /* 1: training forward (fp32 mode): keep what snerf_render_rays_bwd needs in the workspace */
  const float* t_vals;    /* [n_samples]  torch.linspace(0,1,n_samples), required  */
  const float* u_vals;    /* [n_importance] deterministic u (perturb == 0)         */
  const float* t_rand;    /* [n_rays, n_samples] stratified jitter, NULL = none    */
  const float* u_rand;    /* [n_rays, n_importance] random u, NULL = use u_vals    */
  const float* noise0;    /* [n_rays, n_samples] sigma noise (already * raw_noise_std), NULL = none */
  const float* noise1;    /* [n_rays, n_samples+n_importance] likewise for the fine pass */
  /* NeRF_RGB (run_nerf_helpers.py:157-212): packed images (fp32 mode) of the frozen `alpha_model` whose sigma
   * replaces the missing alpha head of the coarse / fine network; NULL = the network has its own alpha_linear. */
  const void* packed_alpha_coarse;
  const void* packed_alpha_fine;
  /* Optional pinhole camera (HOST pointer): with it the kernel builds its rays itself -- get_rays
   * (run_nerf_helpers.py:247-258) and the view-direction normalisation of render() (render.py:56-63) run in the
   * renderer's prologue and the ONLY ray input is this struct.  rays->ray_batch must then be NULL and rays->n_rays is the
   * number of consecutive pixels rendered, starting at camera->first_pixel (row-major over H x W); rays->width = 11.
   * Inference only. */
  const struct SnerfCamera* camera;
  /* Architecture of the FINE network when it differs from `desc` (the coarse one): create_nerf builds the two from
   * netdepth/netwidth and netdepth_fine/netwidth_fine (render.py:176-201; the shipped configs set netdepth = 4 against
   * netdepth_fine = 8).  D, W and skip may differ; input_ch, input_ch_views, use_viewdirs, output_ch must agree.
   * NULL = same as `desc`.  fp32 mode / fp32 train precision; the tensor-core inference modes take 8x256 pairs and the
   * pair coarse NeRF(D=4, W=256, no live skip) + fine NeRF(D=8, W=256, skips=[4]) of the shipped configs. */
  const SnerfNetDesc* desc_fine;
} SnerfOpts;

typedef struct SnerfCamera {
  int32_t H, W;
  float focal, cx, cy;   /* pixel (i, j) -> ((i + .5 - cx) / focal, -(j + .5 - cy) / focal, -1) rotated by c2w[:3,:3] */
  float near, far;
  float c2w[12];         /* [3, 4] row-major */
  int64_t first_pixel;
} SnerfCamera;

/* Outputs = the dict render_rays returns (render.py:394-401).  Any pointer may be
 * NULL to skip that output.  `weights` / `z_vals_map` are the COARSE ones, as in
 * the reference.  The trailing members are extra intermediates for tests. */
typedef struct SnerfOut {
  float* rgb_map;    /* [N,3] */
  float* disp_map;   /* [N]   */
  float* acc_map;    /* [N]   */
  float* depth_map;  /* [N]   */
  float* z_vals_map; /* [N,n_samples] */
  float* weights;    /* [N,n_samples] */
  float* rgb0;       /* [N,3] (n_importance > 0) */
  float* disp0;      /* [N]   */
  float* acc0;       /* [N]   */
  float* z_std;      /* [N]   */
  float* raw;        /* [N,S,4], S = n_samples + n_importance (retraw=True)        */
  float* depth0;     /* [N]   coarse depth (DepthLoss uses it; not in the dict)    */
  float* z_samples;  /* [N,n_importance] */
  float* z_all;      /* [N,S] sorted union */
  float* raw_coarse; /* [N,n_samples,4] */
  float* weights_fine; /* [N,S] */
} SnerfOut;

/* Training: upstream gradients dL/d(output) of the differentiable outputs (same shapes as SnerfOut; NULL = zero).
 * z_vals_map / z_std carry no gradient: the resampled depths are detached (render.py:381). */
typedef struct SnerfOutGrad {
  const float* rgb_map;   /* [N,3] */
  const float* disp_map;  /* [N]   */
  const float* acc_map;   /* [N]   */
  const float* depth_map; /* [N]   */
  const float* weights;   /* [N,n_samples] (coarse weights) */
  const float* rgb0;      /* [N,3] */
  const float* disp0;     /* [N]   */
  const float* acc0;      /* [N]   */
  const float* depth0;    /* [N]   */
  const float* raw;       /* [N,S,4] */
} SnerfOutGrad;

/* Gradient buffers of one `NeRF` module, member for member like SnerfNetF32 (weight[out,in] row-major).
 * The backward ACCUMULATES into them (+=), as autograd does into .grad; NULL members are skipped. */
typedef struct SnerfNetGradF32 {
  float* pts_w[SNERF_MAX_TRUNK_LAYERS];
  float* pts_b[SNERF_MAX_TRUNK_LAYERS];
  float* views_w;   float* views_b;
  float* feature_w; float* feature_b;
  float* alpha_w;   float* alpha_b;
  float* rgb_w;     float* rgb_b;
  float* output_w;  float* output_b;
} SnerfNetGradF32;

/* ---- library ------------------------------------------------------------------ */
int snerf_version(void);
const char* snerf_last_error(void);
/* 0 if device `dev` is an sm_100 part this library can run on. */
int snerf_device_check(int dev);

/* ---- weights -------------------------------------------------------------------
 * The kernels read a packed image of the network (fp32: K-major transposed tiles
 * streamed by the bulk-copy engine; bf16: 128B-swizzled UMMA operand tiles in
 * consumption order).  The fp32 master copy stays in the caller's nn.Linear tensors
 * (same state_dict names as the reference, render.py:245-247); re-pack after every
 * optimizer step.  Replaces: NeRF.__init__ / state_dict (run_nerf_helpers.py:75-101). */
size_t snerf_packed_bytes(const SnerfNetDesc* desc, int mode);
int snerf_pack_weights(const SnerfNetDesc* desc, const SnerfNetF32* src, void* packed,
                       size_t packed_bytes, int mode, void* stream);

/* Several images in one call (the training step re-packs the forward and the backward image of both networks every
 * iteration): the tensor-core images (SNERF_MODE_BF16 / FP16 / FP16X3, SNERF_PACK_BF16_BWD) of all items are written by ONE
 * kernel launch; items of any other mode go through snerf_pack_weights.  Same arguments as snerf_pack_weights, item-wise. */
int snerf_pack_weights_batch(int32_t n, const SnerfNetDesc* const* descs, const SnerfNetF32* const* srcs,
                             void* const* packed, const size_t* packed_bytes, const int32_t* modes, void* stream);

/* ---- the hot path ---------------------------------------------------------------
 * render_rays (render.py:281-409): stratified sampling -> encode -> coarse MLP ->
 * composite -> inverse-CDF resampling -> sort -> fine MLP -> composite, ONE kernel.
 * packed_fine may be NULL (then the coarse network serves both passes, render.py:387). */
size_t snerf_query_workspace(const SnerfNetDesc* desc, const SnerfOpts* opts, int64_t n_rays);
int snerf_render_rays_fwd(const SnerfRays* rays, const SnerfNetDesc* desc,
                          const void* packed_coarse, const void* packed_fine,
                          const SnerfOpts* opts, const SnerfOut* out,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- training ---------------------------------------------------------------------
 * Replaces torch autograd over render_rays (the reference's train step, train.py + render.py:281-409):
 *   1. snerf_render_rays_fwd with opts->save_for_backward = 1, mode fp32 and a workspace of
 *      snerf_train_workspace_bytes(): same outputs as inference, plus every layer's activations in the workspace;
 *   2. the caller evaluates its loss on the outputs (any torch code) and obtains dL/d(outputs);
 *   3. snerf_render_rays_bwd with the same rays / opts (opts->mode may be SNERF_MODE_FP32 or SNERF_MODE_TF32 here) /
 *      workspace and the SNERF_PACK_FP32_BWD images of the networks: accumulates dL/d(parameter) into grad_coarse / grad_fine (grad_fine NULL when packed_bwd_fine is
 *      NULL, i.e. one network serves both passes).
 * Supported at the fp32 level: every network the forward accepts -- view-dependent heads or output_linear, NeRF_RGB with
 * its frozen alpha_model (opts->packed_alpha_*), a fine architecture of its own (opts->desc_fine); W in {64,128,256}.
 * The resampled depths are not differentiated (z_samples.detach(), render.py:381), nor are the rays.
 *
 * Tensor-core training step (opts->mode = SNERF_MODE_BF16 in BOTH calls; NeRF 8x256, skips=[4], view directions, the
 * sample counts the tensor-core renderer is built for): the forward is the fused inference kernel that additionally
 * writes every layer's bf16 output to the workspace (packed images: the ordinary SNERF_MODE_BF16 ones); the backward is
 * one fused dX-chain kernel + one grouped weight-gradient GEMM over bf16 activation / gradient stores with fp32
 * accumulation, fp32 parameter gradients (packed images: SNERF_PACK_BF16_BWD).  Workspace size:
 * snerf_train_workspace_bytes_mode(..., mode). */
size_t snerf_train_workspace_bytes(const SnerfNetDesc* desc, int32_t n_samples, int32_t n_importance, int64_t n_rays);
size_t snerf_train_workspace_bytes_mode(const SnerfNetDesc* desc, int32_t n_samples, int32_t n_importance, int64_t n_rays,
                                        int32_t mode);
/* (desc_fine: architecture of the fine network, NULL = same as desc; see SnerfOpts.desc_fine) */
size_t snerf_train_workspace_bytes_pair(const SnerfNetDesc* desc, const SnerfNetDesc* desc_fine, int32_t n_samples,
                                        int32_t n_importance, int64_t n_rays, int32_t mode);
int snerf_render_rays_bwd(const SnerfRays* rays, const SnerfNetDesc* desc,
                          const void* packed_bwd_coarse, const void* packed_bwd_fine,
                          const SnerfOpts* opts, const SnerfOutGrad* grad_out,
                          const SnerfNetGradF32* grad_coarse, const SnerfNetGradF32* grad_fine,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Debug aid of tools/dw_balance.py: per-CTA cycle counts of the last weight-gradient launch made with SNERF_DW_TIMING=1 in
 * the environment; out_host[cta][4] = total cycles, cycles spent flushing accumulators, flushes, 8 KiB units streamed. */
int snerf_debug_dw_timing(int64_t* out_host, int32_t n_cta);
/* Host-only (no GPU needed): the per-SM cuts the weight-gradient launch would use for rows_c coarse / rows_f fine rows of the
 * stores (multiples of 64) on n_cta SMs.  out_cut[n_cta + 1]: positions in the linearised (problem, 64-row block) space;
 * out_first[29]: first block of each problem, then the total; out_stream_units[28]: 64-channel column blocks a 64-row block of
 * each problem streams; *makespan: modelled cost of the slowest SM.  Returns the number of problems (28) or a negative status. */
int snerf_debug_dw_cuts(int64_t rows_c, int64_t rows_f, int32_t n_cta, int64_t* out_cut, int64_t* out_first,
                        double* out_stream_units, double* makespan);

/* ---- stage entry points (also used on their own by the Python mirror) ------------ */

/* network_query_fn / run_network (run_nerf_helpers.py:460-474): encode pts[N,S,3]
 * (+ per-ray viewdirs[N,3] broadcast over S) and run the MLP -> raw[N,S,4]. */
int snerf_query_network(const SnerfNetDesc* desc, const void* packed, int mode,
                        int multires, int multires_views,
                        const float* pts, const float* viewdirs, int64_t n_rays, int32_t n_samples,
                        float* raw, void* stream);

/* NeRF.forward (run_nerf_helpers.py:103-126) on already-encoded rows
 * x[M, input_ch + input_ch_views] -> out[M, 4] (rgb, sigma). */
int snerf_nerf_forward(const SnerfNetDesc* desc, const void* packed, int mode,
                       const float* x, int64_t n_rows, int32_t row_stride, float* out, void* stream);

/* Embedder.embed (run_nerf_helpers.py:22-52): x[M,3] -> [M, 3+6*n_freqs]. */
int snerf_posenc(const float* x, int64_t n_rows, int32_t n_freqs, float* out, void* stream);

/* raw2outputs (run_nerf_helpers.py:381-424).  noise may be NULL. */
int snerf_composite_fwd(const float* raw, const float* z_vals, const float* rays_d, const float* noise,
                        int64_t n_rays, int32_t n_samples, int32_t white_bkgd,
                        float* rgb_map, float* disp_map, float* acc_map, float* weights,
                        float* depth_map, void* stream);

/* sample_pdf (run_nerf_helpers.py:336-379).  bins[N,B], weights[N,B-1].
 * u: [n_out] shared by all rays (u_per_ray=0) or [N,n_out] (u_per_ray=1).
 * cdf_in (optional, [N,B]): use this cdf instead of deriving it from `weights`
 * (bit-exactness test given an identical cdf).  Outputs: samples[N,n_out],
 * inds int64 [N,n_out] (= torch.searchsorted(cdf,u,right=True)), cdf_out[N,B]; any may be NULL. */
int snerf_sample_pdf_fwd(const float* bins, const float* weights, const float* cdf_in,
                         const float* u, int32_t u_per_ray, int64_t n_rays, int32_t n_bins,
                         int32_t n_out, float* samples, int64_t* inds, float* cdf_out, void* stream);

/* get_rays (run_nerf_helpers.py:247-258): pinhole rays for an H x W image.
 * c2w: HOST pointer to 12 floats (3x4 row-major).  rays_o / rays_d: [H*W,3]. */
int snerf_get_rays(int32_t H, int32_t W, float focal, const float* c2w_host, float cx, float cy,
                   float* rays_o, float* rays_d, void* stream);

/* ---- multi-resolution hash-grid encoder (BASELINE configs[3], zip-NeRF path) ---------------------------------
 * Replaces the reference's torch extension s-nerfpp/zipnerf/gridencoder (src/gridencoder.h:8-15, gridencoder.cu):
 * same operands and semantics as grid_encode_forward / grid_encode_backward / grad_total_variation, as plain
 * pointers.  `dtype` 0 = fp32, 1 = fp16 embeddings / outputs / gradients (inputs are always fp32, as in the
 * reference).  S = log2(per_level_scale), H = base_resolution (grid.py:36-37).  offsets: int32 [L+1], DEVICE.
 *
 * Layout freedom the reference does not have: outputs / grad are addressed as
 *     element(level l, point b, channel c) = base[l * stride_l + b * stride_b + c]        (strides in elements)
 * so the kernels read and write the [B, L*C] tensor the network consumes (stride_l = C, stride_b = L*C) directly;
 * the reference's own layout [L, B, C] (gridencoder.cu:379-383) is stride_l = B*C, stride_b = C. */
typedef struct SnerfGridDesc {
  int32_t D;             /* input_dim: 2, 3 or 4                              */
  int32_t C;             /* level_dim: 1, 2, 4 or 8                           */
  int32_t L;             /* num_levels                                        */
  int32_t H;             /* base_resolution                                   */
  int32_t gridtype;      /* 0 = hash, 1 = tiled        (grid.py:14-17)        */
  int32_t align_corners;
  int32_t interp;        /* 0 = linear, 1 = smoothstep (grid.py:19-22)        */
  int32_t dtype;         /* 0 = fp32, 1 = fp16                                */
  float S;               /* log2(per_level_scale)                             */
} SnerfGridDesc;

/* grid_encode_forward (gridencoder.cu:448-474): inputs[B,D] in [0,1] -> outputs (layout above); dy_dx (optional,
 * [B, L*D*C]) receives d(outputs)/d(inputs) for the input gradient. */
int snerf_grid_encode_fwd(const SnerfGridDesc* desc, const float* inputs, const void* embeddings,
                          const int32_t* offsets, void* outputs, int64_t out_stride_l, int64_t out_stride_b,
                          void* dy_dx, int64_t n_points, void* stream);
/* grid_encode_backward (gridencoder.cu:476-504): ACCUMULATES into grad_embeddings[sO,C] (caller zero-fills, as
 * grid.py:77 does); grad_inputs[B,D] (optional, with dy_dx) is overwritten. */
int snerf_grid_encode_bwd(const SnerfGridDesc* desc, const void* grad, int64_t grad_stride_l, int64_t grad_stride_b,
                          const float* inputs, const int32_t* offsets, void* grad_embeddings,
                          const void* dy_dx, void* grad_inputs, int64_t n_points, void* stream);
/* grad_total_variation (gridencoder.cu:630-644): adds the TV gradient at the cells of `inputs` to `grad` (fp32). */
int snerf_grid_grad_tv(const SnerfGridDesc* desc, const float* inputs, const void* embeddings, void* grad,
                       const int32_t* offsets, float weight, int64_t n_points, void* stream);

/* zip-NeRF multisample featurisation fused with the encoder (s-nerfpp/zipnerf/internal/models.py:481-507,
 * MLP.predict_density): for N samples of M multisample points each (means [N,M,3] in [-bound, bound], stds [N,M])
 *     out[n, l*C + c] = mean_m( encoder((means + bound) / (2 bound))[n,m,l,c] * w[n,m,l] ),
 *     w[n,m,l]        = erf(1 / sqrt(8 stds[n,m]^2 grid_sizes[l]^2)),
 *     out[n, L*C + l] = (2 mean_m(w) - 1) * level_gain[l]          (only when level_gain != NULL: scale_featurization)
 * in ONE kernel: the [N*M, L*C] feature tensor, the weights and their product never reach HBM.  fp32, input_dim 3.
 * out rows are out_stride_n floats apart.  grid_sizes: int32 [L], DEVICE (GridEncoder.grid_sizes, grid.py:139). */
int snerf_grid_encode_ms_fwd(const SnerfGridDesc* desc, const float* means, const float* stds, float bound,
                             const void* embeddings, const int32_t* offsets, const int32_t* grid_sizes,
                             const float* level_gain, float* out, int64_t out_stride_n, int64_t n_samples,
                             int32_t n_multi, void* stream);
/* its gradient w.r.t. the table (what zip-NeRF trains: the sample positions are detached, models.py:208-209):
 * ACCUMULATES w / M * grad[n, l*C + c] * (trilinear corner weight) into grad_embeddings[sO, C]. */
int snerf_grid_encode_ms_bwd(const SnerfGridDesc* desc, const float* grad, int64_t grad_stride_n, const float* means,
                             const float* stds, float bound, const int32_t* offsets, const int32_t* grid_sizes,
                             float* grad_embeddings, int64_t n_samples, int32_t n_multi, void* stream);
/* level_gain[l] = sqrt(init_std^2 + mean_{cells of level l} |embedding|^2) (models.py:496-503; the reference uses
 * torch_scatter.segment_coo).  scratch: double [L], ZERO-FILLED by the caller. */
int snerf_grid_level_gain(const SnerfGridDesc* desc, const void* embeddings, const int32_t* offsets, float init_std,
                          double* scratch, float* level_gain, void* stream);

/* ---- zip-NeRF proposal resampling (BASELINE configs[3]) ---------------------------------------------------------
 * One pass of the sampling loop of Model.forward (s-nerfpp/zipnerf/internal/models.py:156-213) in ONE kernel, one
 * warp per ray:   [dilate]  stepfun.max_dilate_weights(t, w, dilation, domain, renormalize) and the [1:-1] slices
 *                           (stepfun.py:75-105, models.py:174-182)
 *                 logits  = where(t[1:] > t[:-1], anneal * log(w + resample_padding), -inf)   (models.py:193-196)
 *                           (skipped when weights_are_logits: `w` then holds w_logits, as stepfun.sample_intervals takes)
 *                 out     = stepfun.sample_intervals(rand, t, logits, n_samples, single_jitter, domain)
 *                           (stepfun.py:251-294 -> sample :175-218 -> invert_cdf :154-161 -> math.sorted_interp)
 * t [n_rays, n_bins+1] sorted, w [n_rays, n_bins]; n_bins <= 128 with dilate (383 without), n_samples <= 256.
 * u_base [n_samples]: the linspace term of `u` (stepfun.py:205-216, computed by the caller with torch.linspace);
 * jitter: the torch.rand draw [n_rays, jitter_cols] (1 column = single_jitter) or NULL (rand = None);
 * u = u_base + jitter * max_jitter.
 * out [n_rays, n_samples+1] interval edges; centers [n_rays, n_samples] (optional): the sampled points before the
 * midpoint step; t_dilate [n_rays, 3*n_bins+1], w_dilate [n_rays, 3*n_bins] (optional): max_dilate_weights' outputs.
 * n_samples = 0: only the dilation. */
typedef struct SnerfStepfunOpts {
  int32_t dilate;              /* run max_dilate_weights first and drop the first / last dilated bin */
  int32_t renormalize;         /* max_dilate_weights(renormalize=...)                                */
  int32_t weights_are_logits;  /* sample_intervals' own signature: w holds w_logits                  */
  float dilation;
  float domain_lo, domain_hi;
  float anneal;                /* models.py:186-191 */
  float resample_padding;      /* Model.resample_padding */
  float max_jitter;            /* stepfun.py:212 */
} SnerfStepfunOpts;
int snerf_stepfun_resample(const SnerfStepfunOpts* opts, const float* t, const float* w, int64_t n_rays,
                           int32_t n_bins, const float* u_base, const float* jitter, int32_t jitter_cols,
                           int32_t n_samples, float* out, float* centers, float* t_dilate, float* w_dilate,
                           void* stream);

/* ---- training objective ---------------------------------------------------------------
 * The loss of s-nerf/train.py:149-209 evaluated straight from the renderer's outputs:
 *     img_loss   = RgbLoss:   mean((rgb - target)^2)                                   (model/loss_factory.py:5-11)
 *     depth_loss = calc_depth_loss (model/confidence.py:211-226) over DepthLoss (loss_factory.py:26-37):
 *                  mean over rays with target_depth != 0 of  conf * (|f(depth) - f(t)| + coarse_depth_mult * |f(depth0) - f(t)|),
 *                  f(x) = 1/x when disparity (args.disparity_depth) else x;  conf = NULL: no confidence weighting
 *     loss       = img_loss + rgb0_weight * mean((rgb0 - target)^2) + depth_lambda * depth_loss      (train.py:150,209)
 * (rgb0_weight = 0 and rgb0 = NULL reproduce the reference, which has no coarse colour term; depth = NULL drops the depth
 * term.)  One reduction kernel forward, one elementwise kernel backward.
 * scratch: 5 doubles, ZERO-FILLED by the caller.  out: float[5] = {loss, img_loss, depth_loss, masked count, img_loss0}. */
typedef struct SnerfLossOpts {
  float depth_lambda;
  float coarse_depth_mult;
  float rgb0_weight;
  int32_t disparity;
} SnerfLossOpts;
int snerf_loss_fwd(const SnerfLossOpts* opts, const float* rgb, const float* rgb0, const float* target, const float* depth,
                   const float* depth0, const float* target_depth, const float* confidence, int64_t n_rays, double* scratch,
                   float* out, void* stream);
/* gradients w.r.t. rgb [N,3], rgb0 [N,3], depth [N], depth0 [N], confidence [N] (any may be NULL), scaled by the device
 * scalar grad_loss[0]; stats = the `out` of the forward call. */
int snerf_loss_bwd(const SnerfLossOpts* opts, const float* rgb, const float* rgb0, const float* target, const float* depth,
                   const float* depth0, const float* target_depth, const float* confidence, int64_t n_rays,
                   const float* stats, const float* grad_loss, float* g_rgb, float* g_rgb0, float* g_depth, float* g_depth0,
                   float* g_confidence, void* stream);

/* torch.optim.Adam (the reference's optimizer: s-nerf/model/render.py:222, train.py:100) on ONE flat fp32 parameter buffer
 * with its flat gradient buffer: same update as torch's Adam (no amsgrad), one launch.  `lr` (float) and `step` (int64,
 * incremented by the call) are DEVICE scalars so the call can be captured in a CUDA graph. */
int snerf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, const float* lr,
                    float beta1, float beta2, float eps, float weight_decay, int64_t* step, void* stream);

/* ProposalLoss (s-nerf/model/loss_factory.py:54-73): weight * mean over rays of sum_i max(w_f[i] - bound[i], 0)^2 /
 * (w_f[i] + 1e-8), bound = W_c[inds[1:]-1] - W_c[inds[:-1]-1], inds = searchsorted(s_vals_c, s_vals_f, right=True),
 * W_c = cumsum(weights_c) (gather indices clamped into [0, n_coarse-1]: identical wherever the reference's gathers are in
 * range).  s_vals_f [N, n_fine+1], weights_f [N, n_fine], s_vals_c [N, n_coarse+1], weights_c [N, n_coarse]; <= 256
 * intervals.  The fine histogram is detached in the reference: grad_weights_c [N, n_coarse] (optional) receives
 * d loss / d weights_c from the same launch.  scratch: 2 doubles, ZERO-FILLED by the caller; loss_out: float[1]. */
int snerf_proposal_loss(const float* s_vals_f, const float* weights_f, const float* s_vals_c, const float* weights_c,
                        int64_t n_rays, int32_t n_fine, int32_t n_coarse, float weight, double* scratch, float* loss_out,
                        float* grad_weights_c, void* stream);

/* ---- bring-up diagnostics ------------------------------------------------------- */
/* ---- mip-NeRF path (what the reference's train.py / eval.py run: s-nerf/model/models.py:72-187 on the warp path of
 * configs/nuScenes_depth_6cams).  The Python mirror snerf_b200.models.MipNerfModel strings
 * these together per level: encode -> layers (snerf_linear_tc) -> composite (+ resample).  Rows: sample i of ray n is row
 * n * rows_per_ray + i of every [M_pad, ...] buffer (rows_per_ray >= n_samples; M_pad a multiple of 128). */

/* warp_sample_along_rays / sample2enc / integrated_pos_enc(diag=False) (mip.py:268-291, 375-390, 94-118): one bf16 row
 * [128] = [IPE (6 * max_deg) | zeros] per sample, the K-major A operand of the first layer. */
typedef struct SnerfMipEncode {
  const float* rays;       /* [n_rays, 9]: origin, direction, radius, near, far                            */
  int64_t n_rays;
  int32_t n_samples;       /* intervals per ray (level 0: N_samples; level 1: N_fine - 1)                  */
  int32_t rows_per_ray;
  const float* s_lin;      /* [n_samples + 1] torch.linspace(0, 1, n_samples + 1); used when s_in == NULL   */
  const float* s_rand;     /* [n_rays, n_samples + 1] jitter (randomized) or NULL                          */
  const float* s_in;       /* [n_rays, n_samples + 1] given s_vals (resampled level) or NULL               */
  float* s_out;            /* [n_rays, n_samples + 1] the s_vals built from s_lin / s_rand (s_in == NULL)   */
  int32_t transform_idx;   /* 0 log, 1 disparity, 2 linear (mip.py:393-400)                                */
  int32_t max_deg;         /* max_deg_point (16)                                                           */
  int32_t ray_cone;        /* 1 = 'cone', 0 = 'cylinder'                                                   */
  float radius;            /* contraction radius (the reference hard-codes 3, mip.py:379)                  */
  void* enc;               /* [m_pad, 128] bf16                                                            */
  float* enc_f32;          /* optional [n_rays * n_samples, 6 * max_deg] fp32 copy (tests)                 */
  int64_t m_pad;
} SnerfMipEncode;
int snerf_mip_encode(const SnerfMipEncode* e, void* stream);

/* One DenseBlock / nn.Linear (models.py:200-215) on the tensor cores: out = act(A . W^T + bias), A = [a0 | a1] (two
 * K segments, e.g. the skip layer's [x, inputs], models.py:274-275), W packed [n_pad, k0 + k1] bf16 row-major. */
typedef struct SnerfLinear {
  const void* a0; int64_t lda0; int32_t k0;     /* bf16 [m_pad, lda0], first k0 columns used (multiple of 64)     */
  const void* a1; int64_t lda1; int32_t k1;     /* optional second segment                                       */
  const void* w;                                /* bf16 [n_pad, k0 + k1]                                         */
  int32_t n, n_pad;                             /* valid outputs (multiple of 32), padded rows of w (multiple of 128) */
  const float* bias;                            /* [n] fp32 or NULL                                              */
  const float* ray_bias; int32_t rows_per_ray;  /* [rays, n] fp32 added to every row of the ray, or NULL         */
  int32_t relu;
  void* out; int64_t ldo;                       /* bf16 [m_pad, ldo] or NULL (heads only)                        */
  const float* head_w; int32_t n_heads;         /* fp32 [n_heads <= 3, n]: head_out[row * head_ld + h] += act(row) . head_w[h] */
  float* head_out;                              /* fp32, pre-initialised by the caller (zeros or the head biases)  */
  int32_t head_ld;                              /* row pitch of head_out in floats (0 = n_heads)                   */
  int64_t m_rows, m_pad;
} SnerfLinear;
int snerf_linear_tc(const SnerfLinear* l, void* stream);

/* x[rows, row_stride] fp32, columns [col0, col0 + ncols) -> bf16 [m_pad, out_cols] zero-padded (rows past `rows`, columns
 * past ncols): the operand rows snerf_linear_tc consumes.  `repeat` > 1 writes every input row `repeat` times (a per-ray
 * vector broadcast over the ray's samples, run_nerf_helpers.py:467-470). */
int snerf_rows_to_bf16(const float* x, int64_t rows, int32_t row_stride, int32_t col0, int32_t ncols, int32_t repeat,
                       void* out, int32_t out_cols, int64_t m_pad, void* stream);

/* W_cond0[:, k0 : k0 + 3 + 6 deg_view] . pos_enc(viewdirs) + b -> [n_rays, n_out]: the view-direction half of the first
 * condition layer as a per-ray bias (models.py:283-288, mip.py:12-21). */
int snerf_mip_cond_bias(const float* viewdirs, int64_t n_rays, int32_t deg_view, const float* w, int32_t ldw, int32_t k0,
                        const float* b, int32_t n_out, float* out, void* stream);

/* softplus / sigmoid heads + real_volumetric_rendering (models.py:166-178, mip.py:151-189) and, when s_new != NULL, the
 * resampling of warp_resample_along_rays (mip.py:294-313, math_ops.py:19-76). */
typedef struct SnerfMipComposite {
  const float* rays; int64_t n_rays;
  int32_t n_samples, rows_per_ray;
  const float* s_vals;          /* [n_rays, n_samples + 1]                                                 */
  const float* raw_density;     /* [m_pad] density-head dot products without the head bias                 */
  const float* raw_rgb;         /* [m_pad, 3] likewise, or NULL (proposal level: comp_rgb is None)          */
  const float* noise;           /* [n_rays, n_samples] density noise or NULL                               */
  float density_head_bias, density_bias, rgb_padding;
  float rgb_head_bias[3];
  int32_t transform_idx, white_bkgd;
  float* comp_rgb;              /* [n_rays, 3] or NULL                                                     */
  float* distance; float* acc;  /* [n_rays]                                                                */
  float* weights;               /* [n_rays, n_samples] or NULL                                             */
  int32_t n_fine;
  const float* u_lin;           /* [n_fine] torch.linspace(0, 1 - eps, n_fine)                             */
  const float* u_rand;          /* [n_rays, n_fine] uniform(0, 1 / n_fine - eps) draw, or NULL             */
  float resample_padding;
  float* s_new;                 /* [n_rays, n_fine] resampled s_vals, or NULL                              */
} SnerfMipComposite;
int snerf_mip_composite(const SnerfMipComposite* c, void* stream);

/* One 128x128x64 bf16 tcgen05.mma on device-resident row-major A[128,64], B[128,64]
 * (fp32 in, rounded to bf16 inside): D[128,128] = A * B^T.  Validates the UMMA
 * descriptor / swizzle / TMEM plumbing in isolation.  variant 0: A operand from shared
 * memory; variant 1: A operand staged in tensor memory (tcgen05.st + the TS form), the
 * path the renderer's hidden layers use. */
int snerf_selftest_umma(const float* a, const float* b, float* d, int32_t variant, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SNERF_B200_H_ */
