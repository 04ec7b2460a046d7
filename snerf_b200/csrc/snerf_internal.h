// snerf_internal.h -- host-side declarations shared by the translation units of libsnerf_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/snerf_b200.h"

namespace snerf {

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int sm_count();

// which rows feed the MLP tiles of the fp32 kernel
enum Frontend { FE_RAYS = 0, FE_QUERY = 1, FE_ROWS = 2 };

struct RenderParams {
  // rays front-end (render_rays)
  // camera mode (SnerfOpts.camera): rays are generated in the kernel prologue, ray_batch is null
  int cam_on, cam_W;
  float cam_focal, cam_cx, cam_cy, cam_near, cam_far;
  float cam_m[12];
  long long cam_first;
  const float* ray_batch;
  long long n_rays;
  int width, row_stride, has_vd;
  int Nc, Nf, lindisp, white_bkgd, L, Lv;
  const float *t_vals, *u_vals, *t_rand, *u_rand, *noise0, *noise1;
  SnerfOut out;
  const unsigned char* img_coarse;
  const unsigned char* img_fine;
  int tc_op;                              // tensor-core kernel: operand arithmetic (OP_BF16 / OP_F16 / OP_F16X3)
  int coarse_depth;   // tensor-core modes: trunk depth of the COARSE network, 8 or 4 (0 = 8); the fine network is always 8x256
  const unsigned char* img_alpha_coarse;  // optional frozen sigma network evaluated before img_coarse (NeRF_RGB)
  const unsigned char* img_alpha_fine;    // likewise for the fine pass
  // query front-end (network_query_fn): pts[n_rays, S, 3], viewdirs[n_rays, 3]
  const float* pts;
  const float* viewdirs;
  int S;
  // rows front-end (NeRF.forward): x[n_rows, x_stride]
  const float* x;
  long long n_rows;
  int x_stride, in_ch, in_ch_views;
  float* out_raw;  // [rows, 4] for the query / rows front-ends
  // training (fp32 rays front-end): activation stores [channel][R] of the coarse / fine pass, null = inference
  float* save_c;
  float* save_f;
  long long Rc, Rf;
  // training (tensor-core rays front-end): 16-bit activation stores + relu' mask bits (layout: snerf_packed.h), null = inference
  unsigned char* act_c;
  unsigned char* act_f;
  unsigned long long* bits_c;
  unsigned long long* bits_f;
  long long act_rows_c, act_rows_f;
  int stage;       // 0 = whole pipeline; 1 = coarse inputs only; 2 = from stored coarse raw: composite, resample, fine inputs
  int round_tf32;  // store the encoded inputs rounded to tf32
};

// ---- training (snerf_train.cu)
struct TrainChannels { int trunk[SNERF_MAX_TRUNK_LAYERS]; int feature, views, total; };
struct TrainLayout {   // float offsets into the training workspace
  int TC, TF;
  long long Rc, Rf;
  size_t save_c, save_f, dz_c, dz_f, draw_c, draw_f, raw_c, raw_f, z_c, z_f, total_floats;
};
struct TrainParams {
  const float* ray_batch;
  long long n_rays;
  int width, row_stride;
  int Nc, Nf, white_bkgd, TC, TF;
  long long Rc, Rf;
  const float *noise0, *noise1;
  const float *raw_c, *raw_f, *z_c, *z_f;          // saved by the forward
  SnerfOutGrad g;                                   // upstream gradients
  const float *save_c, *save_f;                     // activation stores [channel][R]
  float *dz_c, *dz_f;                               // gradient stores, same channel numbering (pre-shifted base)
  float *draw_c, *draw_f;                           // d_raw [4][R]
  float4 *draw4_c, *draw4_f;                        // tensor-core path instead: d_raw [n_rays * S][4]
  const unsigned char *bwd_c, *bwd_f;               // backward images
  int dw_tf32;                                      // weight gradients on the tensor cores (kind::tf32)
};
constexpr int kMaxDwProblems = 32;
constexpr int kMaxSkProblems = 40;
struct DwProblem {
  const float* A; const float* B; float* C;
  long long R;
  int M, N, ldc, mt, nt, splits, rows_per_split, first;
};
struct DwTable { int n; DwProblem p[kMaxDwProblems]; };
struct SkProblem {
  const float* X; const float* G; float* out;
  long long R;
  int K, C, ldo, first;
};
struct SkTable { int n; SkProblem p[kMaxSkProblems]; };

TrainChannels train_channels(const SnerfNetDesc* d);
TrainLayout train_layout(const SnerfNetDesc* d, const SnerfNetDesc* df, int Nc, int Nf, long long n_rays);
bool train_supported(const SnerfNetDesc* d, int tf32);
struct Fp32BwdHeader;
size_t plan_bwd(const SnerfNetDesc* d, Fp32BwdHeader* h);
int pack_bwd(const SnerfNetDesc* d, const SnerfNetF32* src, void* packed, int tf32, cudaStream_t stream);
struct RenderParams;
struct Fp32Header;
size_t plan_fp32(const SnerfNetDesc* d, Fp32Header* h, bool with_alpha);
int launch_train_forward_tf32(const SnerfNetDesc* d, RenderParams p, const TrainLayout& L, float* ws, cudaStream_t stream);
int launch_train_backward(const SnerfNetDesc* d, const SnerfNetDesc* d_fine, const TrainParams& p,
                          const SnerfNetGradF32* grad_coarse, const SnerfNetGradF32* grad_fine, cudaStream_t stream);

int launch_fp32(int frontend, int W, const RenderParams& p, cudaStream_t stream);
int launch_bf16_render(const RenderParams& p, cudaStream_t stream);
int launch_x3_render(const RenderParams& p, cudaStream_t stream);
int launch_bf16_render_d4(const RenderParams& p, cudaStream_t stream);   // coarse network 4x256 (snerf_bf16_d4.cu)
int launch_x3_render_d4(const RenderParams& p, cudaStream_t stream);     // (snerf_x3_d4.cu)
bool bf16_geometry_supported(int n_samples, int n_importance);
int launch_bf16_query(const RenderParams& p, cudaStream_t stream);
int launch_selftest_umma(const float* a, const float* b, float* d, int variant, cudaStream_t stream);

// ---- batched weight packing (snerf_api.cu): every block / transposed-segment copy of one image in ONE launch
struct PackJob {
  const float* w;    // source matrix [n_out or rows, ld]
  float* dst;
  int kind;          // 0 = K-major with up to three padded column segments (dst[k][n]); 1 = row block (dst[n][k])
  int ld, n_out;
  int rows_pad[3], rows_real[3], col0[3];   // kind 0
  int rows, cols, col_first;                // kind 1: dst[n * cols + k] = w[n * ld + col_first + k]
  int round_tf32;
};
constexpr int kMaxPackJobs = 28;
struct PackJobs { int n; PackJob j[kMaxPackJobs]; };
int launch_pack_jobs(const PackJobs& jobs, cudaStream_t stream);

// ---- hash-grid encoder (snerf_grid.cu)
int grid_check_desc(const SnerfGridDesc* d);
int grid_fwd(const SnerfGridDesc* d, const float* inputs, const void* emb, const int32_t* offsets, void* out,
             long long sl, long long sb, void* dy_dx, long long B, cudaStream_t st);
int grid_bwd(const SnerfGridDesc* d, const void* grad, long long sl, long long sb, const float* inputs,
             const int32_t* offsets, void* grad_emb, const void* dy_dx, void* grad_inputs, long long B, cudaStream_t st);
int grid_tv(const SnerfGridDesc* d, const float* inputs, const void* emb, void* grad, const int32_t* offsets,
            float weight, long long B, cudaStream_t st);
int grid_ms_fwd(const SnerfGridDesc* d, const float* means, const float* stds, float bound, const void* emb,
                const int32_t* offsets, const int32_t* grid_sizes, const float* level_gain, float* out, long long sn,
                long long N, int M, cudaStream_t st);
int grid_ms_bwd(const SnerfGridDesc* d, const float* grad, long long sn, const float* means, const float* stds, float bound,
                const int32_t* offsets, const int32_t* grid_sizes, float* grad_emb, long long N, int M, cudaStream_t st);
int grid_level_gain(const SnerfGridDesc* d, const void* emb, const int32_t* offsets, float init_std, double* scratch,
                    float* gain, cudaStream_t st);

// ---- mip-NeRF path (snerf_mip.cu)
int linear_tc(const SnerfLinear* L, cudaStream_t stream);
int mip_encode(const SnerfMipEncode* e, cudaStream_t stream);
int rows_to_bf16(const float* x, long long rows, int row_stride, int col0, int ncols, int repeat, void* out, int out_cols,
                 long long m_pad, cudaStream_t stream);
int mip_cond_bias(const float* viewdirs, long long n_rays, int deg_view, const float* w, int ldw, int k0, const float* b, int n_out,
                  float* out, cudaStream_t stream);
int mip_composite(const SnerfMipComposite* c, cudaStream_t stream);

// ---- training objective (snerf_loss.cu)
int loss_fwd(const SnerfLossOpts* o, const float* rgb, const float* rgb0, const float* target, const float* depth,
             const float* depth0, const float* tdepth, const float* conf, long long N, double* scratch, float* out,
             cudaStream_t st);
int loss_bwd(const SnerfLossOpts* o, const float* rgb, const float* rgb0, const float* target, const float* depth,
             const float* depth0, const float* tdepth, const float* conf, long long N, const float* stats, const float* g,
             float* g_rgb, float* g_rgb0, float* g_depth, float* g_depth0, float* g_conf, cudaStream_t st);

int adam_step(float* p, const float* g, float* m, float* v, long long n, const float* lr_dev, float b1, float b2, float eps,
              float wd, long long* step_dev, cudaStream_t st);

int proposal_loss(const float* s_f, const float* w_f, const float* s_c, const float* w_c, long long N, int Sf, int Sc,
                  float weight, double* scratch, float* loss_out, float* grad_wc, cudaStream_t st);

// ---- proposal resampling (snerf_stepfun.cu)
int stepfun_resample(const float* t, const float* w, long long N, int S, int dilate, int renormalize, int logits_in,
                     float dilation, float lo, float hi, float anneal, float padding, const float* u_base,
                     const float* jitter, int jd, float max_jitter, int n, float* out, float* centers, float* t_dil,
                     float* w_dil, cudaStream_t st);

}  // namespace snerf
