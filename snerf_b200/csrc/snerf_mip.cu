// snerf_mip.cu -- the mip-NeRF path the reference's train.py / eval.py actually run:
// MipNerfModel.forward on the warp path of configs/nuScenes_depth_6cams (s-nerf/model/models.py:72-187; no_warp_sample = 0,
// fn = 1, ray_shape = 'cone', transform_idx = 0, hidden_layer = 1024, rgb_layer = 3).
//
//   mip_encode_kernel     s_vals (stratified, jittered: mip.py:268-288, or the resampled ones) -> t = T(s) (mip.py:7-9)
//                         -> conical frustum -> Gaussian (mip.py:56-70, 31-45) -> mip-360 contraction + Jacobian
//                         (mip.py:339-367) -> diag(J cov J^T) (mip.py:375-390) -> integrated positional encoding
//                         (mip.py:94-118, 24-28) -> one bf16 row of 128 (96 + padding) per sample: the K-major A operand
//                         of the first network layer.  fp32 CUDA-core math, one thread per sample.
//   lin_tc_kernel         every DenseBlock of `proposal` (4 x 256) and `MLP` (8 x 1024, [x, inputs] skip, bottleneck,
//                         condition layers; models.py:217-325) as a persistent tcgen05 GEMM
//                             out[M, N] = act(A[M, K] . W[N, K]^T + bias)          bf16 operands, fp32 accumulate in TMEM
//                         128 x 256 tiles, TMA-fed 4-stage ring (both operands K-major, 128B swizzle), two TMEM
//                         accumulators so the epilogue of tile i overlaps the MMAs of tile i + 1, A from up to two
//                         buffers (the skip layer contracts over [h4 | enc] without materialising the concatenation),
//                         per-ray bias (the view-direction part of the first condition layer), and the narrow heads
//                         (density_layer, rgb_layer) folded into the epilogue as fp32 dot products.
//                         A 1024-wide layer does not fit one SM (128 x 1024 fp32 accumulators = 2x TMEM), so unlike the
//                         vanilla renderer the activations travel through HBM / L2 as bf16: 4 KB per sample and layer
//                         against 2.1 MFLOP -> 512 FLOP/B, compute-bound on the tensor pipe.
//   mip_cond_bias_kernel  pos_enc(viewdirs) (mip.py:12-21) . W_cond0[:, hidden:] + b: the per-ray bias of that layer.
//   mip_composite_kernel  softplus(raw_density + density_bias), sigmoid + rgb_padding (models.py:166-175),
//                         real_volumetric_rendering (mip.py:151-189) and -- for the proposal level -- the blurred max-pool +
//                         sorted_piecewise_constant_pdf resampling (mip.py:294-313, math_ops.py:19-76): one warp per ray.
#include <cuda.h>  // CUtensorMap + cuTensorMapEncodeTiled prototype only; resolved at run time (no libcuda link)
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "snerf_internal.h"
#include "snerf_umma.cuh"

namespace snerf {

// ------------------------------------------------------------------------------------
// sampling + encoding
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float mip_transform(float s, float near, float far, int idx) {
  if (idx == 0) return __fmul_rn(near, expf(__fmul_rn(s, logf(__fdiv_rn(far, near)))));
  if (idx == 1) return __fdiv_rn(1.f, __fadd_rn(__fdiv_rn(__fsub_rn(1.f, s), near), __fdiv_rn(s, far)));
  return __fadd_rn(__fmul_rn(near, __fsub_rn(1.f, s)), __fmul_rn(far, s));
}
// math_ops.safe_sin (math_ops.py:6-16): sin(x) with |x| >= 100 pi folded by torch's remainder (sign of the divisor)
__device__ __forceinline__ float mip_safe_sin(float x) {
  const float t = 314.159265358979f;
  if (!(fabsf(x) < t)) {
    float r = fmodf(x, t);
    if (r != 0.f && ((r < 0.f) != (t < 0.f))) r += t;
    x = r;
  }
  return sinf(x);
}

struct MipEncodeParams {
  const float* rays;        // [N, 9]: origin, direction, radius, near, far
  long long n_rays;
  int S;                    // samples (intervals) per ray
  int rows_per_ray;         // >= S, rows [S, rows_per_ray) are zero padding
  const float* s_lin;       // [S + 1] linspace(0, 1, S + 1): level 0 builds its own s_vals ...
  const float* s_rand;      // [N, S + 1] jitter or null
  const float* s_in;        // ... or [N, S + 1] given s_vals (resampled level)
  float* s_out;             // [N, S + 1] written when s_in is null
  int transform_idx, max_deg, ray_cone;
  float radius;
  __nv_bfloat16* enc;       // [M_pad, 128]
  float* enc_f32;           // optional [N * S, 6 * max_deg] (tests)
  long long m_pad;
};

__global__ void __launch_bounds__(128) mip_encode_kernel(const MipEncodeParams p) {
  const long long row = blockIdx.x * 128ll + threadIdx.x;
  if (row >= p.m_pad) return;
  const long long ray = row / p.rows_per_ray;
  const int i = (int)(row - ray * p.rows_per_ray);
  uint4* dst = reinterpret_cast<uint4*>(p.enc + row * 128);
  if (ray >= p.n_rays || i >= p.S) {
#pragma unroll
    for (int q = 0; q < 16; ++q) dst[q] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const float* rr = p.rays + ray * 9;
  const float o[3] = {rr[0], rr[1], rr[2]}, d[3] = {rr[3], rr[4], rr[5]};
  const float rad = rr[6], near = rr[7], far = rr[8];
  float s0, s1;
  if (p.s_in) {
    s0 = p.s_in[ray * (p.S + 1) + i];
    s1 = p.s_in[ray * (p.S + 1) + i + 1];
  } else {
    // stratified s in [0, 1] (mip.py:276-288): edge k jittered between the mid-points around it
    auto edge = [&](int k) {
      float s = p.s_lin[k];
      if (p.s_rand) {
        const float lo = k > 0 ? __fmul_rn(0.5f, __fadd_rn(p.s_lin[k], p.s_lin[k - 1])) : s;
        const float hi = k < p.S ? __fmul_rn(0.5f, __fadd_rn(p.s_lin[k + 1], p.s_lin[k])) : s;
        s = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), p.s_rand[ray * (p.S + 1) + k]));
      }
      return s;
    };
    s0 = edge(i);
    s1 = edge(i + 1);
    p.s_out[ray * (p.S + 1) + i] = s0;
    if (i == p.S - 1) p.s_out[ray * (p.S + 1) + p.S] = s1;
  }
  const float t0 = mip_transform(s0, near, far, p.transform_idx), t1 = mip_transform(s1, near, far, p.transform_idx);
  float t_mean, t_var, r_var;
  if (p.ray_cone) {  // conical_frustum_to_gaussian, stable branch (mip.py:56-64)
    const float mu = (t0 + t1) / 2.f, hw = (t1 - t0) / 2.f;
    const float mu2 = mu * mu, hw2 = hw * hw, hw4 = hw2 * hw2;
    const float den = 3.f * mu2 + hw2;
    t_mean = mu + (2.f * mu * hw2) / den;
    t_var = hw2 / 3.f - (4.f / 15.f) * ((hw4 * (12.f * mu2 - hw2)) / (den * den));
    r_var = rad * rad * (mu2 / 4.f + (5.f / 12.f) * hw2 - (4.f / 15.f) * hw4 / den);
  } else {           // cylinder_to_gaussian (mip.py:73-77)
    t_mean = (t0 + t1) / 2.f;
    r_var = rad * rad / 4.f;
    t_var = (t1 - t0) * (t1 - t0) / 12.f;
  }
  // lift_gaussian, diagonal (mip.py:31-45)
  const float dmag = fmaxf(1e-10f, d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  float x[3], cov[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    x[j] = d[j] * t_mean + o[j];
    const float dd = d[j] * d[j];
    cov[j] = t_var * dd + r_var * (1.f - dd / dmag);
  }
  // contraction fn2 + Jacobi_g (mip.py:339-367); J = a I + b x x^T outside the radius, I / radius inside
  const float nrm = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  float f[3], fc[3];
  {
    const float ln = 1.f / (nrm + 1e-5f);
    const bool outside_j = (nrm + 1e-5f) >= p.radius;
    const float a = outside_j ? (-p.radius * ln * ln + 2.f * ln) : 1.f / p.radius;
    const float b = outside_j ? (2.f * p.radius * ln * ln * ln * ln - 2.f * ln * ln * ln) : 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float J = (r == c ? a : 0.f) + b * x[r] * x[c];
        acc += J * J * cov[c];
      }
      fc[r] = acc;
    }
    const float l = nrm + 1e-8f;
    const bool outside_f = l > p.radius;
#pragma unroll
    for (int j = 0; j < 3; ++j) f[j] = outside_f ? (2.f - p.radius / l) * x[j] / l : x[j] / p.radius;
  }
  // integrated positional encoding, full-covariance form reduced to its diagonal (mip.py:104-118): y = 2^i x, var = 4^i cov_jj
  float e[128];
  const int nd = p.max_deg * 3;
#pragma unroll 1
  for (int deg = 0; deg < p.max_deg; ++deg) {
    const float sc = __int_as_float((127 + deg) << 23);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float y = f[j] * sc, var = fc[j] * sc * sc;
      const float damp = expf(-0.5f * var);
      e[deg * 3 + j] = damp * mip_safe_sin(y);
      e[nd + deg * 3 + j] = damp * mip_safe_sin(__fadd_rn(y, 1.5707963267948966f));
    }
  }
  for (int k = 2 * nd; k < 128; ++k) e[k] = 0.f;
  if (p.enc_f32) {
    float* ef = p.enc_f32 + (ray * p.S + i) * (2 * nd);
    for (int k = 0; k < 2 * nd; ++k) ef[k] = e[k];
  }
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      __nv_bfloat162 h = __floats2bfloat162_rn(e[q * 8 + 2 * k], e[q * 8 + 2 * k + 1]);
      w[k] = *reinterpret_cast<uint32_t*>(&h);
    }
    dst[q] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// per-ray bias of the first condition layer: W[:, k0 : k0 + 3 + 6 deg] . pos_enc(viewdir) + b   (models.py:283-288, mip.py:12-21)
__global__ void __launch_bounds__(128) mip_cond_bias_kernel(const float* __restrict__ viewdirs, long long n_rays, int deg_view,
                                                            const float* __restrict__ w, int ldw, int k0, const float* __restrict__ b,
                                                            int n_out, float* __restrict__ out) {
  const long long idx = blockIdx.x * 128ll + threadIdx.x;
  if (idx >= n_rays * n_out) return;
  const long long ray = idx / n_out;
  const int j = (int)(idx - ray * n_out);
  const float v[3] = {viewdirs[ray * 3], viewdirs[ray * 3 + 1], viewdirs[ray * 3 + 2]};
  const float* wr = w + (long long)j * ldw + k0;
  float acc = b[j];
#pragma unroll
  for (int c = 0; c < 3; ++c) acc = fmaf(wr[c], v[c], acc);
  const int nd = 3 * deg_view;
  for (int dgr = 0; dgr < deg_view; ++dgr) {
    const float sc = __int_as_float((127 + dgr) << 23);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float xb = v[c] * sc;
      acc = fmaf(wr[3 + dgr * 3 + c], sinf(xb), acc);
      acc = fmaf(wr[3 + nd + dgr * 3 + c], sinf(__fadd_rn(xb, 1.5707963267948966f)), acc);
    }
  }
  out[idx] = acc;
}

// ------------------------------------------------------------------------------------
// compositing (+ resampling): one warp per ray
// ------------------------------------------------------------------------------------
struct MipCompositeParams {
  const float* rays;          // [N, 9]
  long long n_rays;
  int S, rows_per_ray;
  const float* s_vals;        // [N, S + 1]
  const float* raw_density;   // [M_pad] head dot products WITHOUT the head bias
  const float* raw_rgb;       // [M_pad, 3] or null (proposal level)
  const float* noise;         // [N, S] or null (density_noise * randn)
  float density_head_bias, density_bias, rgb_padding;
  float rgb_head_bias[3];
  int transform_idx, white_bkgd;
  float* comp_rgb;            // [N, 3] or null
  float* distance;            // [N]
  float* acc;                 // [N]
  float* weights;             // [N, S]
  // resampling (null s_new = none)
  int n_fine;
  const float* u_lin;         // [n_fine] linspace(0, 1 - eps, n_fine)
  const float* u_rand;        // [N, n_fine] uniform(0, 1 / n_fine - eps) or null
  float resample_padding;
  float* s_new;               // [N, n_fine]
};

constexpr int kMipMaxS = 256;

__global__ void __launch_bounds__(128) mip_composite_kernel(const MipCompositeParams p) {
  __shared__ float sh_w[4][kMipMaxS + 2];
  __shared__ float sh_cdf[4][kMipMaxS + 2];
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const long long ray = blockIdx.x * 4ll + wl;
  if (ray >= p.n_rays) return;
  const int S = p.S;
  const float* rr = p.rays + ray * 9;
  const float dnorm = sqrtf(rr[3] * rr[3] + rr[4] * rr[4] + rr[5] * rr[5]);
  const float near = rr[7], far = rr[8];
  const float* sv = p.s_vals + ray * (S + 1);
  float* w = sh_w[wl];
  // each lane owns a contiguous run of samples
  const int C = (S + 31) >> 5, i0 = lane * C;
  float dd[8], tm[8], col[8][3];
  double run = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    dd[j] = 0.f; tm[j] = 0.f; col[j][0] = col[j][1] = col[j][2] = 0.f;
    const int s = i0 + j;
    if (j < C && s < S) {
      const float t0 = mip_transform(sv[s], near, far, p.transform_idx), t1 = mip_transform(sv[s + 1], near, far, p.transform_idx);
      tm[j] = 0.5f * (t0 + t1);
      const float delta = (t1 - t0) * dnorm;
      float raw = p.raw_density[ray * p.rows_per_ray + s] + p.density_head_bias;
      if (p.noise) raw += p.noise[ray * S + s];
      const float xsp = raw + p.density_bias;
      const float dens = xsp > 20.f ? xsp : log1pf(expf(xsp));   // F.softplus
      dd[j] = dens * delta;
      run += (double)dd[j];
      if (p.raw_rgb) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float r = p.raw_rgb[(ray * p.rows_per_ray + s) * 3 + c] + p.rgb_head_bias[c];
          col[j][c] = (1.f / (1.f + expf(-r))) * (1.f + 2.f * p.rgb_padding) - p.rgb_padding;
        }
      }
    }
  }
  const double incl = warp_scan_sum_d(run, lane);
  double excl = incl - run;   // sum of density_delta over the lanes below
  float a_r = 0.f, a_g = 0.f, a_b = 0.f, a_acc = 0.f, a_dist = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int s = i0 + j;
    if (j < C && s < S) {
      const float alpha = 1.f - expf(-dd[j]);
      const float trans = expf(-(float)excl);
      const float wt = alpha * trans;
      excl += (double)dd[j];
      w[s] = wt;
      if (p.weights) p.weights[ray * S + s] = wt;
      a_r += wt * col[j][0]; a_g += wt * col[j][1]; a_b += wt * col[j][2];
      a_acc += wt;
      a_dist += wt * tm[j];
    }
  }
  a_r = warp_sum(a_r); a_g = warp_sum(a_g); a_b = warp_sum(a_b); a_acc = warp_sum(a_acc); a_dist = warp_sum(a_dist);
  if (lane == 0) {
    const float tlo = mip_transform(sv[0], near, far, p.transform_idx), thi = mip_transform(sv[S], near, far, p.transform_idx);
    float dist = a_dist;
    if (dist != dist) dist = INFINITY;
    dist = fminf(fmaxf(dist, tlo), thi);
    p.distance[ray] = dist;
    p.acc[ray] = a_acc;
    if (p.comp_rgb) {
      const float wb = p.white_bkgd ? 1.f - a_acc : 0.f;
      p.comp_rgb[ray * 3] = a_r + wb; p.comp_rgb[ray * 3 + 1] = a_g + wb; p.comp_rgb[ray * 3 + 2] = a_b + wb;
    }
  }
  if (!p.s_new) return;
  __syncwarp();
  // ---- warp_resample_along_rays (mip.py:294-313): blurred max-pool + padding ...
  float* cdf = sh_cdf[wl];
  float blur[8];
  float part = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    blur[j] = 0.f;
    const int s = i0 + j;
    if (j < C && s < S) {
      const float wm1 = w[s > 0 ? s - 1 : 0], w0 = w[s], wp1 = w[s < S - 1 ? s + 1 : S - 1];
      blur[j] = 0.5f * (fmaxf(wm1, w0) + fmaxf(w0, wp1)) + p.resample_padding;
      part += blur[j];
    }
  }
  // ... sorted_piecewise_constant_pdf (math_ops.py:19-76)
  float wsum = warp_sum(part);
  const float padding = fmaxf(0.f, 1e-5f - wsum);
  wsum += padding;
  double prun = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int s = i0 + j;
    if (j < C && s < S) {
      blur[j] = (blur[j] + padding / (float)S) / wsum;   // pdf
      if (s < S - 1) prun += (double)blur[j];
    }
  }
  const double pincl = warp_scan_sum_d(prun, lane);
  double pex = pincl - prun;
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int s = i0 + j;
    if (j < C && s < S - 1) {
      pex += (double)blur[j];
      cdf[s + 1] = fminf(1.f, (float)pex);
    }
  }
  if (lane == 0) { cdf[0] = 0.f; cdf[S] = 1.f; }
  __syncwarp();
  const float step = 1.f / (float)p.n_fine;
  for (int k = lane; k < p.n_fine; k += 32) {
    float u;
    if (p.u_rand) u = fminf(__fadd_rn(__fmul_rn((float)k, step), p.u_rand[ray * p.n_fine + k]), 1.f - 1.1920929e-07f);
    else u = p.u_lin[k];
    // last index with cdf[idx] <= u (cdf[0] = 0 <= u always; cdf[S] = 1 > u always)
    int lo = 0, hi = S;   // invariant: cdf[lo] <= u < cdf[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= u) lo = mid; else hi = mid;
    }
    const float c0 = cdf[lo], c1 = cdf[lo + 1];
    float t = (u - c0) / (c1 - c0);
    if (t != t) t = 0.f;
    t = fminf(fmaxf(t, 0.f), 1.f);
    const float b0 = sv[lo], b1 = sv[lo + 1];
    p.s_new[ray * p.n_fine + k] = b0 + t * (b1 - b0);
  }
}

// ------------------------------------------------------------------------------------
// persistent tcgen05 GEMM:  out = act(A . W^T + bias)
// ------------------------------------------------------------------------------------
constexpr int kLinStages = 4;
constexpr int kLinABytes = 128 * 64 * 2;   // 16 KiB
constexpr int kLinBBytes = 256 * 64 * 2;   // 32 KiB
constexpr int kLinThreads = 192;

struct alignas(1024) LinSmem {
  uint8_t a[kLinStages][kLinABytes];
  uint8_t b[kLinStages][kLinBBytes];
  uint8_t stage_out[4][32 * 128];    // per epilogue warp: 32 rows x 64 bf16 columns, 16-byte pieces XOR-swizzled by row
  uint64_t full[kLinStages], empty[kLinStages];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

struct LinParams {
  int kb0, kb1;               // 64-wide k-blocks of the two A segments
  int tile_n, n_tiles_n, m_tiles;
  long long M;                // valid rows
  int N;                      // valid output columns (multiple of 8)
  const float* bias;          // [N] or null
  const float* ray_bias;      // [rays, N] or null
  int rows_per_ray, relu;
  __nv_bfloat16* out;         // [M_pad, ldo] or null
  long long ldo;
  const float* head_w;        // [n_heads, N] fp32 or null
  int n_heads, head_ld;
  float* head_out;            // [M_pad, head_ld], accumulated with atomics
};

__device__ __forceinline__ void tma_load_2d_lin(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// Epilogue of one 128-row x tile_n accumulator (thread = row = TMEM lane): + bias / per-ray bias, ReLU, fused heads, bf16
// conversion and the transposed, coalesced write-out through the warp's staging buffer.
__device__ __forceinline__ void lin_epilogue_tile(const LinParams& p, uint32_t acc_tmem, uint8_t* stg, int lg, int lane, int m0,
                                              int n0, long long row, bool row_ok, const float* rb, float (&hs)[3]) {
  for (int c0 = 0; c0 < p.tile_n; c0 += 64) {      // 64 columns per round: one 128-byte run per row
    const int n = n0 + c0;
    if (n >= p.N) break;                           // (warp-uniform) padded output columns; N is a multiple of 32
    const int halves = (n + 64 <= p.N) ? 2 : 1;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      if (hh >= halves) break;
      uint32_t v[32];
      tmem_ld32(acc_tmem + ((uint32_t)(lg * 32) << 16) + c0 + 32 * hh, v);
      tmem_ld_wait();
      const int nn = n + 32 * hh;
      float f[32];
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {   // bias / per-ray bias as 16-byte loads (the same address in every lane: one transaction)
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nn) + q4);
        if (rb) {
          const float4 r4 = __ldg(reinterpret_cast<const float4*>(rb + nn) + q4);
          b4.x += r4.x; b4.y += r4.y; b4.z += r4.z; b4.w += r4.w;
        }
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float x = __uint_as_float(v[q4 * 4 + k]) + bb[k];
          f[q4 * 4 + k] = p.relu ? fmaxf(x, 0.f) : x;
        }
      }
      if (p.head_w) {
        for (int h = 0; h < p.n_heads; ++h) {
          const float4* hw = reinterpret_cast<const float4*>(p.head_w + (long long)h * p.N + nn);
          float a = 0.f;
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 w4 = __ldg(hw + q4);
            a = fmaf(f[q4 * 4], w4.x, a); a = fmaf(f[q4 * 4 + 1], w4.y, a);
            a = fmaf(f[q4 * 4 + 2], w4.z, a); a = fmaf(f[q4 * 4 + 3], w4.w, a);
          }
          hs[h] += a;
        }
      }
      if (p.out) {   // this lane's row -> staging buffer (piece index XOR row: conflict-free 16-byte stores and loads)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t w4[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(f[q * 8 + 2 * k], f[q * 8 + 2 * k + 1]);
            w4[k] = *reinterpret_cast<uint32_t*>(&h2);
          }
          const int piece = hh * 4 + q;
          *reinterpret_cast<uint4*>(stg + lane * 128 + ((piece ^ (lane & 7)) << 4)) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
      }
    }
    if (p.out) {
      // transposed write-out: 8 lanes cover one row's 128 bytes, so every store instruction writes 4 full lines
      // (a lane-per-row store would touch 32 different lines per instruction)
      __syncwarp();
      const int piece = lane & 7, npieces = halves * 4;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + (lane >> 3);
        const long long grow = (long long)m0 + lg * 32 + r;
        if (piece < npieces && grow < p.M) {
          const uint4 val = *reinterpret_cast<const uint4*>(stg + r * 128 + ((piece ^ (r & 7)) << 4));
          *reinterpret_cast<uint4*>(p.out + grow * p.ldo + n + piece * 8) = val;
        }
      }
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(kLinThreads, 1)
lin_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
              const __grid_constant__ CUtensorMap mapW, const LinParams p) {
  extern __shared__ __align__(1024) unsigned char smem_lin[];
  LinSmem& sm = *reinterpret_cast<LinSmem*>(smem_lin);
  if ((smem_u32(smem_lin) & 1023u) != 0) __trap();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kLinStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sm.acc_full[i], 1); mbar_init(&sm.acc_empty[i], 4); }
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&sm.tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;
  const int n_kb = p.kb0 + p.kb1;
  const long long n_tiles = (long long)p.m_tiles * p.n_tiles_n;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t bytes = (uint32_t)kLinABytes + (uint32_t)p.tile_n * 128u;
      for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int m0 = (int)(t / p.n_tiles_n) * 128, n0 = (int)(t % p.n_tiles_n) * p.tile_n;
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(&sm.empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&sm.full[stage], bytes);
          if (kb < p.kb0) tma_load_2d_lin(sm.a[stage], &mapA0, kb * 64, m0, &sm.full[stage]);
          else tma_load_2d_lin(sm.a[stage], &mapA1, (kb - p.kb0) * 64, m0, &sm.full[stage]);
          tma_load_2d_lin(sm.b[stage], &mapW, kb * 64, n0, &sm.full[stage]);
          if (++stage == kLinStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(128, p.tile_n);
    int stage = 0;
    uint32_t phase = 0;
    long long it = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int buf = (int)(it & 1);
      mbar_wait(&sm.acc_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));   // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + 256 * buf;
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(&sm.full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t abase = smem_u32(sm.a[stage]), bbase = smem_u32(sm.b[stage]);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            tc_mma_ss(d_tmem, umma_desc_sw128(abase + ks * 32), umma_desc_sw128(bbase + ks * 32), idesc, (kb | ks) != 0 ? 1u : 0u);
          tc_commit(&sm.empty[stage]);
          if (kb == n_kb - 1) tc_commit(&sm.acc_full[buf]);
        }
        __syncwarp();
        if (++stage == kLinStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // epilogue: thread = accumulator row (TMEM lane group = warp % 4)
    const int lg = warp & 3;
    const int r_in_tile = lg * 32 + lane;
    long long it = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int buf = (int)(it & 1);
      const int m0 = (int)(t / p.n_tiles_n) * 128, n0 = (int)(t % p.n_tiles_n) * p.tile_n;
      const long long row = (long long)m0 + r_in_tile;
      const bool row_ok = row < p.M;
      const float* rb = p.ray_bias ? p.ray_bias + (row_ok ? row / p.rows_per_ray : 0) * p.N : nullptr;
      mbar_wait(&sm.acc_full[buf], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      float hs[3] = {0.f, 0.f, 0.f};
      lin_epilogue_tile(p, tmem_base + 256 * buf, sm.stage_out[warp - 2], lg, lane, m0, n0, row, row_ok, rb, hs);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.acc_empty[buf]);
      if (p.head_w && row_ok)
        for (int h = 0; h < p.n_heads; ++h) atomicAdd(p.head_out + row * p.head_ld + h, hs[h]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------------------------
// The same GEMM on CTA PAIRS (tcgen05 cta_group::2): one 256 x 256 tile per pair of SMs.  Each CTA loads its own 128 rows of
// A and its own HALF of the B tile (128 of the 256 output columns' weight rows), the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256, N = 256) which reads B from both CTAs' shared memory and writes each CTA's 128
// accumulator rows into that CTA's TMEM: per SM and k-block 32 KiB of operands instead of 48 KiB for the same FLOPs.
// TMA loads of both CTAs signal the LEADER's `full` barrier (cp.async.bulk.tensor ... cta_group::2), tcgen05.commit
// multicasts `empty` / `acc_full` to both CTAs, the peer's epilogue warps arrive on the leader's `acc_empty` remotely.
// ------------------------------------------------------------------------------------
constexpr int kLin2Stages = 6;
struct alignas(1024) Lin2Smem {
  uint8_t a[kLin2Stages][kLinABytes];      // this CTA's 128 rows x 64 K
  uint8_t b[kLin2Stages][kLinABytes];      // this CTA's 128 weight rows (N half) x 64 K
  uint8_t stage_out[4][32 * 128];
  uint64_t full[kLin2Stages], empty[kLin2Stages];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* leader_bar) {
  // the barrier address with the peer bit cleared = the same barrier in CTA 0 of the pair (shared::cluster window)
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(leader_bar) & 0xFEFFFFFFu)
      : "memory");
}
__device__ __forceinline__ void tc_mma_ss_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {   // arrives on the barrier at this offset in BOTH CTAs
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {   // arrive on `bar` of CTA `cta` of the cluster
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

__global__ void __launch_bounds__(kLinThreads, 1)
lin_tc2_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapW, const LinParams p) {
  extern __shared__ __align__(1024) unsigned char smem_lin2[];
  Lin2Smem& sm = *reinterpret_cast<Lin2Smem*>(smem_lin2);
  if ((smem_u32(smem_lin2) & 1023u) != 0) __trap();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (tid == 0) {
    for (int s = 0; s < kLin2Stages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sm.acc_full[i], 1); mbar_init(&sm.acc_empty[i], 8); }   // 4 epilogue warps x 2 CTAs
    mbar_fence_init();
  }
  if (warp == 1) {   // the same warp of both CTAs allocates the pair's tensor memory
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&sm.tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;
  const int n_kb = p.kb0 + p.kb1;
  const int pair_tiles_m = (p.m_tiles + 1) / 2;
  const long long n_tiles = (long long)pair_tiles_m * p.n_tiles_n;       // tile_n == 256 here
  const long long cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = cluster_id; t < n_tiles; t += n_clusters) {
        const int m0 = (int)(t / p.n_tiles_n) * 256 + (int)rank * 128, n0 = (int)(t % p.n_tiles_n) * 256 + (int)rank * 128;
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(&sm.empty[stage], phase ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&sm.full[stage], 4u * kLinABytes);   // both CTAs' A and B halves
          if (kb < p.kb0) tma_load_2d_pair(sm.a[stage], &mapA0, kb * 64, m0, &sm.full[stage]);
          else tma_load_2d_pair(sm.a[stage], &mapA1, (kb - p.kb0) * 64, m0, &sm.full[stage]);
          tma_load_2d_pair(sm.b[stage], &mapW, kb * 64, n0, &sm.full[stage]);
          if (++stage == kLin2Stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {   // the leader issues every MMA of the pair
      const uint32_t idesc = umma_idesc_bf16(256, 256);
      int stage = 0;
      uint32_t phase = 0;
      long long it = 0;
      for (long long t = cluster_id; t < n_tiles; t += n_clusters, ++it) {
        const int buf = (int)(it & 1);
        mbar_wait(&sm.acc_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + 256 * buf;
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(&sm.full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t abase = smem_u32(sm.a[stage]), bbase = smem_u32(sm.b[stage]);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc_mma_ss_pair(d_tmem, umma_desc_sw128(abase + ks * 32), umma_desc_sw128(bbase + ks * 32), idesc, (kb | ks) != 0 ? 1u : 0u);
            tc_commit_pair(&sm.empty[stage]);
            if (kb == n_kb - 1) tc_commit_pair(&sm.acc_full[buf]);
          }
          __syncwarp();
          if (++stage == kLin2Stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    const int lg = warp & 3;
    const int r_in_tile = lg * 32 + lane;
    long long it = 0;
    for (long long t = cluster_id; t < n_tiles; t += n_clusters, ++it) {
      const int buf = (int)(it & 1);
      const int m0 = (int)(t / p.n_tiles_n) * 256 + (int)rank * 128, n0 = (int)(t % p.n_tiles_n) * 256;
      const long long row = (long long)m0 + r_in_tile;
      const bool row_ok = row < p.M;
      const float* rb = p.ray_bias ? p.ray_bias + (row_ok ? row / p.rows_per_ray : 0) * p.N : nullptr;
      mbar_wait(&sm.acc_full[buf], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      float hs[3] = {0.f, 0.f, 0.f};
      lin_epilogue_tile(p, tmem_base + 256 * buf, sm.stage_out[warp - 2], lg, lane, m0, n0, row, row_ok, rb, hs);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&sm.acc_empty[buf], 0);   // the leader's MMA warp waits for both CTAs' epilogues
      if (p.head_w && row_ok)
        for (int h = 0; h < p.n_heads; ++h) atomicAdd(p.head_out + row * p.head_ld + h, hs[h]);
    }
  }
  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---- host
typedef CUresult (*EncodeTiledFnMip)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnMip encode_tiled_mip() {
  static EncodeTiledFnMip fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFnMip>(ptr);
  }
  return fn;
}
// [rows, cols] bf16, row pitch `ld` elements -> boxes of 64 columns x box_rows rows, 128B swizzle
static int make_map_bf16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFnMip fn = encode_tiled_mip();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return SNERF_ERR_CUDA; }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return SNERF_ERR_CUDA; }
  return 0;
}

int linear_tc(const SnerfLinear* L, cudaStream_t stream) {
  if (!L || !L->a0 || !L->w || L->m_rows <= 0) { set_error("snerf_linear_tc: bad argument"); return SNERF_ERR_BAD_ARG; }
  if (L->k0 % 64 || L->k1 % 64 || L->k0 <= 0 || (L->k1 > 0 && !L->a1)) { set_error("snerf_linear_tc: K segments must be multiples of 64"); return SNERF_ERR_BAD_ARG; }
  if (L->n % 32 || L->n <= 0 || L->n_pad % 128 || L->n_pad < L->n) { set_error("snerf_linear_tc: N must be a multiple of 32, n_pad of 128"); return SNERF_ERR_BAD_ARG; }
  if (L->m_pad % 128 || L->m_pad < L->m_rows) { set_error("snerf_linear_tc: m_pad must be a multiple of 128"); return SNERF_ERR_BAD_ARG; }
  if (L->out && (L->ldo % 8 || (reinterpret_cast<uintptr_t>(L->out) & 15))) { set_error("snerf_linear_tc: output must be 16-byte aligned"); return SNERF_ERR_BAD_ARG; }
  if (L->n_heads < 0 || L->n_heads > 3 || (L->n_heads > 0 && (!L->head_w || !L->head_out))) { set_error("snerf_linear_tc: bad heads"); return SNERF_ERR_BAD_ARG; }
  // the epilogue reads bias / per-ray bias / head weights as 16-byte vectors
  if ((reinterpret_cast<uintptr_t>(L->bias) | reinterpret_cast<uintptr_t>(L->ray_bias) | reinterpret_cast<uintptr_t>(L->n_heads ? L->head_w : nullptr)) & 15) {
    set_error("snerf_linear_tc: bias, ray_bias and head_w must be 16-byte aligned"); return SNERF_ERR_BAD_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(L->a0) | reinterpret_cast<uintptr_t>(L->a1) | reinterpret_cast<uintptr_t>(L->w)) & 15 || L->lda0 % 8 || (L->k1 > 0 && L->lda1 % 8)) {
    set_error("snerf_linear_tc: operands must be 16-byte aligned with row pitches that are multiples of 8 elements"); return SNERF_ERR_BAD_ARG;
  }
  LinParams p{};
  p.kb0 = L->k0 / 64; p.kb1 = L->k1 / 64;
  p.tile_n = (L->n_pad % 256 == 0) ? 256 : 128;
  p.n_tiles_n = L->n_pad / p.tile_n;
  p.m_tiles = (int)(L->m_pad / 128);
  p.M = L->m_rows; p.N = L->n;
  p.bias = L->bias; p.ray_bias = L->ray_bias; p.rows_per_ray = L->rows_per_ray > 0 ? L->rows_per_ray : 1; p.relu = L->relu;
  p.out = reinterpret_cast<__nv_bfloat16*>(L->out); p.ldo = L->ldo;
  p.head_w = L->n_heads ? L->head_w : nullptr; p.n_heads = L->n_heads; p.head_out = L->head_out;
  p.head_ld = L->head_ld > 0 ? L->head_ld : L->n_heads;
  CUtensorMap mA0, mA1, mW;
  if (int e = make_map_bf16(&mA0, L->a0, L->m_pad, L->k0, L->lda0, 128)) return e;
  if (L->k1 > 0) { if (int e = make_map_bf16(&mA1, L->a1, L->m_pad, L->k1, L->lda1, 128)) return e; }
  else mA1 = mA0;
  if (int e = make_map_bf16(&mW, L->w, L->n_pad, L->k0 + L->k1, L->k0 + L->k1, p.tile_n)) return e;
  // CTA pairs (cta_group::2, 256 x 256 tiles) where the problem has full 256-column tiles and at least two row tiles
  static const int pair_env = [] { const char* e = getenv("SNERF_LIN_2SM"); return e ? atoi(e) : 1; }();
  if (pair_env && p.tile_n == 256 && p.m_tiles >= 2) {
    // the pair kernel loads 128-row boxes of the weights (each CTA its half of the 256-column tile)
    if (int e = make_map_bf16(&mW, L->w, L->n_pad, L->k0 + L->k1, L->k0 + L->k1, 128)) return e;
    const size_t smem2 = sizeof(Lin2Smem);
    if (check_cuda(cudaFuncSetAttribute(lin_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2),
                   "cudaFuncSetAttribute(lin_tc2 smem)"))
      return SNERF_ERR_CUDA;
    const long long pair_tiles = (long long)((p.m_tiles + 1) / 2) * p.n_tiles_n;
    const long long max_clusters = sm_count() / 2;
    const int clusters = (int)(pair_tiles < max_clusters ? pair_tiles : max_clusters);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2 * clusters));
    cfg.blockDim = dim3(kLinThreads);
    cfg.dynamicSmemBytes = smem2;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return check_cuda(cudaLaunchKernelEx(&cfg, lin_tc2_kernel, mA0, mA1, mW, p), "launch lin_tc2_kernel");
  }
  const size_t smem = sizeof(LinSmem);
  if (check_cuda(cudaFuncSetAttribute(lin_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                 "cudaFuncSetAttribute(lin_tc smem)"))
    return SNERF_ERR_CUDA;
  const long long tiles = (long long)p.m_tiles * p.n_tiles_n;
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  lin_tc_kernel<<<grid, kLinThreads, smem, stream>>>(mA0, mA1, mW, p);
  return check_cuda(cudaGetLastError(), "launch lin_tc_kernel");
}

__global__ void __launch_bounds__(256) rows_to_bf16_kernel(const float* __restrict__ x, long long rows, int row_stride, int col0,
                                                           int ncols, int repeat, __nv_bfloat16* __restrict__ out, int out_cols,
                                                           long long m_pad) {
  const long long total = m_pad * (out_cols / 2);
  for (long long idx = blockIdx.x * 256ll + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    const long long r = idx / (out_cols / 2);
    const int c = (int)(idx % (out_cols / 2)) * 2;
    float a = 0.f, b = 0.f;
    if (r < rows * repeat) {
      const float* src = x + (r / repeat) * row_stride + col0;
      if (c < ncols) a = src[c];
      if (c + 1 < ncols) b = src[c + 1];
    }
    reinterpret_cast<__nv_bfloat162*>(out)[idx] = __floats2bfloat162_rn(a, b);
  }
}

int rows_to_bf16(const float* x, long long rows, int row_stride, int col0, int ncols, int repeat, void* out, int out_cols,
                 long long m_pad, cudaStream_t stream) {
  if (!x || !out || rows < 0 || ncols < 1 || out_cols < ncols || out_cols % 2 || repeat < 1 || m_pad < rows * repeat) {
    set_error("snerf_rows_to_bf16: bad argument"); return SNERF_ERR_BAD_ARG;
  }
  if (m_pad == 0) return SNERF_OK;
  const long long total = m_pad * (out_cols / 2);
  const long long want = (total + 255) / 256;
  const int grid = (int)(want < 148ll * 16 ? want : 148ll * 16);
  rows_to_bf16_kernel<<<grid, 256, 0, stream>>>(x, rows, row_stride, col0, ncols, repeat, reinterpret_cast<__nv_bfloat16*>(out), out_cols, m_pad);
  return check_cuda(cudaGetLastError(), "launch rows_to_bf16_kernel");
}

int mip_encode(const SnerfMipEncode* e, cudaStream_t stream) {
  if (!e || !e->rays || !e->enc || e->n_rays < 0) { set_error("snerf_mip_encode: bad argument"); return SNERF_ERR_BAD_ARG; }
  if (e->n_samples < 1 || e->n_samples > kMipMaxS || e->rows_per_ray < e->n_samples || e->max_deg < 1 || e->max_deg * 6 > 128) {
    set_error("snerf_mip_encode: need 1 <= n_samples <= %d <= rows_per_ray and 6 * max_deg <= 128", kMipMaxS); return SNERF_ERR_UNSUPPORTED;
  }
  if (!e->s_in && (!e->s_lin || !e->s_out)) { set_error("snerf_mip_encode: s_lin / s_out missing"); return SNERF_ERR_BAD_ARG; }
  if (e->m_pad < e->n_rays * e->rows_per_ray) { set_error("snerf_mip_encode: m_pad too small"); return SNERF_ERR_BAD_ARG; }
  if (e->m_pad == 0) return SNERF_OK;
  MipEncodeParams p{};
  p.rays = e->rays; p.n_rays = e->n_rays; p.S = e->n_samples; p.rows_per_ray = e->rows_per_ray;
  p.s_lin = e->s_lin; p.s_rand = e->s_rand; p.s_in = e->s_in; p.s_out = e->s_out;
  p.transform_idx = e->transform_idx; p.max_deg = e->max_deg; p.ray_cone = e->ray_cone; p.radius = e->radius;
  p.enc = reinterpret_cast<__nv_bfloat16*>(e->enc); p.enc_f32 = e->enc_f32; p.m_pad = e->m_pad;
  mip_encode_kernel<<<(unsigned)((e->m_pad + 127) / 128), 128, 0, stream>>>(p);
  return check_cuda(cudaGetLastError(), "launch mip_encode_kernel");
}

int mip_cond_bias(const float* viewdirs, long long n_rays, int deg_view, const float* w, int ldw, int k0, const float* b, int n_out,
                  float* out, cudaStream_t stream) {
  if (!viewdirs || !w || !b || !out || n_rays < 0 || n_out <= 0) { set_error("snerf_mip_cond_bias: bad argument"); return SNERF_ERR_BAD_ARG; }
  if (n_rays == 0) return SNERF_OK;
  mip_cond_bias_kernel<<<(unsigned)((n_rays * n_out + 127) / 128), 128, 0, stream>>>(viewdirs, n_rays, deg_view, w, ldw, k0, b, n_out, out);
  return check_cuda(cudaGetLastError(), "launch mip_cond_bias_kernel");
}

int mip_composite(const SnerfMipComposite* c, cudaStream_t stream) {
  if (!c || !c->rays || !c->s_vals || !c->raw_density || !c->distance || !c->acc) { set_error("snerf_mip_composite: bad argument"); return SNERF_ERR_BAD_ARG; }
  if (c->n_samples < 1 || c->n_samples > kMipMaxS || c->rows_per_ray < c->n_samples) { set_error("snerf_mip_composite: bad sample count"); return SNERF_ERR_UNSUPPORTED; }
  if (c->s_new && (c->n_fine < 1 || (!c->u_lin && !c->u_rand))) { set_error("snerf_mip_composite: resampling needs n_fine and u"); return SNERF_ERR_BAD_ARG; }
  if (c->n_rays == 0) return SNERF_OK;
  MipCompositeParams p{};
  p.rays = c->rays; p.n_rays = c->n_rays; p.S = c->n_samples; p.rows_per_ray = c->rows_per_ray; p.s_vals = c->s_vals;
  p.raw_density = c->raw_density; p.raw_rgb = c->raw_rgb; p.noise = c->noise;
  p.density_head_bias = c->density_head_bias; p.density_bias = c->density_bias; p.rgb_padding = c->rgb_padding;
  for (int i = 0; i < 3; ++i) p.rgb_head_bias[i] = c->rgb_head_bias[i];
  p.transform_idx = c->transform_idx; p.white_bkgd = c->white_bkgd;
  p.comp_rgb = c->comp_rgb; p.distance = c->distance; p.acc = c->acc; p.weights = c->weights;
  p.n_fine = c->n_fine; p.u_lin = c->u_lin; p.u_rand = c->u_rand; p.resample_padding = c->resample_padding; p.s_new = c->s_new;
  mip_composite_kernel<<<(unsigned)((c->n_rays + 3) / 4), 128, 0, stream>>>(p);
  return check_cuda(cudaGetLastError(), "launch mip_composite_kernel");
}

}  // namespace snerf
