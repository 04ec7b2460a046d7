// snerf_train.cu -- training path (fp32): d(render_rays outputs)/d(network parameters).
//
// The reference differentiates its eager ops with torch autograd (train step of BASELINE config 3:
// render.py:281-409 under perturb=1 / raw_noise_std=1, z_samples detached at render.py:381).  Here the forward
// kernel (snerf_fp32.cu, save_for_backward) keeps each tile's layer outputs in a channel-major store [channel][R]
// and the backward is four launches:
//   1. composite_bwd_kernel  one warp per (ray, pass): d(rgb/disp/acc/depth/weights) -> d_raw[4][R]
//      (run_nerf_helpers.py:381-424 differentiated by hand; transmittance and suffix sums in fp64);
//   2. mlp_bwd_kernel        per 64-row tile, the chain dZ_l -> dZ_{l-1} = (dZ_l . W_l) * relu'(h_{l-1}) with the same
//      FFMA tile GEMM as the forward, un-transposed weights streamed through the ring; every dZ_l goes to the
//      gradient store [channel][R];
//   3. dw_gemm_kernel        grouped split-K GEMM  dW_l[n][k] += sum_r dZ_l[n][r] * X_l[k][r]  (both operands are
//      r-contiguous in the stores), 128x128x16 tiles, 8x8 per thread, atomicAdd into the nn.Linear-shaped gradients;
//   4. skinny_kernel         bias gradients (row sums of dZ) and the 1- and 3-row heads (alpha_linear, rgb_linear).
// Gradients are ACCUMULATED into the caller's buffers (like autograd's .grad), summation order over rays is not fixed.
#include <cuda.h>  // CUtensorMap + cuTensorMapEncodeTiled prototype only; resolved at run time (no libcuda link)

#include "snerf_fp32_core.cuh"
#include <stdlib.h>

#include "snerf_internal.h"
#include "snerf_umma.cuh"

namespace snerf {

// ------------------------------------------------------------------------------------
// host: channel numbering, workspace layout, backward plan
// ------------------------------------------------------------------------------------
static int round_up_i(int x, int m) { return (x + m - 1) / m * m; }

TrainChannels train_channels(const SnerfNetDesc* d) {
  TrainChannels c{};
  for (int i = 0; i < d->D; ++i) c.trunk[i] = kSaveActCh + i * d->W;
  c.feature = kSaveActCh + d->D * d->W;
  c.views = c.feature + d->W;
  c.total = d->use_viewdirs ? c.views + d->W / 2 : c.feature;   // output_linear reads the trunk directly
  return c;
}

TrainLayout train_layout(const SnerfNetDesc* d, const SnerfNetDesc* df, int Nc, int Nf, long long n_rays) {
  TrainLayout L{};
  const TrainChannels ch = train_channels(d), chf = train_channels(df ? df : d);
  L.TC = (Nc + kTileRows - 1) / kTileRows;
  L.TF = Nf > 0 ? (Nc + Nf + kTileRows - 1) / kTileRows : 0;
  L.Rc = n_rays * L.TC * kTileRows;
  L.Rf = n_rays * L.TF * kTileRows;
  const long long S = Nc + Nf;
  size_t off = 0;
  auto take = [&](long long n) { size_t o = off; off += ((size_t)n + 31) / 32 * 32; return o; };  // 128-byte granules
  L.save_c = take((long long)ch.total * L.Rc);
  L.save_f = take((long long)chf.total * L.Rf);
  L.dz_c = take((long long)(ch.total - kSaveActCh) * L.Rc);
  L.dz_f = take((long long)(chf.total - kSaveActCh) * L.Rf);
  L.draw_c = take(4 * L.Rc);
  L.draw_f = take(4 * L.Rf);
  L.raw_c = take(n_rays * Nc * 4);
  L.raw_f = take(Nf > 0 ? n_rays * S * 4 : 0);
  L.z_c = take(n_rays * Nc);
  L.z_f = take(Nf > 0 ? n_rays * S : 0);
  L.total_floats = off;
  return L;
}

bool train_supported(const SnerfNetDesc* d, int tf32) {
  if (!d->use_viewdirs && tf32) {
    set_error("tf32 training needs use_viewdirs=True (alpha/feature/views/rgb heads); use train precision fp32");
    return false;
  }
  return true;
}

size_t plan_bwd(const SnerfNetDesc* d, Fp32BwdHeader* h) {
  memset(h, 0, sizeof(*h));
  const TrainChannels ch = train_channels(d);
  h->magic = kFp32BwdMagic;
  h->W = d->W;
  h->n_channels = ch.total;
  uint32_t off = kFp32DataOffset / 4;
  int ns = 0, buf = 0;
  if (!d->use_viewdirs) {
    // output_linear[:4]^T : d_raw[0:4] -> d(h_{D-1}), masked by the last trunk ReLU (run_nerf_helpers.py:124)
    BwdStep& s = h->steps[ns++];
    s.kind = 1; s.K = 4; s.n_out = d->W; s.src = buf; s.dst = buf; s.raw_col = 0;
    s.mask_ch = ch.trunk[d->D - 1]; s.dz_ch = ch.trunk[d->D - 1]; s.w_off = off; s.add_col = -1;
    off += (uint32_t)(4 * d->W);
    auto wide = [&](int mask_ch) {
      BwdStep& t = h->steps[ns++];
      off = (uint32_t)round_up_i((int)off, d->W);
      t.kind = 0; t.K = d->W; t.n_out = d->W; t.src = buf; t.dst = buf ^ 1; t.raw_col = 0;
      t.mask_ch = mask_ch; t.dz_ch = mask_ch; t.w_off = off; t.add_col = -1;
      off += (uint32_t)d->W * d->W;
      buf ^= 1;
    };
    for (int l = d->D - 1; l >= 1; --l) wide(ch.trunk[l - 1]);
    h->n_steps = ns;
    return (size_t)off * 4;
  }
  {  // rgb_linear^T : d_raw[0:3] -> d(views output), masked by the views ReLU
    BwdStep& s = h->steps[ns++];
    s.kind = 1; s.K = 3; s.n_out = d->W / 2; s.src = buf; s.dst = buf; s.raw_col = 0;
    s.mask_ch = ch.views; s.dz_ch = ch.views; s.w_off = off; s.add_col = -1;
    off += (uint32_t)round_up_i(3 * (d->W / 2), 4);
  }
  auto add_wide = [&](int K, int mask_ch, int dz_ch) -> BwdStep& {
    BwdStep& s = h->steps[ns++];
    off = (uint32_t)round_up_i((int)off, d->W);  // whole rows of the [rows][W] view the TMA descriptor uses
    s.kind = 0; s.K = K; s.n_out = d->W; s.src = buf; s.dst = buf ^ 1; s.raw_col = 0;
    s.mask_ch = mask_ch; s.dz_ch = dz_ch; s.w_off = off; s.add_col = -1;
    off += (uint32_t)K * d->W;
    buf ^= 1;
    return s;
  };
  add_wide(d->W / 2, -1, ch.feature);  // views_linears.0[:, :W]^T : d(views pre-act) -> d(feature)
  {  // feature_linear^T (+ alpha_linear^T on d_raw[3]) -> d(h_{D-1}), masked by the last trunk ReLU
    BwdStep& s = add_wide(d->W, ch.trunk[d->D - 1], ch.trunk[d->D - 1]);
    s.add_col = 3; s.add_w_off = off; off += (uint32_t)d->W;
  }
  for (int l = d->D - 1; l >= 1; --l) add_wide(d->W, ch.trunk[l - 1], ch.trunk[l - 1]);
  h->n_steps = ns;
  return (size_t)off * 4;
}

// round-to-nearest onto the tf32 grid (10 mantissa bits).  The tensor core TRUNCATES fp32 operands to tf32, which
// biases every product towards zero (~2^-11 each); operands that are already tf32 values pass through exactly.
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__global__ void write_bwd_header_kernel(Fp32BwdHeader h, Fp32BwdHeader* dst) {
  const int n = sizeof(Fp32BwdHeader) / 4;
  const int* s = reinterpret_cast<const int*>(&h);
  int* d = reinterpret_cast<int*>(dst);
  for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
}

int pack_bwd(const SnerfNetDesc* d, const SnerfNetF32* src, void* packed, int tf32, cudaStream_t stream) {
  if (!train_supported(d, tf32)) return SNERF_ERR_UNSUPPORTED;
  if (d->use_viewdirs && !src->alpha_w && tf32) {
    set_error("tf32 training of a network without alpha_linear (NeRF_RGB) is not supported; use train precision fp32");
    return SNERF_ERR_UNSUPPORTED;
  }
  if (!d->use_viewdirs && !src->output_w) { set_error("output_linear weights missing"); return SNERF_ERR_BAD_ARG; }
  Fp32BwdHeader h;
  plan_bwd(d, &h);
  float* base = reinterpret_cast<float*>(packed);
  write_bwd_header_kernel<<<1, 128, 0, stream>>>(h, reinterpret_cast<Fp32BwdHeader*>(packed));
  int s = 0;
  const int W = d->W;
  PackJobs jobs{};
  auto block = [&](const float* w, int ld, int col0, int rows, int cols, uint32_t off, int round = -1) {
    if (jobs.n == kMaxPackJobs) { launch_pack_jobs(jobs, stream); jobs.n = 0; }   // table full: flush
    PackJob& J = jobs.j[jobs.n++];
    J.w = w; J.dst = base + off; J.kind = 1; J.ld = ld; J.col_first = col0; J.rows = rows; J.cols = cols;
    J.round_tf32 = round < 0 ? tf32 : round;
  };
  if (!d->use_viewdirs) {
    block(src->output_w, W, 0, 4, W, h.steps[s++].w_off, 0);
    for (int l = d->D - 1; l >= 1; --l, ++s) {
      const bool has_enc = d->skip >= 0 && l - 1 == d->skip;
      block(src->pts_w[l], (has_enc ? d->input_ch : 0) + W, has_enc ? d->input_ch : 0, W, W, h.steps[s].w_off);
    }
    if (int e = launch_pack_jobs(jobs, stream)) return e;
    return check_cuda(cudaGetLastError(), "pack backward image");
  }
  block(src->rgb_w, W / 2, 0, 3, W / 2, h.steps[s++].w_off, 0);   // the two narrow heads stay on CUDA cores
  block(src->views_w, W + d->input_ch_views, 0, W / 2, W, h.steps[s++].w_off);
  block(src->feature_w, W, 0, W, W, h.steps[s].w_off);
  block(src->alpha_w, W, 0, 1, W, h.steps[s].add_w_off, 0);   // (NeRF_RGB: no alpha head -> zeros, d_raw[3] stops here)
  ++s;
  for (int l = d->D - 1; l >= 1; --l, ++s) {
    const bool has_enc = d->skip >= 0 && l - 1 == d->skip;
    block(src->pts_w[l], (has_enc ? d->input_ch : 0) + W, has_enc ? d->input_ch : 0, W, W, h.steps[s].w_off);
  }
  if (int e = launch_pack_jobs(jobs, stream)) return e;
  return check_cuda(cudaGetLastError(), "pack backward image");
}

// ------------------------------------------------------------------------------------
// 1. compositing backward
// ------------------------------------------------------------------------------------
__device__ __forceinline__ double shfl_down_d(double v, int delta) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_down_sync(0xffffffffu, lo, delta);
  hi = __shfl_down_sync(0xffffffffu, hi, delta);
  return __hiloint2double(hi, lo);
}

__global__ void __launch_bounds__(128) composite_bwd_kernel(const TrainParams p) {
  const int lane = threadIdx.x & 31;
  const long long wid = blockIdx.x * 4ll + (threadIdx.x >> 5);
  const int passes = p.Nf > 0 ? 2 : 1;
  if (wid >= p.n_rays * passes) return;
  const int pass = (int)(wid / p.n_rays);
  const long long ray = wid % p.n_rays;
  const int S = pass ? p.Nc + p.Nf : p.Nc;
  const int T = pass ? p.TF : p.TC;
  const long long R = pass ? p.Rf : p.Rc;
  const float4* raw = reinterpret_cast<const float4*>(pass ? p.raw_f : p.raw_c) + ray * S;
  const float* z = (pass ? p.z_f : p.z_c) + ray * S;
  const float* noise = pass ? p.noise1 : p.noise0;
  if (noise) noise += ray * S;
  float4* draw4 = pass ? p.draw4_f : p.draw4_c;   // tensor-core path: d_raw [ray * S + s][4], no tile padding
  if (draw4) draw4 += ray * S;
  float* draw = draw4 ? nullptr : (pass ? p.draw_f : p.draw_c) + ray * T * kTileRows;
  const bool final_pass = pass == 1 || p.Nf == 0;
  const float* g_rgb = final_pass ? p.g.rgb_map : p.g.rgb0;
  const float* g_disp = final_pass ? p.g.disp_map : p.g.disp0;
  const float* g_acc = final_pass ? p.g.acc_map : p.g.acc0;
  const float* g_depth = final_pass ? p.g.depth_map : p.g.depth0;
  const float* g_w = pass == 0 ? p.g.weights : nullptr;
  const float* g_raw = final_pass ? p.g.raw : nullptr;
  const Ray rayv = load_ray(p.ray_batch + ray * p.row_stride, p.width, 0);

  const int C = (S + 31) >> 5;
  const int i0 = lane * C;
  float alpha[8], ex[8], dist[8], w[8];
  double Tj[8];
  double prod = 1.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    alpha[j] = 0.f; ex[j] = 1.f; dist[j] = 0.f; w[j] = 0.f; Tj[j] = 0.0;
    const int s = i0 + j;
    if (j < C && s < S) {
      float dd = (s == S - 1) ? kHuge : __fsub_rn(z[s + 1], z[s]);
      dd = __fmul_rn(dd, rayv.dnorm);
      float sig = raw[s].w;
      if (noise) sig = __fadd_rn(sig, noise[s]);
      const float e = expf(__fmul_rn(-fmaxf(sig, 0.f), dd));
      dist[j] = sig > 0.f ? dd : 0.f;  // relu'(sigma) folded into the distance
      ex[j] = e;
      alpha[j] = __fsub_rn(1.f, e);
      prod *= (double)__fadd_rn(__fsub_rn(1.f, alpha[j]), 1e-10f);
    }
  }
  const double incl = warp_scan_prod_d(prod, lane);
  double excl = shfl_up_d(incl, 1);
  if (lane == 0) excl = 1.0;
  double Trun = excl;
  float sd = 0.f, sa = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int s = i0 + j;
    if (j < C && s < S) {
      Tj[j] = Trun;
      w[j] = __fmul_rn(alpha[j], (float)Trun);
      Trun *= (double)__fadd_rn(__fsub_rn(1.f, alpha[j]), 1e-10f);
      sd += w[j] * z[s];
      sa += w[j];
    }
  }
  const float depth = warp_sum(sd), acc = warp_sum(sa);
  // upstream gradients of this ray
  float gr[3] = {0.f, 0.f, 0.f};
  if (g_rgb) { gr[0] = g_rgb[ray * 3 + 0]; gr[1] = g_rgb[ray * 3 + 1]; gr[2] = g_rgb[ray * 3 + 2]; }
  float gd = g_depth ? g_depth[ray] : 0.f;
  float ga = g_acc ? g_acc[ray] : 0.f;
  if (g_disp) {  // disp = 1 / max(1e-10, depth / acc)
    const float q = depth / acc;
    if (q > 1e-10f) {
      const float dq = -g_disp[ray] / (q * q);
      gd += dq / acc;
      ga += -dq * depth / (acc * acc);
    }
  }
  if (p.white_bkgd) ga -= (gr[0] + gr[1]) + gr[2];
  // G_i = dL/dw_i, local sums of G_i w_i
  float G[8], cr[8][3];
  double loc = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    G[j] = 0.f; cr[j][0] = cr[j][1] = cr[j][2] = 0.f;
    const int s = i0 + j;
    if (j < C && s < S) {
      const float4 q = raw[s];
      cr[j][0] = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-q.x)));
      cr[j][1] = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-q.y)));
      cr[j][2] = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-q.z)));
      float g = gr[0] * cr[j][0] + gr[1] * cr[j][1] + gr[2] * cr[j][2] + gd * z[s] + ga;
      if (g_w) g += g_w[ray * S + s];
      G[j] = g;
      loc += (double)g * (double)w[j];
    }
  }
  const double pre = warp_scan_sum_d(loc, lane);
  double suffix = shfl_d(pre, 31) - pre;  // sum over higher lanes
#pragma unroll
  for (int j = 7; j >= 0; --j) {
    const int s = i0 + j;
    if (j < C && s < S) {
      const double f = (double)__fadd_rn(__fsub_rn(1.f, alpha[j]), 1e-10f);
      const double dLda = (double)G[j] * Tj[j] - suffix / f;
      suffix += (double)G[j] * (double)w[j];
      float ds = (float)(dLda * (double)dist[j] * (double)ex[j]);
      float d0 = w[j] * gr[0] * cr[j][0] * (1.f - cr[j][0]);
      float d1 = w[j] * gr[1] * cr[j][1] * (1.f - cr[j][1]);
      float d2 = w[j] * gr[2] * cr[j][2] * (1.f - cr[j][2]);
      if (g_raw) {
        const float4 t = reinterpret_cast<const float4*>(g_raw)[ray * S + s];
        d0 += t.x; d1 += t.y; d2 += t.z; ds += t.w;
      }
      if (draw4) draw4[s] = make_float4(d0, d1, d2, ds);
      else { draw[0 * R + s] = d0; draw[1 * R + s] = d1; draw[2 * R + s] = d2; draw[3 * R + s] = ds; }
    }
  }
  if (draw4) return;
  for (int s = S + lane; s < T * kTileRows; s += 32) {  // padding rows of the last tile
    draw[0 * R + s] = 0.f; draw[1 * R + s] = 0.f; draw[2 * R + s] = 0.f; draw[3 * R + s] = 0.f;
  }
}

// ------------------------------------------------------------------------------------
// 2. MLP backward (d_raw -> dZ of every layer), one 64-row tile at a time
// ------------------------------------------------------------------------------------
template <int W>
struct alignas(128) BwdSmem {
  float wstage[kStages][kFp32ChunkRows * W];
  float actX[W * kLd];
  float actY[W * kLd];
  float draw[4][kTileRows];
  BwdStep steps[2][kBwdMaxSteps];
  int n_steps[2];
  uint64_t full[kStages];
  uint64_t empty[kStages];
};

// (NJ = n_out / 32 of THIS network: a CTA sized for the wider of two architectures also runs the narrower one)
template <int W, int NJ>
__device__ __forceinline__ void bwd_wide(BwdSmem<W>& sm, const Fp32Ring& rg, const BwdStep& S, const float* imgf,
                                         const float* save, float* dz, long long R, int& stage, uint32_t& phase,
                                         int warp, int lane) {
  float acc[8][NJ];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;
  const int r0 = warp * 8;
  const float* src = S.src ? sm.actY : sm.actX;
  float* dst = S.dst ? sm.actY : sm.actX;
  ring_gemm<NJ>(rg, src + r0, S.K, acc, stage, phase, lane);
  float dr[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) dr[i] = S.add_col >= 0 ? sm.draw[S.add_col][r0 + i] : 0.f;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int k = lane + 32 * j;
    float v[8];
    const float aw = S.add_col >= 0 ? __ldg(imgf + S.add_w_off + k) : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaf(dr[i], aw, acc[i][j]);
    if (S.mask_ch >= 0) {
      const float* m = save + (long long)(S.mask_ch + k) * R + r0;
      const float4 m0 = __ldg(reinterpret_cast<const float4*>(m));
      const float4 m1 = __ldg(reinterpret_cast<const float4*>(m + 4));
      v[0] = m0.x > 0.f ? v[0] : 0.f; v[1] = m0.y > 0.f ? v[1] : 0.f;
      v[2] = m0.z > 0.f ? v[2] : 0.f; v[3] = m0.w > 0.f ? v[3] : 0.f;
      v[4] = m1.x > 0.f ? v[4] : 0.f; v[5] = m1.y > 0.f ? v[5] : 0.f;
      v[6] = m1.z > 0.f ? v[6] : 0.f; v[7] = m1.w > 0.f ? v[7] : 0.f;
    }
    float* d = dst + k * kLd + r0;
    *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(d + 4) = make_float4(v[4], v[5], v[6], v[7]);
    float* g = dz + (long long)(S.dz_ch + k) * R + r0;
    *reinterpret_cast<float4*>(g) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(g + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

template <int W>
__global__ void __launch_bounds__(kFp32Threads, 1) mlp_bwd_kernel(const TrainParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BwdSmem<W>& sm = *reinterpret_cast<BwdSmem<W>*>(smem_raw);
  const int tid = threadIdx.x;
  const unsigned char* img[2] = {p.bwd_c, p.bwd_f ? p.bwd_f : p.bwd_c};
  Fp32Ring rg;
  rg.wstage = &sm.wstage[0][0]; rg.full = sm.full; rg.empty = sm.empty; rg.stage_floats = kFp32ChunkRows * W;
  {
    constexpr int nint = kBwdMaxSteps * (int)sizeof(BwdStep) / 4;
    for (int net = 0; net < 2; ++net) {
      const Fp32BwdHeader* h = reinterpret_cast<const Fp32BwdHeader*>(img[net]);
      const int* src = reinterpret_cast<const int*>(h->steps);
      int* dst = reinterpret_cast<int*>(sm.steps[net]);
      for (int i = tid; i < nint; i += kFp32Threads) dst[i] = src[i];
      if (tid == 0) {
        if (h->magic != kFp32BwdMagic) __trap();
        sm.n_steps[net] = h->n_steps;
      }
    }
    if (tid == 0) ring_init(rg);
  }
  __syncthreads();
  int stage = 0;
  uint32_t phase = 0;
  const long long tiles_c = p.n_rays * p.TC, tiles = tiles_c + p.n_rays * p.TF;

  if (tid >= kComputeThreads) {
    if (tid == kComputeThreads) {
      for (long long u = blockIdx.x; u < tiles; u += gridDim.x) {
        const int net = u >= tiles_c;
        const float* imgf = reinterpret_cast<const float*>(img[net]);
        for (int s = 0; s < sm.n_steps[net]; ++s) {
          const BwdStep& S = sm.steps[net][s];
          if (S.kind == 0) ring_stream(rg, imgf + S.w_off, S.K, S.n_out, stage, phase);
        }
      }
    }
    return;
  }
  const int warp = tid >> 5, lane = tid & 31;
  for (long long u = blockIdx.x; u < tiles; u += gridDim.x) {
    const int net = u >= tiles_c;
    const long long R = net ? p.Rf : p.Rc;
    const long long row0 = (net ? u - tiles_c : u) * kTileRows;
    const float* save = (net ? p.save_f : p.save_c) + row0;
    float* dz = (net ? p.dz_f : p.dz_c) + row0;
    const float* drawg = (net ? p.draw_f : p.draw_c) + row0;
    const float* imgf = reinterpret_cast<const float*>(img[net]);
    sm.draw[tid >> 6][tid & 63] = drawg[(long long)(tid >> 6) * R + (tid & 63)];
    named_bar_sync(1, kComputeThreads);
    for (int s = 0; s < sm.n_steps[net]; ++s) {
      const BwdStep& S = sm.steps[net][s];
      if (S.kind == 1) {
        float* dst = S.dst ? sm.actY : sm.actX;
        const int r = tid & 63;
        const float* w = imgf + S.w_off;
        for (int k = tid >> 6; k < S.n_out; k += 4) {
          float v = 0.f;
          for (int c = 0; c < S.K; ++c) v = fmaf(sm.draw[S.raw_col + c][r], __ldg(w + c * S.n_out + k), v);
          if (S.mask_ch >= 0 && !(save[(long long)(S.mask_ch + k) * R + r] > 0.f)) v = 0.f;
          dst[k * kLd + r] = v;
          dz[(long long)(S.dz_ch + k) * R + r] = v;
        }
      } else {
        if (S.n_out == W) bwd_wide<W, W / 32>(sm, rg, S, imgf, save, dz, R, stage, phase, warp, lane);
        else if (W >= 128 && S.n_out == W / 2) bwd_wide<W, (W >= 128 ? W / 64 : 1)>(sm, rg, S, imgf, save, dz, R, stage, phase, warp, lane);
        else if (W >= 256 && S.n_out == W / 4) bwd_wide<W, (W >= 256 ? W / 128 : 1)>(sm, rg, S, imgf, save, dz, R, stage, phase, warp, lane);
        else __trap();
      }
      named_bar_sync(1, kComputeThreads);
    }
  }
}

// ------------------------------------------------------------------------------------
// 3. grouped split-K weight-gradient GEMM:  C[m][n] += sum_r A[m][r] * B[n][r]
// ------------------------------------------------------------------------------------
constexpr int kDwTile = 128;
constexpr int kDwLd = kDwTile + 4;

__global__ void __launch_bounds__(256) dw_gemm_kernel(const DwTable tab) {
  __shared__ __align__(16) float As[16][kDwLd];
  __shared__ __align__(16) float Bs[16][kDwLd];
  const int tid = threadIdx.x;
  int pi = 0;
  while (pi + 1 < tab.n && (int)blockIdx.x >= tab.p[pi + 1].first) ++pi;
  const DwProblem& P = tab.p[pi];
  int local = blockIdx.x - P.first;
  const int per_split = P.mt * P.nt;
  const int split = local / per_split;
  local -= split * per_split;
  const int m0 = (local / P.nt) * kDwTile, n0 = (local % P.nt) * kDwTile;
  const long long r_begin = (long long)split * P.rows_per_split;
  const long long r_end = r_begin + P.rows_per_split < P.R ? r_begin + P.rows_per_split : P.R;

  const int lr = tid >> 2, rq = (tid & 3) * 4;   // loader: row lr (+64), r offset rq
  const float* a0p = P.A + (long long)(m0 + lr) * P.R + rq;
  const float* a1p = P.A + (long long)(m0 + lr + 64) * P.R + rq;
  const float* b0p = P.B + (long long)(n0 + lr) * P.R + rq;
  const float* b1p = P.B + (long long)(n0 + lr + 64) * P.R + rq;
  const bool va0 = m0 + lr < P.M, va1 = m0 + lr + 64 < P.M, vb0 = n0 + lr < P.N, vb1 = n0 + lr + 64 < P.N;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra0 = zero4, ra1 = zero4, rb0 = zero4, rb1 = zero4;
  if (r_begin < r_end) {
    ra0 = va0 ? __ldg(reinterpret_cast<const float4*>(a0p + r_begin)) : zero4;
    ra1 = va1 ? __ldg(reinterpret_cast<const float4*>(a1p + r_begin)) : zero4;
    rb0 = vb0 ? __ldg(reinterpret_cast<const float4*>(b0p + r_begin)) : zero4;
    rb1 = vb1 ? __ldg(reinterpret_cast<const float4*>(b1p + r_begin)) : zero4;
  }
  for (long long r = r_begin; r < r_end; r += 16) {
    As[rq + 0][lr] = ra0.x; As[rq + 1][lr] = ra0.y; As[rq + 2][lr] = ra0.z; As[rq + 3][lr] = ra0.w;
    As[rq + 0][lr + 64] = ra1.x; As[rq + 1][lr + 64] = ra1.y; As[rq + 2][lr + 64] = ra1.z; As[rq + 3][lr + 64] = ra1.w;
    Bs[rq + 0][lr] = rb0.x; Bs[rq + 1][lr] = rb0.y; Bs[rq + 2][lr] = rb0.z; Bs[rq + 3][lr] = rb0.w;
    Bs[rq + 0][lr + 64] = rb1.x; Bs[rq + 1][lr + 64] = rb1.y; Bs[rq + 2][lr + 64] = rb1.z; Bs[rq + 3][lr + 64] = rb1.w;
    __syncthreads();
    const long long rn = r + 16;
    if (rn < r_end) {
      ra0 = va0 ? __ldg(reinterpret_cast<const float4*>(a0p + rn)) : zero4;
      ra1 = va1 ? __ldg(reinterpret_cast<const float4*>(a1p + rn)) : zero4;
      rb0 = vb0 ? __ldg(reinterpret_cast<const float4*>(b0p + rn)) : zero4;
      rb1 = vb1 ? __ldg(reinterpret_cast<const float4*>(b1p + rn)) : zero4;
    }
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) {
      const float4 x0 = *reinterpret_cast<const float4*>(&As[rr][ty * 4]);
      const float4 x1 = *reinterpret_cast<const float4*>(&As[rr][64 + ty * 4]);
      const float4 y0 = *reinterpret_cast<const float4*>(&Bs[rr][tx * 4]);
      const float4 y1 = *reinterpret_cast<const float4*>(&Bs[rr][64 + tx * 4]);
      const float a[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      const float b[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (m >= P.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
      if (n < P.N) atomicAdd(P.C + (long long)m * P.ldc + n, acc[i][j]);
    }
  }
}

// ------------------------------------------------------------------------------------
// 3b. the same weight-gradient GEMM on the tensor cores (tcgen05, kind::tf32), opt-in (SNERF_MODE_TF32):
//   C[m][n] += sum_r A[m][r] * B[n][r]  with both operands r-contiguous in the stores = K-major for the MMA.
//   One CTA owns the whole [M <= 256] x [N <= 256] block of one layer for one slice of rows: 512 TMEM columns hold
//   the two 128 x N fp32 accumulators.  Operand tiles are 128 channels x 32 rows of fp32 (128-byte lines), fetched
//   by TMA (cp.async.bulk.tensor.2d, 128B swizzle) straight from the [channel][R] stores into a 3-stage ring;
//   warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue (TMEM -> red.global.add into the gradients).
//   Arithmetic intensity is 65 FLOP/B, so this GEMM is HBM-bound: its floor is (activation + gradient stores) / HBM
//   bandwidth.
// ------------------------------------------------------------------------------------
constexpr int kTfStages = 3;
constexpr int kTfBoxRows = 128;                        // channels per TMA box
constexpr int kTfBoxCols = 32;                         // rows (K) per TMA box: 32 fp32 = 128 B
constexpr int kTfBoxBytes = kTfBoxRows * kTfBoxCols * 4;  // 16 KiB
constexpr int kTfThreads = 192;

struct alignas(1024) TfSmem {
  uint8_t stage[kTfStages][4 * kTfBoxBytes];  // A0 | A1 | B0 | B1
  uint64_t full[kTfStages];
  uint64_t empty[kTfStages];
  uint64_t done;
  uint32_t tmem_base;
};

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
// instruction descriptor, kind::tf32: D=f32, A=B=tf32 (format code 2), K-major both
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct TfProblem {
  int mapA, mapB;       // which tensor map (0..3) the operands live in
  int chA, chB;         // first channel (row of the tensor map) of each operand
  int M, N;             // output rows (128 or 256; 64 -> one box, upper half ignored) / columns actually wanted
  int Nmma;             // N of the MMA: multiple of 16, >= N
  int ldc, vec4;        // leading dimension of C; 1 = rows of C are 16-byte aligned (vector reductions)
  long long R;
  float* C;
  int rows_per_split, splits, first;
};
constexpr int kMaxTfProblems = 32;
struct TfTable { int n; TfProblem p[kMaxTfProblems]; };

__global__ void __launch_bounds__(kTfThreads, 1)
dw_tf32_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
               const __grid_constant__ CUtensorMap map2, const __grid_constant__ CUtensorMap map3, const TfTable tab) {
  extern __shared__ __align__(1024) unsigned char smem_tf[];
  TfSmem& sm = *reinterpret_cast<TfSmem*>(smem_tf);
  if ((smem_u32(smem_tf) & 1023u) != 0) __trap();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int pi = 0;
  while (pi + 1 < tab.n && (int)blockIdx.x >= tab.p[pi + 1].first) ++pi;
  const TfProblem& P = tab.p[pi];
  const int split = blockIdx.x - P.first;
  const long long r_begin = (long long)split * P.rows_per_split;
  const long long r_end = r_begin + P.rows_per_split < P.R ? r_begin + P.rows_per_split : P.R;
  const int n_kb = (int)((r_end - r_begin) / kTfBoxCols);
  const int mh = (P.M + 127) / 128;           // accumulator halves
  const int nb = (P.Nmma + 127) / 128;        // B boxes per stage

  if (tid == 0) {
    for (int s = 0; s < kTfStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    mbar_init(&sm.done, 1);
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

  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* maps[4] = {&map0, &map1, &map2, &map3};
      const CUtensorMap* mA = maps[P.mapA];
      const CUtensorMap* mB = maps[P.mapB];
      const uint32_t bytes = (uint32_t)(mh + nb) * kTfBoxBytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(&sm.empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&sm.full[stage], bytes);
        const int r = (int)(r_begin + (long long)kb * kTfBoxCols);
        uint8_t* st = sm.stage[stage];
        for (int i = 0; i < mh; ++i) tma_load_2d(st + i * kTfBoxBytes, mA, r, P.chA + 128 * i, &sm.full[stage]);
        for (int j = 0; j < nb; ++j) tma_load_2d(st + (2 + j) * kTfBoxBytes, mB, r, P.chB + 128 * j, &sm.full[stage]);
        if (++stage == kTfStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_tf32(128, P.Nmma);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < n_kb; ++kb) {
      mbar_wait(&sm.full[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t base = smem_u32(sm.stage[stage]);
        const uint64_t bdesc = umma_desc_sw128(base + 2 * kTfBoxBytes);
        for (int i = 0; i < mh; ++i) {
          const uint64_t adesc = umma_desc_sw128(base + i * kTfBoxBytes);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)   // K = 8 tf32 = 32 bytes per instruction
            tc_mma_tf32(tmem_base + 256 * i, adesc + 2 * ks, bdesc + 2 * ks, idesc, (kb | ks) != 0 ? 1u : 0u);
        }
        tc_commit(&sm.empty[stage]);
        if (kb == n_kb - 1) tc_commit(&sm.done);
      }
      __syncwarp();
      if (++stage == kTfStages) { stage = 0; phase ^= 1; }
    }
  } else if (n_kb > 0) {
    // epilogue: thread = accumulator row; 32 columns per TMEM load
    const int erow = (warp & 3) * 32 + lane;   // TMEM lane group of this warp = warp id % 4
    mbar_wait(&sm.done, 0);
    tc_fence_after();
    for (int i = 0; i < mh; ++i) {
      const int m = i * 128 + erow;
      for (int c0 = 0; c0 < P.Nmma; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256 * i + c0, v);
        tmem_ld_wait();
        if (m < P.M) {
          float* crow = P.C + (long long)m * P.ldc + c0;
          if (P.vec4 && c0 + 32 <= P.N) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(crow + 4 * q), "f"(__uint_as_float(v[4 * q])),
                           "f"(__uint_as_float(v[4 * q + 1])), "f"(__uint_as_float(v[4 * q + 2])),
                           "f"(__uint_as_float(v[4 * q + 3])) : "memory");
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (c0 + q < P.N) atomicAdd(crow + q, __uint_as_float(v[q]));
          }
        }
      }
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
// 2b. the dX chain on the tensor cores, one launch per layer (opt-in with SNERF_MODE_TF32):
//   out[m][r] = epilogue( sum_k Wb[k][m] * S[ch0 + k][r] )       m < M <= 256 output channels, r = rows
//   Channels sit on the accumulator LANES and rows on its COLUMNS, so both operands are consumed exactly as they
//   lie in memory -- Wb[k][m] (backward image, m contiguous) and the store S[channel][R] (r contiguous) are
//   "MN-major" UMMA operands: TMA boxes of 32 k-rows x 32 contiguous elements, 128B swizzle -- and every
//   epilogue thread (= one channel) reads its mask and writes its result as contiguous 128-byte runs of the
//   [channel][R] stores.  128-row tiles, 256 TMEM columns, 2 CTAs per SM so one CTA's epilogue overlaps the
//   other's main loop.  Epilogue: + add_w[m] * add_row[r] (the alpha head), * relu'(saved activation).
// ------------------------------------------------------------------------------------
constexpr int kCgStages = 2;
constexpr int kCgTileRows = 128;
constexpr int kCgBox = 32 * 32 * 4;          // 32 k-rows x 32 contiguous fp32 = 4 KiB
constexpr int kCgABytes = 8 * kCgBox;        // 256 channels
constexpr int kCgBBytes = 4 * kCgBox;        // 128 rows
constexpr int kCgThreads = 192;

// A/B switch for the epilogue of chan_gemm_tf32_kernel: SNERF_CG_EPILOGUE = 0 per-lane row stores, 1 block-transposed
// (default), 2 block-transposed for the forward layers only
static int cg_coalesced() {
  static const int v = [] { const char* e = getenv("SNERF_CG_EPILOGUE"); return e ? atoi(e) : 1; }();
  return v;
}

struct alignas(1024) CgSmem {
  uint8_t a[kCgStages][kCgABytes];
  uint8_t b[kCgStages][kCgBBytes];
  uint64_t full[kCgStages];
  uint64_t empty[kCgStages];
  uint64_t done;
  uint32_t tmem_base;
};
struct CgProblem {
  int mapA, mapB;          // tensor maps: weight image [rows][ld]; store [channels][R] (32 x 32 boxes)
  int a_row0;              // row of the image map holding k = 0
  int n_seg;               // the contraction runs over up to three channel ranges of the store
  int seg_ch[3];           //   first channel (map-B coordinates) ...
  int seg_kb[3];           //   ... and length in 32-channel blocks of each range
  int M;                   // output channels
  int epi;                 // 0 = backward link (add_row / mask), 1 = forward layer (bias / relu)
  long long R;
  const float* mask;       // epi 0: [M][R] saved activation of the output channels, null = no ReLU'
  const float* add_row;    // epi 0: [R] or null
  const float* add_w;      // epi 0: [M]
  const float* bias;       // epi 1: [M]
  int relu;                // epi 1
  float* out;              // [M][R]
  float* bias_grad;        // epi 0: [M] += row sums of the stored result (d loss / d bias of the layer it belongs to), or null
  int first;
};
struct CgTable { int n; int coalesced; CgProblem p[2]; };   // coalesced: block-transposed epilogue (cg_coalesced())

// MN-major operand of 32-bit elements: 32 contiguous fp32 (128 B) per k-row, 32-element blocks `lbo` bytes apart.
// Transposing 4-byte elements needs the 128-byte swizzle with 32-byte atoms (layout type 1, SWIZZLE_128B_BASE32B;
// TMA side: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): the pattern repeats every 4 k-rows, so the stride between k-groups
// (SBO) is 512 bytes.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}

__global__ void __launch_bounds__(kCgThreads, 2)
chan_gemm_tf32_kernel(const __grid_constant__ CUtensorMap mapW0, const __grid_constant__ CUtensorMap mapW1,
                      const __grid_constant__ CUtensorMap mapS0, const __grid_constant__ CUtensorMap mapS1,
                      const CgTable tab) {
  extern __shared__ __align__(1024) unsigned char smem_cg[];
  CgSmem& sm = *reinterpret_cast<CgSmem*>(smem_cg);
  if ((smem_u32(smem_cg) & 1023u) != 0) __trap();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pi = (tab.n > 1 && (int)blockIdx.x >= tab.p[1].first) ? 1 : 0;
  const CgProblem& P = tab.p[pi];
  const long long r0 = (long long)(blockIdx.x - P.first) * kCgTileRows;
  const int n_kb = P.seg_kb[0] + (P.n_seg > 1 ? P.seg_kb[1] : 0) + (P.n_seg > 2 ? P.seg_kb[2] : 0);
  const int a_boxes = (P.M + 31) / 32;

  if (tid == 0) {
    for (int s = 0; s < kCgStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    mbar_init(&sm.done, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&sm.tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* mW = P.mapA == 0 ? &mapW0 : &mapW1;
      const CUtensorMap* mS = P.mapB == 0 ? &mapS0 : &mapS1;
      const uint32_t bytes = (uint32_t)(a_boxes + 4) * kCgBox;
      int stage = 0;
      uint32_t phase = 0;
      int seg = 0, in_seg = 0;
      for (int kb = 0; kb < n_kb; ++kb) {
        while (in_seg == P.seg_kb[seg]) { ++seg; in_seg = 0; }
        const int ch = P.seg_ch[seg] + 32 * in_seg++;
        mbar_wait(&sm.empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&sm.full[stage], bytes);
        for (int j = 0; j < a_boxes; ++j) tma_load_2d(sm.a[stage] + j * kCgBox, mW, 32 * j, P.a_row0 + 32 * kb, &sm.full[stage]);
        for (int j = 0; j < 4; ++j) tma_load_2d(sm.b[stage] + j * kCgBox, mS, (int)r0 + 32 * j, ch, &sm.full[stage]);
        if (++stage == kCgStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // both operands MN-major: bits 15 / 16 of the instruction descriptor
    const uint32_t idesc = umma_idesc_tf32(128, kCgTileRows) | (1u << 15) | (1u << 16);
    const int mh = (P.M + 127) / 128;
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < n_kb; ++kb) {
      mbar_wait(&sm.full[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t abase = smem_u32(sm.a[stage]), bbase = smem_u32(sm.b[stage]);
        for (int i = 0; i < mh; ++i) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)   // K = 8 per instruction = one 8-row group = 1024 bytes
            tc_mma_tf32(tmem_base + 128 * i, umma_desc_mn_sw128(abase + i * 4 * kCgBox + ks * 1024, kCgBox),
                        umma_desc_mn_sw128(bbase + ks * 1024, kCgBox), idesc, (kb | ks) != 0 ? 1u : 0u);
        }
        tc_commit(&sm.empty[stage]);
        if (kb == n_kb - 1) tc_commit(&sm.done);
      }
      __syncwarp();
      if (++stage == kCgStages) { stage = 0; phase ^= 1; }
    }
  } else if (tab.coalesced == 1 || (tab.coalesced == 2 && P.epi == 1)) {
    // Epilogue, block-transposed: a lane owns one channel of the accumulator (TMEM lane), but the [channel][R] stores want
    // whole 128-byte lines per instruction.  Each 32 x 32 block goes through shared memory (the operand ring is idle once
    // `done` fires: every MMA has finished reading it), rows padded to 36 floats so the float4 writes (lane = channel) and
    // float4 reads (8 lanes per channel) are both bank-conflict free; every global load / store instruction of the warp
    // then covers 4 channels x 128 contiguous bytes.
    const int lg = warp & 3;                  // TMEM lane group this warp may access
    const int mh = (P.M + 127) / 128;
    mbar_wait(&sm.done, 0);
    tc_fence_after();
    float* tile = reinterpret_cast<float*>(sm.a[0]) + (warp - 2) * (32 * 36);
    const int sub = lane >> 3, r4 = (lane & 7) * 4;
    for (int i = 0; i < mh; ++i) {
      const int m = i * 128 + lg * 32 + lane;
      const float bs = (P.epi == 1 && m < P.M) ? __ldg(P.bias + m) : 0.f;
      float rs[8], awk[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        rs[k] = 0.f;
        const int mm = i * 128 + lg * 32 + 4 * k + sub;
        awk[k] = (P.epi == 0 && P.add_row && mm < P.M) ? __ldg(P.add_w + mm) : 0.f;
      }
#pragma unroll 1
      for (int c0 = 0; c0 < kCgTileRows; c0 += 32) {
        const long long r = r0 + c0;
        if (r >= P.R) continue;               // R is a multiple of 64, tiles are 128 rows: whole 32-row groups are in or out (warp-uniform)
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + 128 * i + c0, v);
        // the backward's global operands, all in flight before anything waits: relu' masks of the 8 channel groups and
        // the alpha-head row
        float4 mk[8];
        float4 ar = make_float4(0.f, 0.f, 0.f, 0.f);
        if (P.epi == 0) {
          if (P.add_row) ar = __ldg(reinterpret_cast<const float4*>(P.add_row + r + r4));
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int mm = i * 128 + lg * 32 + 4 * k + sub;
            mk[k] = (P.mask && mm < P.M) ? __ldg(reinterpret_cast<const float4*>(P.mask + (long long)mm * P.R + r + r4))
                                         : make_float4(1.f, 1.f, 1.f, 1.f);
          }
        }
        tmem_ld_wait();
        float4* trow = reinterpret_cast<float4*>(tile + lane * 36);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 t = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                 __uint_as_float(v[4 * q + 3]));
          if (P.epi == 1) {
            t.x += bs; t.y += bs; t.z += bs; t.w += bs;
            if (P.relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
          }
          trow[q] = t;
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int ch = 4 * k + sub;
          const int mm = i * 128 + lg * 32 + ch;
          float4 t = *reinterpret_cast<const float4*>(tile + ch * 36 + r4);
          if (mm < P.M) {
            if (P.epi == 0) {
              const float aw = awk[k];
              t.x = fmaf(aw, ar.x, t.x); t.y = fmaf(aw, ar.y, t.y); t.z = fmaf(aw, ar.z, t.z); t.w = fmaf(aw, ar.w, t.w);
              const float4 m4 = mk[k];
              t.x = m4.x > 0.f ? t.x : 0.f; t.y = m4.y > 0.f ? t.y : 0.f; t.z = m4.z > 0.f ? t.z : 0.f; t.w = m4.w > 0.f ? t.w : 0.f;
            }
            // stored on the tf32 grid (round to nearest): the next link and the weight-gradient GEMM consume it exactly
            *reinterpret_cast<float4*>(P.out + (long long)mm * P.R + r + r4) =
                make_float4(round_tf32(t.x), round_tf32(t.y), round_tf32(t.z), round_tf32(t.w));
            rs[k] += (t.x + t.y) + (t.z + t.w);
          }
        }
        __syncwarp();
      }
      if (P.epi == 0 && P.bias_grad) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float t = rs[k];
          t += __shfl_xor_sync(0xffffffffu, t, 1);
          t += __shfl_xor_sync(0xffffffffu, t, 2);
          t += __shfl_xor_sync(0xffffffffu, t, 4);
          const int mm = i * 128 + lg * 32 + 4 * k + sub;
          if ((lane & 7) == 0 && mm < P.M) atomicAdd(P.bias_grad + mm, t);
        }
      }
    }
  } else {
    const int lg = warp & 3;                  // TMEM lane group this warp may access
    const int mh = (P.M + 127) / 128;
    mbar_wait(&sm.done, 0);
    tc_fence_after();
    for (int i = 0; i < mh; ++i) {
      const int m = i * 128 + lg * 32 + lane;
      const float aw = (P.epi == 0 && P.add_row && m < P.M) ? __ldg(P.add_w + m) : 0.f;
      const float bs = (P.epi == 1 && m < P.M) ? __ldg(P.bias + m) : 0.f;
      float row_sum = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < kCgTileRows; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + 128 * i + c0, v);
        tmem_ld_wait();
        const long long r = r0 + c0;
        if (m < P.M && r < P.R) {      // R is a multiple of 64, tiles are 128 rows: whole 32-row groups are in or out
          float f[32];
#pragma unroll
          for (int q = 0; q < 32; ++q) f[q] = __uint_as_float(v[q]);
          if (P.epi == 1) {
#pragma unroll
            for (int q = 0; q < 32; ++q) { f[q] += bs; if (P.relu) f[q] = fmaxf(f[q], 0.f); }
          }
          if (P.epi == 0 && P.add_row) {
            const float4* ar = reinterpret_cast<const float4*>(P.add_row + r);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 t = __ldg(ar + q);
              f[4 * q] = fmaf(aw, t.x, f[4 * q]); f[4 * q + 1] = fmaf(aw, t.y, f[4 * q + 1]);
              f[4 * q + 2] = fmaf(aw, t.z, f[4 * q + 2]); f[4 * q + 3] = fmaf(aw, t.w, f[4 * q + 3]);
            }
          }
          if (P.epi == 0 && P.mask) {
            const float4* mk = reinterpret_cast<const float4*>(P.mask + (long long)m * P.R + r);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 t = __ldg(mk + q);
              f[4 * q] = t.x > 0.f ? f[4 * q] : 0.f; f[4 * q + 1] = t.y > 0.f ? f[4 * q + 1] : 0.f;
              f[4 * q + 2] = t.z > 0.f ? f[4 * q + 2] : 0.f; f[4 * q + 3] = t.w > 0.f ? f[4 * q + 3] : 0.f;
            }
          }
          // stored on the tf32 grid (round to nearest): the next link and the weight-gradient GEMM consume it exactly
          float4* o = reinterpret_cast<float4*>(P.out + (long long)m * P.R + r);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            o[q] = make_float4(round_tf32(f[4 * q]), round_tf32(f[4 * q + 1]), round_tf32(f[4 * q + 2]), round_tf32(f[4 * q + 3]));
          if (P.epi == 0 && P.bias_grad) {
            float t = 0.f;
#pragma unroll
            for (int q = 0; q < 32; ++q) t += f[q];
            row_sum += t;
          }
        }
      }
      if (P.epi == 0 && P.bias_grad && m < P.M) atomicAdd(P.bias_grad + m, row_sum);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
  }
}

// rgb_linear^T (3 -> W/2 channels), masked by the views ReLU: the first link of the chain, too narrow for the MMA
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ draw, const float* __restrict__ wr,
                                                       const float* __restrict__ vsave, float* __restrict__ dzv,
                                                       int n_ch, long long R, float* bias_grad) {
  const long long r = (blockIdx.x * 256ll + threadIdx.x) * 4;
  const bool valid = r < R;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 d0 = valid ? *reinterpret_cast<const float4*>(draw + r) : zero;
  const float4 d1 = valid ? *reinterpret_cast<const float4*>(draw + R + r) : zero;
  const float4 d2 = valid ? *reinterpret_cast<const float4*>(draw + 2 * R + r) : zero;
  __shared__ float bsum[8];
  for (int k = blockIdx.y; k < n_ch; k += gridDim.y) {
    const float w0 = __ldg(wr + k), w1 = __ldg(wr + n_ch + k), w2 = __ldg(wr + 2 * n_ch + k);
    float4 o = zero;
    if (valid) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(vsave + (long long)k * R + r));
      o.x = m.x > 0.f ? round_tf32(fmaf(d2.x, w2, fmaf(d1.x, w1, d0.x * w0))) : 0.f;
      o.y = m.y > 0.f ? round_tf32(fmaf(d2.y, w2, fmaf(d1.y, w1, d0.y * w0))) : 0.f;
      o.z = m.z > 0.f ? round_tf32(fmaf(d2.z, w2, fmaf(d1.z, w1, d0.z * w0))) : 0.f;
      o.w = m.w > 0.f ? round_tf32(fmaf(d2.w, w2, fmaf(d1.w, w1, d0.w * w0))) : 0.f;
      *reinterpret_cast<float4*>(dzv + (long long)k * R + r) = o;
    }
    if (bias_grad) {   // d loss / d views bias: row sum of this channel -- ONE atomic per block and channel
      const float t = warp_sum((o.x + o.y) + (o.z + o.w));
      if ((threadIdx.x & 31) == 0) bsum[threadIdx.x >> 5] = t;
      __syncthreads();
      if (threadIdx.x == 0) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += bsum[w];
        if (tot != 0.f) atomicAdd(bias_grad + k, tot);
      }
      __syncthreads();
    }
  }
}

// ---- host: tensor maps of the stores
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}
// [channels][R] fp32 store -> 2D tensor map with 128-channel x 32-row boxes, 128B swizzle
static int make_store_map(CUtensorMap* map, const float* base, long long R, int channels, int box_rows = kTfBoxRows,
                          CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return SNERF_ERR_CUDA; }
  if (R <= 0 || channels <= 0 || !base) { memset(map, 0, sizeof(*map)); return 0; }
  const cuuint64_t dims[2] = {(cuuint64_t)R, (cuuint64_t)channels};
  const cuuint64_t strides[1] = {(cuuint64_t)R * 4};
  const cuuint32_t box[2] = {kTfBoxCols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return SNERF_ERR_CUDA; }
  return 0;
}

// ------------------------------------------------------------------------------------
// 4. skinny reductions:  out[c][k] += sum_r G[c][r] * X[k][r]   (G == null: G = 1, C = 1)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) skinny_kernel(const SkTable tab) {
  __shared__ float red[3][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int pi = 0;
  while (pi + 1 < tab.n && (int)blockIdx.x >= tab.p[pi + 1].first) ++pi;
  const SkProblem& P = tab.p[pi];
  const int k = blockIdx.x - P.first;
  const float4* x = reinterpret_cast<const float4*>(P.X + (long long)k * P.R);
  const long long n4 = P.R >> 2;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  if (!P.G) {
    for (long long i = tid; i < n4; i += 256) { const float4 v = __ldg(x + i); a0 += (v.x + v.y) + (v.z + v.w); }
  } else {
    const float4* g0 = reinterpret_cast<const float4*>(P.G);
    const float4* g1 = reinterpret_cast<const float4*>(P.G + P.R);
    const float4* g2 = reinterpret_cast<const float4*>(P.G + 2 * P.R);
    for (long long i = tid; i < n4; i += 256) {
      const float4 v = __ldg(x + i);
      const float4 u0 = __ldg(g0 + i);
      a0 += (v.x * u0.x + v.y * u0.y) + (v.z * u0.z + v.w * u0.w);
      if (P.C > 1) {
        const float4 u1 = __ldg(g1 + i), u2 = __ldg(g2 + i);
        a1 += (v.x * u1.x + v.y * u1.y) + (v.z * u1.z + v.w * u1.w);
        a2 += (v.x * u2.x + v.y * u2.y) + (v.z * u2.z + v.w * u2.w);
      }
    }
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
  if (lane == 0) { red[0][warp] = a0; red[1][warp] = a1; red[2][warp] = a2; }
  __syncthreads();
  if (tid < 3 && tid < P.C) {
    float s = 0.f;
    for (int wv = 0; wv < 8; ++wv) s += red[tid][wv];
    atomicAdd(P.out + (long long)tid * P.ldo + k, s);
  }
}

// ------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------
template <int W>
static int launch_mlp_bwd(const TrainParams& p, cudaStream_t stream) {
  const size_t smem = sizeof(BwdSmem<W>);
  auto kern = mlp_bwd_kernel<W>;
  if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                 "cudaFuncSetAttribute(mlp_bwd smem)"))
    return SNERF_ERR_CUDA;
  const long long tiles = p.n_rays * (p.TC + p.TF);
  const long long grid = tiles < (long long)sm_count() ? tiles : (long long)sm_count();
  kern<<<(unsigned)grid, kFp32Threads, smem, stream>>>(p);
  return check_cuda(cudaGetLastError(), "launch mlp_bwd_kernel");
}

// ------------------------------------------------------------------------------------
// tensor-core training FORWARD: the fused kernel runs in stages (inputs only), the MLP as layer-batched GEMMs
// ------------------------------------------------------------------------------------
// alpha_linear / rgb_linear on the stored activations -> raw[ray][sample][4]
__global__ void __launch_bounds__(256) heads_fwd_kernel(const float* __restrict__ save, long long R, int T, int S,
                                                        const float* __restrict__ wa, const float* __restrict__ ba, int cha,
                                                        int Ka, const float* __restrict__ wr, const float* __restrict__ br,
                                                        int chr, int Kr, float* __restrict__ raw) {
  const long long row = blockIdx.x * 256ll + threadIdx.x;
  if (row >= R) return;
  const long long ray = row / (T * kTileRows);
  const int s = (int)(row % (T * kTileRows));
  if (s >= S) return;
  float sig = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
  const float* h = save + (long long)cha * R + row;
#pragma unroll 16
  for (int k = 0; k < Ka; ++k) sig = fmaf(__ldg(wa + k), __ldg(h + (long long)k * R), sig);   // same summation order, loads batched
  const float* v = save + (long long)chr * R + row;
#pragma unroll 16
  for (int k = 0; k < Kr; ++k) {
    const float x = __ldg(v + (long long)k * R);
    c0 = fmaf(__ldg(wr + k), x, c0);
    c1 = fmaf(__ldg(wr + Kr + k), x, c1);
    c2 = fmaf(__ldg(wr + 2 * Kr + k), x, c2);
  }
  reinterpret_cast<float4*>(raw)[ray * S + s] = make_float4(c0 + __ldg(br), c1 + __ldg(br + 1), c2 + __ldg(br + 2), sig + __ldg(ba));
}

// raw2outputs of one pass, one warp per ray (run_nerf_helpers.py:381-424)
__global__ void __launch_bounds__(128) composite_rays_kernel(const float* __restrict__ ray_batch, int width, int row_stride,
                                                             long long n_rays, const float* __restrict__ raw,
                                                             const float* __restrict__ z, int S,
                                                             const float* __restrict__ noise, int white, float* rgb,
                                                             float* disp, float* acc, float* depth, float* weights) {
  const int lane = threadIdx.x & 31;
  const long long ray = blockIdx.x * 4ll + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const Ray rv = load_ray(ray_batch + ray * row_stride, width, 0);
  const RayCarry c = composite_segment(reinterpret_cast<const float4*>(raw) + ray * S, z + ray * S, S, 0, S, rv.dnorm,
                                       noise ? noise + ray * S : nullptr, nullptr, weights ? weights + ray * S : nullptr,
                                       carry_init(), lane);
  if (lane == 0) {
    const float wb = white ? (1.f - c.acc) : 0.f;
    if (rgb) { rgb[ray * 3 + 0] = c.r + wb; rgb[ray * 3 + 1] = c.g + wb; rgb[ray * 3 + 2] = c.b + wb; }
    if (disp) disp[ray] = disparity(c.depth, c.acc);
    if (acc) acc[ray] = c.acc;
    if (depth) depth[ray] = c.depth;
  }
}

// the MLP of one pass over its activation store: one chan_gemm launch per wide layer, then the two heads
static int run_fwd_chain(const SnerfNetDesc* d, const unsigned char* img, float* save, long long R, int T, int S,
                         float* raw, cudaStream_t stream) {
  Fp32Header h;
  const size_t img_bytes = plan_fp32(d, &h, true);
  const TrainChannels ch = train_channels(d);
  const float* imgf = reinterpret_cast<const float*>(img);
  const int W = d->W;
  CUtensorMap mWf, mWh, mS;
  if (int e = make_store_map(&mWf, imgf, W, (int)(img_bytes / 4 / W), 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return e;
  if (int e = make_store_map(&mWh, imgf, W / 2, (int)(img_bytes / 4 / (W / 2)), 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return e;
  if (int e = make_store_map(&mS, save, R, ch.total, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return e;
  const size_t smem = sizeof(CgSmem);
  if (check_cuda(cudaFuncSetAttribute(chan_gemm_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                 "cudaFuncSetAttribute(chan_gemm smem)"))
    return SNERF_ERR_CUDA;
  const int blocks = (int)((R + kCgTileRows - 1) / kCgTileRows);
  int prev_ch = -1;
  const Fp32Layer *alpha = nullptr, *rgb = nullptr;
  int alpha_src = -1, rgb_src = -1;
  for (int l = 0; l < h.n_layers; ++l) {
    const Fp32Layer& Ly = h.layers[l];
    if (Ly.kind != 0) {
      if (Ly.n_out == 1) { alpha = &Ly; alpha_src = prev_ch; } else { rgb = &Ly; rgb_src = prev_ch; }
      continue;
    }
    CgTable tab{};
    tab.coalesced = cg_coalesced();
    tab.n = 1;
    CgProblem& q = tab.p[0];
    q.mapA = Ly.n_out == W ? 0 : 1; q.mapB = 0;
    q.a_row0 = (int)(Ly.w_off / Ly.n_out);
    q.n_seg = 0;
    if (Ly.seg_rows[0]) { q.seg_ch[q.n_seg] = kSaveEncCh; q.seg_kb[q.n_seg++] = Ly.seg_rows[0] / 32; }
    if (Ly.seg_rows[1]) { q.seg_ch[q.n_seg] = prev_ch; q.seg_kb[q.n_seg++] = Ly.seg_rows[1] / 32; }
    if (Ly.seg_rows[2]) { q.seg_ch[q.n_seg] = kSaveDirCh; q.seg_kb[q.n_seg++] = Ly.seg_rows[2] / 32; }
    q.M = Ly.n_out; q.epi = 1; q.R = R;
    q.bias = imgf + Ly.b_off; q.relu = Ly.relu;
    q.out = save + (long long)Ly.ch_off * R;
    q.first = 0;
    chan_gemm_tf32_kernel<<<blocks, kCgThreads, smem, stream>>>(mWf, mWh, mS, mS, tab);
    prev_ch = Ly.ch_off;
  }
  if (!alpha || !rgb) { set_error("internal: heads missing from the layer plan"); return SNERF_ERR_BAD_ARG; }
  heads_fwd_kernel<<<(unsigned)((R + 255) / 256), 256, 0, stream>>>(
      save, R, T, S, imgf + alpha->w_off, imgf + alpha->b_off, alpha_src, alpha->seg_rows[1], imgf + rgb->w_off,
      imgf + rgb->b_off, rgb_src, rgb->seg_rows[1], raw);
  return check_cuda(cudaGetLastError(), "forward chain (tf32)");
}

int launch_train_forward_tf32(const SnerfNetDesc* d, RenderParams p, const TrainLayout& L, float* ws, cudaStream_t stream) {
  p.round_tf32 = 1;
  p.stage = 1;
  if (int e = launch_fp32(FE_RAYS, d->W, p, stream)) return e;
  if (int e = run_fwd_chain(d, p.img_coarse, p.save_c, L.Rc, L.TC, p.Nc, ws + L.raw_c, stream)) return e;
  const unsigned cgrid = (unsigned)((p.n_rays + 3) / 4);
  if (p.Nf > 0) {
    p.stage = 2;
    if (int e = launch_fp32(FE_RAYS, d->W, p, stream)) return e;
    if (int e = run_fwd_chain(d, p.img_fine, p.save_f, L.Rf, L.TF, p.Nc + p.Nf, ws + L.raw_f, stream)) return e;
    composite_rays_kernel<<<cgrid, 128, 0, stream>>>(p.ray_batch, p.width, p.row_stride, p.n_rays, ws + L.raw_f, ws + L.z_f,
                                                     p.Nc + p.Nf, p.noise1, p.white_bkgd, p.out.rgb_map, p.out.disp_map,
                                                     p.out.acc_map, p.out.depth_map, p.out.weights_fine);
  } else {
    composite_rays_kernel<<<cgrid, 128, 0, stream>>>(p.ray_batch, p.width, p.row_stride, p.n_rays, ws + L.raw_c, ws + L.z_c,
                                                     p.Nc, p.noise0, p.white_bkgd, p.out.rgb_map, p.out.disp_map,
                                                     p.out.acc_map, p.out.depth_map, p.out.weights);
  }
  return check_cuda(cudaGetLastError(), "training forward (tf32)");
}

// d_raw -> dZ of every layer with one tensor-core launch per layer (see chan_gemm_tf32_kernel)
static int launch_dx_chain_tf32(const SnerfNetDesc* d, const TrainParams& p, const SnerfNetGradF32* gc,
                                const SnerfNetGradF32* gf, cudaStream_t stream) {
  const TrainChannels ch = train_channels(d);
  const int passes = p.Nf > 0 ? 2 : 1, W = d->W;
  Fp32BwdHeader h;
  const size_t img_bytes = plan_bwd(d, &h);
  const long long Rs[2] = {p.Rc, p.Rf};
  const float* saves[2] = {p.save_c, p.save_f};
  float* dzs[2] = {p.dz_c, p.dz_f};                 // pre-shifted by -kSaveActCh channels
  const float* draws[2] = {p.draw_c, p.draw_f};
  const float* imgs[2] = {reinterpret_cast<const float*>(p.bwd_c), reinterpret_cast<const float*>(p.bwd_f)};
  CUtensorMap mW[2], mS[2];
  for (int pass = 0; pass < 2; ++pass) {
    const bool on = pass < passes;
    if (int e = make_store_map(&mW[pass], on ? imgs[pass] : nullptr, W, (int)(img_bytes / 4 / W), 32,
                               CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
      return e;
    if (int e = make_store_map(&mS[pass], on ? dzs[pass] + (long long)kSaveActCh * Rs[pass] : nullptr, on ? Rs[pass] : 0,
                               ch.total - kSaveActCh, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
      return e;
  }
  for (int pass = 0; pass < passes; ++pass) {   // link 0: rgb_linear^T
    const BwdStep& S = h.steps[0];
    const long long R = Rs[pass];
    dim3 grid((unsigned)((R / 4 + 255) / 256), 8);
    const SnerfNetGradF32* g = (pass && gf) ? gf : gc;
    head_bwd_kernel<<<grid, 256, 0, stream>>>(draws[pass], imgs[pass] + S.w_off, saves[pass] + (long long)S.mask_ch * R,
                                              dzs[pass] + (long long)S.dz_ch * R, S.n_out, R, g->views_b);
  }
  const size_t smem = sizeof(CgSmem);
  if (check_cuda(cudaFuncSetAttribute(chan_gemm_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                 "cudaFuncSetAttribute(chan_gemm smem)"))
    return SNERF_ERR_CUDA;
  for (int s = 1; s < h.n_steps; ++s) {
    const BwdStep& S = h.steps[s];
    CgTable tab{};
    tab.coalesced = cg_coalesced();
    int blocks = 0;
    for (int pass = 0; pass < passes; ++pass) {
      const long long R = Rs[pass];
      CgProblem& q = tab.p[tab.n++];
      q.mapA = pass; q.mapB = pass;
      q.a_row0 = (int)(S.w_off / W);
      q.n_seg = 1; q.seg_ch[0] = h.steps[s - 1].dz_ch - kSaveActCh; q.seg_kb[0] = S.K / 32;
      q.M = S.n_out; q.R = R; q.epi = 0;
      q.mask = S.mask_ch >= 0 ? saves[pass] + (long long)S.mask_ch * R : nullptr;
      q.add_row = S.add_col >= 0 ? draws[pass] + (long long)S.add_col * R : nullptr;
      q.add_w = imgs[pass] + S.add_w_off;
      q.out = dzs[pass] + (long long)S.dz_ch * R;
      const SnerfNetGradF32* g = (pass && gf) ? gf : gc;
      q.bias_grad = g->feature_b;                       // which layer's pre-activation gradient this link produces
      for (int l = 0; l < d->D; ++l) if (S.dz_ch == ch.trunk[l]) q.bias_grad = g->pts_b[l];
      q.first = blocks;
      blocks += (int)((R + kCgTileRows - 1) / kCgTileRows);
    }
    chan_gemm_tf32_kernel<<<blocks, kCgThreads, smem, stream>>>(mW[0], mW[1], mS[0], mS[1], tab);
  }
  return check_cuda(cudaGetLastError(), "launch dX chain (tf32)");
}

int launch_composite_bwd_rows(const TrainParams& p, cudaStream_t stream) {
  const int passes = p.Nf > 0 ? 2 : 1;
  composite_bwd_kernel<<<(unsigned)((p.n_rays * passes + 3) / 4), 128, 0, stream>>>(p);
  return check_cuda(cudaGetLastError(), "launch composite_bwd_kernel");
}

int launch_train_backward(const SnerfNetDesc* d, const SnerfNetDesc* d_fine, const TrainParams& p,
                          const SnerfNetGradF32* gc, const SnerfNetGradF32* gf, cudaStream_t stream) {
  if (p.n_rays == 0) return SNERF_OK;
  const SnerfNetDesc* descs[2] = {d, d_fine ? d_fine : d};
  const int Wmax = descs[0]->W > descs[1]->W ? descs[0]->W : descs[1]->W;
  const int passes = p.Nf > 0 ? 2 : 1;
  composite_bwd_kernel<<<(unsigned)((p.n_rays * passes + 3) / 4), 128, 0, stream>>>(p);
  if (check_cuda(cudaGetLastError(), "launch composite_bwd_kernel")) return SNERF_ERR_CUDA;
  int e;
  if (p.dw_tf32) e = launch_dx_chain_tf32(d, p, gc, gf, stream);
  else switch (Wmax) {
    case 64: e = launch_mlp_bwd<64>(p, stream); break;
    case 128: e = launch_mlp_bwd<128>(p, stream); break;
    case 256: e = launch_mlp_bwd<256>(p, stream); break;
    default: set_error("training supports W in {64,128,256}"); return SNERF_ERR_UNSUPPORTED;
  }
  if (e) return e;

  // ---- weight / bias gradient problems of both passes
  DwTable dw{};
  SkTable sk{};
  TfTable tf{};
  int tf_blocks = 0;
  int dw_blocks = 0, sk_blocks = 0;
  TrainChannels ch{};
  for (int pass = 0; pass < passes; ++pass) {
    const SnerfNetDesc* d = descs[pass];     // (shadows the coarse descriptor: everything below is per pass)
    ch = train_channels(d);
    const int W = d->W, ic = d->input_ch, icv = d->input_ch_views;
    const long long R = pass ? p.Rf : p.Rc;
    const float* save = pass ? p.save_f : p.save_c;
    const float* dz = pass ? p.dz_f : p.dz_c;
    const float* draw = pass ? p.draw_f : p.draw_c;
    const SnerfNetGradF32* g = (pass && gf) ? gf : gc;
    // rows per split: aim at ~4 waves of CTAs overall, multiple of 16, at least 1024 rows
    long long rps = 4096;
    while (rps > 1024 && R / rps < 4) rps >>= 1;
    auto gemm = [&](const float* A, int M, const float* B, int N, float* C, int ldc) {
      if (!C) return;
      if (p.dw_tf32) {   // same problem for the tensor-core kernel: operands named by tensor map + channel
        TfProblem& t = tf.p[tf.n++];
        t.mapA = 2 * pass + 1; t.mapB = 2 * pass;
        t.chA = (int)((A - dz) / R) - kSaveActCh; t.chB = (int)((B - save) / R);
        t.M = M; t.N = N; t.Nmma = (N + 15) / 16 * 16; t.ldc = ldc;
        t.vec4 = (ldc % 4 == 0) && (reinterpret_cast<uintptr_t>(C) % 16 == 0);
        t.R = R; t.C = C; t.rows_per_split = (int)rps; t.splits = (int)((R + rps - 1) / rps);
        t.first = tf_blocks; tf_blocks += t.splits;
        return;
      }
      DwProblem& q = dw.p[dw.n++];
      q.A = A; q.B = B; q.C = C; q.M = M; q.N = N; q.ldc = ldc; q.R = R;
      q.mt = (M + kDwTile - 1) / kDwTile; q.nt = (N + kDwTile - 1) / kDwTile;
      q.rows_per_split = (int)rps;
      q.splits = (int)((R + rps - 1) / rps);
      q.first = dw_blocks;
      dw_blocks += q.mt * q.nt * q.splits;
    };
    auto skinny = [&](const float* X, int K, const float* G, int C, float* out, int ldo) {
      if (!out) return;
      if (p.dw_tf32 && !G && K > 4) return;   // bias gradients of the wide layers come out of the dX chain's epilogues
      SkProblem& q = sk.p[sk.n++];
      q.X = X; q.G = G; q.out = out; q.K = K; q.C = C; q.ldo = ldo; q.R = R; q.first = sk_blocks;
      sk_blocks += K;
    };
    for (int l = 0; l < d->D; ++l) {
      const float* A = dz + (long long)ch.trunk[l] * R;
      const bool has_enc = l == 0 || (d->skip >= 0 && l - 1 == d->skip);
      const int ld = (has_enc ? ic : 0) + (l == 0 ? 0 : W);
      if (has_enc) gemm(A, W, save + (long long)kSaveEncCh * R, ic, g->pts_w[l], ld);
      if (l > 0) gemm(A, W, save + (long long)ch.trunk[l - 1] * R, W, g->pts_w[l] ? g->pts_w[l] + (has_enc ? ic : 0) : nullptr, ld);
      skinny(A, W, nullptr, 1, g->pts_b[l], 0);
    }
    const float* hlast = save + (long long)ch.trunk[d->D - 1] * R;
    if (!d->use_viewdirs) {   // output_linear rows 0..3 (rows >= 4 never reach raw2outputs: zero gradient)
      skinny(hlast, W, draw, 3, g->output_w, W);
      skinny(hlast, W, draw + 3 * R, 1, g->output_w ? g->output_w + 3 * W : nullptr, W);
      skinny(draw, 3, nullptr, 1, g->output_b, 0);
      skinny(draw + 3 * R, 1, nullptr, 1, g->output_b ? g->output_b + 3 : nullptr, 0);
      continue;
    }
    {  // feature_linear
      const float* A = dz + (long long)ch.feature * R;
      gemm(A, W, hlast, W, g->feature_w, W);
      skinny(A, W, nullptr, 1, g->feature_b, 0);
    }
    {  // views_linears.0 on [feature, dirs]
      const float* A = dz + (long long)ch.views * R;
      gemm(A, W / 2, save + (long long)ch.feature * R, W, g->views_w, W + icv);
      gemm(A, W / 2, save + (long long)kSaveDirCh * R, icv, g->views_w ? g->views_w + W : nullptr, W + icv);
      skinny(A, W / 2, nullptr, 1, g->views_b, 0);
    }
    skinny(hlast, W, draw + 3 * R, 1, g->alpha_w, W);                        // alpha_linear
    skinny(draw + 3 * R, 1, nullptr, 1, g->alpha_b, 0);
    skinny(save + (long long)ch.views * R, W / 2, draw, 3, g->rgb_w, W / 2);  // rgb_linear
    skinny(draw, 3, nullptr, 1, g->rgb_b, 0);
  }
  if (dw.n > kMaxDwProblems || sk.n > kMaxSkProblems) { set_error("internal: gradient problem table overflow"); return SNERF_ERR_BAD_ARG; }
  if (tf.n > kMaxTfProblems) { set_error("internal: gradient problem table overflow"); return SNERF_ERR_BAD_ARG; }
  if (tf_blocks > 0) {
    CUtensorMap maps[4];
    const long long Rs[2] = {p.Rc, p.Rf};
    const float* saves[2] = {p.save_c, p.save_f};
    const float* dzs[2] = {p.dz_c, p.dz_f};
    for (int pass = 0; pass < 2; ++pass) {
      const bool on = pass < passes;
      if (int e2 = make_store_map(&maps[2 * pass], saves[pass], on ? Rs[pass] : 0, ch.total)) return e2;
      if (int e2 = make_store_map(&maps[2 * pass + 1], on ? dzs[pass] + (long long)kSaveActCh * Rs[pass] : nullptr,
                                  on ? Rs[pass] : 0, ch.total - kSaveActCh)) return e2;
    }
    const size_t smem = sizeof(TfSmem);
    if (check_cuda(cudaFuncSetAttribute(dw_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                   "cudaFuncSetAttribute(dw_tf32 smem)"))
      return SNERF_ERR_CUDA;
    dw_tf32_kernel<<<tf_blocks, kTfThreads, smem, stream>>>(maps[0], maps[1], maps[2], maps[3], tf);
  }
  if (dw_blocks > 0) dw_gemm_kernel<<<dw_blocks, 256, 0, stream>>>(dw);
  if (sk_blocks > 0) skinny_kernel<<<sk_blocks, 256, 0, stream>>>(sk);
  return check_cuda(cudaGetLastError(), "launch gradient kernels");
}

}  // namespace snerf
