// snerf_loss.cu -- the training objective of s-nerf/train.py consumed straight from the renderer's outputs
// (SURVEY section 8 row f-3): RgbLoss (model/loss_factory.py:5-11) + DepthLoss (:26-37) reduced by calc_depth_loss
// (model/confidence.py:211-226: mask target_depth != 0, optional per-ray confidence, mean over the masked rays),
// combined as train.py:149,209:   loss = mean((rgb - tgt)^2) [+ w0 mean((rgb0 - tgt)^2)] + depth_lambda * depth_loss.
//
// The reference evaluates this with ~40 torch launches forward and as many in autograd; here it is one reduction kernel
// forward (per-ray terms -> block sums -> double atomics -> the last block finalises) and one elementwise kernel backward.
// Memory-bound on 40 bytes per ray; nothing to put on tensor cores.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "snerf_internal.h"

namespace snerf {
namespace {

struct LossArgs {
  const float* rgb;        // [N,3]
  const float* rgb0;       // [N,3] or null
  const float* target;     // [N,3]
  const float* depth;      // [N]   or null (no depth term)
  const float* depth0;     // [N]
  const float* tdepth;     // [N]   0 = no LiDAR return
  const float* conf;       // [N]   or null
  long long N;
  float depth_lambda, c_weight, rgb0_weight;
  int disparity;
};

__device__ __forceinline__ float depth_fn(float x, int disparity) { return disparity ? __fdiv_rn(1.0f, x) : x; }

// scratch: double[4] = {sum (rgb - t)^2, sum (rgb0 - t)^2, sum depth terms, masked count}, then an unsigned ticket
__global__ void __launch_bounds__(256) loss_fwd_kernel(LossArgs a, double* __restrict__ scratch, float* __restrict__ out) {
  float s_rgb = 0.f, s_rgb0 = 0.f, s_dep = 0.f, s_cnt = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.N; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float t = a.target[i * 3 + c];
      const float e = a.rgb[i * 3 + c] - t;
      s_rgb = fmaf(e, e, s_rgb);
      if (a.rgb0) { const float e0 = a.rgb0[i * 3 + c] - t; s_rgb0 = fmaf(e0, e0, s_rgb0); }
    }
    if (a.depth) {
      const float td = a.tdepth[i];
      if (td != 0.f) {
        const float ft = depth_fn(td, a.disparity);
        float term = fabsf(depth_fn(a.depth[i], a.disparity) - ft) + a.c_weight * fabsf(depth_fn(a.depth0[i], a.disparity) - ft);
        if (a.conf) term *= a.conf[i];
        s_dep += term;
        s_cnt += 1.f;
      }
    }
  }
  __shared__ double red[4][8];
  double v[4] = {(double)s_rgb, (double)s_rgb0, (double)s_dep, (double)s_cnt};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
  }
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    for (int k = 0; k < 4; ++k) {
      double t = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[k][w];
      if (t != 0) atomicAdd(scratch + k, t);
    }
    __threadfence();
    const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(scratch + 4), 1u);
    last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    const volatile double* s = scratch;
    const double n3 = 3.0 * (double)a.N;
    const float img = (float)(s[0] / n3);
    const float img0 = a.rgb0 ? (float)(s[1] / n3) : 0.f;
    const float dep = a.depth ? (float)(s[2] / s[3]) : 0.f;          // mean over the masked rays (0 / 0 = nan, as torch)
    out[0] = img + a.rgb0_weight * img0 + (a.depth ? a.depth_lambda * dep : 0.f);
    out[1] = img;
    out[2] = dep;
    out[3] = (float)s[3];
    out[4] = img0;
  }
}

// d loss / d inputs, scaled by the upstream gradient g[0] (a device scalar: no host round trip)
__global__ void __launch_bounds__(256) loss_bwd_kernel(LossArgs a, const float* __restrict__ stats, const float* __restrict__ g,
                                                       float* __restrict__ g_rgb, float* __restrict__ g_rgb0,
                                                       float* __restrict__ g_depth, float* __restrict__ g_depth0,
                                                       float* __restrict__ g_conf) {
  const float up = g[0];
  const float k_rgb = up * 2.0f / (3.0f * (float)a.N);
  const float k_dep = a.depth ? up * a.depth_lambda / stats[3] : 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.N; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float t = a.target[i * 3 + c];
      if (g_rgb) g_rgb[i * 3 + c] = k_rgb * (a.rgb[i * 3 + c] - t);
      if (g_rgb0) g_rgb0[i * 3 + c] = a.rgb0 ? a.rgb0_weight * k_rgb * (a.rgb0[i * 3 + c] - t) : 0.f;
    }
    if (a.depth) {
      const float td = a.tdepth[i];
      float gd = 0.f, gd0 = 0.f, gc = 0.f;
      if (td != 0.f) {
        const float ft = depth_fn(td, a.disparity);
        const float d = a.depth[i], d0 = a.depth0[i];
        const float e = depth_fn(d, a.disparity) - ft, e0 = depth_fn(d0, a.disparity) - ft;
        const float cf = a.conf ? a.conf[i] : 1.f;
        const float sg = (e > 0.f) - (e < 0.f), sg0 = (e0 > 0.f) - (e0 < 0.f);       // d|x|/dx, 0 at 0 (torch.abs)
        const float fd = a.disparity ? -1.0f / (d * d) : 1.0f, fd0 = a.disparity ? -1.0f / (d0 * d0) : 1.0f;
        gd = k_dep * cf * sg * fd;
        gd0 = k_dep * cf * a.c_weight * sg0 * fd0;
        gc = k_dep * (fabsf(e) + a.c_weight * fabsf(e0));
      }
      if (g_depth) g_depth[i] = gd;
      if (g_depth0) g_depth0[i] = gd0;
      if (g_conf) g_conf[i] = gc;
    }
  }
}

}  // namespace

int loss_fwd(const SnerfLossOpts* o, const float* rgb, const float* rgb0, const float* target, const float* depth,
             const float* depth0, const float* tdepth, const float* conf, long long N, double* scratch, float* out,
             cudaStream_t st) {
  const LossArgs a{rgb, rgb0, target, depth, depth0, tdepth, conf, N, o->depth_lambda, o->coarse_depth_mult, o->rgb0_weight, o->disparity};
  const unsigned blocks = (unsigned)((N + 255) / 256 < 592 ? (N + 255) / 256 : 592);
  loss_fwd_kernel<<<blocks ? blocks : 1, 256, 0, st>>>(a, scratch, out);
  return check_cuda(cudaGetLastError(), "launch loss_fwd_kernel");
}
int loss_bwd(const SnerfLossOpts* o, const float* rgb, const float* rgb0, const float* target, const float* depth,
             const float* depth0, const float* tdepth, const float* conf, long long N, const float* stats, const float* g,
             float* g_rgb, float* g_rgb0, float* g_depth, float* g_depth0, float* g_conf, cudaStream_t st) {
  const LossArgs a{rgb, rgb0, target, depth, depth0, tdepth, conf, N, o->depth_lambda, o->coarse_depth_mult, o->rgb0_weight, o->disparity};
  const unsigned blocks = (unsigned)((N + 255) / 256 < 592 ? (N + 255) / 256 : 592);
  loss_bwd_kernel<<<blocks ? blocks : 1, 256, 0, st>>>(a, stats, g, g_rgb, g_rgb0, g_depth, g_depth0, g_conf);
  return check_cuda(cudaGetLastError(), "launch loss_bwd_kernel");
}

}  // namespace snerf

// ------------------------------------------------------------------------------------------------------------
// ProposalLoss (model/loss_factory.py:54-73): the coarse histogram must upper-bound the fine one.
//     inds  = searchsorted(s_c, s_f, right=True);  W_c = cumsum(w_c)
//     bound = W_c[inds[1:] - 1] - W_c[inds[:-1] - 1]              (indices clamped into the table)
//     loss  = weight * mean_rays( sum_i max(w_f[i] - bound[i], 0)^2 / (w_f[i] + 1e-8) )
// s_f / w_f are detached in the reference, so the only gradient is d loss / d w_c: each active fine interval adds
// g_i = -2 max(w_f - bound, 0) / (w_f + 1e-8) to the coarse intervals (l_i, r_i] -- a difference array + prefix sum.
// One warp per ray, forward value and gradient in the same launch (the reference: searchsorted + cumsum + two gathers
// + ~10 elementwise launches forward, their autograd backward with scatter-adds).
// ------------------------------------------------------------------------------------------------------------
namespace snerf {
namespace {

constexpr int kPlMax = 256;          // intervals per ray (coarse or fine)
constexpr int kPlWarps = 4;

struct alignas(16) PlSmem {
  float sc[kPlMax + 1];
  float W[kPlMax];
  float diff[kPlMax + 2];
  int ind[kPlMax + 1];
};

__global__ void __launch_bounds__(kPlWarps * 32) proposal_loss_kernel(const float* __restrict__ s_f, const float* __restrict__ w_f,
                                                                      const float* __restrict__ s_c, const float* __restrict__ w_c,
                                                                      long long N, int Sf, int Sc, float weight,
                                                                      double* __restrict__ scratch, float* __restrict__ loss_out,
                                                                      float* __restrict__ grad_wc) {
  __shared__ PlSmem sm_all[kPlWarps];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  PlSmem& sm = sm_all[wid];
  const long long ray = (long long)blockIdx.x * kPlWarps + wid;
  float ray_loss = 0.f;
  if (ray < N) {
    for (int i = lane; i <= Sc; i += 32) sm.sc[i] = s_c[ray * (Sc + 1) + i];
    for (int i = lane; i < Sc; i += 32) sm.W[i] = w_c[ray * Sc + i];
    for (int i = lane; i < Sc + 2; i += 32) sm.diff[i] = 0.f;
    __syncwarp();
    // sequential cumulative sum, lane after lane, accumulated in double and rounded per element (torch.cumsum on the CPU:
    // acc_type<float> is double there).  The loss divides by w_f + 1e-8, so the rounding of W matters for tiny w_f.
    const int chunk = (Sc + 31) / 32, k0 = lane * chunk, k1 = min(Sc, k0 + chunk);
    double acc = 0.0;
    for (int l = 0; l < 32; ++l) {
      if (lane == l)
        for (int k = k0; k < k1; ++k) { acc += (double)sm.W[k]; sm.W[k] = (float)acc; }
      acc = __shfl_sync(0xffffffffu, acc, l);
    }
    // searchsorted(s_c, s_f, right=True)
    for (int i = lane; i <= Sf; i += 32) {
      const float x = s_f[ray * (Sf + 1) + i];
      int lo = 0, hi = Sc + 1;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (sm.sc[mid] <= x) lo = mid + 1; else hi = mid; }
      sm.ind[i] = lo;
    }
    __syncwarp();
    const float scale = weight / (float)N;
    for (int i = lane; i < Sf; i += 32) {
      // loss_factory.py:66-67: left index clamp(min=0), right index clamp(max=weights_f.shape[1]-1) -- the reference clamps
      // the right index with the FINE count; kept as is, then both are forced into the table (where torch would raise)
      const int l = min(max(sm.ind[i] - 1, 0), Sc - 1), r = min(max(min(sm.ind[i + 1] - 1, Sf - 1), 0), Sc - 1);
      const float wf = w_f[ray * Sf + i];
      const float bound = __fsub_rn(sm.W[r], sm.W[l]);
      const float e = fmaxf(__fsub_rn(wf, bound), 0.f);
      const float den = __fadd_rn(wf, 1e-8f);
      ray_loss += __fdiv_rn(__fmul_rn(e, e), den);
      if (grad_wc && e > 0.f && r != l) {      // (r < l only through the reference's fine-count clamp: the range is then negative)
        const float g = -2.0f * e / den * scale;
        // bound = W[r] - W[l] = +sum w_c(l, r]  (or -sum w_c(r, l] when r < l)
        atomicAdd(&sm.diff[min(l, r) + 1], r > l ? g : -g);
        atomicAdd(&sm.diff[max(l, r) + 1], r > l ? -g : g);
      }
    }
    __syncwarp();
    if (grad_wc) {      // prefix sum of the difference array -> d loss / d w_c
      float run = 0.f;
      for (int l = 0; l < 32; ++l) {
        if (lane == l)
          for (int k = k0; k < k1; ++k) { run += sm.diff[k]; grad_wc[ray * Sc + k] = run; }
        run = __shfl_sync(0xffffffffu, run, l);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ray_loss += __shfl_xor_sync(0xffffffffu, ray_loss, o);
  __shared__ float wl[kPlWarps];
  __shared__ bool last;
  if (lane == 0) wl[wid] = ray_loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < kPlWarps; ++w) t += (double)wl[w];
    if (t != 0) atomicAdd(scratch, t);
    __threadfence();
    last = atomicAdd(reinterpret_cast<unsigned*>(scratch + 1), 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    const volatile double* s = scratch;
    loss_out[0] = (float)(s[0] / (double)N * (double)weight);
  }
}

}  // namespace

int proposal_loss(const float* s_f, const float* w_f, const float* s_c, const float* w_c, long long N, int Sf, int Sc,
                  float weight, double* scratch, float* loss_out, float* grad_wc, cudaStream_t st) {
  if (Sf < 1 || Sc < 1 || Sf > kPlMax || Sc > kPlMax) { set_error("proposal loss: 1 <= intervals <= %d (got %d fine, %d coarse)", kPlMax, Sf, Sc); return SNERF_ERR_UNSUPPORTED; }
  proposal_loss_kernel<<<(unsigned)((N + kPlWarps - 1) / kPlWarps), kPlWarps * 32, 0, st>>>(s_f, w_f, s_c, w_c, N, Sf, Sc, weight, scratch,
                                                                                        loss_out, grad_wc);
  return check_cuda(cudaGetLastError(), "launch proposal_loss_kernel");
}


// ------------------------------------------------------------------------------------
// Adam update of ONE flat parameter buffer (train.py:100 / render.py:222: torch.optim.Adam(betas=(0.9, 0.999)), same
// arithmetic as torch's single-tensor Adam without amsgrad / maximize: exp_avg, exp_avg_sq, bias corrections 1 - beta^t,
// p -= lr / bc1 * exp_avg / (sqrt(exp_avg_sq) / sqrt(bc2) + eps); weight_decay adds wd * p to the gradient).
// The step count and the learning rate live in device memory so the launch can sit in a CUDA graph; a one-thread kernel
// advances the count after the update.  1,191,688 parameters = 19 MB of traffic: one launch instead of torch's
// multi-tensor loop over 48 small tensors.
// ------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, const float* __restrict__ lr_dev,
                                                   float b1, float b2, float eps, float wd, const long long* __restrict__ step_dev) {
  const float t = (float)(*step_dev + 1);
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  const float step_size = *lr_dev / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i];
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = fmaf(1.f - b1, gi - m[i], m[i]);           // lerp, as torch: exp_avg.lerp_(grad, 1 - beta1)
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi; v[i] = vi;
    p[i] = pi - step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
  }
}
__global__ void adam_tick_kernel(long long* step_dev) { *step_dev += 1; }
}  // namespace

int adam_step(float* p, const float* g, float* m, float* v, long long n, const float* lr_dev, float b1, float b2, float eps,
              float wd, long long* step_dev, cudaStream_t st) {
  if (n <= 0) return SNERF_OK;
  const int blocks = (int)((n + 255) / 256 < 4 * 148 ? (n + 255) / 256 : 4 * 148);
  adam_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, n, lr_dev, b1, b2, eps, wd, step_dev);
  adam_tick_kernel<<<1, 1, 0, st>>>(step_dev);
  return check_cuda(cudaGetLastError(), "launch adam_kernel");
}

}  // namespace snerf
