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
