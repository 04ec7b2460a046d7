// snerf_common.cuh -- device helpers shared by the fp32 (FFMA) and bf16 (tcgen05) renderers.
//
// Everything that is not the MLP lives here and is shared by both arithmetic modes:
// stratified depths (render.py:330-352), positional encoding (run_nerf_helpers.py:22-52),
// sigma->alpha compositing (run_nerf_helpers.py:381-424), inverse-CDF resampling
// (run_nerf_helpers.py:336-379) and the sorted merge (render.py:383).  All of it is fp32
// with explicit round-to-nearest intrinsics wherever the reference (eager torch) performs
// separate mul/add, so that sample positions are bit-identical to the reference's and the
// high octaves of the encoding (sin(512 x)) see the same argument.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace snerf {

constexpr int kMaxSamples = 256;  // n_samples + n_importance upper bound
constexpr float kHuge = 1e10f;    // last inter-sample distance (run_nerf_helpers.py:397)

// ------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, bulk async copy (TMA engine, non-tensor form), proxies
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (surfaces as a CUDA error) instead of hanging the GPU.  The slow path is
// kept out of line so the hot loops that call this stay small.
static __device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0 && clock64() - t0 > 8000000000ll) __trap();  // ~4 s: protocol bug, not a long wait
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}
// Same, for waits that are expected to be long (a whole tile): back off so the spinning warp does not take issue
// slots from the warps sharing its scheduler.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(200);
    if ((++spins & 1023u) == 0 && clock64() - t0 > 8000000000ll) __trap();
  }
}
// global -> shared bulk copy, completion counted in bytes on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// the same copy delivered to the same shared-memory offset (and signalled on the same barrier offset) of every CTA in
// `cta_mask` of the cluster
__device__ __forceinline__ void bulk_g2s_multicast(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                                   uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
// generic-proxy writes (st.shared) -> visible to the async proxy (tcgen05.mma / bulk copies)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------
// warp helpers
// ------------------------------------------------------------------------------------
__device__ __forceinline__ double shfl_up_d(double v, int delta) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_up_sync(0xffffffffu, lo, delta);
  hi = __shfl_up_sync(0xffffffffu, hi, delta);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_d(double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(0xffffffffu, lo, src);
  hi = __shfl_sync(0xffffffffu, hi, src);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, o);
    hi = __shfl_xor_sync(0xffffffffu, hi, o);
    v += __hiloint2double(hi, lo);
  }
  return v;
}
// inclusive scans across the 32 lanes (fp64: the reference's CPU scans accumulate in fp64)
__device__ __forceinline__ double warp_scan_prod_d(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double n = shfl_up_d(v, o);
    if (lane >= o) v *= n;
  }
  return v;
}
__device__ __forceinline__ double warp_scan_sum_d(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double n = shfl_up_d(v, o);
    if (lane >= o) v += n;
  }
  return v;
}

// ------------------------------------------------------------------------------------
// ray record
// ------------------------------------------------------------------------------------
struct Ray {
  float ox, oy, oz, dx, dy, dz, near, far, vx, vy, vz, dnorm;
};
__device__ __forceinline__ Ray load_ray(const float* __restrict__ rb, int width, int has_vd) {
  Ray r;
  r.ox = rb[0]; r.oy = rb[1]; r.oz = rb[2];
  r.dx = rb[3]; r.dy = rb[4]; r.dz = rb[5];
  r.near = rb[6]; r.far = rb[7];
  if (has_vd) { r.vx = rb[width - 3]; r.vy = rb[width - 2]; r.vz = rb[width - 1]; }
  else { r.vx = r.vy = r.vz = 0.f; }
  // torch.norm(rays_d[..., None, :], dim=-1): sqrt of the fp32 sum of squares
  r.dnorm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(r.dx, r.dx), __fmul_rn(r.dy, r.dy)), __fmul_rn(r.dz, r.dz)));
  return r;
}

// ray `idx` of the call: from the [N, width] batch, or -- camera mode -- built here: get_rays (run_nerf_helpers.py:247-258:
// dirs = ((i+.5-cx)/f, -(j+.5-cy)/f, -1) rotated by c2w[:3,:3], origin = c2w[:3,3]) and viewdirs = rays_d / ||rays_d||
// (render.py:56-63), with the same separate mul / add / div roundings as the eager reference
template <class P>
__device__ __forceinline__ Ray make_ray(const P& p, long long idx) {
  if (!p.cam_on) return load_ray(p.ray_batch + idx * p.row_stride, p.width, p.has_vd);
  const long long pix = p.cam_first + idx;
  const int i = (int)(pix % p.cam_W), j = (int)(pix / p.cam_W);
  const float d0 = __fdiv_rn(__fsub_rn(__fadd_rn((float)i, 0.5f), p.cam_cx), p.cam_focal);
  const float d1 = -__fdiv_rn(__fsub_rn(__fadd_rn((float)j, 0.5f), p.cam_cy), p.cam_focal);
  float d[3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
    d[k] = __fadd_rn(__fadd_rn(__fmul_rn(d0, p.cam_m[k * 4 + 0]), __fmul_rn(d1, p.cam_m[k * 4 + 1])), __fmul_rn(-1.f, p.cam_m[k * 4 + 2]));
  Ray r;
  r.ox = p.cam_m[3]; r.oy = p.cam_m[7]; r.oz = p.cam_m[11];
  r.dx = d[0]; r.dy = d[1]; r.dz = d[2];
  r.near = p.cam_near; r.far = p.cam_far;
  r.dnorm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(r.dx, r.dx), __fmul_rn(r.dy, r.dy)), __fmul_rn(r.dz, r.dz)));
  r.vx = __fdiv_rn(r.dx, r.dnorm); r.vy = __fdiv_rn(r.dy, r.dnorm); r.vz = __fdiv_rn(r.dz, r.dnorm);
  return r;
}

// near*(1-t) + far*t   |   1/(1/near*(1-t) + 1/far*t)       (render.py:331-334)
__device__ __forceinline__ float coarse_depth(float near, float far, float t, int lindisp) {
  const float omt = __fsub_rn(1.f, t);
  if (!lindisp) return __fadd_rn(__fmul_rn(near, omt), __fmul_rn(far, t));
  const float a = __fmul_rn(__fdiv_rn(1.f, near), omt);
  const float b = __fmul_rn(__fdiv_rn(1.f, far), t);
  return __fdiv_rn(1.f, __fadd_rn(a, b));
}
// stratified jitter of sample i given its un-jittered neighbours (render.py:338-352)
__device__ __forceinline__ float jitter_depth(float zm1, float z0, float zp1, bool first, bool last, float tr) {
  const float lower = first ? z0 : __fmul_rn(0.5f, __fadd_rn(z0, zm1));
  const float upper = last ? z0 : __fmul_rn(0.5f, __fadd_rn(zp1, z0));
  return __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), tr));
}
// o + d*z with separate rounding of the product (render.py:354)
__device__ __forceinline__ float ray_point(float o, float d, float z) { return __fadd_rn(o, __fmul_rn(d, z)); }

// ------------------------------------------------------------------------------------
// compositing: one warp, one contiguous segment [s0, s0+cnt) of one ray of S samples.
// Carry (transmittance so far, partial sums) makes it usable tile-by-tile.
// ------------------------------------------------------------------------------------
struct RayCarry {
  double T;  // product of (1 - alpha + 1e-10) over all samples before the segment
  float r, g, b, depth, acc;
};
__device__ __forceinline__ RayCarry carry_init() {
  RayCarry c; c.T = 1.0; c.r = c.g = c.b = c.depth = c.acc = 0.f; return c;
}

// raw4[i] = (r,g,b,sigma) of sample s0+i;  z indexed by absolute sample;  noise / w_out
// indexed by absolute sample (may be null).  All lanes return the updated carry.
__device__ __forceinline__ RayCarry composite_segment(const float4* raw4, const float* z, int S, int s0, int cnt,
                                                      float dnorm, const float* __restrict__ noise,
                                                      float* w_out0, float* w_out1, RayCarry carry, int lane) {
  const int C = (cnt + 31) >> 5;  // contiguous samples per lane (<= 8)
  const int i0 = lane * C;
  float alpha[8];
  double p = 1.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    alpha[j] = 0.f;
    const int i = i0 + j;
    if (j < C && i < cnt) {
      const int s = s0 + i;
      float dist = (s == S - 1) ? kHuge : __fsub_rn(z[s + 1], z[s]);
      dist = __fmul_rn(dist, dnorm);
      float sig = raw4[i].w;
      if (noise) sig = __fadd_rn(sig, noise[s]);
      const float a = __fsub_rn(1.f, expf(__fmul_rn(-fmaxf(sig, 0.f), dist)));
      alpha[j] = a;
      p *= (double)__fadd_rn(__fsub_rn(1.f, a), 1e-10f);
    }
  }
  const double incl = warp_scan_prod_d(p, lane);
  double excl = shfl_up_d(incl, 1);
  if (lane == 0) excl = 1.0;
  double T = carry.T * excl;
  float sr = 0.f, sg = 0.f, sb = 0.f, sd = 0.f, sa = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int i = i0 + j;
    if (j < C && i < cnt) {
      const int s = s0 + i;
      const float w = __fmul_rn(alpha[j], (float)T);
      T *= (double)__fadd_rn(__fsub_rn(1.f, alpha[j]), 1e-10f);
      const float4 q = raw4[i];
      sr += w * __fdiv_rn(1.f, __fadd_rn(1.f, expf(-q.x)));
      sg += w * __fdiv_rn(1.f, __fadd_rn(1.f, expf(-q.y)));
      sb += w * __fdiv_rn(1.f, __fadd_rn(1.f, expf(-q.z)));
      sd += w * z[s];
      sa += w;
      if (w_out0) w_out0[s] = w;
      if (w_out1) w_out1[s] = w;
    }
  }
  RayCarry out;
  out.T = carry.T * shfl_d(incl, 31);
  out.r = carry.r + warp_sum(sr);
  out.g = carry.g + warp_sum(sg);
  out.b = carry.b + warp_sum(sb);
  out.depth = carry.depth + warp_sum(sd);
  out.acc = carry.acc + warp_sum(sa);
  return out;
}

// disp = 1 / max(1e-10, depth/acc) with torch.max's NaN propagation (run_nerf_helpers.py:418)
__device__ __forceinline__ float disparity(float depth, float acc) {
  const float q = __fdiv_rn(depth, acc);
  const float m = (q != q) ? q : fmaxf(1e-10f, q);
  return __fdiv_rn(1.f, m);
}

// ------------------------------------------------------------------------------------
// inverse-CDF resampling: one warp, one ray.
// ------------------------------------------------------------------------------------
// cdf[0..B) from the interior weights w_in[0..B-1) (= weights[1:-1]); bins unused here.
__device__ __forceinline__ void build_cdf(const float* w_in, int B, float* cdf, int lane) {
  const int n = B - 1;
  const int C = (n + 31) >> 5;
  const int i0 = lane * C;
  double loc = 0.0;
  for (int j = 0; j < C; ++j) {
    const int i = i0 + j;
    if (i < n) loc += (double)__fadd_rn(w_in[i], 1e-5f);
  }
  const float total = (float)warp_sum_d(loc);
  double run = 0.0;
  float pdf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    pdf[j] = 0.f;
    const int i = i0 + j;
    if (j < C && i < n) { pdf[j] = __fdiv_rn(__fadd_rn(w_in[i], 1e-5f), total); run += (double)pdf[j]; }
  }
  const double incl = warp_scan_sum_d(run, lane);
  double acc = incl - run;  // exclusive prefix of this lane's chunk
  if (lane == 0) cdf[0] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int i = i0 + j;
    if (j < C && i < n) { acc += (double)pdf[j]; cdf[i + 1] = (float)acc; }
  }
}

// torch.searchsorted(cdf, u, right=True): number of entries <= u.
__device__ __forceinline__ int upper_bound(const float* a, int n, float u) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= u) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int lower_bound(const float* a, int n, float u) {  // entries < u
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < u) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ float invert_cdf_one(const float* bins, const float* cdf, int B, float u, int* ind_out) {
  const int ind = upper_bound(cdf, B, u);
  const int below = max(ind - 1, 0);
  const int above = min(ind, B - 1);
  const float cb = cdf[below], ca = cdf[above];
  const float bb = bins[below], ba = bins[above];
  float den = __fsub_rn(ca, cb);
  if (den < 1e-5f) den = 1.f;
  const float t = __fdiv_rn(__fsub_rn(u, cb), den);
  *ind_out = ind;
  return __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
}

// ascending bitonic sort of a[0..n) in shared memory by one warp; pad[] must hold P>=n floats (P pow2)
__device__ __forceinline__ void warp_sort(float* a, int n, int lane) {
  int P = 1;
  while (P < n) P <<= 1;
  for (int i = n + lane; i < P; i += 32) a[i] = __int_as_float(0x7f800000);  // +inf padding
  __syncwarp();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (P >> 1); t += 32) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const bool up = ((i & k) == 0);
        const float x = a[i], y = a[l];
        if ((x > y) == up) { a[i] = y; a[l] = x; }
      }
      __syncwarp();
    }
  }
}

// merged[rank] of two ascending lists (render.py:383: sort(cat[z_vals, z_samples])), one warp.
__device__ __forceinline__ void merge_sorted(const float* a, int na, const float* b, int nb, float* out, int lane) {
  for (int i = lane; i < na; i += 32) out[i + lower_bound(b, nb, a[i])] = a[i];
  for (int j = lane; j < nb; j += 32) out[j + upper_bound(a, na, b[j])] = b[j];
}

// population std of x[0..n) (torch.std(unbiased=False), render.py:401), one warp, all lanes get it
__device__ __forceinline__ float warp_std(const float* x, int n, int lane) {
  double s = 0.0;
  for (int i = lane; i < n; i += 32) s += (double)x[i];
  const double mean = warp_sum_d(s) / (double)n;
  double v = 0.0;
  for (int i = lane; i < n; i += 32) { const double d = (double)x[i] - mean; v += d * d; }
  return (float)sqrt(warp_sum_d(v) / (double)n);
}

}  // namespace snerf
