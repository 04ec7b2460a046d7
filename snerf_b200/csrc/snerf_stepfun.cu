// snerf_stepfun.cu -- zip-NeRF's proposal resampling step as ONE kernel (BASELINE configs[3]).
//
// Replaces one pass of the sampling loop of s-nerfpp/zipnerf/internal/models.py:156-213, which the reference runs as
// ~60 torch launches over [N, 3S+1, S] and [N, T+1, n] broadcast tensors:
//     stepfun.max_dilate_weights (stepfun.py:91-105; max_dilate :75-88, weight_to_pdf :64-67, pdf_to_weight :70-72)
//     sdist[..., 1:-1], weights[..., 1:-1]                                   (models.py:181-182)
//     logits = where(sdist[1:] > sdist[:-1], anneal * log(weights + padding), -inf)        (models.py:193-196)
//     stepfun.sample_intervals (stepfun.py:251-294; sample :175-218, invert_cdf :154-161, integrate_weights :108-128,
//                               math.sorted_interp math.py:88-107)
//
// One warp per ray, every per-ray array in shared memory; the only HBM traffic is the ray's inputs (S+1 + S floats),
// its jitter draw and the n+1 new interval edges.  No tensor cores: this is sorted-merge / scan / binary-search work.
//   * the sort of cat[t, t - d, t + d] is a 3-way merge of sorted runs: every element finds its rank with two binary
//     searches (stable tie-break by run), no sorting network;
//   * the max-pool over covering intervals scans only the contiguous range of intervals that can cover the point
//     (both interval ends are sorted), instead of the reference's [3S+1, S] mask;
//   * softmax / cumulative sum: lane-contiguous chunks + a serial carry across lanes, so the CDF is monotone by
//     construction like the reference's sequential cumsum;
//   * the inverse CDF is a binary search per sample instead of the [T+1, n] comparison mask.
// fp32 operation order follows the reference (explicit __f*_rn where nvcc would contract); the order-dependent sums
// (softmax denominator, cumsum, renormalisation) agree to rounding.
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "snerf_internal.h"

namespace snerf {
namespace {

constexpr int kSfMaxBins = 128;                      // S (bins of the incoming step function)
constexpr int kSfMaxEdges = 3 * kSfMaxBins + 1;      // 3S + 1 dilated edges
constexpr int kSfMaxOut = 256;                       // n (intervals to sample)
constexpr int kSfWarps = 4;
constexpr float kEps = 1.1920928955078125e-07f;      // torch.finfo(float32).eps

struct SfArgs {
  const float* t;          // [N, S+1] sorted
  const float* w;          // [N, S]   weights, or logits when logits_in
  const float* u_base;     // [n]      linspace part of u (torch.linspace values, stepfun.py:205-216)
  const float* jitter;     // [N, jd]  torch.rand draw, or null (deterministic)
  float* out;              // [N, n+1] or null
  float* centers;          // [N, n]   or null   (the sampled points before the midpoint step; tests)
  float* t_dil;            // [N, 3S+1] or null  (max_dilate_weights outputs)
  float* w_dil;            // [N, 3S]   or null
  long long N;
  int S, n, jd;
  int dilate, renormalize, logits_in;
  float dilation, lo, hi, anneal, padding, max_jitter;
};

struct alignas(16) SfWarpSmem {
  float t[kSfMaxBins + 1];
  float p[kSfMaxBins];
  float e[kSfMaxEdges];      // dilated (or plain) edges; the step function that gets sampled
  float v[kSfMaxEdges];      // dilated weights -> logits -> softmax -> CDF (in place)
  float c[kSfMaxOut];        // sampled centres
};

// first index in sorted a[0..n) with a[i] >= x / a[i] > x
__device__ __forceinline__ int lower_bound(const float* a, int n, float x) {
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < x) lo = mid + 1; else hi = mid; }
  return lo;
}
__device__ __forceinline__ int upper_bound(const float* a, int n, float x) {
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] <= x) lo = mid + 1; else hi = mid; }
  return lo;
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
__device__ __forceinline__ float warp_max(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}

__global__ void __launch_bounds__(kSfWarps * 32) stepfun_resample_kernel(const SfArgs a) {
  __shared__ SfWarpSmem sm_all[kSfWarps];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  SfWarpSmem& sm = sm_all[wid];
  const long long ray = (long long)blockIdx.x * kSfWarps + wid;
  if (ray >= a.N) return;
  const int S = a.S;
  if (a.dilate) {
    for (int i = lane; i <= S; i += 32) sm.t[i] = a.t[ray * (S + 1) + i];
    for (int i = lane; i < S; i += 32) sm.p[i] = a.w[ray * S + i];
  } else {                     // sampled as it is: straight into the edge / value arrays (up to 3 * 128 - 1 bins)
    for (int i = lane; i <= S; i += 32) sm.e[i] = a.t[ray * (S + 1) + i];
    for (int i = lane; i < S; i += 32) sm.v[i] = a.w[ray * S + i];
  }
  __syncwarp();

  int T;                       // bins of the step function (sm.e[0..T], sm.v[0..T-1]) that gets sampled
  int e0 = 0;                  // first edge of that step function inside sm.e
  if (a.dilate) {
    // ---- stepfun.max_dilate_weights: p = w / max(dt, eps); edges = sort(cat[t, t - d, t + d]) clipped to the domain
    const float d = a.dilation;
    for (int j = lane; j < S; j += 32)                        // weight_to_pdf, in place (lane-private slots)
      sm.p[j] = __fdiv_rn(sm.p[j], fmaxf(__fsub_rn(sm.t[j + 1], sm.t[j]), kEps));
    __syncwarp();
    // ranks by binary search.  Runs: A = t[0..S], B = t[j] - d (j < S), C = t[j+1] + d (j < S); x - d and x + d are
    // monotone in x, so B and C are searched through t itself: #{B < e} = #{j < S : t[j] - d < e}
    const int NE = 3 * S + 1;
    auto cnt_b = [&](float e, bool or_equal) {   // #{j < S: t[j] - d < e (or <=)}
      int lo = 0, hi = S;
      while (lo < hi) { const int mid = (lo + hi) >> 1; const float x = __fsub_rn(sm.t[mid], d); if (or_equal ? x <= e : x < e) lo = mid + 1; else hi = mid; }
      return lo;
    };
    auto cnt_c = [&](float e, bool or_equal) {   // #{j < S: t[j+1] + d < e (or <=)}
      int lo = 0, hi = S;
      while (lo < hi) { const int mid = (lo + hi) >> 1; const float x = __fadd_rn(sm.t[mid + 1], d); if (or_equal ? x <= e : x < e) lo = mid + 1; else hi = mid; }
      return lo;
    };
    for (int i = lane; i <= S; i += 32) {                       // run A: ties with B, C go first
      const float e = sm.t[i];
      sm.e[i + cnt_b(e, false) + cnt_c(e, false)] = fminf(fmaxf(e, a.lo), a.hi);
    }
    for (int j = lane; j < S; j += 32) {
      const float eb = __fsub_rn(sm.t[j], d);                   // run B: after equal A, before equal C
      sm.e[j + upper_bound(sm.t, S + 1, eb) + cnt_c(eb, false)] = fminf(fmaxf(eb, a.lo), a.hi);
      const float ec = __fadd_rn(sm.t[j + 1], d);               // run C: after equal A and B
      sm.e[j + upper_bound(sm.t, S + 1, ec) + cnt_b(ec, true)] = fminf(fmaxf(ec, a.lo), a.hi);
    }
    __syncwarp();
    // max-pool: interval j covers x when t0[j] <= x < t1[j]; both ends sorted -> j in [#{t1 <= x}, #{t0 <= x})
    float part = 0.f;
    for (int k = lane; k < NE - 1; k += 32) {
      const float x = sm.e[k];
      const int j_end = cnt_b(x, true), j_beg = cnt_c(x, true);
      float m = 0.f;
      for (int j = j_beg; j < j_end; ++j) m = fmaxf(m, sm.p[j]);
      const float wv = __fmul_rn(m, __fsub_rn(sm.e[k + 1], x));        // pdf_to_weight
      sm.v[k] = wv;
      part += wv;
    }
    if (a.renormalize) {
      const float tot = fmaxf(warp_sum(part), kEps);
      __syncwarp();
      for (int k = lane; k < NE - 1; k += 32) sm.v[k] = __fdiv_rn(sm.v[k], tot);
    }
    __syncwarp();
    if (a.t_dil) for (int k = lane; k < NE; k += 32) a.t_dil[ray * NE + k] = sm.e[k];
    if (a.w_dil) for (int k = lane; k < NE - 1; k += 32) a.w_dil[ray * (NE - 1) + k] = sm.v[k];
    // models.py:181-182: drop the first and last edge / bin
    e0 = 1;
    T = NE - 3;
  } else {
    T = S;
  }
  if (a.n <= 0) return;
  const float* E = sm.e + e0;      // T + 1 edges
  float* V = sm.v + e0;            // T values

  // ---- logits (models.py:193-196), softmax (stepfun.py:157), integrate_weights (:108-128), all in place in V
  float mx = -INFINITY;
  if (!a.logits_in) {
    for (int k = lane; k < T; k += 32) {
      const float l = E[k + 1] > E[k] ? __fmul_rn(a.anneal, logf(__fadd_rn(V[k], a.padding))) : -INFINITY;
      V[k] = l;
      mx = fmaxf(mx, l);
    }
  } else {
    for (int k = lane; k < T; k += 32) mx = fmaxf(mx, V[k]);
  }
  mx = warp_max(mx);
  __syncwarp();
  const int chunk = (T + 31) / 32;
  const int k0 = lane * chunk, k1 = min(T, k0 + chunk);
  float tot = 0.f;
  for (int k = k0; k < k1; ++k) { const float ev = expf(__fsub_rn(V[k], mx)); V[k] = ev; tot += ev; }
  const float denom = warp_sum(tot);
  // CDF edge k+1 = clamp_max(cumsum(w)[k], 1) for k < T-1; edge 0 = 0, edge T = 1.  Stored as cw[k] = CDF edge k.
  for (int k = k0; k < k1; ++k) V[k] = __fdiv_rn(V[k], denom);
  __syncwarp();
  // sequential cumulative sum, lane after lane (the reference's cumsum order; monotone by construction), written in
  // place shifted by one: V[k] <- CDF edge k = min(sum w[0..k), 1), edge 0 = 0
  float acc = 0.f;
  for (int l = 0; l < 32; ++l) {
    if (lane == l) {
      for (int k = k0; k < k1; ++k) {
        const float wv = V[k];
        V[k] = k == 0 ? 0.f : fminf(acc, 1.0f);
        acc = __fadd_rn(acc, wv);
      }
    }
    acc = __shfl_sync(0xffffffffu, acc, l);
  }
  __syncwarp();
  // edge T = 1 (sm.v has at least one slot after V[T-1]: e0 + T <= 3S - 1 < kSfMaxEdges)
  if (lane == 0) V[T] = 1.0f;
  __syncwarp();

  // ---- sample (stepfun.py:199-218) + sorted_interp (math.py:88-107) + interval midpoints (stepfun.py:281-293)
  const int n = a.n;
  for (int i = lane; i < n; i += 32) {
    float u = a.u_base[i];
    if (a.jitter) u = __fadd_rn(u, __fmul_rn(a.jitter[ray * a.jd + (a.jd == 1 ? 0 : i)], a.max_jitter));
    // mask = u >= cw[k]; last true index (cw[0] = 0 <= u always for u >= 0)
    int idx = upper_bound(V, T + 1, u) - 1;
    float x0, x1, f0, f1;
    if (idx < 0) { x0 = V[0]; f0 = E[0]; idx = -1; } else { x0 = V[idx]; f0 = E[idx]; }
    if (idx + 1 <= T) { x1 = V[idx + 1]; f1 = E[idx + 1]; } else { x1 = V[T]; f1 = E[T]; }
    float off = __fdiv_rn(__fsub_rn(u, x0), __fsub_rn(x1, x0));
    if (off != off) off = 0.f;                               // nan_to_num(nan -> 0); +-inf clip to [0, 1] below
    off = fminf(fmaxf(off, 0.f), 1.f);
    sm.c[i] = __fadd_rn(f0, __fmul_rn(off, __fsub_rn(f1, f0)));
  }
  __syncwarp();
  if (a.centers) for (int i = lane; i < n; i += 32) a.centers[ray * n + i] = sm.c[i];
  if (a.out) {
    float* o = a.out + ray * (n + 1);
    for (int i = lane; i < n - 1; i += 32) o[i + 1] = __fdiv_rn(__fadd_rn(sm.c[i + 1], sm.c[i]), 2.0f);
    if (lane == 0) {
      const float mid0 = __fdiv_rn(__fadd_rn(sm.c[1], sm.c[0]), 2.0f);
      o[0] = fmaxf(__fsub_rn(__fmul_rn(2.0f, sm.c[0]), mid0), a.lo);
      const float midl = __fdiv_rn(__fadd_rn(sm.c[n - 1], sm.c[n - 2]), 2.0f);
      o[n] = fminf(__fsub_rn(__fmul_rn(2.0f, sm.c[n - 1]), midl), a.hi);
    }
  }
}

}  // namespace

int stepfun_resample(const float* t, const float* w, long long N, int S, int dilate, int renormalize, int logits_in,
                     float dilation, float lo, float hi, float anneal, float padding, const float* u_base,
                     const float* jitter, int jd, float max_jitter, int n, float* out, float* centers, float* t_dil,
                     float* w_dil, cudaStream_t st) {
  const int max_bins = dilate ? kSfMaxBins : kSfMaxEdges - 2;     // a dilated step function (3S - 2 bins) can be sampled again
  if (S < 1 || S > max_bins) { set_error("stepfun: 1 <= bins <= %d (got %d)", max_bins, S); return SNERF_ERR_UNSUPPORTED; }
  if (n < 0 || n == 1 || n > kSfMaxOut) { set_error("stepfun: num_samples must be 0 or in [2, %d] (got %d)", kSfMaxOut, n); return SNERF_ERR_UNSUPPORTED; }
  SfArgs a;
  a.t = t; a.w = w; a.u_base = u_base; a.jitter = jitter; a.out = out; a.centers = centers; a.t_dil = t_dil; a.w_dil = w_dil;
  a.N = N; a.S = S; a.n = n; a.jd = jd; a.dilate = dilate; a.renormalize = renormalize; a.logits_in = logits_in;
  a.dilation = dilation; a.lo = lo; a.hi = hi; a.anneal = anneal; a.padding = padding; a.max_jitter = max_jitter;
  stepfun_resample_kernel<<<(unsigned)((N + kSfWarps - 1) / kSfWarps), kSfWarps * 32, 0, st>>>(a);
  return check_cuda(cudaGetLastError(), "launch stepfun_resample_kernel");
}

}  // namespace snerf
