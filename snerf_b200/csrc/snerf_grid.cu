// snerf_grid.cu -- multi-resolution hash-grid encoder for sm_100a (BASELINE configs[3]).
//
// Replaces the reference's only native component near the renderer, the torch extension
// s-nerfpp/zipnerf/gridencoder/src/gridencoder.cu (kernel_grid :87-245, kernel_grid_backward :248-340,
// kernel_input_backward :343-369, kernel_grad_tv :506-610) behind the same operator interface
// (grid_encode_forward / grid_encode_backward / grad_total_variation, gridencoder.h).
//
// This is HBM / L2 gather-bound integer + fp32 work, not a contraction: no tensor cores.  What is B200-specific:
//   * one thread encodes a point for a GROUP of consecutive levels and writes its features as ONE aligned 32-byte
//     store into the caller's layout (any level / point strides), so the [B, L*C] tensor the network consumes is
//     produced directly -- the reference writes [L, B, C] and pays a full permute pass (read + write of B*L*C);
//   * the grid y index (slow in block scheduling order) walks the level groups, so at any time the gathers hit one
//     or two level tables (<= 2 x 32 MiB at 2^21 x 4 fp32), which stay resident in the 126 MB L2;
//   * every corner is ONE 16/8-byte read-only gather (LDG.128 / LDG.64) instead of C scalar loads, and every
//     backward corner ONE vector reduction (RED.128 / RED.64 via atomicAdd(float4*/float2*), sm_90+) instead of C
//     scalar atomics from C/2 threads;
//   * streaming stores (st.global.cs) for outputs / dy_dx keep the tables, not the outputs, in L2.
// The arithmetic (FMA contraction pattern, corner order, uint32 index wrap-around) follows the reference kernels so
// the fp32 forward is bit-identical to them; tests compare against the reference's own kernels compiled for sm_100.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "snerf_internal.h"

namespace snerf {

namespace {

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// read-only gather of C consecutive features (16-byte aligned when C * sizeof(T) >= 16)
template <typename T, int C>
__device__ __forceinline__ void gather(const T* __restrict__ p, float (&out)[C]) {
  constexpr int BYTES = C * (int)sizeof(T);
  if constexpr (BYTES % 16 == 0) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < BYTES / 16; ++i) {
      const uint4 v = __ldg(q + i);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if constexpr (sizeof(T) == 4) out[i * 4 + j] = __uint_as_float(w[j]);
        else {
          const __half2 h = *reinterpret_cast<const __half2*>(&w[j]);
          out[i * 8 + 2 * j] = __low2float(h); out[i * 8 + 2 * j + 1] = __high2float(h);
        }
      }
    }
  } else if constexpr (BYTES == 8) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    if constexpr (sizeof(T) == 4) { out[0] = __uint_as_float(v.x); out[1] = __uint_as_float(v.y); }
    else {
      const __half2 a = *reinterpret_cast<const __half2*>(&v.x), b = *reinterpret_cast<const __half2*>(&v.y);
      out[0] = __low2float(a); out[1] = __high2float(a); out[2] = __low2float(b); out[3] = __high2float(b);
    }
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) out[c] = to_f(__ldg(p + c));
  }
}

struct GridArgs {
  const float* inputs;     // [B, D]
  const void* grid;        // [sO, C]
  const int32_t* offsets;  // [L + 1]
  long long B;
  int L, H, gridtype, align_corners, interp;
  float S;
};

struct LevelGeom {
  uint32_t hashmap_size, resolution, offset;
  float scale;
};
__device__ __forceinline__ LevelGeom level_geom(const GridArgs& a, int level) {
  LevelGeom g;
  g.offset = (uint32_t)a.offsets[level];
  g.hashmap_size = (uint32_t)a.offsets[level + 1] - g.offset;
  g.scale = exp2f(level * a.S) * a.H - 1.0f;           // same expression as the reference (nvcc contracts it to an FMA)
  g.resolution = (uint32_t)ceilf(g.scale) + 1;
  return g;
}
// gridencoder.cu:50-84 -- dense index while the running stride fits the table, else the xor-of-primes hash
template <int D>
__device__ __forceinline__ uint32_t cell_index(const uint32_t (&pg)[D], const LevelGeom& g, int gridtype, int align_corners) {
  constexpr uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
  uint32_t stride = 1, index = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    if (stride <= g.hashmap_size) {
      index += pg[d] * stride;
      stride *= align_corners ? g.resolution : (g.resolution + 1);
    }
  }
  if (gridtype == 0 && stride > g.hashmap_size) {
    index = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) index ^= pg[d] * primes[d];
  }
  return index % g.hashmap_size;
}
template <int D>
__device__ __forceinline__ bool load_point(const float* __restrict__ inputs, long long b, float (&x)[D]) {
  bool oob = false;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    x[d] = inputs[b * D + d];
    if (x[d] < 0 || x[d] > 1) oob = true;
  }
  return oob;
}
template <int D>
__device__ __forceinline__ void locate(const float (&x)[D], const LevelGeom& g, int align_corners, int interp,
                                       float (&pos)[D], float (&deriv)[D], uint32_t (&pg)[D]) {
#pragma unroll
  for (int d = 0; d < D; ++d) {
    pos[d] = x[d] * g.scale + (align_corners ? 0.0f : 0.5f);
    const float fl = floorf(pos[d]);
    pg[d] = (uint32_t)fl;
    pos[d] -= (float)pg[d];
    if (interp == 1) {
      deriv[d] = 6 * pos[d] * (1.0f - pos[d]);
      pos[d] = pos[d] * pos[d] * (3.0f - 2.0f * pos[d]);
    } else {
      deriv[d] = 1.0f;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// forward: thread = (point, group of LPT consecutive levels); out[level * sl + b * sb + c]
// ------------------------------------------------------------------------------------------------------------
template <typename T, int D, int C, int LPT, int MINB = 3>
__global__ void __launch_bounds__(256, MINB) grid_fwd_kernel(GridArgs a, T* __restrict__ out, long long sl, long long sb,
                                                       T* __restrict__ dy_dx) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  const int level0 = blockIdx.y * LPT;
  float x[D];
  const bool oob = load_point<D>(a.inputs, b, x);
  const T* grid = reinterpret_cast<const T*>(a.grid);
  alignas(16) T res[LPT][C];
#pragma unroll
  for (int j = 0; j < LPT; ++j) {
    const int level = level0 + j;
#pragma unroll
    for (int c = 0; c < C; ++c) res[j][c] = from_f<T>(0.f);
    if (level >= a.L) continue;
    T* dyp = dy_dx ? dy_dx + ((size_t)b * a.L + level) * D * C : nullptr;  // [B, L, D, C]
    if (oob) {
      if (dyp)
        for (int i = 0; i < D * C; ++i) dyp[i] = from_f<T>(0.f);
      continue;
    }
    const LevelGeom g = level_geom(a, level);
    const T* tab = grid + (size_t)g.offset * C;
    float pos[D], deriv[D];
    uint32_t pg[D];
    locate<D>(x, g, a.align_corners, a.interp, pos, deriv, pg);
#pragma unroll
    for (int idx = 0; idx < (1 << D); ++idx) {
      float w = 1;
      uint32_t loc[D];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        if ((idx & (1 << d)) == 0) { w *= 1 - pos[d]; loc[d] = pg[d]; }
        else { w *= pos[d]; loc[d] = pg[d] + 1; }
      }
      float v[C];
      gather<T, C>(tab + (size_t)cell_index<D>(loc, g, a.gridtype, a.align_corners) * C, v);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        if constexpr (sizeof(T) == 4) res[j][c] += w * v[c];
        else res[j][c] = from_f<T>(to_f(res[j][c]) + to_f(from_f<T>(w * v[c])));  // at::Half: round the product, then the sum
      }
    }
    if (dyp) {  // d(feature) / d(input), gridencoder.cu:200-243
#pragma unroll
      for (int gd = 0; gd < D; ++gd) {
        T rg[C];
#pragma unroll
        for (int c = 0; c < C; ++c) rg[c] = from_f<T>(0.f);
#pragma unroll
        for (int idx = 0; idx < (1 << (D - 1)); ++idx) {
          float w = g.scale;
          uint32_t loc[D];
#pragma unroll
          for (int nd = 0; nd < D - 1; ++nd) {
            const int d = (nd >= gd) ? (nd + 1) : nd;
            if ((idx & (1 << nd)) == 0) { w *= 1 - pos[d]; loc[d] = pg[d]; }
            else { w *= pos[d]; loc[d] = pg[d] + 1; }
          }
          float vl[C], vr[C];
          loc[gd] = pg[gd];
          gather<T, C>(tab + (size_t)cell_index<D>(loc, g, a.gridtype, a.align_corners) * C, vl);
          loc[gd] = pg[gd] + 1;
          gather<T, C>(tab + (size_t)cell_index<D>(loc, g, a.gridtype, a.align_corners) * C, vr);
#pragma unroll
          for (int c = 0; c < C; ++c) {
            if constexpr (sizeof(T) == 4) rg[c] += w * (vr[c] - vl[c]) * deriv[gd];
            else rg[c] = from_f<T>(to_f(rg[c]) + to_f(from_f<T>(w * to_f(from_f<T>(vr[c] - vl[c])) * deriv[gd])));
          }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) dyp[gd * C + c] = rg[c];
      }
    }
  }
  // ---- store: one vector store per thread when the level group is contiguous in the caller's layout
  constexpr int BYTES = LPT * C * (int)sizeof(T);
  constexpr int ALIGN = BYTES >= 16 ? 16 : BYTES;
  T* dst = out + (size_t)level0 * sl + (size_t)b * sb;
  const bool vec_ok = (LPT == 1 || sl == C) && level0 + LPT <= a.L && (reinterpret_cast<uintptr_t>(dst) & (ALIGN - 1)) == 0;
  if (vec_ok && BYTES % 16 == 0) {
    const float4* s = reinterpret_cast<const float4*>(&res[0][0]);
#pragma unroll
    for (int i = 0; i < BYTES / 16; ++i) __stcs(reinterpret_cast<float4*>(dst) + i, s[i]);
  } else if (vec_ok && BYTES == 8) {
    __stcs(reinterpret_cast<float2*>(dst), *reinterpret_cast<const float2*>(&res[0][0]));
  } else if (vec_ok && BYTES == 4) {
    __stcs(reinterpret_cast<float*>(dst), *reinterpret_cast<const float*>(&res[0][0]));
  } else {
#pragma unroll
    for (int j = 0; j < LPT; ++j)
      if (level0 + j < a.L)
#pragma unroll
        for (int c = 0; c < C; ++c) out[(size_t)(level0 + j) * sl + (size_t)b * sb + c] = res[j][c];
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward w.r.t. the table: thread = (point, level); one vector reduction per corner
// ------------------------------------------------------------------------------------------------------------
template <int C>
__device__ __forceinline__ void red_add(float* p, const float (&v)[C]) {
  if constexpr (C % 4 == 0) {
#pragma unroll
    for (int i = 0; i < C / 4; ++i)
      atomicAdd(reinterpret_cast<float4*>(p) + i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
  } else if constexpr (C == 2) {
    atomicAdd(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) atomicAdd(p + c, v[c]);
  }
}
template <int C>
__device__ __forceinline__ void red_add(__half* p, const float (&v)[C]) {
  if constexpr (C % 2 == 0) {
#pragma unroll
    for (int c = 0; c < C; c += 2) atomicAdd(reinterpret_cast<__half2*>(p + c), __floats2half2_rn(v[c], v[c + 1]));
  } else {
    atomicAdd(p, __float2half_rn(v[0]));
  }
}
template <typename T, int D, int C, int LPT>
__global__ void __launch_bounds__(256) grid_bwd_kernel(GridArgs a, const T* __restrict__ grad, long long sl, long long sb,
                                                       T* __restrict__ grad_grid) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  float x[D];
  if (load_point<D>(a.inputs, b, x)) return;  // grad_grid starts at 0: nothing to add
#pragma unroll
  for (int j = 0; j < LPT; ++j) {
    const int level = blockIdx.y * LPT + j;
    if (level >= a.L) break;
    const LevelGeom g = level_geom(a, level);
    float pos[D], deriv[D];
    uint32_t pg[D];
    locate<D>(x, g, a.align_corners, a.interp, pos, deriv, pg);
    float gc[C];
    gather<T, C>(grad + (size_t)level * sl + (size_t)b * sb, gc);
    T* tab = grad_grid + (size_t)g.offset * C;
#pragma unroll
    for (int idx = 0; idx < (1 << D); ++idx) {
      float w = 1;
      uint32_t loc[D];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        if ((idx & (1 << d)) == 0) { w *= 1 - pos[d]; loc[d] = pg[d]; }
        else { w *= pos[d]; loc[d] = pg[d] + 1; }
      }
      float v[C];
#pragma unroll
      for (int c = 0; c < C; ++c) v[c] = w * gc[c];
      red_add<C>(tab + (size_t)cell_index<D>(loc, g, a.gridtype, a.align_corners) * C, v);
    }
  }
}
// grad_inputs[b, d] = sum_l sum_c grad[l, b, c] * dy_dx[b, l, d, c]   (gridencoder.cu:343-369)
template <typename T, int D, int C>
__global__ void __launch_bounds__(256) grid_input_bwd_kernel(const T* __restrict__ grad, long long sl, long long sb,
                                                             const T* __restrict__ dy_dx, T* __restrict__ grad_inputs,
                                                             long long B, int L) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * D) return;
  const long long b = t / D;
  const int d = (int)(t - b * D);
  const T* dy = dy_dx + (size_t)b * L * D * C;
  T result = from_f<T>(0.f);
  for (int l = 0; l < L; ++l)
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float gv = to_f(grad[(size_t)l * sl + (size_t)b * sb + c]), dv = to_f(dy[(size_t)l * D * C + d * C + c]);
      if constexpr (sizeof(T) == 4) result += gv * dv;
      else result = from_f<T>(to_f(result) + to_f(from_f<T>(gv * dv)));
    }
  grad_inputs[t] = result;
}

// total-variation gradient on the cells the sample points fall into (gridencoder.cu:506-610)
template <typename T, int D, int C>
__global__ void __launch_bounds__(256) grid_tv_kernel(GridArgs a, T* __restrict__ grad, float weight) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  const int level = blockIdx.y;
  float x[D];
  if (load_point<D>(a.inputs, b, x)) return;
  const LevelGeom g = level_geom(a, level);
  const T* tab = reinterpret_cast<const T*>(a.grid) + (size_t)g.offset * C;
  uint32_t pg[D];
#pragma unroll
  for (int d = 0; d < D; ++d) pg[d] = (uint32_t)floorf(x[d] * g.scale + (a.align_corners ? 0.0f : 0.5f));
  const uint32_t index = cell_index<D>(pg, g, a.gridtype, a.align_corners);
  float centre[C], results[C], idelta[C];
  gather<T, C>(tab + (size_t)index * C, centre);
#pragma unroll
  for (int c = 0; c < C; ++c) results[c] = idelta[c] = 0.f;
  const float w = weight / (2 * D);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const uint32_t cur = pg[d];
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      if (side == 0 ? (cur < g.resolution) : (cur > 0)) {
        pg[d] = side == 0 ? cur + 1 : cur - 1;
        float other[C];
        gather<T, C>(tab + (size_t)cell_index<D>(pg, g, a.gridtype, a.align_corners) * C, other);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float gv = centre[c] - other[c];
          results[c] += gv;
          idelta[c] += gv * gv;
        }
      }
    }
    pg[d] = cur;
  }
  float v[C];
#pragma unroll
  for (int c = 0; c < C; ++c) v[c] = w * results[c] * rsqrtf(idelta[c] + 1e-9f);
  red_add<C>(grad + ((size_t)g.offset + index) * C, v);
}

// ------------------------------------------------------------------------------------------------------------
// zip-NeRF multisample featurisation fused into the encoder (s-nerfpp/zipnerf/internal/models.py:481-507):
//     x01      = (means + bound) / (2 bound)                                  (grid.py:159)
//     features = encoder(x01)                            [N, M, L, C]
//     w        = erf(1 / sqrt(8 stds^2 grid_sizes^2))    [N, M, L]
//     out[:, l*C + c] = mean_m(features * w)             [N, L*C]
//     out[:, L*C + l] = (2 mean_m(w) - 1) * level_gain[l]                     (scale_featurization)
// The reference materialises features ([N*M, L*C], 6x the result), w, their product and the mean as separate
// tensors; here one thread owns (sample n, LPT consecutive levels), walks the M multisamples and writes only the
// result.  fp32, input_dim 3.
// ------------------------------------------------------------------------------------------------------------
struct GridMsArgs {
  const float* means;          // [N, M, 3] in [-bound, bound]
  const float* stds;           // [N, M]
  const int32_t* grid_sizes;   // [L]  (GridEncoder.grid_sizes)
  const float* level_gain;     // [L] or null (no featurized_w columns)
  long long N;
  int M;
  float bound;
};
// erf(1 / sqrt(8 s^2 g^2)) with torch's operation order and roundings (models.py:493)
__device__ __forceinline__ float ms_weight(float std, int32_t gsize) {
  const float a = __fmul_rn(__fmul_rn(8.0f, __fmul_rn(std, std)), (float)(gsize * gsize));
  return erff(__fdiv_rn(1.0f, __fsqrt_rn(a)));
}
__device__ __forceinline__ bool ms_load_point(const GridMsArgs& m, long long p, float (&x)[3]) {
  bool oob = false;
  const float two_b = __fmul_rn(2.0f, m.bound);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    x[d] = __fdiv_rn(__fadd_rn(m.means[p * 3 + d], m.bound), two_b);
    if (x[d] < 0 || x[d] > 1) oob = true;
  }
  return oob;
}

template <int C, int LPT>
__global__ void __launch_bounds__(256) grid_ms_fwd_kernel(GridArgs a, GridMsArgs m, float* __restrict__ out, long long sn) {
  constexpr int D = 3;
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= m.N) return;
  const int level0 = blockIdx.y * LPT;
  const float* grid = reinterpret_cast<const float*>(a.grid);
  float acc[LPT][C], wsum[LPT];
  LevelGeom g[LPT];
  int32_t gs[LPT];
#pragma unroll
  for (int j = 0; j < LPT; ++j) {
    wsum[j] = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) acc[j][c] = 0.f;
    const int level = level0 + j < a.L ? level0 + j : a.L - 1;
    g[j] = level_geom(a, level);
    gs[j] = m.grid_sizes[level];
  }
  for (int i = 0; i < m.M; ++i) {
    const long long p = n * m.M + i;
    float x[D];
    const bool oob = ms_load_point(m, p, x);
    const float sd = m.stds[p];
#pragma unroll
    for (int j = 0; j < LPT; ++j) {
      const float w_ms = ms_weight(sd, gs[j]);
      wsum[j] += w_ms;
      if (oob) continue;                     // the encoder returns zero features outside [0, 1]
      const float* tab = grid + (size_t)g[j].offset * C;
      float pos[D], deriv[D];
      uint32_t pg[D];
      locate<D>(x, g[j], a.align_corners, a.interp, pos, deriv, pg);
      float res[C];
#pragma unroll
      for (int c = 0; c < C; ++c) res[c] = 0.f;
#pragma unroll
      for (int idx = 0; idx < (1 << D); ++idx) {
        float w = 1;
        uint32_t loc[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
          if ((idx & (1 << d)) == 0) { w *= 1 - pos[d]; loc[d] = pg[d]; }
          else { w *= pos[d]; loc[d] = pg[d] + 1; }
        }
        float v[C];
        gather<float, C>(tab + (size_t)cell_index<D>(loc, g[j], a.gridtype, a.align_corners) * C, v);
#pragma unroll
        for (int c = 0; c < C; ++c) res[c] += w * v[c];
      }
#pragma unroll
      for (int c = 0; c < C; ++c) acc[j][c] = __fadd_rn(acc[j][c], __fmul_rn(res[c], w_ms));
    }
  }
  const float inv_m = 1.0f / (float)m.M;
  float* row = out + (size_t)n * sn;
#pragma unroll
  for (int j = 0; j < LPT; ++j) {
    const int level = level0 + j;
    if (level >= a.L) continue;
    float* dst = row + level * C;
    if constexpr (C % 2 == 0) {
#pragma unroll
      for (int c = 0; c < C; c += 2) __stcs(reinterpret_cast<float2*>(dst + c), make_float2(acc[j][c] * inv_m, acc[j][c + 1] * inv_m));
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) __stcs(dst + c, acc[j][c] * inv_m);
    }
    if (m.level_gain)
      __stcs(row + a.L * C + level, __fmul_rn(__fadd_rn(__fmul_rn(2.0f, wsum[j] * inv_m), -1.0f), m.level_gain[level]));
  }
}

// v2 of the fused forward: thread = one multisample POINT (consecutive lanes = the M neighbouring points of a sample, which
// share most of their cells: their gathers coalesce in the same request, as in grid_fwd_kernel), weighted features staged
// in shared memory, then summed over the M points of each sample by the block.  SPB = 256 / M samples per block.
template <int C, int LPT>
__global__ void __launch_bounds__(256, 3) grid_ms_fwd2_kernel(GridArgs a, GridMsArgs m, float* __restrict__ out, long long sn,
                                                              int spb) {
  constexpr int D = 3;
  constexpr int NQ = LPT * (C + 1);            // values per point: LPT x (C weighted features + the weight)
  __shared__ float sh[NQ * 257];
  const int t = threadIdx.x;
  const int pts = spb * m.M;                   // points handled by this block (<= 256)
  const long long n0 = (long long)blockIdx.x * spb;
  const long long p = n0 * m.M + t;
  const int level0 = blockIdx.y * LPT;
  float val[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) val[q] = 0.f;
  if (t < pts && p < m.N * m.M) {
    const float* grid = reinterpret_cast<const float*>(a.grid);
    float x[D];
    const bool oob = ms_load_point(m, p, x);
    const float sd = m.stds[p];
#pragma unroll
    for (int j = 0; j < LPT; ++j) {
      const int level = level0 + j;
      if (level >= a.L) break;
      const float w_ms = ms_weight(sd, m.grid_sizes[level]);
      val[j * (C + 1) + C] = w_ms;
      if (oob) continue;
      const LevelGeom g = level_geom(a, level);
      const float* tab = grid + (size_t)g.offset * C;
      float pos[D], deriv[D];
      uint32_t pg[D];
      locate<D>(x, g, a.align_corners, a.interp, pos, deriv, pg);
      float res[C];
#pragma unroll
      for (int c = 0; c < C; ++c) res[c] = 0.f;
#pragma unroll
      for (int idx = 0; idx < (1 << D); ++idx) {
        float w = 1;
        uint32_t loc[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
          if ((idx & (1 << d)) == 0) { w *= 1 - pos[d]; loc[d] = pg[d]; }
          else { w *= pos[d]; loc[d] = pg[d] + 1; }
        }
        float v[C];
        gather<float, C>(tab + (size_t)cell_index<D>(loc, g, a.gridtype, a.align_corners) * C, v);
#pragma unroll
        for (int c = 0; c < C; ++c) res[c] += w * v[c];
      }
#pragma unroll
      for (int c = 0; c < C; ++c) val[j * (C + 1) + c] = __fmul_rn(res[c], w_ms);
    }
  }
#pragma unroll
  for (int q = 0; q < NQ; ++q) sh[q * 257 + t] = val[q];
  __syncthreads();
  const float inv_m = 1.0f / (float)m.M;
  for (int k = t; k < spb * NQ; k += blockDim.x) {
    const int s = k / NQ, q = k - s * NQ;
    const long long n = n0 + s;
    if (n >= m.N) break;
    const int j = q / (C + 1), c = q - j * (C + 1);
    const int level = level0 + j;
    if (level >= a.L) continue;
    const float* src = sh + q * 257 + s * m.M;
    float sum = 0.f;
    for (int i = 0; i < m.M; ++i) sum = __fadd_rn(sum, src[i]);
    float* row = out + (size_t)n * sn;
    if (c < C) row[level * C + c] = sum * inv_m;
    else if (m.level_gain) row[a.L * C + level] = __fmul_rn(__fadd_rn(__fmul_rn(2.0f, sum * inv_m), -1.0f), m.level_gain[level]);
  }
}

// gradient w.r.t. the table: thread = (multisample point, level); d out / d feature = w / M
template <int C>
__global__ void __launch_bounds__(256) grid_ms_bwd_kernel(GridArgs a, GridMsArgs m, const float* __restrict__ grad, long long sn,
                                                          float* __restrict__ grad_grid) {
  constexpr int D = 3;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= m.N * m.M) return;
  const int level = blockIdx.y;
  float x[D];
  if (ms_load_point(m, p, x)) return;
  const long long n = p / m.M;
  const LevelGeom g = level_geom(a, level);
  const float w_ms = ms_weight(m.stds[p], m.grid_sizes[level]);
  float pos[D], deriv[D];
  uint32_t pg[D];
  locate<D>(x, g, a.align_corners, a.interp, pos, deriv, pg);
  float gc[C];
  const float* gp = grad + (size_t)n * sn + level * C;      // rows of L*C (+L) floats: 8-byte aligned at best
  if constexpr (C % 2 == 0) {
#pragma unroll
    for (int c = 0; c < C; c += 2) {
      const float2 t = __ldg(reinterpret_cast<const float2*>(gp + c));
      gc[c] = t.x; gc[c + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) gc[c] = __ldg(gp + c);
  }
#pragma unroll
  for (int c = 0; c < C; ++c) gc[c] = __fmul_rn(__fdiv_rn(gc[c], (float)m.M), w_ms);   // mean backward, then * w
  float* tab = grad_grid + (size_t)g.offset * C;
#pragma unroll
  for (int idx = 0; idx < (1 << D); ++idx) {
    float w = 1;
    uint32_t loc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      if ((idx & (1 << d)) == 0) { w *= 1 - pos[d]; loc[d] = pg[d]; }
      else { w *= pos[d]; loc[d] = pg[d] + 1; }
    }
    float v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = w * gc[c];
    red_add<C>(tab + (size_t)cell_index<D>(loc, g, a.gridtype, a.align_corners) * C, v);
  }
}

// level_gain[l] = sqrt(init_std^2 + mean over the level's cells of |embedding|^2)   (models.py:496-503; the
// reference reduces with torch_scatter.segment_coo).  Partial sums go to a zero-filled double scratch [L].
template <int C>
__global__ void __launch_bounds__(256) grid_level_sqsum_kernel(const float* __restrict__ grid, const int32_t* __restrict__ offsets,
                                                               double* __restrict__ scratch) {
  const int level = blockIdx.y;
  const long long first = offsets[level], count = (long long)offsets[level + 1] - first;
  const float* tab = grid + (size_t)first * C;
  float part = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    float v[C];
    gather<float, C>(tab + (size_t)i * C, v);
    float t = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) t = fmaf(v[c], v[c], t);
    part += t;
  }
  double dsum = (double)part;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
  __shared__ double ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = dsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += ws[i];
    if (t != 0) atomicAdd(scratch + level, t);
  }
}
__global__ void grid_level_gain_kernel(const double* __restrict__ scratch, const int32_t* __restrict__ offsets, int L,
                                       float init_std, float* __restrict__ gain) {
  const int l = threadIdx.x;
  if (l >= L) return;
  const float mean = (float)(scratch[l] / (double)(offsets[l + 1] - offsets[l]));
  gain[l] = __fsqrt_rn(__fadd_rn(__fmul_rn(init_std, init_std), mean));
}

// ---------------------------------------------------------------- host dispatch
int check_desc(const SnerfGridDesc* d) {
  if (!d) { set_error("null grid descriptor"); return SNERF_ERR_BAD_ARG; }
  if (d->D < 2 || d->D > 4) { set_error("grid encoder: input_dim must be 2, 3 or 4 (got %d)", d->D); return SNERF_ERR_UNSUPPORTED; }
  if (d->C != 1 && d->C != 2 && d->C != 4 && d->C != 8) { set_error("grid encoder: level_dim must be 1, 2, 4 or 8 (got %d)", d->C); return SNERF_ERR_BAD_ARG; }
  if (d->L < 1 || d->L > 64 || d->H < 1) { set_error("grid encoder: bad level count / base resolution"); return SNERF_ERR_BAD_ARG; }
  if (d->dtype != 0 && d->dtype != 1) { set_error("grid encoder: dtype must be 0 (fp32) or 1 (fp16)"); return SNERF_ERR_BAD_ARG; }
  if (d->gridtype != 0 && d->gridtype != 1) { set_error("grid encoder: gridtype must be 0 (hash) or 1 (tiled)"); return SNERF_ERR_BAD_ARG; }
  if (d->interp != 0 && d->interp != 1) { set_error("grid encoder: interpolation must be 0 (linear) or 1 (smoothstep)"); return SNERF_ERR_BAD_ARG; }
  return SNERF_OK;
}
GridArgs make_args(const SnerfGridDesc* d, const float* inputs, const void* grid, const int32_t* offsets, long long B) {
  GridArgs a;
  a.inputs = inputs; a.grid = grid; a.offsets = offsets; a.B = B;
  a.L = d->L; a.H = d->H; a.gridtype = d->gridtype; a.align_corners = d->align_corners; a.interp = d->interp; a.S = d->S;
  return a;
}
// levels per thread: enough consecutive levels for a 32-byte store (capped at 4)
template <typename T, int C> constexpr int lpt_for() {
  constexpr int n = 32 / (C * (int)sizeof(T));
  return n < 1 ? 1 : (n > 4 ? 4 : n);
}

// bench-only A/B knob (SNERF_GRID_VARIANT bit mask, read once; default 6 = the measured best, profiles/r1c_grid.md):
// 1 = forward kernels take one level per thread (slower: 16-byte stores); 2 = forward at >= 4 CTAs / SM (64 registers,
// +5 %); 4 = backward takes a level group per thread (+3 %); 8 = fused multisample forward v1 (thread = sample,
// sequential multisamples: -32 %, the multisamples' shared cells no longer coalesce)
static int grid_variant() {
  static const int v = [] { const char* e = getenv("SNERF_GRID_VARIANT"); return e ? atoi(e) : 6; }();
  return v;
}
template <typename T, int D, int C>
int run_fwd(const GridArgs& a, void* out, long long sl, long long sb, void* dy_dx, cudaStream_t st) {
  // grouping levels only pays when they are adjacent in the output (the [B, L*C] layout); else one level per thread
  const int variant = grid_variant();
  constexpr int LPT = lpt_for<T, C>();
  const bool group = sl == C && LPT > 1 && !(variant & 1);
  dim3 grid((unsigned)((a.B + 255) / 256), (unsigned)(group ? (a.L + LPT - 1) / LPT : a.L));
  if (group) {
    if (variant & 2) grid_fwd_kernel<T, D, C, LPT, 4><<<grid, 256, 0, st>>>(a, (T*)out, sl, sb, (T*)dy_dx);
    else grid_fwd_kernel<T, D, C, LPT><<<grid, 256, 0, st>>>(a, (T*)out, sl, sb, (T*)dy_dx);
  } else {
    if (variant & 2) grid_fwd_kernel<T, D, C, 1, 4><<<grid, 256, 0, st>>>(a, (T*)out, sl, sb, (T*)dy_dx);
    else grid_fwd_kernel<T, D, C, 1><<<grid, 256, 0, st>>>(a, (T*)out, sl, sb, (T*)dy_dx);
  }
  return check_cuda(cudaGetLastError(), "launch grid_fwd_kernel");
}
template <typename T, int D, int C>
int run_bwd(const GridArgs& a, const void* grad, long long sl, long long sb, void* grad_grid, const void* dy_dx,
            void* grad_inputs, cudaStream_t st) {
  constexpr int LPT = lpt_for<T, C>();
  if (sl == C && LPT > 1 && (grid_variant() & 4)) {   // level groups: each thread reads a full 32-byte sector of its grad row
    dim3 grid((unsigned)((a.B + 255) / 256), (unsigned)((a.L + LPT - 1) / LPT));
    grid_bwd_kernel<T, D, C, LPT><<<grid, 256, 0, st>>>(a, (const T*)grad, sl, sb, (T*)grad_grid);
  } else {
    dim3 grid((unsigned)((a.B + 255) / 256), (unsigned)a.L);
    grid_bwd_kernel<T, D, C, 1><<<grid, 256, 0, st>>>(a, (const T*)grad, sl, sb, (T*)grad_grid);
  }
  if (dy_dx && grad_inputs)
    grid_input_bwd_kernel<T, D, C><<<(unsigned)((a.B * D + 255) / 256), 256, 0, st>>>((const T*)grad, sl, sb, (const T*)dy_dx,
                                                                                      (T*)grad_inputs, a.B, a.L);
  return check_cuda(cudaGetLastError(), "launch grid_bwd_kernel");
}
template <typename T, int D, int C>
int run_tv(const GridArgs& a, void* grad, float weight, cudaStream_t st) {
  dim3 grid((unsigned)((a.B + 255) / 256), (unsigned)a.L);
  grid_tv_kernel<T, D, C><<<grid, 256, 0, st>>>(a, (T*)grad, weight);
  return check_cuda(cudaGetLastError(), "launch grid_tv_kernel");
}

#define SNERF_GRID_DISPATCH(FN, ...)                                                                       \
  do {                                                                                                       \
    const int key = d->dtype * 1000 + d->D * 10 + d->C;                                                      \
    switch (key) {                                                                                           \
      case 21: return FN<float, 2, 1>(__VA_ARGS__);   case 22: return FN<float, 2, 2>(__VA_ARGS__);          \
      case 24: return FN<float, 2, 4>(__VA_ARGS__);   case 28: return FN<float, 2, 8>(__VA_ARGS__);          \
      case 31: return FN<float, 3, 1>(__VA_ARGS__);   case 32: return FN<float, 3, 2>(__VA_ARGS__);          \
      case 34: return FN<float, 3, 4>(__VA_ARGS__);   case 38: return FN<float, 3, 8>(__VA_ARGS__);          \
      case 41: return FN<float, 4, 1>(__VA_ARGS__);   case 42: return FN<float, 4, 2>(__VA_ARGS__);          \
      case 44: return FN<float, 4, 4>(__VA_ARGS__);   case 48: return FN<float, 4, 8>(__VA_ARGS__);          \
      case 1022: return FN<__half, 2, 2>(__VA_ARGS__); case 1024: return FN<__half, 2, 4>(__VA_ARGS__);      \
      case 1028: return FN<__half, 2, 8>(__VA_ARGS__); case 1032: return FN<__half, 3, 2>(__VA_ARGS__);      \
      case 1034: return FN<__half, 3, 4>(__VA_ARGS__); case 1038: return FN<__half, 3, 8>(__VA_ARGS__);      \
      default: break;                                                                                        \
    }                                                                                                        \
    set_error("grid encoder: (dtype=%d, input_dim=%d, level_dim=%d) is not instantiated", d->dtype, d->D, d->C); \
    return SNERF_ERR_UNSUPPORTED;                                                                            \
  } while (0)

}  // namespace

int grid_fwd(const SnerfGridDesc* d, const float* inputs, const void* emb, const int32_t* offsets, void* out,
             long long sl, long long sb, void* dy_dx, long long B, cudaStream_t st) {
  const GridArgs a = make_args(d, inputs, emb, offsets, B);
  SNERF_GRID_DISPATCH(run_fwd, a, out, sl, sb, dy_dx, st);
}
int grid_bwd(const SnerfGridDesc* d, const void* grad, long long sl, long long sb, const float* inputs,
             const int32_t* offsets, void* grad_emb, const void* dy_dx, void* grad_inputs, long long B, cudaStream_t st) {
  const GridArgs a = make_args(d, inputs, nullptr, offsets, B);
  SNERF_GRID_DISPATCH(run_bwd, a, grad, sl, sb, grad_emb, dy_dx, grad_inputs, st);
}
int grid_tv(const SnerfGridDesc* d, const float* inputs, const void* emb, void* grad, const int32_t* offsets,
            float weight, long long B, cudaStream_t st) {
  if (d->dtype != 0) { set_error("grad_total_variation runs in fp32 (as the reference: autocast disabled)"); return SNERF_ERR_UNSUPPORTED; }
  const GridArgs a = make_args(d, inputs, emb, offsets, B);
  SNERF_GRID_DISPATCH(run_tv, a, grad, weight, st);
}
int grid_check_desc(const SnerfGridDesc* d) { return check_desc(d); }
namespace {
template <int C>
int run_ms_fwd(const GridArgs& a, const GridMsArgs& m, float* out, long long sn, cudaStream_t st) {
  constexpr int LPT = C >= 8 ? 1 : 2;
  if (!(grid_variant() & 8) && m.M <= 256) {       // v2: thread = point, block reduction over the multisamples
    const int spb = 256 / m.M;
    const unsigned bx = (unsigned)((m.N + spb - 1) / spb);
    if (grid_variant() & 1) grid_ms_fwd2_kernel<C, 1><<<dim3(bx, (unsigned)a.L), 256, 0, st>>>(a, m, out, sn, spb);
    else grid_ms_fwd2_kernel<C, LPT><<<dim3(bx, (unsigned)((a.L + LPT - 1) / LPT)), 256, 0, st>>>(a, m, out, sn, spb);
  } else if (grid_variant() & 1) {
    grid_ms_fwd_kernel<C, 1><<<dim3((unsigned)((m.N + 255) / 256), (unsigned)a.L), 256, 0, st>>>(a, m, out, sn);
  } else {
    dim3 grid((unsigned)((m.N + 255) / 256), (unsigned)((a.L + LPT - 1) / LPT));
    grid_ms_fwd_kernel<C, LPT><<<grid, 256, 0, st>>>(a, m, out, sn);
  }
  return check_cuda(cudaGetLastError(), "launch grid_ms_fwd_kernel");
}
template <int C>
int run_ms_bwd(const GridArgs& a, const GridMsArgs& m, const float* grad, long long sn, float* grad_grid, cudaStream_t st) {
  dim3 grid((unsigned)((m.N * m.M + 255) / 256), (unsigned)a.L);
  grid_ms_bwd_kernel<C><<<grid, 256, 0, st>>>(a, m, grad, sn, grad_grid);
  return check_cuda(cudaGetLastError(), "launch grid_ms_bwd_kernel");
}
template <int C>
int run_level_gain(const float* emb, const int32_t* offsets, int L, float init_std, double* scratch, float* gain, cudaStream_t st) {
  grid_level_sqsum_kernel<C><<<dim3(148 * 2, (unsigned)L), 256, 0, st>>>(emb, offsets, scratch);
  grid_level_gain_kernel<<<1, 64, 0, st>>>(scratch, offsets, L, init_std, gain);
  return check_cuda(cudaGetLastError(), "launch grid_level_gain_kernel");
}
int check_ms(const SnerfGridDesc* d) {
  if (int e = check_desc(d)) return e;
  if (d->D != 3 || d->dtype != 0) { set_error("multisample grid encode: input_dim 3 and fp32 only (got D=%d, dtype=%d)", d->D, d->dtype); return SNERF_ERR_UNSUPPORTED; }
  return SNERF_OK;
}
#define SNERF_GRID_C_DISPATCH(FN, ...)                     \
  switch (d->C) {                                          \
    case 1: return FN<1>(__VA_ARGS__);                     \
    case 2: return FN<2>(__VA_ARGS__);                     \
    case 4: return FN<4>(__VA_ARGS__);                     \
    default: return FN<8>(__VA_ARGS__);                    \
  }
}  // namespace

int grid_ms_fwd(const SnerfGridDesc* d, const float* means, const float* stds, float bound, const void* emb,
                const int32_t* offsets, const int32_t* grid_sizes, const float* level_gain, float* out, long long sn,
                long long N, int M, cudaStream_t st) {
  if (int e = check_ms(d)) return e;
  const GridArgs a = make_args(d, nullptr, emb, offsets, N);
  const GridMsArgs m{means, stds, grid_sizes, level_gain, N, M, bound};
  SNERF_GRID_C_DISPATCH(run_ms_fwd, a, m, out, sn, st);
}
int grid_ms_bwd(const SnerfGridDesc* d, const float* grad, long long sn, const float* means, const float* stds, float bound,
                const int32_t* offsets, const int32_t* grid_sizes, float* grad_emb, long long N, int M, cudaStream_t st) {
  if (int e = check_ms(d)) return e;
  const GridArgs a = make_args(d, nullptr, nullptr, offsets, N);
  const GridMsArgs m{means, stds, grid_sizes, nullptr, N, M, bound};
  SNERF_GRID_C_DISPATCH(run_ms_bwd, a, m, grad, sn, grad_emb, st);
}
int grid_level_gain(const SnerfGridDesc* d, const void* emb, const int32_t* offsets, float init_std, double* scratch,
                    float* gain, cudaStream_t st) {
  if (int e = check_ms(d)) return e;
  if (d->L > 64) { set_error("level gain: at most 64 levels"); return SNERF_ERR_BAD_ARG; }
  SNERF_GRID_C_DISPATCH(run_level_gain, (const float*)emb, offsets, d->L, init_std, scratch, gain, st);
}


}  // namespace snerf
