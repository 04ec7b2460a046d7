// snerf_fp32.cu -- reference-accurate renderer: the whole render_rays pipeline in one kernel
// with the MLP on CUDA-core FFMA, fp32 activations kept in shared memory.
//
//   * one CTA renders one ray at a time (grid-stride over rays): coarse depths -> encode ->
//     coarse MLP -> composite -> inverse-CDF -> merge -> encode -> fine MLP -> composite;
//     nothing but the per-ray outputs goes to HBM (render.py:281-409);
//   * MLP tile = 64 samples.  Activations live transposed ([channel][row], row pitch 68) in two
//     ping-pong buffers; each of the 8 compute warps owns 8 rows x all columns and accumulates
//     8 x (n_out/32) outputs per lane; the K loop broadcasts the 8 activations (2 LDS.128) and
//     reads n_out/32 weights per lane (conflict-free LDS.32);
//   * weights stream from L2 through a 3-stage ring of 16-row K-chunks filled by a dedicated
//     producer warp with bulk async copies (cp.async.bulk, SASS UBLKCP) and mbarriers;
//   * two more front-ends reuse the tile routine: network_query_fn (pts + viewdirs -> raw) and
//     NeRF.forward on pre-encoded rows (run_nerf_helpers.py:103-126,460-474).
//
// This is the parity mode (fp32 end to end, 1e-4 relative vs the reference).  The throughput mode
// is the tcgen05 kernel in snerf_bf16.cu.
#include "snerf_fp32_core.cuh"
#include "snerf_internal.h"

namespace snerf {

constexpr int kNarrowMax = 4 * 256;

template <int W>
struct alignas(128) Fp32Smem {
  float wstage[kStages][kFp32ChunkRows * W];
  float encT[kEncRows * kLd];
  float dirT[kDirRows * kLd];
  float actX[W * kLd];
  float actY[W * kLd];
  float raw[kMaxSamples * 4];
  float zc[kMaxSamples];
  float zf[kMaxSamples];
  float wts[kMaxSamples];
  float cdf[kMaxSamples];
  float zs[kMaxSamples];
  float bins[kMaxSamples];
  float narrow_w[kNarrowMax];
  float direnc[kDirRows];
  Fp32Layer layers[4][kFp32MaxLayers];  // coarse, fine, frozen-sigma coarse, frozen-sigma fine
  int n_layers[4];
  uint64_t full[kStages];
  uint64_t empty[kStages];
};

template <int W>
__device__ __forceinline__ Fp32Ring ring_of(Fp32Smem<W>& sm) {
  Fp32Ring rg;
  rg.wstage = &sm.wstage[0][0]; rg.full = sm.full; rg.empty = sm.empty; rg.stage_floats = kFp32ChunkRows * W;
  return rg;
}

__device__ __forceinline__ int tiles_of(int n) { return (n + kTileRows - 1) / kTileRows; }
__device__ __forceinline__ float pow2i(int o) { return __int_as_float((127 + o) << 23); }

// ------------------------------------------------------------------------------------
// producer: stream one network's wide-layer chunks for one tile
// ------------------------------------------------------------------------------------
template <int W>
__device__ __forceinline__ void stream_net(Fp32Smem<W>& sm, int net, const unsigned char* img, int& stage,
                                           uint32_t& phase) {
  const int nl = sm.n_layers[net];
  for (int l = 0; l < nl; ++l) {
    const Fp32Layer& L = sm.layers[net][l];
    if (L.kind != 0) continue;
    const int K = L.seg_rows[0] + L.seg_rows[1] + L.seg_rows[2];
    ring_stream(ring_of(sm), reinterpret_cast<const float*>(img) + L.w_off, K, L.n_out, stage, phase);
  }
}

// ------------------------------------------------------------------------------------
// consumer: one wide layer  dst[n][r] = act(bias[n] + sum_k in[k][r] * Wt[k][n])
// ------------------------------------------------------------------------------------
// `gsave` (training): global [channel][R] slice this tile's outputs are also written to (row r0 of channel 0), or null
template <int W, int NJ>
__device__ __forceinline__ void wide_layer(Fp32Smem<W>& sm, const Fp32Layer& L, const float* __restrict__ bias,
                                           int& stage, uint32_t& phase, int warp, int lane, float* gsave,
                                           long long R) {
  float acc[8][NJ];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;
  const int r0 = warp * 8;
  const float* src = L.src ? sm.actY : sm.actX;
  float* dst = L.dst ? sm.actY : sm.actX;
  const Fp32Ring rg = ring_of(sm);
  for (int seg = 0; seg < 3; ++seg) {
    const int rows = L.seg_rows[seg];
    if (rows == 0) continue;
    const float* abase = (seg == 0 ? sm.encT : (seg == 1 ? src : sm.dirT)) + r0;
    ring_gemm<NJ>(rg, abase, rows, acc, stage, phase, lane);
  }
  const bool relu = L.relu != 0;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int n = lane + 32 * j;
    const float b = __ldg(bias + n);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = acc[i][j] + b;
      if (relu) v[i] = fmaxf(v[i], 0.f);
    }
    float* d = dst + n * kLd + r0;
    *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(d + 4) = make_float4(v[4], v[5], v[6], v[7]);
    if (gsave) {
      float* g = gsave + (long long)n * R + r0;
      *reinterpret_cast<float4*>(g) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(g + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
}

// narrow head (<= 4 outputs): raw[row0 + r][dst + n] = bias[n] + sum_k src[k][r] * w[n][k]
template <int W>
__device__ __forceinline__ void narrow_layer(Fp32Smem<W>& sm, const Fp32Layer& L, const unsigned char* img, int row0,
                                             int tid) {
  const int K = L.seg_rows[1];
  const float* wg = reinterpret_cast<const float*>(img) + L.w_off;
  const float* bg = reinterpret_cast<const float*>(img) + L.b_off;
  for (int i = tid; i < L.n_out * K; i += kComputeThreads) sm.narrow_w[i] = __ldg(wg + i);
  named_bar_sync(1, kComputeThreads);
  const float* src = L.src ? sm.actY : sm.actX;
  const int r = tid & 63;
  for (int n = tid >> 6; n < L.n_out; n += 4) {
    const float* w = sm.narrow_w + n * K;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int k = 0; k < K; k += 4) {
      a0 = fmaf(src[(k + 0) * kLd + r], w[k + 0], a0);
      a1 = fmaf(src[(k + 1) * kLd + r], w[k + 1], a1);
      a2 = fmaf(src[(k + 2) * kLd + r], w[k + 2], a2);
      a3 = fmaf(src[(k + 3) * kLd + r], w[k + 3], a3);
    }
    sm.raw[(row0 + r) * 4 + L.dst + n] = ((a0 + a1) + (a2 + a3)) + __ldg(bg + n);
  }
}

// `save` (training): this pass's activation store [channel][R] offset to the tile's first row, or null
template <int W>
__device__ __forceinline__ void mlp_tile(Fp32Smem<W>& sm, int net, const unsigned char* img, int row0, int& stage,
                                         uint32_t& phase, int tid, float* save = nullptr, long long R = 0) {
  const int warp = tid >> 5, lane = tid & 31;
  const int nl = sm.n_layers[net];
  for (int l = 0; l < nl; ++l) {
    const Fp32Layer& L = sm.layers[net][l];
    if (L.kind == 0) {
      const float* bias = reinterpret_cast<const float*>(img) + L.b_off;
      float* gs = save ? save + (long long)L.ch_off * R : nullptr;
      // lane l owns columns l + 32 j, j < n_out / 32: the network's own width W' in {W, W/2, W/4} (a kernel sized for the
      // wider of the coarse / fine architectures also runs the narrower one) or W'/2 for its views layer
      switch (L.n_out >> 5) {
        case 8: if constexpr (W >= 256) wide_layer<W, 8>(sm, L, bias, stage, phase, warp, lane, gs, R); break;
        case 4: if constexpr (W >= 128) wide_layer<W, 4>(sm, L, bias, stage, phase, warp, lane, gs, R); break;
        case 2: wide_layer<W, 2>(sm, L, bias, stage, phase, warp, lane, gs, R); break;
        default: wide_layer<W, 1>(sm, L, bias, stage, phase, warp, lane, gs, R); break;
      }
    } else {
      narrow_layer<W>(sm, L, img, row0, tid);
    }
    named_bar_sync(1, kComputeThreads);
  }
}

// encode the 64 rows of one tile of a ray: sample s = tile*64 + r at depth zbuf[s]
template <int W>
__device__ __forceinline__ void encode_ray_tile(Fp32Smem<W>& sm, const Ray& ray, const float* zbuf, int count,
                                                int tile, int L, int tid) {
  const int r = tid & 63, part = tid >> 6;
  const int s = min(tile * kTileRows + r, count - 1);
  const float z = zbuf[s];
  const float p[3] = {ray_point(ray.ox, ray.dx, z), ray_point(ray.oy, ray.dy, z), ray_point(ray.oz, ray.dz, z)};
  if (part == 3) { sm.encT[0 * kLd + r] = p[0]; sm.encT[1 * kLd + r] = p[1]; sm.encT[2 * kLd + r] = p[2]; }
  for (int o = part; o < L; o += 4) {
    const float f = pow2i(o);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float sn, cs;
      sincosf(p[c] * f, &sn, &cs);
      sm.encT[(3 + 6 * o + c) * kLd + r] = sn;
      sm.encT[(6 + 6 * o + c) * kLd + r] = cs;
    }
  }
  for (int k = 3 + 6 * L + part; k < kEncRows; k += 4) sm.encT[k * kLd + r] = 0.f;
}

// training: copy the tile's encoded inputs (64 point + 32 direction channels) to the activation store
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// (tf32 != 0: values go to the store rounded to the nearest tf32, the operand format of the tensor-core GEMMs)
template <int W>
__device__ __forceinline__ void save_inputs(const Fp32Smem<W>& sm, float* save, long long R, int tid, int tf32 = 0) {
  for (int i = tid; i < (kEncRows + kDirRows) * (kTileRows / 4); i += kComputeThreads) {
    const int ch = i >> 4, r4 = (i & 15) * 4;
    const float* src = ch < kEncRows ? sm.encT + ch * kLd + r4 : sm.dirT + (ch - kEncRows) * kLd + r4;
    float4 v = *reinterpret_cast<const float4*>(src);
    if (tf32) v = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
    *reinterpret_cast<float4*>(save + (long long)ch * R + r4) = v;
  }
}

// value k of the encoding of a 3-vector (k < 3+6L), 0 beyond
__device__ __forceinline__ float posenc_value(float x, float y, float z, int k, int L) {
  if (k < 3) return k == 0 ? x : (k == 1 ? y : z);
  if (k >= 3 + 6 * L) return 0.f;
  const int o = (k - 3) / 6, j = (k - 3) % 6, c = j % 3;
  const float a = (c == 0 ? x : (c == 1 ? y : z)) * pow2i(o);
  return j < 3 ? sinf(a) : cosf(a);
}

// ------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------
template <int W, int FE>
__global__ void __launch_bounds__(kFp32Threads, 1) snerf_fp32_kernel(const RenderParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Fp32Smem<W>& sm = *reinterpret_cast<Fp32Smem<W>*>(smem_raw);  // keeps the shared address space (LDS/STS)
  const int tid = threadIdx.x;

  // ---- one-time setup: layer tables, barriers
  const unsigned char* img[4] = {p.img_coarse, p.img_fine ? p.img_fine : p.img_coarse, p.img_alpha_coarse,
                                 p.img_alpha_fine};
  {
    constexpr int nint = kFp32MaxLayers * (int)sizeof(Fp32Layer) / 4;
    for (int net = 0; net < 4; ++net) {
      if (!img[net]) { if (tid == 0) sm.n_layers[net] = 0; continue; }
      const Fp32Header* h = reinterpret_cast<const Fp32Header*>(img[net]);
      const int* src = reinterpret_cast<const int*>(h->layers);
      int* dst = reinterpret_cast<int*>(sm.layers[net]);
      for (int i = tid; i < nint; i += kFp32Threads) dst[i] = src[i];
      if (tid == 0) sm.n_layers[net] = h->n_layers;
    }
    if (tid == 0) ring_init(ring_of(sm));
  }
  __syncthreads();

  int stage = 0;
  uint32_t phase = 0;

  if (tid >= kComputeThreads) {
    // =============================== producer warp ===============================
    if (tid == kComputeThreads && !(FE == FE_RAYS && p.stage != 0)) {
      if (FE == FE_RAYS) {
        const int TC = tiles_of(p.Nc), TF = p.Nf > 0 ? tiles_of(p.Nc + p.Nf) : 0;
        for (long long ray = blockIdx.x; ray < p.n_rays; ray += gridDim.x) {
          for (int t = 0; t < TC; ++t) {
            if (img[2]) stream_net<W>(sm, 2, img[2], stage, phase);
            stream_net<W>(sm, 0, img[0], stage, phase);
          }
          for (int t = 0; t < TF; ++t) {
            if (img[3]) stream_net<W>(sm, 3, img[3], stage, phase);
            stream_net<W>(sm, 1, img[1], stage, phase);
          }
        }
      } else {
        const long long M = (FE == FE_QUERY) ? p.n_rays * p.S : p.n_rows;
        const long long tiles = (M + kTileRows - 1) / kTileRows;
        for (long long t = blockIdx.x; t < tiles; t += gridDim.x) stream_net<W>(sm, 0, img[0], stage, phase);
      }
    }
    return;
  }

  // ================================ compute warps =================================
  const int warp = tid >> 5, lane = tid & 31;

  if (FE == FE_RAYS) {
    const int Nc = p.Nc, Nf = p.Nf, S = Nc + Nf;
    const int TC = tiles_of(Nc), TF = Nf > 0 ? tiles_of(S) : 0;
    for (long long ray_i = blockIdx.x; ray_i < p.n_rays; ray_i += gridDim.x) {
      const Ray ray = make_ray(p, ray_i);
      // ---- coarse depths (render.py:330-352) and the per-ray direction encoding
      if (tid < Nc) {
        float z = coarse_depth(ray.near, ray.far, p.t_vals[tid], p.lindisp);
        if (p.t_rand) {
          const float zm1 = tid > 0 ? coarse_depth(ray.near, ray.far, p.t_vals[tid - 1], p.lindisp) : z;
          const float zp1 = tid < Nc - 1 ? coarse_depth(ray.near, ray.far, p.t_vals[tid + 1], p.lindisp) : z;
          z = jitter_depth(zm1, z, zp1, tid == 0, tid == Nc - 1, p.t_rand[ray_i * Nc + tid]);
        }
        sm.zc[tid] = z;
        if (p.out.z_vals_map) p.out.z_vals_map[ray_i * Nc + tid] = z;
      }
      if (tid >= 64 && tid < 64 + kDirRows) sm.direnc[tid - 64] = posenc_value(ray.vx, ray.vy, ray.vz, tid - 64, p.has_vd ? p.Lv : 0);
      named_bar_sync(1, kComputeThreads);
      for (int i = tid; i < kDirRows * kTileRows; i += kComputeThreads) sm.dirT[(i >> 6) * kLd + (i & 63)] = sm.direnc[i >> 6];

      // ---- coarse network.  Training on the tensor cores splits the kernel in stages (p.stage): 1 = stop after
      // the coarse inputs are in the activation store (the MLP runs as layer-batched tcgen05 GEMMs over the store);
      // 2 = resume from the coarse raw those GEMMs produced: composite, resample, fine inputs.
      if (p.stage == 2) {
        for (int i = tid; i < Nc * 4; i += kComputeThreads) sm.raw[i] = p.out.raw_coarse[ray_i * Nc * 4 + i];
      }
      for (int t = 0; t < TC && p.stage != 2; ++t) {
        encode_ray_tile<W>(sm, ray, sm.zc, Nc, t, p.L, tid);
        named_bar_sync(1, kComputeThreads);
        float* save = nullptr;
        if (p.save_c) {
          save = p.save_c + (ray_i * TC + t) * kTileRows;
          save_inputs<W>(sm, save, p.Rc, tid, p.round_tf32);
        }
        if (p.stage == 1) { named_bar_sync(1, kComputeThreads); continue; }
        // (NeRF_RGB: the frozen sigma network first -- it fills all four raw columns -- then the rgb network,
        //  which has no alpha head and overwrites r,g,b only; run_nerf_helpers.py:189-206)
        if (img[2]) mlp_tile<W>(sm, 2, img[2], t * kTileRows, stage, phase, tid);
        mlp_tile<W>(sm, 0, img[0], t * kTileRows, stage, phase, tid, save, p.Rc);
      }
      if (p.stage == 1) continue;
      if (p.stage == 2) named_bar_sync(1, kComputeThreads);
      // ---- composite + hierarchical resampling (one warp; tiny next to the MLP)
      if (warp == 0) {
        const RayCarry c = composite_segment(reinterpret_cast<const float4*>(sm.raw), sm.zc, Nc, 0, Nc, ray.dnorm,
                                             p.noise0 ? p.noise0 + ray_i * Nc : nullptr, sm.wts,
                                             p.out.weights ? p.out.weights + ray_i * Nc : nullptr, carry_init(), lane);
        if (lane == 0) {
          const float wb = p.white_bkgd ? (1.f - c.acc) : 0.f;
          float* rgb = Nf > 0 ? p.out.rgb0 : p.out.rgb_map;
          float* disp = Nf > 0 ? p.out.disp0 : p.out.disp_map;
          float* acc = Nf > 0 ? p.out.acc0 : p.out.acc_map;
          float* depth = Nf > 0 ? p.out.depth0 : p.out.depth_map;
          if (rgb) { rgb[ray_i * 3 + 0] = c.r + wb; rgb[ray_i * 3 + 1] = c.g + wb; rgb[ray_i * 3 + 2] = c.b + wb; }
          if (disp) disp[ray_i] = disparity(c.depth, c.acc);
          if (acc) acc[ray_i] = c.acc;
          if (depth) depth[ray_i] = c.depth;
        }
        float* rawc = Nf > 0 ? p.out.raw_coarse : (p.out.raw ? p.out.raw : p.out.raw_coarse);
        if (rawc) for (int i = lane; i < Nc * 4; i += 32) rawc[ray_i * Nc * 4 + i] = sm.raw[i];
        if (Nf > 0) {
          const int B = Nc - 1;
          for (int i = lane; i < B; i += 32) sm.bins[i] = __fmul_rn(0.5f, __fadd_rn(sm.zc[i + 1], sm.zc[i]));
          __syncwarp();
          build_cdf(sm.wts + 1, B, sm.cdf, lane);
          __syncwarp();
          for (int j = lane; j < Nf; j += 32) {
            const float u = p.u_rand ? p.u_rand[ray_i * Nf + j] : p.u_vals[j];
            int ind;
            const float zs = invert_cdf_one(sm.bins, sm.cdf, B, u, &ind);
            sm.zs[j] = zs;
            if (p.out.z_samples) p.out.z_samples[ray_i * Nf + j] = zs;
          }
          __syncwarp();
          const float sd = warp_std(sm.zs, Nf, lane);
          if (lane == 0 && p.out.z_std) p.out.z_std[ray_i] = sd;
          if (p.u_rand) warp_sort(sm.zs, Nf, lane);
          __syncwarp();
          merge_sorted(sm.zc, Nc, sm.zs, Nf, sm.zf, lane);
          __syncwarp();
          if (p.out.z_all) for (int i = lane; i < S; i += 32) p.out.z_all[ray_i * S + i] = sm.zf[i];
        }
      }
      named_bar_sync(1, kComputeThreads);
      // ---- fine network on the sorted union
      if (Nf > 0) {
        for (int t = 0; t < TF; ++t) {
          encode_ray_tile<W>(sm, ray, sm.zf, S, t, p.L, tid);
          named_bar_sync(1, kComputeThreads);
          float* save = nullptr;
          if (p.save_f) {
            save = p.save_f + (ray_i * TF + t) * kTileRows;
            save_inputs<W>(sm, save, p.Rf, tid, p.round_tf32);
          }
          if (p.stage == 2) { named_bar_sync(1, kComputeThreads); continue; }
          if (img[3]) mlp_tile<W>(sm, 3, img[3], t * kTileRows, stage, phase, tid);
          mlp_tile<W>(sm, 1, img[1], t * kTileRows, stage, phase, tid, save, p.Rf);
        }
        if (warp == 0 && p.stage == 0) {
          const RayCarry c = composite_segment(reinterpret_cast<const float4*>(sm.raw), sm.zf, S, 0, S, ray.dnorm,
                                               p.noise1 ? p.noise1 + ray_i * S : nullptr, nullptr,
                                               p.out.weights_fine ? p.out.weights_fine + ray_i * S : nullptr,
                                               carry_init(), lane);
          if (lane == 0) {
            const float wb = p.white_bkgd ? (1.f - c.acc) : 0.f;
            if (p.out.rgb_map) {
              p.out.rgb_map[ray_i * 3 + 0] = c.r + wb; p.out.rgb_map[ray_i * 3 + 1] = c.g + wb;
              p.out.rgb_map[ray_i * 3 + 2] = c.b + wb;
            }
            if (p.out.disp_map) p.out.disp_map[ray_i] = disparity(c.depth, c.acc);
            if (p.out.acc_map) p.out.acc_map[ray_i] = c.acc;
            if (p.out.depth_map) p.out.depth_map[ray_i] = c.depth;
          }
          if (p.out.raw) for (int i = lane; i < S * 4; i += 32) p.out.raw[ray_i * S * 4 + i] = sm.raw[i];
        }
        named_bar_sync(1, kComputeThreads);
      }
    }
  } else {
    // ---- stage front-ends: flat rows -> raw[rows, 4]
    const long long M = (FE == FE_QUERY) ? p.n_rays * p.S : p.n_rows;
    const long long tiles = (M + kTileRows - 1) / kTileRows;
    const int r = tid & 63, part = tid >> 6;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
      const long long m = min(t * kTileRows + r, M - 1);
      if (FE == FE_QUERY) {
        const float* q = p.pts + m * 3;
        const float px = q[0], py = q[1], pz = q[2];
        for (int k = part; k < kEncRows; k += 4) sm.encT[k * kLd + r] = posenc_value(px, py, pz, k, p.L);
        float vx = 0.f, vy = 0.f, vz = 0.f;
        if (p.viewdirs) { const float* v = p.viewdirs + (m / p.S) * 3; vx = v[0]; vy = v[1]; vz = v[2]; }
        for (int k = part; k < kDirRows; k += 4) sm.dirT[k * kLd + r] = p.viewdirs ? posenc_value(vx, vy, vz, k, p.Lv) : 0.f;
      } else {
        const float* xr = p.x + m * p.x_stride;
        for (int k = part; k < kEncRows; k += 4) sm.encT[k * kLd + r] = k < p.in_ch ? xr[k] : 0.f;
        for (int k = part; k < kDirRows; k += 4) sm.dirT[k * kLd + r] = k < p.in_ch_views ? xr[p.in_ch + k] : 0.f;
      }
      named_bar_sync(1, kComputeThreads);
      mlp_tile<W>(sm, 0, img[0], 0, stage, phase, tid);
      const long long m0 = t * kTileRows;
      for (int i = tid; i < kTileRows * 4; i += kComputeThreads)
        if (m0 + (i >> 2) < M) p.out_raw[m0 * 4 + i] = sm.raw[i];
      named_bar_sync(1, kComputeThreads);
    }
  }
}

// ------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------
template <int W, int FE>
static int launch_one(const RenderParams& p, long long units, cudaStream_t stream) {
  const size_t smem = sizeof(Fp32Smem<W>);
  auto kern = snerf_fp32_kernel<W, FE>;
  if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                 "cudaFuncSetAttribute(fp32 kernel smem)"))
    return SNERF_ERR_CUDA;
  long long grid = units < (long long)sm_count() ? units : (long long)sm_count();
  if (grid < 1) return SNERF_OK;
  kern<<<(unsigned)grid, kFp32Threads, smem, stream>>>(p);
  return check_cuda(cudaGetLastError(), "launch snerf_fp32_kernel");
}

template <int W>
static int launch_w(int fe, const RenderParams& p, cudaStream_t stream) {
  switch (fe) {
    case FE_RAYS: return launch_one<W, FE_RAYS>(p, p.n_rays, stream);
    case FE_QUERY: return launch_one<W, FE_QUERY>(p, (p.n_rays * p.S + kTileRows - 1) / kTileRows, stream);
    case FE_ROWS: return launch_one<W, FE_ROWS>(p, (p.n_rows + kTileRows - 1) / kTileRows, stream);
  }
  return SNERF_ERR_BAD_ARG;
}

int launch_fp32(int frontend, int W, const RenderParams& p, cudaStream_t stream) {
  switch (W) {
    case 64: return launch_w<64>(frontend, p, stream);
    case 128: return launch_w<128>(frontend, p, stream);
    case 256: return launch_w<256>(frontend, p, stream);
  }
  set_error("fp32 mode supports W in {64,128,256}, got %d", W);
  return SNERF_ERR_UNSUPPORTED;
}

}  // namespace snerf
