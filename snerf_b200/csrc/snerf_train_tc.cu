// snerf_train_tc.cu -- training step on the tensor cores with 16-bit operand stores (SNERF_MODE_BF16 / FP16 + save_for_backward).
//
// The reference trains by torch autograd over its eager ops (train.py:110-221; render.py:281-409 under perturb=1 /
// raw_noise_std=1, z_samples detached at render.py:381).  Here one training step of render_rays is FOUR launches:
//
//   1. forward   snerf_bf16_render_kernel<..., kSave = true> (snerf_tc_kernel.cuh): the inference kernel; its epilogues
//                additionally write every layer's operand-precision output to the activation store
//                  act[slot][rows][256] (16-bit): slot 0 = [enc 64 | dir 32 | 0 | views 128], 1..8 = h0..h7, 9 = feature
//                (row = ray * X + sample, X = Nc / Nc + Nf; padding ray of an odd batch included; stored as 4 KiB
//                blocks of 32 rows x 64 channels in UMMA canonical order, snerf_packed.h) and one relu' bit per activation;
//   2. composite_bwd_kernel (snerf_train.cu): d(outputs) -> d_raw[rows][4] fp32 (raw2outputs differentiated by hand);
//   3. dx_chain_tc_kernel (this file): per 128-row tile the whole chain
//                  dvp = (rgb_w^T d_rgb) * relu'(views)                         CUDA cores (front-end warpgroup)
//                  dfeature = dvp . Wviews[:, :256]                             tcgen05, A from shared memory
//                  dz7 = (dfeature . Wfeature + dsigma alpha_w) * relu'(h7)     tcgen05, A from TMEM
//                  dz_{l-1} = (dz_l . W_l[:, hidden]) * relu'(h_{l-1}), l = 7..1
//                stays in TMEM / shared memory exactly like the forward (persistent CTA per SM, weights streamed as
//                pre-swizzled 16 KiB chunks through a bulk-copy ring, multicast across a 2-CTA cluster); every dz goes to
//                the gradient store dz[slot][rows][256] bf16 (slot 0 = [dvp 128 | d_raw 4 | 0], 1..8 = dz0..dz7,
//                9 = dfeature), the relu' masks are read back from the activation store (128 bytes per thread and step);
//   4. dw_tc_kernel (this file): EVERY parameter gradient as one grouped, HBM-streaming GEMM
//                  dW[n][k] += sum_r dz[r][n] * x[r][k]
//                Both operands are [row][channel] = "MN-major" UMMA operands (the reduction runs over rows): the 4 KiB
//                blocks of the two stores are bulk-copied as they lie (they ARE the no-swizzle canonical layout) and
//                multiplied in place, M = 256 x N <= 256 fp32 accumulators (all 512 TMEM columns), fp32 reductions into
//                the nn.Linear-shaped gradients.  The work
//                (problem, 64-row block) is cut into equal-byte contiguous ranges, one per SM, so every SM streams
//                the same number of HBM bytes and runs at most a few accumulator flushes.  Bias gradients (column sums
//                of dz) are summed from the shared-memory tiles by the otherwise idle epilogue warps; the narrow heads
//                (alpha_linear, rgb_linear) are rows of the same GEMM (A = the d_raw columns of slot 0).
//
// Arithmetic: 16-bit operands (activations in the forward's operand type, gradients bf16), fp32 accumulation, fp32
// gradients.  HBM bytes per row and step: 5 KB activations + 5 KB gradients, each written once and read about twice.
#include <stdlib.h>
#include <string.h>

#include "snerf_tc_kernel.cuh"
#include <cstdlib>
#include "snerf_train_tc.h"

namespace snerf {

// (the backward weight image 'SBWB' is packed in snerf_api.cu, next to the forward images: pack_bwd_tc)

// ------------------------------------------------------------------------------------
// 3. dX chain
// ------------------------------------------------------------------------------------
constexpr int kBwRing = 9;
struct alignas(1024) BwSmem {
  uint8_t a0[2][2][kBfChunkBytes];   // A operand of step 0 of tile n in a0[n & 1]: dvp, two 64-wide k-blocks
  uint8_t ring[kBwRing][kBfChunkBytes];
  float4 draw[2][128];               // d_raw of tile n in draw[n & 1]
  float alpha_w[2][256];             // [network]
  float rgb_w[2][384];
  uint64_t w_full[kBwRing], w_empty[kBwRing];
  uint64_t a0_full[2];               // front-end -> MMA
  uint64_t tile_started;             // MMA -> front-end
  uint64_t acc_ready[2];             // MMA -> epilogue
  uint64_t a_ready[4];               // epilogue -> MMA
  uint32_t tmem_base;
};
static_assert(sizeof(BwSmem) <= 232448, "shared memory budget");

// schedule: tile index space [0, tcp) coarse (tcp = coarse tiles rounded up to the cluster size), then fine; indices
// that fall into the padding or past the end are dummies (they run the protocol on zeros and store nothing)
struct BwTile { int net; long long tile; bool real; };
__device__ __forceinline__ BwTile bw_tile(const BwdTcParams& p, long long u, int tcp) {
  BwTile t;
  if (u < tcp) { t.net = 0; t.tile = u; t.real = u < p.tiles[0]; }
  else { t.net = 1; t.tile = u - tcp; t.real = t.tile < p.tiles[1]; }
  if (!t.real) { t.tile = 0; if (p.tiles[t.net] == 0) t.net ^= 1; }
  return t;
}

enum { BW_LINEAR = 0, BW_ALPHA = 1, BW_MASK = 2 };

// one 32-column chunk of the chain epilogue: v = accumulator columns, m = relu' bits of these 32 columns (bit w = column
// 2 w, bit 16 + w = column 2 w + 1), aw = alpha_w of these columns (BW_ALPHA), ds = dsigma of this row
template <int KIND>
__device__ __forceinline__ void bw_chunk(const uint32_t (&v)[32], uint32_t m, const float* __restrict__ aw, float ds,
                                         uint32_t (&packed)[16]) {
#pragma unroll
  for (int w = 0; w < 16; ++w) {
    float a = __uint_as_float(v[2 * w]), b = __uint_as_float(v[2 * w + 1]);
    if (KIND == BW_ALPHA) {
      a = fmaf(ds, aw[2 * w], a);
      b = fmaf(ds, aw[2 * w + 1], b);
    }
    if (KIND != BW_LINEAR) {
      a = (m & (1u << w)) ? a : 0.f;
      b = (m & (0x10000u << w)) ? b : 0.f;
    }
    packed[w] = cvt_bf16x2(a, b);
  }
}

// Epilogue group e of one chain step (same protocol and column assignment as the forward's `epilogue`): of accumulator
// half h it drains chunk 4h + e, then chunk 4h + 2 + e -- the 32-column half e of column blocks 2h and 2h + 1 -- writes the
// bf16 results into the next A operand (not after the last step), signals a_ready[2h] after the first chunk and
// a_ready[2h + 1] after the second (both groups arrive on both), then stores to the gradient store.
// `mask[cb]` = relu' bits of the 32-column half e of column block cb.
template <int KIND, bool kLast>
__device__ __forceinline__ void bw_epilogue(uint64_t* acc_ready, uint64_t* a_ready, uint32_t acc_addr, uint32_t anext_addr,
                                            uint32_t acc_phase, const uint32_t (&mask)[4], const float* __restrict__ aw,
                                            float ds, int e, uint4* save) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int ja = 4 * h + e, jb = 4 * h + 2 + e;
    mbar_wait(&acc_ready[h], acc_phase);
    tc_fence_after();
    uint32_t va[32], vb[32];
    tmem_ld32(acc_addr + (uint32_t)(ja * 32), va);
    tmem_ld32(acc_addr + (uint32_t)(jb * 32), vb);
    tmem_ld_wait_dep(va);
    uint32_t pa[16], pb[16];
    bw_chunk<KIND>(va, mask[2 * h], aw + ja * 32, ds, pa);
    if (!kLast) {
      tmem_st16(anext_addr + (uint32_t)(ja * 16), pa);
      tmem_st_wait();
    }
    tc_fence_before();
    mbar_arrive(&a_ready[2 * h]);
    tmem_ld_wait_dep(vb);
    bw_chunk<KIND>(vb, mask[2 * h + 1], aw + jb * 32, ds, pb);
    if (!kLast) {
      tmem_st16(anext_addr + (uint32_t)(jb * 16), pb);
      tmem_st_wait();
    }
    tc_fence_before();
    mbar_arrive(&a_ready[2 * h + 1]);
    if (save) {
      save_words(save + (2 * h) * 256 + e * 128, pa);
      save_words(save + (2 * h + 1) * 256 + e * 128, pb);
    }
  }
}

template <int kCluster>
__global__ void __launch_bounds__(kBfThreads, 1) dx_chain_tc_kernel(const BwdTcParams p, const int T, const int tcp) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  BwSmem& sm = *reinterpret_cast<BwSmem*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    if (reinterpret_cast<const Bf16Header*>(p.img[0])->magic != kBwMagic ||
        reinterpret_cast<const Bf16Header*>(p.img[1])->magic != kBwMagic) __trap();
    for (int s = 0; s < kBwRing; ++s) { mbar_init(&sm.w_full[s], 1); mbar_init(&sm.w_empty[s], kCluster); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sm.a0_full[i], kGroup); mbar_init(&sm.acc_ready[i], 1); }
    for (int i = 0; i < 4; ++i) mbar_init(&sm.a_ready[i], 2 * kGroup);   // both epilogue groups contribute half of each k-block
    mbar_init(&sm.tile_started, 1);
    mbar_fence_init();
  }
  for (int i = tid; i < 2 * kBwParamFloats; i += kBfThreads) {
    const int net = i / kBwParamFloats, j = i % kBwParamFloats;
    const float v = __ldg(reinterpret_cast<const float*>(p.img[net] + kBwParamsOffset) + j);
    if (j < 256) sm.alpha_w[net][j] = v; else sm.rgb_w[net][j - 256] = v;
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&sm.tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;
  uint32_t cta_rank = 0;
  if (kCluster > 1) {
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  // tile n of this CTA = schedule index n * gridDim.x + blockIdx.x: the CTAs of a cluster always work on the same network
  // (the coarse range is padded to the cluster size), so they consume the same weight chunks
  auto tile_of = [&](int n) { return bw_tile(p, (long long)n * gridDim.x + blockIdx.x, tcp); };

  if (warp == 0) {
    // ================================ weight producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, nchunk = 0;
      for (int n = 0; n < T; ++n) {
        const unsigned char* im = p.img[tile_of(n).net] + kBwChunksOffset;
        for (int c = 0; c < kBwChunks; ++c) {
          mbar_wait(&sm.w_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&sm.w_full[stage], kBfChunkBytes);
          const unsigned char* src = im + (size_t)c * kBfChunkBytes;
          if (kCluster == 1) bulk_g2s(sm.ring[stage], src, kBfChunkBytes, &sm.w_full[stage]);
          else if (nchunk % kCluster == cta_rank)
            bulk_g2s_multicast(sm.ring[stage], src, kBfChunkBytes, &sm.w_full[stage], (uint16_t)((1 << kCluster) - 1));
          ++nchunk;
          if (++stage == kBwRing) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================== MMA issuer ==================================
    const uint32_t leader = elect_one() ? 1u : 0u;
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
    constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);  // SBO | version | SWIZZLE_128B
    const uint32_t ring_lo0 = ((smem_u32(sm.ring[0]) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t empty0 = smem_u32(&sm.w_empty[0]);
    const uint32_t accr0 = smem_u32(&sm.acc_ready[0]), accr1 = smem_u32(&sm.acc_ready[1]);
    int stage = 0;
    uint32_t phase = 0, aphase = 1;
    const uint32_t acc_h0 = tmem_base + kAccCol, acc_h1 = tmem_base + kAccCol + 128;
#define BW_KB_TS(D_TMEM, A_TMEM, ACCUM)                                                                       \
    do {                                                                                                        \
      mbar_wait(&sm.w_full[stage], phase);                                                                      \
      issue_kblock_ts<kCluster>(leader, (D_TMEM), (A_TMEM), ring_lo0 + (uint32_t)stage * (kBfChunkBytes >> 4), kDescHi, \
                                idesc, (ACCUM), empty0 + (uint32_t)stage * 8);                                  \
      if (++stage == kBwRing) { stage = 0; phase ^= 1; }                                                        \
    } while (0)
#define BW_KB_SS(D_TMEM, A_LO, ACCUM)                                                                         \
    do {                                                                                                        \
      mbar_wait(&sm.w_full[stage], phase);                                                                      \
      issue_kblock_ss<kCluster>(leader, (D_TMEM), (A_LO), ring_lo0 + (uint32_t)stage * (kBfChunkBytes >> 4), kDescHi,   \
                                idesc, (ACCUM), empty0 + (uint32_t)stage * 8);                                  \
      if (++stage == kBwRing) { stage = 0; phase ^= 1; }                                                        \
    } while (0)
    for (int n = 0; n < T; ++n) {
      mbar_wait(&sm.a0_full[n & 1], (n >> 1) & 1);
      const uint32_t a0_lo = ((smem_u32(sm.a0[n & 1][0]) & 0x3FFFFu) >> 4) | (1u << 16);
      for (int step = 0; step < kBwSteps; ++step) {
        const uint32_t a_tmem = tmem_base + (((step - 1) & 1) ? kAbufCol1 : kAbufCol0);
        // a_ready[kb] = k-block kb of A is in TMEM; a_ready[0] / [2] also mean accumulator half 0 / 1 has been drained
        mbar_wait(&sm.a_ready[0], aphase);
        if (step == 0) mbar_wait(&sm.a_ready[1], aphase);
        tc_fence_after();
        if (step == 0) {
          BW_KB_SS(acc_h0, a0_lo, 0u);
          BW_KB_SS(acc_h0, a0_lo + (kBfChunkBytes >> 4), 1u);
          commit_if(leader, smem_u32(&sm.tile_started));
        } else {
          BW_KB_TS(acc_h0, a_tmem, 0u);
          mbar_wait(&sm.a_ready[1], aphase);
          tc_fence_after();
          BW_KB_TS(acc_h0, a_tmem + 32, 1u);
          mbar_wait(&sm.a_ready[2], aphase);
          tc_fence_after();
          BW_KB_TS(acc_h0, a_tmem + 64, 1u);
          mbar_wait(&sm.a_ready[3], aphase);
          tc_fence_after();
          BW_KB_TS(acc_h0, a_tmem + 96, 1u);
        }
        commit_if(leader, accr0);
        if (step == 0) {
          mbar_wait(&sm.a_ready[2], aphase);
          mbar_wait(&sm.a_ready[3], aphase);
          tc_fence_after();
          BW_KB_SS(acc_h1, a0_lo, 0u);
          BW_KB_SS(acc_h1, a0_lo + (kBfChunkBytes >> 4), 1u);
        } else {
          BW_KB_TS(acc_h1, a_tmem, 0u);
          BW_KB_TS(acc_h1, a_tmem + 32, 1u);
          BW_KB_TS(acc_h1, a_tmem + 64, 1u);
          BW_KB_TS(acc_h1, a_tmem + 96, 1u);
        }
        commit_if(leader, accr1);
        aphase ^= 1;
      }
    }
#undef BW_KB_TS
#undef BW_KB_SS
  } else if (warp < 10) {
    // ============================= epilogue warpgroups (2) ============================
    const int e = (warp - 2) >> 2;
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const uint32_t acc_addr = tmem_base + lane_base + kAccCol;
    uint32_t acc_phase = 0;
    for (int n = 0; n < T; ++n) {
      const BwTile t = tile_of(n);
      const long long rows = p.rows[t.net];
      const long long rb = t.tile * 4 + wq;            // 32-row block of this warp
      unsigned char* dz = p.dz[t.net];
      const unsigned long long* bits = p.bits[t.net];
      const float* aw = sm.alpha_w[t.net];
      // relu' bits of the 32-column half e of every column block, fetched one step ahead: step k >= 1 is masked by
      // h_{8-k} (mask slot 8 - k)
      auto half_bits = [&](int slot, int cb) {
        return __ldg(reinterpret_cast<const uint32_t*>(bits + tc_mask_index(rows, slot, rb, cb, lane)) + e);
      };
      uint32_t mnext[4];
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) mnext[cb] = half_bits(7, cb);
      for (int step = 0; step < kBwSteps; ++step) {
        const uint32_t anext = tmem_base + lane_base + ((step & 1) ? kAbufCol1 : kAbufCol0);
        // step 0 -> dfeature (slot 9); step k >= 1 -> dz_{8-k} (slot 9 - k)
        uint4* save = t.real ? reinterpret_cast<uint4*>(dz + tc_block_offset(rows, 9 - step, rb, 0)) + lane : nullptr;
        const uint32_t mask[4] = {mnext[0], mnext[1], mnext[2], mnext[3]};
        if (step >= 1 && step + 1 < kBwSteps) {
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) mnext[cb] = half_bits(7 - step, cb);
        }
        if (step == 0) {
          bw_epilogue<BW_LINEAR, false>(sm.acc_ready, sm.a_ready, acc_addr, anext, acc_phase, mask, aw, 0.f, e, save);
        } else if (step == 1) {
          // d_raw of this tile was staged by the front-end before it released the tile's first operand (step 0 ran since)
          const float ds = sm.draw[n & 1][row].w;
          bw_epilogue<BW_ALPHA, false>(sm.acc_ready, sm.a_ready, acc_addr, anext, acc_phase, mask, aw, ds, e, save);
        } else if (step < kBwSteps - 1) {
          bw_epilogue<BW_MASK, false>(sm.acc_ready, sm.a_ready, acc_addr, anext, acc_phase, mask, aw, 0.f, e, save);
        } else {
          bw_epilogue<BW_MASK, true>(sm.acc_ready, sm.a_ready, acc_addr, anext, acc_phase, mask, aw, 0.f, e, save);
        }
        acc_phase ^= 1;
      }
    }
  } else {
    // ============================== front-end warpgroup =============================
    // row wt of tile n: d_raw -> dvp = (rgb_w^T d_rgb) * relu'(views) as the swizzled A operand of step 0 and as slot 0 of
    // the gradient store ([dvp 128 | d_r d_g d_b d_sigma | 0 ...])
    const int wt = tid - 320;
    for (int n = 0; n < T; ++n) {
      const BwTile t = tile_of(n);
      const long long rows = p.rows[t.net];
      const long long r = t.tile * 128 + wt;
      const long long rb = r >> 5;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t.real && r < p.valid_rows[t.net]) g = __ldg(p.draw[t.net] + r);
      // relu' bits of the 128 views channels (mask slot 8, column blocks 0 and 1)
      const unsigned long long vb0 = __ldg(p.bits[t.net] + tc_mask_index(rows, 8, rb, 0, lane));
      const unsigned long long vb1 = __ldg(p.bits[t.net] + tc_mask_index(rows, 8, rb, 1, lane));
      if (n >= 2) mbar_wait_relaxed(&sm.tile_started, (n - 1) & 1);  // tile n-1 has started => tile n-2 is done with a0 / draw[n & 1]
      sm.draw[n & 1][wt] = g;
      const float* rw = sm.rgb_w[t.net];
      uint4* save = t.real ? reinterpret_cast<uint4*>(p.dz[t.net] + tc_block_offset(rows, 0, rb, 0)) + lane : nullptr;
#pragma unroll
      for (int c8 = 0; c8 < 16; ++c8) {  // 8 channels per 16-byte chunk; column block c8 / 8, chunk c8 % 8
        const unsigned long long vb = c8 < 8 ? vb0 : vb1;
        const uint32_t half = (uint32_t)(((c8 & 7) < 4) ? vb : (vb >> 32));
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int k = c8 * 8 + 2 * i;
          const int w = (c8 & 3) * 4 + i;      // word index inside the 32-column half
          float a = fmaf(g.z, rw[256 + k], fmaf(g.y, rw[128 + k], g.x * rw[k]));
          float b = fmaf(g.z, rw[256 + k + 1], fmaf(g.y, rw[128 + k + 1], g.x * rw[k + 1]));
          a = (half & (1u << w)) ? a : 0.f;
          b = (half & (0x10000u << w)) ? b : 0.f;
          o[i] = cvt_bf16x2(a, b);
        }
        const uint4 v = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(sm.a0[n & 1][c8 >> 3] + sw128_offset(wt, c8 & 7)) = v;
        if (save) save[(c8 >> 3) * 256 + (c8 & 7) * 32] = v;
      }
      if (save) {  // column block 2: [d_r d_g d_b d_sigma | 0 ...], column block 3: zeros
        save[2 * 256] = make_uint4(cvt_bf16x2(g.x, g.y), cvt_bf16x2(g.z, g.w), 0u, 0u);
#pragma unroll
        for (int i = 1; i < 16; ++i) save[2 * 256 + i * 32] = make_uint4(0u, 0u, 0u, 0u);
      }
      fence_proxy_async();
      mbar_arrive(&sm.a0_full[n & 1]);
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1)
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

template <int kCluster>
static int launch_dx_chain_t(const BwdTcParams& p, long long grid, int T, int tcp, cudaStream_t stream) {
  const size_t smem = sizeof(BwSmem);
  auto kern = dx_chain_tc_kernel<kCluster>;
  if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                 "cudaFuncSetAttribute(dx_chain smem)"))
    return SNERF_ERR_CUDA;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kBfThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return check_cuda(cudaLaunchKernelEx(&cfg, kern, p, T, tcp), "launch dx_chain_tc_kernel");
}

int launch_dx_chain_tc(const BwdTcParams& p, cudaStream_t stream) {
  const long long total = (long long)p.tiles[0] + p.tiles[1];
  if (total == 0) return SNERF_OK;
  long long grid = total < (long long)sm_count() ? total : (long long)sm_count();
  static const int cluster_env = [] { const char* e = getenv("SNERF_B200_CLUSTER"); return e ? atoi(e) : 2; }();
  const bool use_cluster = cluster_env == 2 && grid >= 2;
  if (use_cluster) grid &= ~1ll;
  const int cl = use_cluster ? 2 : 1;
  const int tcp = (p.tiles[0] + cl - 1) / cl * cl;
  const long long sched = (long long)tcp + p.tiles[1];
  const int T = (int)((sched + grid - 1) / grid);
  return use_cluster ? launch_dx_chain_t<2>(p, grid, T, tcp, stream) : launch_dx_chain_t<1>(p, grid, T, tcp, stream);
}

// ------------------------------------------------------------------------------------
// 4. weight gradients
// ------------------------------------------------------------------------------------
// One pipeline stage = 64 rows (two 32-row blocks) of both operands, exactly as the stores hold them: per 32-row block
// the operand's (up to four) column blocks are adjacent 4 KiB blocks in global memory = one bulk copy.  Shared-memory image
// of an operand: [rb (2)][cb (4)][chunk (8)][row (32)][16 bytes] -- the no-swizzle MN-major canonical layout: 8 channels
// contiguous, 16 bytes between rows of a core matrix, 128 bytes between 8-row groups (LBO), 512 bytes between chunks (SBO).
constexpr int kDwStages = 3;
constexpr int kDwRbBytes = 4 * kTcBlockBytes;        // one 32-row block of up to 256 channels: 16 KiB
constexpr int kDwOperandBytes = 2 * kDwRbBytes;      // 64 rows
constexpr int kDwThreads = 192;

struct alignas(1024) DwSmem {
  uint8_t a[kDwStages][kDwOperandBytes];
  uint8_t b[kDwStages][kDwOperandBytes];
  uint64_t full[kDwStages];
  uint64_t empty[kDwStages];
  uint64_t done;      // MMA -> epilogue: the accumulators of the current problem are complete
  uint64_t flushed;   // epilogue -> MMA: they have been read out, the next problem may overwrite them
  uint32_t tmem_base;
};

__device__ __forceinline__ uint64_t umma_desc_mn_noswizzle(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(128 >> 4) << 16;   // LBO: between 8-row (K) groups
  d |= (uint64_t)(512 >> 4) << 32;   // SBO: between 8-channel (MN) chunks
  d |= (uint64_t)1 << 46;            // descriptor version; layout type 0 = no swizzle
  return d;
}

// the range of 64-row blocks of problem `pi` that belongs to CTA `cta`: the intersection of the CTA's cut of the
// linearised (problem, block) space with the problem's own blocks
__device__ __forceinline__ void dw_range(const DwTcTable& tab, int pi, int cta, int& kb0, int& kb1) {
  const DwTcProblem& P = tab.p[pi];
  const long long nb = P.R / 64;
  long long b0 = tab.cut[cta] - P.first_block, b1 = tab.cut[cta + 1] - P.first_block;
  b0 = b0 < 0 ? 0 : (b0 > nb ? nb : b0);
  b1 = b1 < 0 ? 0 : (b1 > nb ? nb : b1);
  kb0 = (int)b0; kb1 = (int)b1;
}

__device__ long long g_dw_timing[kMaxDwCtas][4];   // debug (tab.timing): total cycles, flush cycles, flushes, blocks

__global__ void __launch_bounds__(kDwThreads, 1) dw_tc_kernel(const __grid_constant__ DwTcTable tab) {
  extern __shared__ __align__(1024) unsigned char smem_dw[];
  DwSmem& sm = *reinterpret_cast<DwSmem*>(smem_dw);
  if ((smem_u32(smem_dw) & 1023u) != 0) __trap();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kDwStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1 + 4); }
    mbar_init(&sm.done, 1);
    mbar_init(&sm.flushed, 4);
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
  const int cta = blockIdx.x;
  const long long t_start = clock64();

  if (warp == 0) {
    // ---- producer: two bulk copies per operand and stage
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pi = 0; pi < tab.n; ++pi) {
        const DwTcProblem& P = tab.p[pi];
        int kb0, kb1;
        dw_range(tab, pi, cta, kb0, kb1);
        const uint32_t abytes = (uint32_t)(P.M / 64) * kTcBlockBytes, bbytes = (uint32_t)((P.Nmma + 63) / 64) * kTcBlockBytes;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&sm.empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&sm.full[stage], 2 * (abytes + bbytes));
#pragma unroll
          for (int rb = 0; rb < 2; ++rb) {
            const long long blk = (2ll * kb + rb) * 4 * kTcBlockBytes;   // 32-row block `2 kb + rb` of the slot
            bulk_g2s(sm.a[stage] + rb * kDwRbBytes, P.A + blk, abytes, &sm.full[stage]);
            bulk_g2s(sm.b[stage] + rb * kDwRbBytes, P.B + blk, bbytes, &sm.full[stage]);
          }
          if (++stage == kDwStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: D[m][n] (+)= sum_r A[r][m] * B[r][n], both operands MN-major (descriptor bits 15 / 16)
    int stage = 0;
    uint32_t phase = 0, fphase = 0;
    bool dirty = false;
    for (int pi = 0; pi < tab.n; ++pi) {
      const DwTcProblem& P = tab.p[pi];
      int kb0, kb1;
      dw_range(tab, pi, cta, kb0, kb1);
      if (kb1 <= kb0) continue;
      if (dirty) {  // the previous problem's accumulators must have been read out
        mbar_wait(&sm.flushed, fphase);
        fphase ^= 1;
        tc_fence_after();
      }
      dirty = true;
      const uint32_t idesc = umma_idesc_bf16(128, P.Nmma) | (1u << 15) | (1u << 16);
      const int mh = P.M / 128;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&sm.full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t abase = smem_u32(sm.a[stage]), bbase = smem_u32(sm.b[stage]);
          for (int i = 0; i < mh; ++i) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {  // K = 16 rows per instruction: two 8-row groups of 32-row block ks / 2
              const uint32_t koff = (uint32_t)(ks >> 1) * kDwRbBytes + (uint32_t)(ks & 1) * 256;
              tc_mma_ss(tmem_base + 256 * i, umma_desc_mn_noswizzle(abase + koff + i * 2 * kTcBlockBytes),
                        umma_desc_mn_noswizzle(bbase + koff), idesc, (kb != kb0 || ks != 0) ? 1u : 0u);
            }
          }
          tc_commit(&sm.empty[stage]);
          if (kb == kb1 - 1) tc_commit(&sm.done);
        }
        __syncwarp();
        if (++stage == kDwStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ---- warps 2..5: bias sums from the shared-memory A tiles during the main loop, accumulator flush at the end.
    // Sums: warp w owns the 8-channel chunks {w, w + 4, ...} of the A operand; a lane reads rows lane and 32 + lane of each
    // (16-byte loads, a warp reads 512 contiguous bytes: conflict free) and keeps 8 partial sums per chunk, reduced over
    // the lanes once per problem.
    const int ew = warp - 2;                       // 0..3
    const int lg = warp & 3;                       // TMEM lane group this warp may access
    int stage = 0;
    uint32_t phase = 0, dphase = 0;
    long long t_flush = 0, n_flush = 0, n_blocks = 0;
    for (int pi = 0; pi < tab.n; ++pi) {
      const DwTcProblem& P = tab.p[pi];
      int kb0, kb1;
      dw_range(tab, pi, cta, kb0, kb1);
      if (kb1 <= kb0) continue;
      n_blocks += (long long)(kb1 - kb0) * P.weight;
      const bool sum_on = P.bias != nullptr;
      const int n_chunks = P.M / 8;                // 16 or 32: chunks ew, ew + 4, ... -> 4 or 8 per warp
      float acc[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&sm.full[stage], phase);
        if (sum_on) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int ch = ew + 4 * i;
            if (ch < n_chunks) {
#pragma unroll
              for (int rb = 0; rb < 2; ++rb) {
                const uint4 q = *reinterpret_cast<const uint4*>(sm.a[stage] + rb * kDwRbBytes + ch * 512 + lane * 16);
                const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  acc[i][2 * j] += __uint_as_float(w4[j] << 16);
                  acc[i][2 * j + 1] += __uint_as_float(w4[j] & 0xFFFF0000u);
                }
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[stage]);
        if (++stage == kDwStages) { stage = 0; phase ^= 1; }
      }
      if (sum_on) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int ch = ew + 4 * i;
          if (ch < n_chunks) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float t = warp_sum(acc[i][j]);
              const int m = ch * 8 + j;
              if (lane == 0 && m >= P.m_lo && m < P.m_hi) atomicAdd(P.bias + (m - P.m_lo), t);
            }
          }
        }
      }
      // flush: thread = accumulator row (TMEM lane); 32 columns per load
      mbar_wait(&sm.done, dphase);
      dphase ^= 1;
      tc_fence_after();
      const long long t_f0 = clock64();
      const int mh = P.M / 128;
      for (int i = 0; i < mh; ++i) {
        const int m = i * 128 + lg * 32 + lane;
        const bool row_ok = m >= P.m_lo && m < P.m_hi;
        float* crow = P.C + (long long)(m - P.m_lo) * P.ldc;
        for (int c0 = 0; c0 < P.Nmma; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + 256 * i + c0, v);
          tmem_ld_wait();
          if (!row_ok || !P.C) continue;
          if (P.vec4 && c0 + 32 <= P.N) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(crow + c0 + 4 * q), "f"(__uint_as_float(v[4 * q])),
                           "f"(__uint_as_float(v[4 * q + 1])), "f"(__uint_as_float(v[4 * q + 2])),
                           "f"(__uint_as_float(v[4 * q + 3])) : "memory");
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (c0 + q < P.N) atomicAdd(crow + c0 + q, __uint_as_float(v[q]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.flushed);
      t_flush += clock64() - t_f0; ++n_flush;
    }
    if (tab.timing && warp == 2 && lane == 0) {
      g_dw_timing[cta][0] = clock64() - t_start; g_dw_timing[cta][1] = t_flush; g_dw_timing[cta][2] = n_flush;
      g_dw_timing[cta][3] = n_blocks;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// the 28-problem table of one step (host; pointers are only offset, never dereferenced here)
static void dw_build_table(DwTcTable& tab, const BwdTcParams& p, const unsigned char* const act[2], const SnerfNetGradF32* gc,
                           const SnerfNetGradF32* gf) {
  for (int net = 0; net < 2; ++net) {
    const long long R = p.rows[net];
    if (R == 0) continue;
    const SnerfNetGradF32* g = (net && gf) ? gf : gc;
    // A = slot / first channel of the gradient store, B = slot / first channel of the activation store
    auto add = [&](int slotA, int chA, int M, int m_lo, int m_hi, int slotB, int chB, int N, float* C, int ldc, float* bias) {
      if (!C && !bias) return;
      DwTcProblem& P = tab.p[tab.n++];
      P.A = p.dz[net] + tc_block_offset(R, slotA, 0, chA / 64);
      P.B = act[net] + tc_block_offset(R, slotB, 0, chB / 64);
      P.M = M; P.m_lo = m_lo; P.m_hi = m_hi; P.N = N; P.Nmma = (N + 15) / 16 * 16;
      P.C = C; P.ldc = ldc; P.bias = bias; P.R = R;
      P.vec4 = (ldc % 4 == 0) && (reinterpret_cast<uintptr_t>(C) % 16 == 0);
      P.weight = M / 64 + (P.Nmma + 63) / 64;
    };
    add(1, 0, 256, 0, 256, 0, 0, 63, g->pts_w[0], 63, g->pts_b[0]);
    for (int l = 1; l < 8; ++l)
      add(1 + l, 0, 256, 0, 256, l, 0, 256, g->pts_w[l] ? g->pts_w[l] + (l == 5 ? 63 : 0) : nullptr, l == 5 ? 319 : 256, g->pts_b[l]);
    add(6, 0, 256, 0, 256, 0, 0, 63, g->pts_w[5], 319, nullptr);
    add(9, 0, 256, 0, 256, 8, 0, 256, g->feature_w, 256, g->feature_b);
    add(0, 0, 128, 0, 128, 9, 0, 256, g->views_w, 283, g->views_b);
    add(0, 0, 128, 0, 128, 0, 64, 27, g->views_w ? g->views_w + 256 : nullptr, 283, nullptr);
    add(0, 128, 128, 0, 3, 0, 128, 128, g->rgb_w, 128, g->rgb_b);
    add(0, 128, 128, 3, 4, 8, 0, 256, g->alpha_w, 256, g->alpha_b);
  }
}

// contiguous per-CTA cuts of the linearised (problem, block) space; returns the modelled makespan (cost units)
static int dw_compute_cuts(DwTcTable& tab, int n_cta, double* makespan = nullptr) {
  if (tab.n > kMaxDwTcProblems) { set_error("internal: gradient problem table overflow"); return SNERF_ERR_BAD_ARG; }
  if (tab.n == 0) { tab.n_cta = 0; return SNERF_OK; }
  // ---- contiguous cuts of the linearised (problem, 64-row block) space, one range per CTA, balanced under a cost model
  // fitted to per-CTA cycle counts (tools/dw_balance.py, SNERF_DW_TIMING=1; unit = the time one 64-channel column block of
  // 64 rows takes to stream, ~150 clk):  a 64-row block of problem p costs  (A blocks + B blocks) + 5.9 (+ 1.3 when the
  // reducer warps also sum the A operand for a bias gradient);  every problem a CTA touches costs one accumulator flush:
  // 13 per 64 x 64 block of C written with vector reductions, 38 per block with scalar atomics (unaligned rows).
  // Minimal makespan by bisection over the per-CTA budget with greedy packing.
  long long nb_total = 0;
  for (int i = 0; i < tab.n; ++i) { tab.p[i].first_block = nb_total; nb_total += tab.p[i].R / 64; }
  if (n_cta > kMaxDwCtas) n_cta = kMaxDwCtas;
  if (n_cta < 1) n_cta = 1;
  static const double flush_scale = [] { const char* e = getenv("SNERF_DW_FLUSH_COST"); return e ? atof(e) : 1.0; }();
  static const double block_fixed = [] { const char* e = getenv("SNERF_DW_BLOCK_COST"); return e ? atof(e) : 5.9; }();
  auto block_cost = [&](const DwTcProblem& P) {   // (+0.5: a single-block B operand streams in 4 KiB copies)
    return (double)P.weight + block_fixed + (P.bias ? 1.3 : 0.0) + (P.Nmma <= 64 ? 0.5 : 0.0);
  };
  auto flush_cost = [&](const DwTcProblem& P) {
    if (!P.C) return flush_scale * 2.0;
    const int rows = P.m_hi - P.m_lo;
    const double blocks = (double)((rows + 63) / 64) * ((P.N + 63) / 64);
    return flush_scale * (2.0 + (P.vec4 ? 13.0 : 38.0) * blocks);
  };
  auto pack = [&](double budget, long long* cut) -> bool {     // greedy: fill each CTA up to `budget`
    int pi = 0;
    long long blk = 0;      // next unassigned block of problem pi
    for (int c = 0; c < n_cta; ++c) {
      if (cut) cut[c] = pi < tab.n ? tab.p[pi].first_block + blk : nb_total;
      double left = budget;
      while (pi < tab.n) {
        const DwTcProblem& P = tab.p[pi];
        const long long nb = P.R / 64;
        const double room = left - flush_cost(P);
        const long long take = room <= 0 ? 0 : (long long)(room / block_cost(P));
        if (take <= 0) break;
        const long long got = take < nb - blk ? take : nb - blk;
        left -= flush_cost(P) + (double)got * block_cost(P);
        blk += got;
        if (blk < nb) break;
        ++pi; blk = 0;
      }
    }
    if (cut) cut[n_cta] = nb_total;
    return pi >= tab.n;
  };
  double lo = 0, hi = 0;
  for (int i = 0; i < tab.n; ++i) hi += flush_cost(tab.p[i]) + (double)(tab.p[i].R / 64) * block_cost(tab.p[i]);
  for (int it = 0; it < 40; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (pack(mid, nullptr)) hi = mid; else lo = mid;
  }
  pack(hi, tab.cut);
  tab.n_cta = n_cta;
  if (makespan) *makespan = hi;
  return SNERF_OK;
}

int launch_dw_tc(const BwdTcParams& p, const unsigned char* const act[2], const SnerfNetGradF32* gc, const SnerfNetGradF32* gf,
                 cudaStream_t stream) {
  DwTcTable tab{};
  dw_build_table(tab, p, act, gc, gf);
  if (int e = dw_compute_cuts(tab, sm_count())) return e;
  if (tab.n == 0) return SNERF_OK;
  const int n_cta = tab.n_cta;
  static const int timing_env = [] { const char* e = getenv("SNERF_DW_TIMING"); return e ? atoi(e) : 0; }();
  tab.timing = timing_env;
  const size_t smem = sizeof(DwSmem);
  if (check_cuda(cudaFuncSetAttribute(dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                 "cudaFuncSetAttribute(dw_tc smem)"))
    return SNERF_ERR_CUDA;
  dw_tc_kernel<<<n_cta, kDwThreads, smem, stream>>>(tab);
  return check_cuda(cudaGetLastError(), "launch dw_tc_kernel");
}


// debug / test aid (host only, no GPU): the cuts launch_dw_tc would use for `rows_c` coarse and `rows_f` fine rows on n_cta
// SMs: out_cut[n_cta + 1] block positions, out_first[n problems + 1] first block of each problem (+ total); returns the
// number of problems, or a negative status
int debug_dw_cuts(long long rows_c, long long rows_f, int n_cta, long long* out_cut, long long* out_first, double* out_cost,
                  double* makespan) {
  if (rows_c % 64 || rows_f % 64 || rows_c < 0 || rows_f < 0 || n_cta < 1 || n_cta > kMaxDwCtas) { set_error("bad argument"); return SNERF_ERR_BAD_ARG; }
  BwdTcParams p{};
  p.rows[0] = rows_c; p.rows[1] = rows_f;
  static unsigned char fake_store;                   // addresses are offset, never dereferenced
  p.dz[0] = p.dz[1] = &fake_store;
  const unsigned char* act[2] = {&fake_store, &fake_store};
  alignas(16) static float fake_grad[4];
  SnerfNetGradF32 g{};
  for (int i = 0; i < 8; ++i) { g.pts_w[i] = fake_grad; g.pts_b[i] = fake_grad; }
  g.views_w = g.views_b = g.feature_w = g.feature_b = g.alpha_w = g.alpha_b = g.rgb_w = g.rgb_b = fake_grad;
  DwTcTable tab{};
  dw_build_table(tab, p, act, &g, &g);
  if (int e = dw_compute_cuts(tab, n_cta, makespan)) return e;
  for (int c = 0; c <= tab.n_cta; ++c) out_cut[c] = tab.cut[c];
  for (int i = 0; i < tab.n; ++i) { out_first[i] = tab.p[i].first_block; if (out_cost) out_cost[i] = (double)tab.p[i].weight; }
  out_first[tab.n] = tab.n ? tab.p[tab.n - 1].first_block + tab.p[tab.n - 1].R / 64 : 0;
  return tab.n;
}

// debug: per-CTA cycle counts of the last dw_tc_kernel launch run with SNERF_DW_TIMING=1 ([cta][total, flush, flushes, units])
int debug_dw_timing(long long* out_host, int n_cta) {
  if (n_cta > kMaxDwCtas) n_cta = kMaxDwCtas;
  return check_cuda(cudaMemcpyFromSymbol(out_host, g_dw_timing, sizeof(long long) * 4 * n_cta), "read dw timing");
}

}  // namespace snerf
