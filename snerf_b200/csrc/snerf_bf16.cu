// snerf_bf16.cu -- throughput renderer: render_rays as ONE persistent sm_100a kernel with the MLP on
// tcgen05 tensor cores (bf16 operands, fp32 accumulators in TMEM).
//
// Specialised for the configuration BASELINE.json's metric is quoted on: NeRF D=8, W=256, skip=4,
// 63/27 encoded inputs, view directions, 64 coarse + 128 fine samples (render.py:281-409,
// run_nerf_helpers.py:74-126).
//
// Work unit = a PAIR of rays = 4 MLP tiles of 128 sample-rows:
//     tile 0  coarse   ray0[0:64]  | ray1[0:64]           (coarse network)
//     tile 1  fine     ray0[0:128]                        (fine network, sorted union of 192 depths)
//     tile 2  fine     ray0[128:192] | ray1[0:64]
//     tile 3  fine     ray1[64:192]
// Each tile runs ten tensor-core steps (L0..L7, feature, views); alpha/rgb heads and the direction
// half of the views layer are folded into the epilogues (CUDA cores, fp32).
//
// CTA = 320 threads, 1 CTA / SM, persistent:
//     warp 0      weight producer : bulk async copies (cp.async.bulk / UBLKCP) of pre-swizzled 16 KiB
//                                   B-operand chunks + per-step parameter packets into a 3-deep ring
//     warp 1      MMA issuer      : one thread issues tcgen05.mma (M=128, N=128, K=16), commits to mbarriers
//     warps 2-5   chain 0 \  each chain = 128 threads = 128 TMEM lanes = the 128 rows of its tile:
//     warps 6-9   chain 1 /  sample -> encode -> per-step epilogue (TMEM -> +bias, ReLU -> bf16 -> smem A
//                            operand, 128B swizzle) -> composite / inverse-CDF / merge
// The two chains ping-pong on the tensor core: while chain 0's accumulator (TMEM cols 0-255) is being
// drained by its epilogue, the MMA thread runs chain 1's step into cols 256-511, and vice versa.
// Per-sample activations never leave the SM; HBM sees only the ray batch and the per-ray outputs.
#include "snerf_common.cuh"
#include "snerf_internal.h"
#include "snerf_packed.h"

namespace snerf {

constexpr int kBfThreads = 320;
constexpr int kRing = 3;
constexpr int kChainThreads = 128;

// ------------------------------------------------------------------------------------
// tcgen05 wrappers
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T ; M=128, N from idesc, K=16 (bf16)
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, 128-byte-swizzled operand tile: rows of 64 bf16 (128 B), 8-row atoms 1024 B apart.
// Field layout follows the sm_100 shared-memory matrix descriptor: start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout type [61,64) with 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;            // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset between 8-row atoms
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}
// instruction descriptor, kind::f16: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, K-major both,
// N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// byte offset of the 16-byte chunk `chunk` (8 bf16) of row `row` inside a [128 x 64] bf16 k-block
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}

// ---- packed fp32x2 arithmetic (sm_100: FADD2/FFMA2) and fused ReLU+bf16 conversion ----
__device__ __forceinline__ uint64_t pack2u(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ uint64_t pack2f(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2f(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// {bf16(max(hi,0)), bf16(max(lo,0))} -- ReLU folded into the conversion
__device__ __forceinline__ uint32_t cvt_relu_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// tcgen05.wait::ld that also names the destination registers, so no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.wait::ld.sync.aligned;"
      : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
        "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),
        "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
        "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
      :
      : "memory");
}

// Per-step epilogue of one tile row (thread = TMEM lane = row): accumulator -> +bias -> activation ->
//   EPI_RELU   bf16 A operand of the next layer (steps 0..6)
//   EPI_ALPHA  same, plus sigma = alpha_linear(h) on the fp32 hidden state (step 7)
//   EPI_LINEAR no activation (feature_linear, step 8)
//   EPI_RGB    views layer (N=128): +per-ray direction bias, ReLU, rgb_linear; nothing stored (step 9)
// The TMEM load of chunk j+1 is in flight while chunk j is processed.
enum { EPI_RELU = 0, EPI_ALPHA = 1, EPI_LINEAR = 2, EPI_RGB = 3 };

template <int KIND>
__device__ __forceinline__ void epilogue(uint32_t taddr, const float* __restrict__ bias, const float* __restrict__ aux,
                                         uint8_t* act, int row, float& o0, float& o1, float& o2) {
  constexpr int NCH = (KIND == EPI_RGB) ? 4 : 8;
  const uint32_t row_off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
  const uint32_t r7s = (uint32_t)((row & 7) << 4);
  uint64_t acc0 = 0, acc1 = 0, acc2 = 0;  // packed partial sums of the head dot products
  uint32_t va[32], vb[32];
  tmem_ld32(taddr, va);
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    uint32_t(&v)[32] = (j & 1) ? vb : va;
    tmem_ld_wait_dep(v);
    if (j + 1 < NCH) tmem_ld32(taddr + (uint32_t)((j + 1) * 32), (j & 1) ? va : vb);
#pragma unroll
    for (int q8 = 0; q8 < 4; ++q8) {
      const int col = j * 32 + q8 * 8;
      const float4 b0 = *reinterpret_cast<const float4*>(bias + col);
      const float4 b1 = *reinterpret_cast<const float4*>(bias + col + 4);
      uint64_t s[4];
      s[0] = add2(pack2u(v[q8 * 8 + 0], v[q8 * 8 + 1]), pack2f(b0.x, b0.y));
      s[1] = add2(pack2u(v[q8 * 8 + 2], v[q8 * 8 + 3]), pack2f(b0.z, b0.w));
      s[2] = add2(pack2u(v[q8 * 8 + 4], v[q8 * 8 + 5]), pack2f(b1.x, b1.y));
      s[3] = add2(pack2u(v[q8 * 8 + 6], v[q8 * 8 + 7]), pack2f(b1.z, b1.w));
      float f[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) unpack2f(s[i], f[2 * i], f[2 * i + 1]);
      if (KIND == EPI_ALPHA || KIND == EPI_RGB) {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
      }
      if (KIND == EPI_ALPHA) {
        const float4 w0 = *reinterpret_cast<const float4*>(aux + col);
        const float4 w1 = *reinterpret_cast<const float4*>(aux + col + 4);
        acc0 = fma2(pack2f(f[0], f[1]), pack2f(w0.x, w0.y), acc0);
        acc1 = fma2(pack2f(f[2], f[3]), pack2f(w0.z, w0.w), acc1);
        acc0 = fma2(pack2f(f[4], f[5]), pack2f(w1.x, w1.y), acc0);
        acc1 = fma2(pack2f(f[6], f[7]), pack2f(w1.z, w1.w), acc1);
      }
      if (KIND == EPI_RGB) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float4 w0 = *reinterpret_cast<const float4*>(aux + k * 128 + col);
          const float4 w1 = *reinterpret_cast<const float4*>(aux + k * 128 + col + 4);
          uint64_t& a = k == 0 ? acc0 : (k == 1 ? acc1 : acc2);
          a = fma2(pack2f(f[0], f[1]), pack2f(w0.x, w0.y), a);
          a = fma2(pack2f(f[2], f[3]), pack2f(w0.z, w0.w), a);
          a = fma2(pack2f(f[4], f[5]), pack2f(w1.x, w1.y), a);
          a = fma2(pack2f(f[6], f[7]), pack2f(w1.z, w1.w), a);
        }
      } else {
        uint4 o;
        if (KIND == EPI_RELU) {
          o.x = cvt_relu_bf16x2(f[0], f[1]); o.y = cvt_relu_bf16x2(f[2], f[3]);
          o.z = cvt_relu_bf16x2(f[4], f[5]); o.w = cvt_relu_bf16x2(f[6], f[7]);
        } else {
          o.x = cvt_bf16x2(f[0], f[1]); o.y = cvt_bf16x2(f[2], f[3]);
          o.z = cvt_bf16x2(f[4], f[5]); o.w = cvt_bf16x2(f[6], f[7]);
        }
        const uint32_t chunk = (uint32_t)((j & 1) * 4 + q8);
        *reinterpret_cast<uint4*>(act + (j >> 1) * kBfChunkBytes + row_off + ((chunk << 4) ^ r7s)) = o;
      }
    }
  }
  if (KIND == EPI_ALPHA) {
    float a, b, c, d;
    unpack2f(acc0, a, b); unpack2f(acc1, c, d);
    o0 = (a + b) + (c + d);
  }
  if (KIND == EPI_RGB) {
    float a, b;
    unpack2f(acc0, a, b); o0 = a + b;
    unpack2f(acc1, a, b); o1 = a + b;
    unpack2f(acc2, a, b); o2 = a + b;
  }
}

// ------------------------------------------------------------------------------------
// shared memory
// ------------------------------------------------------------------------------------
struct ChainScratch {  // aliases the chain's enc buffer once the skip layer (step 5) has consumed it
  float4 raw[128];     // (r,g,b,sigma) of the tile's rows
  float wts[2][64];    // coarse weights
  float cdf[2][64];
  float bins[2][64];
  float zs[2][128];    // importance samples
};
struct alignas(1024) BfSmem {
  uint8_t act[2][4 * kBfChunkBytes];  // per chain: hidden activations, 4 k-blocks of [128 x 64] bf16
  uint8_t enc[2][kBfChunkBytes];      // per chain: encoded points, 1 k-block
  uint8_t ring[kRing][kBfChunkBytes]; // weight chunks
  float packet[2][2][kBfPacketFloats];
  float dirbias[2][2][128];
  float direnc[2][2][32];
  float zc[2][2][64];
  float zf[2][2][192];
  float rayrec[2][2][12];
  RayCarry carry[2][2];
  uint64_t w_full[kRing], w_empty[kRing];
  uint64_t a_ready[2], acc_ready[2], pk_full[2][2];
  uint32_t tmem_base;
};
static_assert(sizeof(ChainScratch) <= kBfChunkBytes, "scratch must fit in the enc buffer");
static_assert(sizeof(BfSmem) <= 232448, "shared memory budget");

__device__ __forceinline__ Ray ray_from_rec(const float* r) {
  Ray q;
  q.ox = r[0]; q.oy = r[1]; q.oz = r[2]; q.dx = r[3]; q.dy = r[4]; q.dz = r[5];
  q.near = r[6]; q.far = r[7]; q.vx = r[8]; q.vy = r[9]; q.vz = r[10]; q.dnorm = r[11];
  return q;
}

// tile/row -> (ray in pair, sample index)
__device__ __forceinline__ void row_to_sample(int tile, int row, int& ray, int& s) {
  if (tile == 0) { ray = row >> 6; s = row & 63; }
  else if (tile == 1) { ray = 0; s = row; }
  else if (tile == 2) { ray = row >> 6; s = (row < 64) ? 128 + row : row - 64; }
  else { ray = 1; s = 64 + row; }
}

// ------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBfThreads, 1) snerf_bf16_render_kernel(const RenderParams p, const int iters) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  BfSmem& sm = *reinterpret_cast<BfSmem*>(smem_raw);  // stays in the shared address space (LDS/STS, not generic LD/ST)
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();     // 128B-swizzled UMMA tiles need 1024-byte alignment
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned char* img[2] = {p.img_coarse, p.img_fine ? p.img_fine : p.img_coarse};

  if (tid == 0) {
    for (int s = 0; s < kRing; ++s) { mbar_init(&sm.w_full[s], 1); mbar_init(&sm.w_empty[s], 1); }
    for (int c = 0; c < 2; ++c) {
      mbar_init(&sm.a_ready[c], kChainThreads);
      mbar_init(&sm.acc_ready[c], 1);
      mbar_init(&sm.pk_full[c][0], 1);
      mbar_init(&sm.pk_full[c][1], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) {  // allocate all 512 TMEM columns: two 128x256 fp32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&sm.tmem_base))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    // ================================ weight producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t nstep[2] = {0, 0};
      for (int it = 0; it < iters; ++it)
        for (int tile = 0; tile < 4; ++tile) {
          const unsigned char* im = img[tile == 0 ? 0 : 1];
          for (int step = 0; step < kBfSteps; ++step) {
            const int first = bf_step_first_chunk(step), n = bf_step_chunks(step);
            for (int c = 0; c < 2; ++c) {
              for (int i = 0; i < n; ++i) {
                mbar_wait(&sm.w_empty[stage], phase ^ 1);
                if (i == 0) {
                  const int par = nstep[c] & 1;
                  mbar_arrive_expect_tx(&sm.pk_full[c][par], kBfPacketBytes);
                  bulk_g2s(sm.packet[c][par], im + kBfPacketsOffset + step * kBfPacketBytes, kBfPacketBytes,
                           &sm.pk_full[c][par]);
                  ++nstep[c];
                }
                mbar_arrive_expect_tx(&sm.w_full[stage], kBfChunkBytes);
                bulk_g2s(sm.ring[stage], im + kBfChunksOffset + (size_t)(first + i) * kBfChunkBytes, kBfChunkBytes,
                         &sm.w_full[stage]);
                if (++stage == kRing) { stage = 0; phase ^= 1; }
              }
            }
          }
        }
    }
  } else if (warp == 1) {
    // ================================== MMA issuer ==================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t aphase[2] = {0, 0};
      const uint32_t enc_addr[2] = {smem_u32(sm.enc[0]), smem_u32(sm.enc[1])};
      const uint32_t act_addr[2] = {smem_u32(sm.act[0]), smem_u32(sm.act[1])};
      for (int it = 0; it < iters; ++it)
        for (int tile = 0; tile < 4; ++tile)
          for (int step = 0; step < kBfSteps; ++step)
            for (int c = 0; c < 2; ++c) {
              mbar_wait(&sm.a_ready[c], aphase[c]);
              aphase[c] ^= 1;
              tc_fence_after();
              const int nhalf = (step == 9) ? 1 : 2;
              const int nkb = (step == 0) ? 1 : (step == 5 ? 5 : 4);
              for (int nh = 0; nh < nhalf; ++nh) {
                const uint32_t d_tmem = tmem_base + (uint32_t)(c * 256 + nh * 128);
                for (int kb = 0; kb < nkb; ++kb) {
                  uint32_t a_addr;
                  if (step == 0 || (step == 5 && kb == 0)) a_addr = enc_addr[c];
                  else a_addr = act_addr[c] + (uint32_t)((step == 5 ? kb - 1 : kb) * kBfChunkBytes);
                  mbar_wait(&sm.w_full[stage], phase);
                  tc_fence_after();
                  const uint32_t b_addr = smem_u32(sm.ring[stage]);
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk)
                    tc_mma_bf16(d_tmem, umma_desc_sw128(a_addr + kk * 32), umma_desc_sw128(b_addr + kk * 32), idesc,
                                (kb | kk) != 0 ? 1u : 0u);
                  tc_commit(&sm.w_empty[stage]);
                  if (++stage == kRing) { stage = 0; phase ^= 1; }
                }
              }
              tc_commit(&sm.acc_ready[c]);
            }
    }
  } else {
    // ============================ the two sample chains ============================
    const int c = (warp - 2) >> 2;             // chain 0 / 1
    const int wt = tid - 64 - c * kChainThreads;  // 0..127 within the chain
    const int wq = warp & 3;                   // TMEM lane quarter this warp may read
    const int row = wq * 32 + lane;            // tile row owned by this thread
    const int wl = wt >> 5;                    // warp index within the chain (0..3)
    const int bar_id = 2 + c;
    const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(c * 256);
    uint8_t* act = sm.act[c];
    uint8_t* enc = sm.enc[c];
    ChainScratch& sc = *reinterpret_cast<ChainScratch*>(sm.enc[c]);
    const int chains_total = 2 * gridDim.x;
    const long long n_pairs = (p.n_rays + 1) >> 1;
    uint32_t acc_phase = 0;
    uint32_t nstep = 0;
    const int Nc = 64, Nf = 128, S = 192;

    for (int it = 0; it < iters; ++it) {
      const long long pair = (long long)it * chains_total + (long long)blockIdx.x * 2 + c;
      const bool pair_valid = pair < n_pairs;
      long long ray_idx[2];
      bool ray_valid[2];
      for (int r = 0; r < 2; ++r) {
        const long long ri = pair * 2 + r;
        ray_valid[r] = pair_valid && ri < p.n_rays;
        ray_idx[r] = ri < p.n_rays ? ri : p.n_rays - 1;
      }
      // ---- (A) ray records + direction encodings
      if (wt < 2) {
        const Ray q = load_ray(p.ray_batch + ray_idx[wt] * p.row_stride, p.width, p.has_vd);
        float* rr = sm.rayrec[c][wt];
        rr[0] = q.ox; rr[1] = q.oy; rr[2] = q.oz; rr[3] = q.dx; rr[4] = q.dy; rr[5] = q.dz;
        rr[6] = q.near; rr[7] = q.far; rr[8] = q.vx; rr[9] = q.vy; rr[10] = q.vz; rr[11] = q.dnorm;
      }
      named_bar_sync(bar_id, kChainThreads);
      if (wt < 64) {
        const int r = wt >> 5, k = wt & 31;
        const float* rr = sm.rayrec[c][r];
        float v = 0.f;
        if (k < 3) v = rr[8 + k];
        else if (k < 27) {
          const int o = (k - 3) / 6, j = (k - 3) % 6;
          const float a = rr[8 + j % 3] * __int_as_float((127 + o) << 23);
          v = j < 3 ? sinf(a) : cosf(a);
        }
        sm.direnc[c][r][k] = v;
      }
      // ---- (B) coarse depths (render.py:330-352)
      {
        const int r = row >> 6, i = row & 63;
        const float* rr = sm.rayrec[c][r];
        const float near = rr[6], far = rr[7];
        float z = coarse_depth(near, far, p.t_vals[i], p.lindisp);
        if (p.t_rand) {
          const float zm1 = i > 0 ? coarse_depth(near, far, p.t_vals[i - 1], p.lindisp) : z;
          const float zp1 = i < Nc - 1 ? coarse_depth(near, far, p.t_vals[i + 1], p.lindisp) : z;
          z = jitter_depth(zm1, z, zp1, i == 0, i == Nc - 1, p.t_rand[ray_idx[r] * Nc + i]);
        }
        sm.zc[c][r][i] = z;
        if (ray_valid[r] && p.out.z_vals_map) p.out.z_vals_map[ray_idx[r] * Nc + i] = z;
      }
      named_bar_sync(bar_id, kChainThreads);

      for (int tile = 0; tile < 4; ++tile) {
        const int net = tile == 0 ? 0 : 1;
        // ---- per-ray bias of the views layer: b_views + W_views[:, 256:283] . direnc   (fp32)
        if (tile < 2) {
          const float* wd = reinterpret_cast<const float*>(img[net] + kBfDirWOffset) + wt * 32;
          const float bv = __ldg(reinterpret_cast<const float*>(img[net] + kBfPacketsOffset + 9 * kBfPacketBytes) + wt);
          float a0 = bv, a1 = bv;
#pragma unroll
          for (int k = 0; k < 27; ++k) {
            const float w = __ldg(wd + k);
            a0 = fmaf(w, sm.direnc[c][0][k], a0);
            a1 = fmaf(w, sm.direnc[c][1][k], a1);
          }
          sm.dirbias[c][0][wt] = a0;
          sm.dirbias[c][1][wt] = a1;
          named_bar_sync(bar_id, kChainThreads);
        }
        // ---- encode this thread's sample into the A operand of step 0 / step 5
        int ray, s;
        row_to_sample(tile, row, ray, s);
        {
          const Ray q = ray_from_rec(sm.rayrec[c][ray]);
          const float z = tile == 0 ? sm.zc[c][ray][s] : sm.zf[c][ray][s];
          const float pt[3] = {ray_point(q.ox, q.dx, z), ray_point(q.oy, q.dy, z), ray_point(q.oz, q.dz, z)};
          float e[64];
          e[0] = pt[0]; e[1] = pt[1]; e[2] = pt[2];
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            float sn, cs;
            sincosf(pt[a], &sn, &cs);
            e[3 + a] = sn; e[6 + a] = cs;
#pragma unroll
            for (int o = 1; o < 10; ++o) {  // sin 2x = 2 sin x cos x ; cos 2x = 1 - 2 sin^2 x
              const float s2 = 2.f * sn * cs;
              cs = fmaf(-2.f * sn, sn, 1.f);
              sn = s2;
              e[3 + 6 * o + a] = sn; e[6 + 6 * o + a] = cs;
            }
          }
          e[63] = 0.f;
#pragma unroll
          for (int q8 = 0; q8 < 8; ++q8) {
            uint4 v;
            v.x = pack_bf16x2(e[q8 * 8 + 0], e[q8 * 8 + 1]);
            v.y = pack_bf16x2(e[q8 * 8 + 2], e[q8 * 8 + 3]);
            v.z = pack_bf16x2(e[q8 * 8 + 4], e[q8 * 8 + 5]);
            v.w = pack_bf16x2(e[q8 * 8 + 6], e[q8 * 8 + 7]);
            *reinterpret_cast<uint4*>(enc + sw128_offset(row, q8)) = v;
          }
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(&sm.a_ready[c]);

        float sigma = 0.f;
        for (int step = 0; step < kBfSteps; ++step) {
          const int par = nstep & 1;
          mbar_wait(&sm.pk_full[c][par], (nstep >> 1) & 1);
          ++nstep;
          mbar_wait(&sm.acc_ready[c], acc_phase);
          acc_phase ^= 1;
          tc_fence_after();
          const float* pk = sm.packet[c][par];
          float h0 = 0.f, h1 = 0.f, h2 = 0.f;
          if (step < 7) epilogue<EPI_RELU>(taddr, pk, pk, act, row, h0, h1, h2);
          else if (step == 7) { epilogue<EPI_ALPHA>(taddr, pk, pk + 256, act, row, h0, h1, h2); sigma = h0 + pk[512]; }
          else if (step == 8) epilogue<EPI_LINEAR>(taddr, pk, pk, act, row, h0, h1, h2);
          else {
            epilogue<EPI_RGB>(taddr, sm.dirbias[c][ray], pk + 128, act, row, h0, h1, h2);
            const float4 rv = make_float4(h0 + pk[512], h1 + pk[513], h2 + pk[514], sigma);
            sc.raw[row] = rv;  // enc buffer is dead since step 5
            float* rawg = tile == 0 ? p.out.raw_coarse : p.out.raw;
            if (rawg && ray_valid[ray])
              *reinterpret_cast<float4*>(rawg + (ray_idx[ray] * (tile == 0 ? Nc : S) + s) * 4) = rv;
          }
          if (step < 9) {
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&sm.a_ready[c]);
          }
        }
        tc_fence_before();
        named_bar_sync(bar_id, kChainThreads);

        // ---- composite this tile's segments (one warp per segment)
        if (tile == 0) {
          if (wl < 2) {
            const int r = wl;
            const Ray q = ray_from_rec(sm.rayrec[c][r]);
            const RayCarry cc = composite_segment(sc.raw + r * 64, sm.zc[c][r], Nc, 0, Nc, q.dnorm,
                                                  p.noise0 ? p.noise0 + ray_idx[r] * Nc : nullptr, sc.wts[r],
                                                  (ray_valid[r] && p.out.weights) ? p.out.weights + ray_idx[r] * Nc : nullptr,
                                                  carry_init(), lane);
            if (lane == 0 && ray_valid[r]) {
              const long long ri = ray_idx[r];
              const float wb = p.white_bkgd ? (1.f - cc.acc) : 0.f;
              if (p.out.rgb0) { p.out.rgb0[ri * 3] = cc.r + wb; p.out.rgb0[ri * 3 + 1] = cc.g + wb; p.out.rgb0[ri * 3 + 2] = cc.b + wb; }
              if (p.out.disp0) p.out.disp0[ri] = disparity(cc.depth, cc.acc);
              if (p.out.acc0) p.out.acc0[ri] = cc.acc;
              if (p.out.depth0) p.out.depth0[ri] = cc.depth;
            }
            // hierarchical resampling (run_nerf_helpers.py:336-379) + merge (render.py:383)
            const int B = Nc - 1;
            for (int i = lane; i < B; i += 32) sc.bins[r][i] = __fmul_rn(0.5f, __fadd_rn(sm.zc[c][r][i + 1], sm.zc[c][r][i]));
            __syncwarp();
            build_cdf(sc.wts[r] + 1, B, sc.cdf[r], lane);
            __syncwarp();
            for (int j = lane; j < Nf; j += 32) {
              const float u = p.u_rand ? p.u_rand[ray_idx[r] * Nf + j] : p.u_vals[j];
              int ind;
              const float zs = invert_cdf_one(sc.bins[r], sc.cdf[r], B, u, &ind);
              sc.zs[r][j] = zs;
              if (ray_valid[r] && p.out.z_samples) p.out.z_samples[ray_idx[r] * Nf + j] = zs;
            }
            __syncwarp();
            const float sd = warp_std(sc.zs[r], Nf, lane);
            if (lane == 0 && ray_valid[r] && p.out.z_std) p.out.z_std[ray_idx[r]] = sd;
            if (p.u_rand) warp_sort(sc.zs[r], Nf, lane);
            __syncwarp();
            merge_sorted(sm.zc[c][r], Nc, sc.zs[r], Nf, sm.zf[c][r], lane);
            __syncwarp();
            if (ray_valid[r] && p.out.z_all)
              for (int i = lane; i < S; i += 32) p.out.z_all[ray_idx[r] * S + i] = sm.zf[c][r][i];
            if (lane == 0) sm.carry[c][r] = carry_init();
          }
        } else {
          // segments of fine tiles: tile1 = ray0[0:128]; tile2 = ray0[128:192], ray1[0:64]; tile3 = ray1[64:192]
          int seg_ray = -1, seg_s0 = 0, seg_cnt = 0, seg_row0 = 0;
          if (tile == 1 && wl == 0) { seg_ray = 0; seg_s0 = 0; seg_cnt = 128; seg_row0 = 0; }
          if (tile == 2 && wl == 0) { seg_ray = 0; seg_s0 = 128; seg_cnt = 64; seg_row0 = 0; }
          if (tile == 2 && wl == 1) { seg_ray = 1; seg_s0 = 0; seg_cnt = 64; seg_row0 = 64; }
          if (tile == 3 && wl == 0) { seg_ray = 1; seg_s0 = 64; seg_cnt = 128; seg_row0 = 0; }
          if (seg_ray >= 0) {
            const int r = seg_ray;
            const Ray q = ray_from_rec(sm.rayrec[c][r]);
            const RayCarry cc = composite_segment(sc.raw + seg_row0, sm.zf[c][r], S, seg_s0, seg_cnt, q.dnorm,
                                                  p.noise1 ? p.noise1 + ray_idx[r] * S : nullptr, nullptr,
                                                  (ray_valid[r] && p.out.weights_fine) ? p.out.weights_fine + ray_idx[r] * S : nullptr,
                                                  sm.carry[c][r], lane);
            if (lane == 0) sm.carry[c][r] = cc;
            if (seg_s0 + seg_cnt == S && lane == 0 && ray_valid[r]) {
              const long long ri = ray_idx[r];
              const float wb = p.white_bkgd ? (1.f - cc.acc) : 0.f;
              if (p.out.rgb_map) { p.out.rgb_map[ri * 3] = cc.r + wb; p.out.rgb_map[ri * 3 + 1] = cc.g + wb; p.out.rgb_map[ri * 3 + 2] = cc.b + wb; }
              if (p.out.disp_map) p.out.disp_map[ri] = disparity(cc.depth, cc.acc);
              if (p.out.acc_map) p.out.acc_map[ri] = cc.acc;
              if (p.out.depth_map) p.out.depth_map[ri] = cc.depth;
            }
          }
        }
        named_bar_sync(bar_id, kChainThreads);  // scratch (enc buffer) free again, zf / carry visible
      }
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------------------------
// bring-up self test: one 128x128x64 tile through the same descriptors / swizzle / TMEM path
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) snerf_selftest_umma_kernel(const float* __restrict__ a,
                                                                     const float* __restrict__ b,
                                                                     float* __restrict__ d) {
  __shared__ alignas(1024) uint8_t sa[kBfChunkBytes];
  __shared__ alignas(1024) uint8_t sb[kBfChunkBytes];
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&done, 1); mbar_fence_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // row `tid` of A and of B -> swizzled bf16
  for (int q8 = 0; q8 < 8; ++q8) {
    uint4 va, vb;
    const float* pa = a + tid * 64 + q8 * 8;
    const float* pb = b + tid * 64 + q8 * 8;
    va.x = pack_bf16x2(pa[0], pa[1]); va.y = pack_bf16x2(pa[2], pa[3]);
    va.z = pack_bf16x2(pa[4], pa[5]); va.w = pack_bf16x2(pa[6], pa[7]);
    vb.x = pack_bf16x2(pb[0], pb[1]); vb.y = pack_bf16x2(pb[2], pb[3]);
    vb.z = pack_bf16x2(pb[4], pb[5]); vb.w = pack_bf16x2(pb[6], pb[7]);
    *reinterpret_cast<uint4*>(sa + sw128_offset(tid, q8)) = va;
    *reinterpret_cast<uint4*>(sb + sw128_offset(tid, q8)) = vb;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (tid == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
      tc_mma_bf16(tmem_base, umma_desc_sw128(smem_u32(sa) + kk * 32), umma_desc_sw128(smem_u32(sb) + kk * 32), idesc,
                  kk != 0 ? 1u : 0u);
    tc_commit(&done);
  }
  mbar_wait(&done, 0);
  tc_fence_after();
  const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
  for (int j = 0; j < 4; ++j) {
    uint32_t v[32];
    tmem_ld32(taddr + j * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) d[(warp * 32 + lane) * 128 + j * 32 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------
int launch_bf16_render(const RenderParams& p, cudaStream_t stream) {
  const size_t smem = sizeof(BfSmem);
  if (check_cuda(cudaFuncSetAttribute(snerf_bf16_render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                 "cudaFuncSetAttribute(bf16 kernel smem)"))
    return SNERF_ERR_CUDA;
  if (p.n_rays <= 0) return SNERF_OK;
  const long long n_pairs = (p.n_rays + 1) / 2;
  long long grid = (n_pairs + 1) / 2;
  if (grid > sm_count()) grid = sm_count();
  const long long chains = 2 * grid;
  const int iters = (int)((n_pairs + chains - 1) / chains);
  snerf_bf16_render_kernel<<<(unsigned)grid, kBfThreads, smem, stream>>>(p, iters);
  return check_cuda(cudaGetLastError(), "launch snerf_bf16_render_kernel");
}

int launch_bf16_query(const RenderParams&, cudaStream_t) {
  set_error("network_query_fn in bf16 mode is not available on its own; use mode fp32 or the fused renderer");
  return SNERF_ERR_UNSUPPORTED;
}

int launch_selftest_umma(const float* a, const float* b, float* d, cudaStream_t stream) {
  snerf_selftest_umma_kernel<<<1, 128, 0, stream>>>(a, b, d);
  return check_cuda(cudaGetLastError(), "launch snerf_selftest_umma_kernel");
}

}  // namespace snerf
