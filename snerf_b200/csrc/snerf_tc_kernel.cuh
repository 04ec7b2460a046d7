// snerf_tc_kernel.cuh -- tensor-core renderer (kernel template; instantiated by snerf_bf16.cu and snerf_x3.cu): render_rays as ONE persistent sm_100a kernel with the MLP on
// tcgen05 tensor cores (bf16 operands, fp32 accumulators in TMEM).
//
// Specialised for the configuration BASELINE.json's metric is quoted on: NeRF D=8, W=256, skip=4,
// 63/27 encoded inputs, view directions, 64 coarse + 128 fine samples (render.py:281-409,
// run_nerf_helpers.py:74-126).
//
// Work unit = a PAIR of rays = 4 MLP tiles of 128 sample-rows:
//     C   coarse   ray0[0:64]  | ray1[0:64]           (coarse network)
//     F1  fine     ray0[0:128]                        (fine network, sorted union of 192 depths)
//     F2  fine     ray0[128:192] | ray1[0:64]
//     F3  fine     ray1[64:192]
// A CTA walks the tile sequence  C(0) | C(1) F1(0) F2(0) F3(0) | C(2) F1(1) F2(1) F3(1) | ...  so the
// coarse->fine dependency (composite, inverse-CDF, merge) of a pair has three tiles of slack.
// Each tile runs ten tensor-core steps (L0..L7, feature, views); the alpha/rgb heads and the direction
// half of the views layer are folded into the epilogues (CUDA cores, fp32).
//
// Where the data lives (per SM):
//   TMEM (all 512 columns)   cols   0-255  fp32 accumulator of the current step (two 128-col halves)
//                            cols 256-383  A operand "ping": hidden activations, bf16, 2 per column
//                            cols 384-511  A operand "pong"
//     -> layer l+1's tcgen05.mma reads its A operand straight from TMEM (the epilogue of layer l wrote it
//        there with tcgen05.st); k-block kb of layer l+1 can issue as soon as the epilogue has produced
//        columns [64kb, 64kb+64), so the tensor core keeps running while the second accumulator half drains.
//   shared memory            10-stage ring of 16 KiB pre-swizzled weight chunks (B operand) fed by bulk async
//                            copies (cp.async.bulk / UBLKCP); two 16 KiB buffers for the encoded points
//                            (A operand of L0 and of the skip layer); per-step parameter packets; per-pair
//                            depths / carries.  Nothing per-sample ever goes to HBM.
// CTA = 448 threads, 1 CTA / SM, persistent:
//     warp 0       weight producer      warp 1      MMA issuer (single thread) + TMEM allocator
//     warps 2-9    two epilogue warpgroups: thread = TMEM lane = tile row; warpgroup e drains 32-column chunks
//                  {2e, 2e+1} of each accumulator half: TMEM -> +bias -> ReLU -> bf16 -> TMEM (next A operand)
//     warps 10-13  front-end warpgroup  : rays, stratified depths, encoding of the NEXT tile, compositing /
//                                         inverse-CDF / merge of the PREVIOUS tile, output writes
#pragma once
#include <cstdlib>
#include <cuda_fp16.h>

#include "snerf_common.cuh"
#include "snerf_internal.h"
#include "snerf_packed.h"
#include "snerf_umma.cuh"

namespace snerf {

constexpr int kBfThreads = 448;

// operand arithmetic of the tensor-core MLP (template parameter kOp of everything below)
//   OP_BF16   bf16 operands                              1 MMA pass   (throughput mode)
//   OP_F16    fp16 operands                              1 MMA pass
//   OP_F16X3  fp32-class: every operand split into fp16 hi + fp16 lo, products hi*hi + lo*hi + hi*lo accumulated
//             in the fp32 TMEM accumulator (the dropped lo*lo term is 2^-22 relative)   3 MMA passes
enum { OP_BF16 = 0, OP_F16 = 1, OP_F16X3 = 2 };

// packet buffers: 4 small ones in the split mode; 2 (each with its bias tile) otherwise -- the producer then loads step
// g + 1's packet while step g runs, which is as far ahead as the weight ring lets it run anyway
template <bool kSplit> struct PkRing { static constexpr int kBufs = kSplit ? 4 : 2, kShift = kSplit ? 2 : 1; };
constexpr int kGroup = 128;  // threads per warpgroup (epilogue / front-end)

// TMEM column map
constexpr uint32_t kAccCol = 0;
constexpr uint32_t kAbufCol0 = 256;  // "ping"
constexpr uint32_t kAbufCol1 = 384;  // "pong"
// OP_F16X3: no ping/pong -- cols 256-383 hold the hi halves of the A operand, cols 384-511 the lo halves; the epilogue
// keeps the first accumulator half's results in registers until every MMA of the step has read the old operand.
constexpr uint32_t kAhiCol = 256;
constexpr uint32_t kAloCol = 384;

// Four K=16 MMAs over one 64-wide k-block + the commit that frees the weight stage, as ONE predicated block
// (issued by the elected lane only; no divergent branch around it).  b_lo = low word of the B descriptor;
// successive K slices advance it by 2 (32 bytes >> 4).  TS form: A from TMEM (8 columns per K slice).
// Frees a weight stage: with clusters the stage is shared (multicast) by all CTAs of the cluster, so the commit
// arrives on the same barrier in every CTA.
// kPair: the two CTAs of the cluster work as ONE tensor-core unit (tcgen05 cta_group::2): M = 256 rows (128 per CTA),
// each CTA's shared memory holds HALF of the B tile (64 of the 128 output channels of the k-block), the leader CTA issues
// every MMA and its commits arrive on the same barrier in both CTAs.
template <int kCluster, bool kPair = false>
__device__ __forceinline__ void commit_stage(uint32_t leader, uint32_t empty_bar) {
  if (kPair) {
    asm volatile(
        "{\n\t.reg .pred pl;\n\t.reg .b16 m;\n\tsetp.ne.b32 pl, %0, 0;\n\tmov.b16 m, 3;\n\t"
        "@pl tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%1], m;\n\t}"
        ::"r"(leader), "r"(empty_bar)
        : "memory");
  } else if (kCluster == 1) {
    asm volatile(
        "{\n\t.reg .pred pl;\n\tsetp.ne.b32 pl, %0, 0;\n\t"
        "@pl tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%1];\n\t}"
        ::"r"(leader), "r"(empty_bar)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred pl;\n\t.reg .b16 m;\n\tsetp.ne.b32 pl, %0, 0;\n\tmov.b16 m, %2;\n\t"
        "@pl tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%1], m;\n\t}"
        ::"r"(leader), "r"(empty_bar), "n"((1 << kCluster) - 1)
        : "memory");
  }
}
#define SNERF_ISSUE4_TS(GROUP)                                                                                    \
  asm volatile(                                                                                                   \
      "{\n\t.reg .pred pl, pa;\n\t.reg .b64 b0, b1, b2, b3;\n\t.reg .b32 t1, t2, t3, a1, a2, a3;\n\t"               \
      "setp.ne.b32 pl, %0, 0;\n\t"                                                                                \
      "setp.ne.b32 pa, %6, 0;\n\t"                                                                                \
      "add.u32 t1, %3, 2;\n\tadd.u32 t2, %3, 4;\n\tadd.u32 t3, %3, 6;\n\t"                                        \
      "add.u32 a1, %2, 8;\n\tadd.u32 a2, %2, 16;\n\tadd.u32 a3, %2, 24;\n\t"                                      \
      "mov.b64 b0, {%3, %4};\n\tmov.b64 b1, {t1, %4};\n\tmov.b64 b2, {t2, %4};\n\tmov.b64 b3, {t3, %4};\n\t"      \
      "@pl tcgen05.mma.cta_group::" GROUP ".kind::f16 [%1], [%2], b0, %5, pa;\n\t"                                \
      "@pl tcgen05.mma.cta_group::" GROUP ".kind::f16 [%1], [a1], b1, %5, pl;\n\t"                                \
      "@pl tcgen05.mma.cta_group::" GROUP ".kind::f16 [%1], [a2], b2, %5, pl;\n\t"                                \
      "@pl tcgen05.mma.cta_group::" GROUP ".kind::f16 [%1], [a3], b3, %5, pl;\n\t}"                               \
      ::"r"(leader), "r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)               \
      : "memory")
template <bool kPair = false>
__device__ __forceinline__ void issue4_ts(uint32_t leader, uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo,
                                          uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
  if (kPair) SNERF_ISSUE4_TS("2"); else SNERF_ISSUE4_TS("1");
}
#undef SNERF_ISSUE4_TS
template <int kCluster, bool kPair = false>
__device__ __forceinline__ void issue_kblock_ts(uint32_t leader, uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo,
                                                uint32_t desc_hi, uint32_t idesc, uint32_t accumulate,
                                                uint32_t empty_bar) {
  issue4_ts<kPair>(leader, d_tmem, a_tmem, b_lo, desc_hi, idesc, accumulate);
  commit_stage<kCluster, kPair>(leader, empty_bar);
}
// SS form: A from shared memory (the encoded points)
#define SNERF_ISSUE4_SS(GROUP)                                                                                    \
  asm volatile(                                                                                                   \
      "{\n\t.reg .pred pl, pa;\n\t.reg .b64 a0, a1, a2, a3, b0, b1, b2, b3;\n\t.reg .b32 t1, t2, t3, u1, u2, u3;\n\t" \
      "setp.ne.b32 pl, %0, 0;\n\t"                                                                                \
      "setp.ne.b32 pa, %6, 0;\n\t"                                                                                \
      "add.u32 t1, %3, 2;\n\tadd.u32 t2, %3, 4;\n\tadd.u32 t3, %3, 6;\n\t"                                        \
      "add.u32 u1, %2, 2;\n\tadd.u32 u2, %2, 4;\n\tadd.u32 u3, %2, 6;\n\t"                                        \
      "mov.b64 b0, {%3, %4};\n\tmov.b64 b1, {t1, %4};\n\tmov.b64 b2, {t2, %4};\n\tmov.b64 b3, {t3, %4};\n\t"      \
      "mov.b64 a0, {%2, %4};\n\tmov.b64 a1, {u1, %4};\n\tmov.b64 a2, {u2, %4};\n\tmov.b64 a3, {u3, %4};\n\t"      \
      "@pl tcgen05.mma.cta_group::" GROUP ".kind::f16 [%1], a0, b0, %5, pa;\n\t"                                  \
      "@pl tcgen05.mma.cta_group::" GROUP ".kind::f16 [%1], a1, b1, %5, pl;\n\t"                                  \
      "@pl tcgen05.mma.cta_group::" GROUP ".kind::f16 [%1], a2, b2, %5, pl;\n\t"                                  \
      "@pl tcgen05.mma.cta_group::" GROUP ".kind::f16 [%1], a3, b3, %5, pl;\n\t}"                                 \
      ::"r"(leader), "r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)                 \
      : "memory")
template <bool kPair = false>
__device__ __forceinline__ void issue4_ss(uint32_t leader, uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo,
                                          uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
  if (kPair) SNERF_ISSUE4_SS("2"); else SNERF_ISSUE4_SS("1");
}
#undef SNERF_ISSUE4_SS
// one K=16 MMA, both operands from shared memory, accumulator overwritten (the bias MMA; single-CTA form)
__device__ __forceinline__ void issue1_ss(uint32_t leader, uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                          uint32_t b_hi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred pl, pz;\n\t.reg .b64 a0, b0;\n\t"
      "setp.ne.b32 pl, %0, 0;\n\t"
      "setp.ne.b32 pz, %0, %0;\n\t"
      "mov.b64 a0, {%2, %3};\n\tmov.b64 b0, {%4, %5};\n\t"
      "@pl tcgen05.mma.cta_group::1.kind::f16 [%1], a0, b0, %6, pz;\n\t}"
      ::"r"(leader), "r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
      : "memory");
}
template <int kCluster, bool kPair = false>
__device__ __forceinline__ void issue_kblock_ss(uint32_t leader, uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo,
                                                uint32_t desc_hi, uint32_t idesc, uint32_t accumulate,
                                                uint32_t empty_bar) {
  issue4_ss<kPair>(leader, d_tmem, a_lo, b_lo, desc_hi, idesc, accumulate);
  commit_stage<kCluster, kPair>(leader, empty_bar);
}
template <bool kPair = false>
__device__ __forceinline__ void commit_if(uint32_t leader, uint32_t bar) {
  if (kPair) {
    asm volatile(
        "{\n\t.reg .pred pl;\n\t.reg .b16 m;\n\tsetp.ne.b32 pl, %0, 0;\n\tmov.b16 m, 3;\n\t"
        "@pl tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%1], m;\n\t}"
        ::"r"(leader), "r"(bar)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred pl;\n\tsetp.ne.b32 pl, %0, 0;\n\t"
        "@pl tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%1];\n\t}"
        ::"r"(leader), "r"(bar)
        : "memory");
  }
}
// arrive on the barrier at `bar`'s offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// epilogue -> MMA hand-off.  kPair: the leader CTA's MMA warp waits for BOTH CTAs' epilogues, one arrival per warp.
// (relaxed: what the barrier orders is TMEM, written by tcgen05.st and completed by tcgen05.wait::st + the tcgen05 fence
//  before this call -- no generic-memory release is needed)
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint64_t* bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
template <bool kPair>
__device__ __forceinline__ void arrive_a_ready(uint64_t* bar) {
  if (kPair) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive_remote_relaxed(bar, 0);
  } else {
    mbar_arrive(bar);
  }
}

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
// operand-type generic versions (kF16: fp16 operands instead of bf16)
template <bool kF16>
__device__ __forceinline__ uint32_t cvt_relu_x2(float lo, float hi) {
  if (!kF16) return cvt_relu_bf16x2(lo, hi);
  uint32_t r;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <bool kF16>
__device__ __forceinline__ uint32_t cvt_x2(float lo, float hi) {
  if (!kF16) return cvt_bf16x2(lo, hi);
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// OP_F16X3: fp32 pair -> fp16x2 "hi" word + fp16x2 "lo" word, x = hi + lo up to 2^-22 |x| (the residual x - hi is
// exact in fp32; lo is its fp16 rounding)
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = cvt_x2<true>(a, b);
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = cvt_x2<true>(__fsub_rn(a, hf.x), __fsub_rn(b, hf.y));
}

// ------------------------------------------------------------------------------------
// shared memory
// ------------------------------------------------------------------------------------
// Sample geometry (compile-time): NC coarse + NF importance samples per ray, both multiples of 64 so that every
// 64-row half tile belongs to exactly one ray.  A pair of rays = RC coarse tiles + RF fine tiles of 128 rows:
//   coarse rows of the pair: [ray0: NC][ray1: NC]      fine rows: [ray0: S][ray1: S],  S = NC + NF
template <int NC, int NF>
struct Geo {
  static constexpr int Nc = NC, Nf = NF, S = NC + NF;
  static constexpr int RC = 2 * NC / 128;                   // coarse tiles per pair
  static constexpr int RF = NF > 0 ? 2 * (NC + NF) / 128 : 0;  // fine tiles per pair
  static constexpr int P = RC + RF;
  static constexpr int SS = S > 0 ? S : 1;
  static_assert(NC == 64 || NC == 128, "N_samples must be 64 or 128 in the tensor-core kernel");
  static_assert(NF % 64 == 0 && NC + NF <= 256, "N_samples + N_importance must be a multiple of 64, at most 256");
  __host__ __device__ static constexpr int n_tiles(int T) { return NF > 0 ? RC + T * P : T * RC; }
};
struct TileId {
  int fine;  // 0 = coarse network, 1 = fine network
  int q;     // local pair index
  int t;     // tile index inside the pair's coarse / fine block
};
// tile sequence of a CTA:  C(0) | C(1) F(0) | C(2) F(1) | ...  (C(q) = RC coarse tiles, F(q) = RF fine tiles of pair q;
// the last coarse block is a dummy); coarse-only (NF == 0): C(0) C(1) ...
template <class G>
__device__ __forceinline__ TileId tile_info(int n) {
  TileId id;
  if (G::Nf == 0) { id.fine = 0; id.q = n / G::RC; id.t = n % G::RC; return id; }
  if (n < G::RC) { id.fine = 0; id.q = 0; id.t = n; return id; }
  const int m = n - G::RC, grp = m / G::P, r = m % G::P;
  if (r < G::RC) { id.fine = 0; id.q = grp + 1; id.t = r; }
  else { id.fine = 1; id.q = grp; id.t = r - G::RC; }
  return id;
}
// tile row -> (ray in pair, sample index)
template <class G>
__device__ __forceinline__ void row_to_sample(const TileId& id, int row, int& ray, int& s) {
  const int X = id.fine ? G::S : G::Nc;
  const int idx = id.t * 128 + row;
  ray = idx / X;
  s = idx - ray * X;
}

template <class G>
struct alignas(16) PairData {  // everything about one ray pair that outlives a tile (triple buffered)
  float rayrec[2][12];
  float direnc[2][32];
  float dirbias[2][2][128];  // [network][ray]: b_views + W_views[:, 256:283] . direnc
  float zc[2][G::Nc];
  float zf[2][G::SS];
  RayCarry carry_c[2], carry_f[2];
  long long ray_idx[2];
  long long pair_index;  // global pair index, -1 for a padding pair (training: rows 2 * pair_index * X ... of the stores)
  int ray_valid[2];
};
template <class G, int kRing, bool kSplit>
struct alignas(1024) BfSmemT {
  // encoded points of tile n in enc[n & 1] (128B-swizzled A operand; OP_F16X3: [0] = hi halves, [1] = lo halves)
  uint8_t enc[2][kSplit ? 2 : 1][kBfChunkBytes];
  uint8_t ring[kRing][kBfChunkBytes];  // weight chunks (B operand)
  float packet[PkRing<kSplit>::kBufs][kSplit ? kBfPacketHeadFloats : kBfPacketFloats];   // (the split mode has no bias tile)
  uint8_t ones[4][128];                // bias MMA: A operand cores [1,1,1,0..] | zeros | [0,0,0,1,1,1,0,0] | zeros (8 rows x 16 B each)
  float4 raw[2][2][128];               // partial (r,g,b,sigma) of tile n from epilogue group e in raw[n & 1][e]
  PairData<G> pair[3];
  float wts[2][G::Nc], cdf[2][G::Nc], bins[2][G::Nc], zs[2][G::Nf > 0 ? G::Nf : 1];  // inverse-CDF scratch
  uint64_t w_full[kRing], w_empty[kRing];
  uint64_t pk_full[PkRing<kSplit>::kBufs];   // producer -> epilogue / MMA : packet of step g is in packet[g % kBufs]
  uint64_t pk_empty[PkRing<kSplit>::kBufs];  // epilogue -> producer
  uint64_t enc_full[2];        // front-end -> MMA : encoding of tile n is in enc[n & 1]
  uint64_t tile_started;       // MMA -> front-end  : first MMAs of tile n completed (tile n-1 no longer reads its enc)
  uint64_t acc_ready[2];       // MMA -> epilogue   : accumulator half h of the current step is complete
  uint64_t a_ready[4];         // epilogue -> MMA   : k-block kb of the next A operand is in TMEM (and, for kb 1 / 3,
                               //                     accumulator half 0 / 1 has been drained)
  uint64_t raw_full[2];        // epilogue -> front-end
  uint64_t raw_free[2];        // front-end -> epilogue
  uint64_t w_peer[kRing];      // kPair, leader CTA: the peer CTA's half of the weight chunk is in ITS ring slot
  uint64_t enc_peer[2];        // kPair, leader CTA: the peer CTA's encoding of tile n is in ITS enc[n & 1]
  uint32_t tmem_base;
};
// deepest weight ring that fits the 227 KB of shared memory for this geometry
template <class G, bool kSplit>
struct RingFor {
  static constexpr int value = sizeof(BfSmemT<G, 10, kSplit>) <= 232448 ? 10
                               : (sizeof(BfSmemT<G, 9, kSplit>) <= 232448 ? 9
                               : (sizeof(BfSmemT<G, 8, kSplit>) <= 232448 ? 8 : (sizeof(BfSmemT<G, 7, kSplit>) <= 232448 ? 7 : 6)));
  static_assert(sizeof(BfSmemT<G, value, kSplit>) <= 232448, "shared memory budget");
};

__device__ __forceinline__ Ray ray_from_rec(const float* r) {
  Ray q;
  q.ox = r[0]; q.oy = r[1]; q.oz = r[2]; q.dx = r[3]; q.dy = r[4]; q.dz = r[5];
  q.near = r[6]; q.far = r[7]; q.vx = r[8]; q.vy = r[9]; q.vz = r[10]; q.dnorm = r[11];
  return q;
}

// ------------------------------------------------------------------------------------
// epilogue of one step for one tile row (thread = TMEM lane = row)
//   EPI_RELU   +bias, ReLU, bf16 -> next A operand (steps 0..6)
//   EPI_ALPHA  same, plus sigma = alpha_linear(h) on the fp32 hidden state (step 7)
//   EPI_LINEAR no activation (feature_linear, step 8)
//   EPI_RGB    views layer (N=128): +per-ray direction bias, ReLU, rgb_linear; nothing stored (step 9)
// The accumulator is drained in 32-column chunks with the TMEM load of chunk j+1 in flight while chunk j is
// processed; a_ready[kb] is signalled as soon as k-block kb of the next A operand is complete.
// ------------------------------------------------------------------------------------
enum { EPI_RELU = 0, EPI_ALPHA = 1, EPI_LINEAR = 2, EPI_RGB = 3 };

// one 32-column chunk: v = accumulator columns [col, col+32) of this thread's row
// kNoBias: the accumulator already holds the bias (bias tile MMA, snerf_packed.h) -- every kind except EPI_RGB, whose
// bias is per ray (direction half of the views layer)
template <int KIND, bool kF16, bool kSave = false, bool kNoBias = false>
__device__ __forceinline__ void epi_chunk(const uint32_t (&v)[32], int col, const float* __restrict__ bias,
                                          const float* __restrict__ aux, uint32_t (&packed)[16], uint64_t& acc0,
                                          uint64_t& acc1, uint64_t& acc2) {
#pragma unroll
  for (int q8 = 0; q8 < 4; ++q8) {
    const int c = col + q8 * 8;
    float f[8];
    if (kNoBias && KIND != EPI_RGB) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[q8 * 8 + i]);
    } else {
      const float4 b0 = *reinterpret_cast<const float4*>(bias + c);
      const float4 b1 = *reinterpret_cast<const float4*>(bias + c + 4);
      uint64_t s[4];
      s[0] = add2(pack2u(v[q8 * 8 + 0], v[q8 * 8 + 1]), pack2f(b0.x, b0.y));
      s[1] = add2(pack2u(v[q8 * 8 + 2], v[q8 * 8 + 3]), pack2f(b0.z, b0.w));
      s[2] = add2(pack2u(v[q8 * 8 + 4], v[q8 * 8 + 5]), pack2f(b1.x, b1.y));
      s[3] = add2(pack2u(v[q8 * 8 + 6], v[q8 * 8 + 7]), pack2f(b1.z, b1.w));
#pragma unroll
      for (int i = 0; i < 4; ++i) unpack2f(s[i], f[2 * i], f[2 * i + 1]);
    }
    if (KIND == EPI_ALPHA || KIND == EPI_RGB) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
    }
    if (KIND == EPI_ALPHA) {
      const float4 w0 = *reinterpret_cast<const float4*>(aux + c);
      const float4 w1 = *reinterpret_cast<const float4*>(aux + c + 4);
      acc0 = fma2(pack2f(f[0], f[1]), pack2f(w0.x, w0.y), acc0);
      acc1 = fma2(pack2f(f[2], f[3]), pack2f(w0.z, w0.w), acc1);
      acc0 = fma2(pack2f(f[4], f[5]), pack2f(w1.x, w1.y), acc0);
      acc1 = fma2(pack2f(f[6], f[7]), pack2f(w1.z, w1.w), acc1);
    }
    if (KIND == EPI_RGB) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float4 w0 = *reinterpret_cast<const float4*>(aux + k * 128 + c);
        const float4 w1 = *reinterpret_cast<const float4*>(aux + k * 128 + c + 4);
        uint64_t& a = k == 0 ? acc0 : (k == 1 ? acc1 : acc2);
        a = fma2(pack2f(f[0], f[1]), pack2f(w0.x, w0.y), a);
        a = fma2(pack2f(f[2], f[3]), pack2f(w0.z, w0.w), a);
        a = fma2(pack2f(f[4], f[5]), pack2f(w1.x, w1.y), a);
        a = fma2(pack2f(f[6], f[7]), pack2f(w1.z, w1.w), a);
      }
      if (kSave) {  // training: the (ReLU'd) views activation is rgb_linear's input -> activation store
        packed[q8 * 4 + 0] = cvt_x2<kF16>(f[0], f[1]); packed[q8 * 4 + 1] = cvt_x2<kF16>(f[2], f[3]);
        packed[q8 * 4 + 2] = cvt_x2<kF16>(f[4], f[5]); packed[q8 * 4 + 3] = cvt_x2<kF16>(f[6], f[7]);
      }
    } else if (KIND == EPI_LINEAR) {
      packed[q8 * 4 + 0] = cvt_x2<kF16>(f[0], f[1]); packed[q8 * 4 + 1] = cvt_x2<kF16>(f[2], f[3]);
      packed[q8 * 4 + 2] = cvt_x2<kF16>(f[4], f[5]); packed[q8 * 4 + 3] = cvt_x2<kF16>(f[6], f[7]);
    } else {
      packed[q8 * 4 + 0] = cvt_relu_x2<kF16>(f[0], f[1]); packed[q8 * 4 + 1] = cvt_relu_x2<kF16>(f[2], f[3]);
      packed[q8 * 4 + 2] = cvt_relu_x2<kF16>(f[4], f[5]); packed[q8 * 4 + 3] = cvt_relu_x2<kF16>(f[6], f[7]);
    }
  }
}

// Epilogue group e (0/1) of one step: for each accumulator half h it drains chunks j = 4h + 2e, 4h + 2e + 1
// (both TMEM loads in flight at once), writes the bf16 result as k-block (2h + e) of the next A operand and
// signals a_ready[2h + e].  Head partial sums (over this group's columns) come back in o0..o2.
// 16 packed words (32 operand values of this row = four 8-channel chunks) -> activation store: chunk i of the row goes
// to dst[32 i] (snerf_packed.h: lanes = consecutive rows are adjacent, so a warp writes 512 contiguous bytes per chunk)
__device__ __forceinline__ void save_words(uint4* dst, const uint32_t (&w)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) dst[32 * i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}
// relu' mask of 32 stored (non-negative) 16-bit values: bit k = word k's low half is non-zero, bit 16 + k = its high half
__device__ __forceinline__ uint32_t nonzero_bits(const uint32_t (&w)[16]) {
  // h + 0x7FFF sets bit 15 iff h > 0 (h <= 0x7F80: no carry into the other half).  Word k's two flags enter at bits 15 / 31
  // and are shifted down once per later word: they end at bits k / 16 + k (shift, add, and-or: three instructions a word).
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) m = (m >> 1) | ((w[k] + 0x7FFF7FFFu) & 0x80008000u);
  return m;
}
// kSave (training forward): `save` = this row's position in column block 0 of the step's slot of the activation store
// (null for rows of padding pairs), `bits` = likewise in the mask store (null: the slot has no ReLU); the 64 columns this
// call produces per accumulator half are column block 2 h + e.
template <int KIND, bool kF16, bool kSave = false, bool kPair = false, bool kNoBias = false>
__device__ __forceinline__ void epilogue(uint64_t* acc_ready, uint64_t* a_ready, uint32_t acc_addr, uint32_t anext_addr,
                                         uint32_t acc_phase,
                                         const float* __restrict__ bias, const float* __restrict__ aux, int e,
                                         float& o0, float& o1, float& o2, uint4* save = nullptr,
                                         unsigned long long* bits = nullptr) {
  // Column assignment: of accumulator half h (32-column chunks 4h .. 4h+3) group e drains chunk 4h + e, then chunk
  // 4h + 2 + e.  The two groups together therefore finish k-block 2h of the next A operand (chunks 4h, 4h+1) half an
  // epilogue before k-block 2h+1: the MMA warp can issue the next layer's k-block 2 while k-block 3 is still being
  // produced, and only ONE k-block (256 tensor cycles) separates the end of this epilogue from the next accumulator
  // half -- the TMEM read path (64 B/clk: 1024 clk per half) is what paces the layer chain, so that gap is what counts.
  constexpr int NHALF = (KIND == EPI_RGB) ? 1 : 2;
  uint64_t acc0 = 0, acc1 = 0, acc2 = 0;  // packed partial sums of the head dot products
#pragma unroll
  for (int h = 0; h < NHALF; ++h) {
    mbar_wait(&acc_ready[h], acc_phase);
    tc_fence_after();
    const int ja = 4 * h + e, jb = 4 * h + 2 + e;
    uint32_t va[32], vb[32];
    tmem_ld32(acc_addr + (uint32_t)(ja * 32), va);
    tmem_ld32(acc_addr + (uint32_t)(jb * 32), vb);
    tmem_ld_wait_dep(va);    // (tcgen05.wait::ld covers both loads: from here on this thread has drained its part of the half)
    uint32_t pa[16], pb[16];
    epi_chunk<KIND, kF16, kSave, kNoBias>(va, ja * 32, bias, aux, pa, acc0, acc1, acc2);
    if (KIND != EPI_RGB) {
      tmem_st16(anext_addr + (uint32_t)(ja * 16), pa);
      tmem_st_wait();
    }
    tc_fence_before();
    arrive_a_ready<kPair>(&a_ready[2 * h]);
    tmem_ld_wait_dep(vb);
    epi_chunk<KIND, kF16, kSave, kNoBias>(vb, jb * 32, bias, aux, pb, acc0, acc1, acc2);
    if (KIND != EPI_RGB) {
      tmem_st16(anext_addr + (uint32_t)(jb * 16), pb);
      tmem_st_wait();
    }
    tc_fence_before();
    arrive_a_ready<kPair>(&a_ready[2 * h + 1]);
    if (kSave && save) {  // after the barriers: the tensor core does not wait for the global stores
      // chunk j = 32-column half (j & 1) of column block j >> 1 of `save` (EPI_RGB: save points at block 2)
      save_words(save + (ja >> 1) * 256 + (ja & 1) * 128, pa);
      save_words(save + (jb >> 1) * 256 + (jb & 1) * 128, pb);
      if (bits) {   // one 64-bit word per (row, column block): low half = its first 32 columns
        reinterpret_cast<uint32_t*>(bits + (ja >> 1) * 32)[ja & 1] = nonzero_bits(pa);
        reinterpret_cast<uint32_t*>(bits + (jb >> 1) * 32)[jb & 1] = nonzero_bits(pb);
      }
    }
  }
  if (KIND == EPI_RGB) {  // N=128 step: no second half; keep every barrier's phase count uniform
    mbar_wait(&acc_ready[1], acc_phase);
    arrive_a_ready<kPair>(&a_ready[2]);
    arrive_a_ready<kPair>(&a_ready[3]);
    float a, b;
    unpack2f(acc0, a, b); o0 = a + b;
    unpack2f(acc1, a, b); o1 = a + b;
    unpack2f(acc2, a, b); o2 = a + b;
  }
  if (KIND == EPI_ALPHA) {
    float a, b, c, d;
    unpack2f(acc0, a, b); unpack2f(acc1, c, d);
    o0 = (a + b) + (c + d);
  }
}

// ---- OP_F16X3 epilogue -------------------------------------------------------------------------------------------
// one 32-column chunk -> 16 packed hi words + 16 packed lo words of the next A operand (fp32 heads as above)
template <int KIND>
__device__ __forceinline__ void epi_chunk_x3(const uint32_t (&v)[32], int col, const float* __restrict__ bias,
                                             const float* __restrict__ aux, uint32_t (&phi)[16], uint32_t (&plo)[16],
                                             uint64_t& acc0, uint64_t& acc1, uint64_t& acc2) {
#pragma unroll
  for (int q8 = 0; q8 < 4; ++q8) {
    const int c = col + q8 * 8;
    const float4 b0 = *reinterpret_cast<const float4*>(bias + c);
    const float4 b1 = *reinterpret_cast<const float4*>(bias + c + 4);
    // the accumulator holds (2^4 x) . (2^8 w): scale back (exact) while adding the bias
    const uint64_t inv = pack2f(kX3InvScale, kX3InvScale);
    uint64_t s[4];
    s[0] = fma2(pack2u(v[q8 * 8 + 0], v[q8 * 8 + 1]), inv, pack2f(b0.x, b0.y));
    s[1] = fma2(pack2u(v[q8 * 8 + 2], v[q8 * 8 + 3]), inv, pack2f(b0.z, b0.w));
    s[2] = fma2(pack2u(v[q8 * 8 + 4], v[q8 * 8 + 5]), inv, pack2f(b1.x, b1.y));
    s[3] = fma2(pack2u(v[q8 * 8 + 6], v[q8 * 8 + 7]), inv, pack2f(b1.z, b1.w));
    float f[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) unpack2f(s[i], f[2 * i], f[2 * i + 1]);
    if (KIND != EPI_LINEAR) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
    }
    if (KIND == EPI_ALPHA) {
      const float4 w0 = *reinterpret_cast<const float4*>(aux + c);
      const float4 w1 = *reinterpret_cast<const float4*>(aux + c + 4);
      acc0 = fma2(pack2f(f[0], f[1]), pack2f(w0.x, w0.y), acc0);
      acc1 = fma2(pack2f(f[2], f[3]), pack2f(w0.z, w0.w), acc1);
      acc0 = fma2(pack2f(f[4], f[5]), pack2f(w1.x, w1.y), acc0);
      acc1 = fma2(pack2f(f[6], f[7]), pack2f(w1.z, w1.w), acc1);
    }
    if (KIND == EPI_RGB) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float4 w0 = *reinterpret_cast<const float4*>(aux + k * 128 + c);
        const float4 w1 = *reinterpret_cast<const float4*>(aux + k * 128 + c + 4);
        uint64_t& a = k == 0 ? acc0 : (k == 1 ? acc1 : acc2);
        a = fma2(pack2f(f[0], f[1]), pack2f(w0.x, w0.y), a);
        a = fma2(pack2f(f[2], f[3]), pack2f(w0.z, w0.w), a);
        a = fma2(pack2f(f[4], f[5]), pack2f(w1.x, w1.y), a);
        a = fma2(pack2f(f[6], f[7]), pack2f(w1.z, w1.w), a);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        split_f16x2(f[2 * i] * kX3ActScale, f[2 * i + 1] * kX3ActScale, phi[q8 * 4 + i], plo[q8 * 4 + i]);
    }
  }
}

// Epilogue group e of one OP_F16X3 step.  There is ONE A-operand buffer (hi + lo fill the 256 columns the ping/pong
// pair occupies otherwise), so the results of accumulator half 0 -- drained while the tensor core works on half 1 --
// wait in registers until acc_ready[1] says that every MMA of the step has consumed the old operand; then they are
// stored as k-block e (a_ready[e]), and half 1 is drained straight into k-block 2 + e (a_ready[2 + e]) while the
// next step's first k-blocks already run.
template <int KIND>
__device__ __forceinline__ void epilogue_x3(uint64_t* acc_ready, uint64_t* a_ready, uint32_t acc_addr, uint32_t ahi_addr,
                                            uint32_t alo_addr, uint32_t acc_phase, const float* __restrict__ bias,
                                            const float* __restrict__ aux, int e, float& o0, float& o1, float& o2) {
  uint64_t acc0 = 0, acc1 = 0, acc2 = 0;
  {
    mbar_wait(&acc_ready[0], acc_phase);
    tc_fence_after();
    const int j0 = 2 * e;
    uint32_t v[32], h0[16], l0[16], h1[16], l1[16];
    tmem_ld32(acc_addr + (uint32_t)(j0 * 32), v);
    tmem_ld_wait_dep(v);
    epi_chunk_x3<KIND>(v, j0 * 32, bias, aux, h0, l0, acc0, acc1, acc2);
    tmem_ld32(acc_addr + (uint32_t)(j0 * 32 + 32), v);
    tmem_ld_wait_dep(v);
    epi_chunk_x3<KIND>(v, j0 * 32 + 32, bias, aux, h1, l1, acc0, acc1, acc2);
    mbar_wait(&acc_ready[1], acc_phase);  // all MMAs of this step are complete: the A operand may be overwritten
    tc_fence_after();
    if (KIND != EPI_RGB) {
      tmem_st16(ahi_addr + (uint32_t)(j0 * 16), h0);
      tmem_st16(alo_addr + (uint32_t)(j0 * 16), l0);
      tmem_st16(ahi_addr + (uint32_t)(j0 * 16 + 16), h1);
      tmem_st16(alo_addr + (uint32_t)(j0 * 16 + 16), l1);
      tmem_st_wait();
    }
    tc_fence_before();
    mbar_arrive(&a_ready[e]);
  }
  if (KIND == EPI_RGB) {  // N=128 step: no second half
    mbar_arrive(&a_ready[2 + e]);
    float a, b;
    unpack2f(acc0, a, b); o0 = a + b;
    unpack2f(acc1, a, b); o1 = a + b;
    unpack2f(acc2, a, b); o2 = a + b;
    return;
  }
  {
    const int j0 = 4 + 2 * e;
    uint32_t va[32], vb[32], ph[16], pl[16];
    tmem_ld32(acc_addr + (uint32_t)(j0 * 32), va);
    tmem_ld32(acc_addr + (uint32_t)(j0 * 32 + 32), vb);
    tmem_ld_wait_dep(va);
    epi_chunk_x3<KIND>(va, j0 * 32, bias, aux, ph, pl, acc0, acc1, acc2);
    tmem_st16(ahi_addr + (uint32_t)(j0 * 16), ph);
    tmem_st16(alo_addr + (uint32_t)(j0 * 16), pl);
    tmem_ld_wait_dep(vb);
    epi_chunk_x3<KIND>(vb, j0 * 32 + 32, bias, aux, ph, pl, acc0, acc1, acc2);
    tmem_st16(ahi_addr + (uint32_t)(j0 * 16 + 16), ph);
    tmem_st16(alo_addr + (uint32_t)(j0 * 16 + 16), pl);
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(&a_ready[2 + e]);
  }
  if (KIND == EPI_ALPHA) {
    float a, b, c, d;
    unpack2f(acc0, a, b); unpack2f(acc1, c, d);
    o0 = (a + b) + (c + d);
  }
}

// ------------------------------------------------------------------------------------
// front-end pieces (128 threads, thread index wt)
// ------------------------------------------------------------------------------------
// load the pair's rays, direction encodings, per-network direction biases and coarse depths
template <class G, bool kSplit>
__device__ __forceinline__ void frontend_load_pair(const RenderParams& p, const unsigned char* const* img,
                                                   PairData<G>& pd, long long gpair, bool pair_valid, int wt,
                                                   int bar_id) {
  if (wt < 2) {
    const long long ri = gpair * 2 + wt;
    const bool valid = pair_valid && ri < p.n_rays;
    const long long rc = (pair_valid && ri < p.n_rays) ? ri : p.n_rays - 1;
    pd.ray_idx[wt] = rc;
    pd.ray_valid[wt] = valid ? 1 : 0;
    if (wt == 0) pd.pair_index = pair_valid ? gpair : -1;
    const Ray q = make_ray(p, rc);
    float* rr = pd.rayrec[wt];
    rr[0] = q.ox; rr[1] = q.oy; rr[2] = q.oz; rr[3] = q.dx; rr[4] = q.dy; rr[5] = q.dz;
    rr[6] = q.near; rr[7] = q.far; rr[8] = q.vx; rr[9] = q.vy; rr[10] = q.vz; rr[11] = q.dnorm;
    pd.carry_c[wt] = carry_init();
    pd.carry_f[wt] = carry_init();
  }
  named_bar_sync(bar_id, kGroup);
  if (wt < 64) {
    const int r = wt >> 5, k = wt & 31;
    const float* rr = pd.rayrec[r];
    float v = 0.f;
    if (k < 3) v = rr[8 + k];
    else if (k < 27) {
      const int o = (k - 3) / 6, j = (k - 3) % 6;
      const float a = rr[8 + j % 3] * __int_as_float((127 + o) << 23);
      v = j < 3 ? sinf(a) : cosf(a);
    }
    pd.direnc[r][k] = v;
  }
  for (int k = wt; k < 2 * G::Nc; k += kGroup) {  // coarse depths (render.py:330-352): k -> (ray, i)
    const int r = k / G::Nc, i = k - r * G::Nc;
    const float near = pd.rayrec[r][6], far = pd.rayrec[r][7];
    float z = coarse_depth(near, far, p.t_vals[i], p.lindisp);
    if (p.t_rand) {
      const float zm1 = i > 0 ? coarse_depth(near, far, p.t_vals[i - 1], p.lindisp) : z;
      const float zp1 = i < G::Nc - 1 ? coarse_depth(near, far, p.t_vals[i + 1], p.lindisp) : z;
      z = jitter_depth(zm1, z, zp1, i == 0, i == G::Nc - 1, p.t_rand[pd.ray_idx[r] * G::Nc + i]);
    }
    pd.zc[r][i] = z;
    if (pd.ray_valid[r] && p.out.z_vals_map) p.out.z_vals_map[pd.ray_idx[r] * G::Nc + i] = z;
  }
  named_bar_sync(bar_id, kGroup);
#pragma unroll
  for (int net = 0; net < 2; ++net) {  // per-ray bias of the views layer, fp32
    const float* wd = reinterpret_cast<const float*>(img[net] + BfImage<kSplit>::kDirWOffset) + wt * 32;
    const float bv = __ldg(reinterpret_cast<const float*>(img[net] + BfImage<kSplit>::kPacketsOffset + 9 * kBfPacketBytes) + wt);
    float a0 = bv, a1 = bv;
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      const float w = __ldg(wd + k);
      a0 = fmaf(w, pd.direnc[0][k], a0);
      a1 = fmaf(w, pd.direnc[1][k], a1);
    }
    pd.dirbias[net][0][wt] = a0;
    pd.dirbias[net][1][wt] = a1;
  }
}

// encode row `wt` of tile (kind, pair) into the 128B-swizzled A-operand buffer `enc` (OP_F16X3: hi halves into enc,
// lo halves into enc + kBfChunkBytes)
// `save` (training): this row's position in column block 0 of slot 0 of the activation store: [enc 64 | dir 32 | zero 32 | views 128]
template <class G, int kOp>
__device__ __forceinline__ void frontend_encode(const PairData<G>& pd, const TileId& id, uint8_t* enc, int wt,
                                                uint4* save = nullptr) {
  constexpr bool kF16 = kOp != OP_BF16;
  int ray, s;
  row_to_sample<G>(id, wt, ray, s);
  const Ray q = ray_from_rec(pd.rayrec[ray]);
  const float z = id.fine ? pd.zf[ray][s] : pd.zc[ray][s];
  const float pt[3] = {ray_point(q.ox, q.dx, z), ray_point(q.oy, q.dy, z), ray_point(q.oz, q.dz, z)};
  float e[64];
  e[0] = pt[0]; e[1] = pt[1]; e[2] = pt[2];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (kOp == OP_F16X3) {
      // fp32-class mode: every octave from the accurate sincosf of the exactly scaled argument, as the reference
      // (and the fp32 kernel) evaluate sin(2^o x)
#pragma unroll
      for (int o = 0; o < 10; ++o) {
        float sn, cs;
        sincosf(pt[a] * __int_as_float((127 + o) << 23), &sn, &cs);
        e[3 + 6 * o + a] = sn; e[6 + 6 * o + a] = cs;
      }
    } else {
      // octave 0 with the accurate sincosf; octaves 1..9 by angle doubling (error <= 2^9 ulp ~ 3e-5, far below
      // the 2^-9 relative rounding of the bf16 operand it feeds)
      float sn, cs;
      sincosf(pt[a], &sn, &cs);
      e[3 + a] = sn; e[6 + a] = cs;
#pragma unroll
      for (int o = 1; o < 10; ++o) {
        const float s2 = 2.f * sn * cs;
        cs = fmaf(-2.f * sn, sn, 1.f);
        sn = s2;
        e[3 + 6 * o + a] = sn; e[6 + 6 * o + a] = cs;
      }
    }
  }
  e[63] = 0.f;
#pragma unroll
  for (int q8 = 0; q8 < 8; ++q8) {
    if (kOp == OP_F16X3) {
      uint4 vh, vl;
      constexpr float S = kX3ActScale;
      split_f16x2(e[q8 * 8 + 0] * S, e[q8 * 8 + 1] * S, vh.x, vl.x);
      split_f16x2(e[q8 * 8 + 2] * S, e[q8 * 8 + 3] * S, vh.y, vl.y);
      split_f16x2(e[q8 * 8 + 4] * S, e[q8 * 8 + 5] * S, vh.z, vl.z);
      split_f16x2(e[q8 * 8 + 6] * S, e[q8 * 8 + 7] * S, vh.w, vl.w);
      *reinterpret_cast<uint4*>(enc + sw128_offset(wt, q8)) = vh;
      *reinterpret_cast<uint4*>(enc + kBfChunkBytes + sw128_offset(wt, q8)) = vl;
    } else {
      uint4 v;
      v.x = cvt_x2<kF16>(e[q8 * 8 + 0], e[q8 * 8 + 1]);
      v.y = cvt_x2<kF16>(e[q8 * 8 + 2], e[q8 * 8 + 3]);
      v.z = cvt_x2<kF16>(e[q8 * 8 + 4], e[q8 * 8 + 5]);
      v.w = cvt_x2<kF16>(e[q8 * 8 + 6], e[q8 * 8 + 7]);
      *reinterpret_cast<uint4*>(enc + sw128_offset(wt, q8)) = v;
      if (save) save[32 * q8] = v;                       // column block 0, chunk q8
    }
  }
  if (kOp != OP_F16X3 && save) {                         // column block 1: direction encoding (4 chunks) + zeros
    const float* de = pd.direnc[ray];
#pragma unroll
    for (int q8 = 0; q8 < 4; ++q8)
      save[256 + 32 * q8] = make_uint4(cvt_x2<kF16>(de[q8 * 8 + 0], de[q8 * 8 + 1]), cvt_x2<kF16>(de[q8 * 8 + 2], de[q8 * 8 + 3]),
                                       cvt_x2<kF16>(de[q8 * 8 + 4], de[q8 * 8 + 5]), cvt_x2<kF16>(de[q8 * 8 + 6], de[q8 * 8 + 7]));
#pragma unroll
    for (int q8 = 4; q8 < 8; ++q8) save[256 + 32 * q8] = make_uint4(0u, 0u, 0u, 0u);
  }
}

// Composite the finished tile from its raw buffer, one 64-sample segment (= half tile) at a time with the ray's
// carry in between, so a ray's arithmetic never depends on where it sits in its pair (results are bit-identical
// under any split of the batch).  When a ray's last coarse segment is done: coarse outputs, then inverse-CDF
// resampling + merge (run_nerf_helpers.py:336-379, render.py:383); last fine segment: the final outputs.
template <class G, class Smem>
__device__ __forceinline__ void frontend_composite(Smem& sm, const RenderParams& p, PairData<G>& pd, const TileId& id,
                                                   const float4* raw, int wl, int lane) {
  constexpr int Nc = G::Nc, Nf = G::Nf, S = G::S;
  const int X = id.fine ? S : Nc;
  const int ray_h0 = (id.t * 128) / X, ray_h1 = (id.t * 128 + 64) / X;
  int h_first, h_cnt;  // which half tiles this warp composites
  if (ray_h0 == ray_h1) { if (wl != 0) return; h_first = 0; h_cnt = 2; }   // same ray: in order, one warp
  else { if (wl > 1) return; h_first = wl; h_cnt = 1; }                     // two rays: one warp each
  for (int h = h_first; h < h_first + h_cnt; ++h) {
    const int idx = id.t * 128 + 64 * h;
    const int r = idx / X, s0 = idx - r * X;
    const bool valid = pd.ray_valid[r] != 0;
    const long long ri = pd.ray_idx[r];
    const float dnorm = pd.rayrec[r][11];
    const float4* seg = raw + 64 * h;
    if (!id.fine) {
      const RayCarry cc = composite_segment(seg, pd.zc[r], Nc, s0, 64, dnorm, p.noise0 ? p.noise0 + ri * Nc : nullptr,
                                            sm.wts[r], (valid && p.out.weights) ? p.out.weights + ri * Nc : nullptr,
                                            pd.carry_c[r], lane);
      if (lane == 0) pd.carry_c[r] = cc;
      __syncwarp();  // the carry written by lane 0 is read back by every lane for the ray's next segment
      float* rawg = Nf > 0 ? p.out.raw_coarse : (p.out.raw ? p.out.raw : p.out.raw_coarse);
      if (valid && rawg)
        for (int i = lane; i < 64; i += 32) reinterpret_cast<float4*>(rawg)[ri * Nc + s0 + i] = seg[i];
      if (s0 + 64 != Nc) continue;
      // ---- the ray's coarse pass is complete
      if (lane == 0 && valid) {
        const float wb = p.white_bkgd ? (1.f - cc.acc) : 0.f;
        float* rgb = Nf > 0 ? p.out.rgb0 : p.out.rgb_map;
        float* disp = Nf > 0 ? p.out.disp0 : p.out.disp_map;
        float* acc = Nf > 0 ? p.out.acc0 : p.out.acc_map;
        float* depth = Nf > 0 ? p.out.depth0 : p.out.depth_map;
        if (rgb) { rgb[ri * 3] = cc.r + wb; rgb[ri * 3 + 1] = cc.g + wb; rgb[ri * 3 + 2] = cc.b + wb; }
        if (disp) disp[ri] = disparity(cc.depth, cc.acc);
        if (acc) acc[ri] = cc.acc;
        if (depth) depth[ri] = cc.depth;
      }
      if (Nf > 0) {
        __syncwarp();
        constexpr int B = Nc - 1;
        for (int i = lane; i < B; i += 32) sm.bins[r][i] = __fmul_rn(0.5f, __fadd_rn(pd.zc[r][i + 1], pd.zc[r][i]));
        __syncwarp();
        build_cdf(sm.wts[r] + 1, B, sm.cdf[r], lane);
        __syncwarp();
        for (int j = lane; j < Nf; j += 32) {
          const float u = p.u_rand ? p.u_rand[ri * Nf + j] : p.u_vals[j];
          int ind;
          const float zs = invert_cdf_one(sm.bins[r], sm.cdf[r], B, u, &ind);
          sm.zs[r][j] = zs;
          if (valid && p.out.z_samples) p.out.z_samples[ri * Nf + j] = zs;
        }
        __syncwarp();
        const float sd = warp_std(sm.zs[r], Nf, lane);
        if (lane == 0 && valid && p.out.z_std) p.out.z_std[ri] = sd;
        if (p.u_rand) warp_sort(sm.zs[r], Nf, lane);
        __syncwarp();
        merge_sorted(pd.zc[r], Nc, sm.zs[r], Nf, pd.zf[r], lane);
        __syncwarp();
        if (valid && p.out.z_all)
          for (int i = lane; i < S; i += 32) p.out.z_all[ri * S + i] = pd.zf[r][i];
      }
    } else {
      const RayCarry cc = composite_segment(seg, pd.zf[r], S, s0, 64, dnorm, p.noise1 ? p.noise1 + ri * S : nullptr,
                                            nullptr, (valid && p.out.weights_fine) ? p.out.weights_fine + ri * S : nullptr,
                                            pd.carry_f[r], lane);
      if (lane == 0) pd.carry_f[r] = cc;
      __syncwarp();
      if (valid && p.out.raw)
        for (int i = lane; i < 64; i += 32) reinterpret_cast<float4*>(p.out.raw)[ri * S + s0 + i] = seg[i];
      if (s0 + 64 == S && lane == 0 && valid) {
        const float wb = p.white_bkgd ? (1.f - cc.acc) : 0.f;
        if (p.out.rgb_map) { p.out.rgb_map[ri * 3] = cc.r + wb; p.out.rgb_map[ri * 3 + 1] = cc.g + wb; p.out.rgb_map[ri * 3 + 2] = cc.b + wb; }
        if (p.out.disp_map) p.out.disp_map[ri] = disparity(cc.depth, cc.acc);
        if (p.out.acc_map) p.out.acc_map[ri] = cc.acc;
        if (p.out.depth_map) p.out.depth_map[ri] = cc.depth;
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// the kernel.  T = ray pairs per CTA; the CTA runs the tile sequence of G (tile_info).
// ------------------------------------------------------------------------------------
// training forward (kSave): every layer's operand-precision output also goes to the activation store (block layout:
// snerf_packed.h): slot 0 = [enc 64 | dir 32 | 0 | views 128], slots 1..8 = h0..h7, slot 9 = feature; row =
// 2 * pair * X + position of the row in the pair's coarse (X = Nc) / fine (X = Nc + Nf) block; the relu' bits of
// h0..h7 / views go to the mask store.
template <class G>
__device__ __forceinline__ uint4* act_row(const RenderParams& p, const TileId& id, long long pair_index, int slot, int row) {
  if (pair_index < 0) return nullptr;
  const long long rows = id.fine ? p.act_rows_f : p.act_rows_c;
  const long long r = pair_index * 2 * (id.fine ? G::S : G::Nc) + id.t * 128 + row;
  unsigned char* base = id.fine ? p.act_f : p.act_c;
  if (!base) return nullptr;
  return reinterpret_cast<uint4*>(base + tc_block_offset(rows, slot, r >> 5, 0)) + (r & 31);
}
template <class G>
__device__ __forceinline__ unsigned long long* mask_row(const RenderParams& p, const TileId& id, long long pair_index, int slot,
                                                        int row) {
  if (pair_index < 0) return nullptr;
  const long long rows = id.fine ? p.act_rows_f : p.act_rows_c;
  const long long r = pair_index * 2 * (id.fine ? G::S : G::Nc) + id.t * 128 + row;
  unsigned long long* base = id.fine ? p.bits_f : p.bits_c;
  if (!base) return nullptr;
  return base + tc_mask_index(rows, slot, r >> 5, 0, (int)(r & 31));
}

// kCoarseD = 4: the coarse network is NeRF(D=4, W=256) without a live skip (create_nerf with netdepth = 4 as in the shipped
// configs; the fine network stays 8x256).  Its four trunk layers are steps 0, 1, 2, 7 of the image (first / hidden /
// hidden / last-with-alpha): a coarse tile simply leaves steps 3..6 out, in every warp role alike.
template <int kCoarseD>
__device__ __forceinline__ bool tc_skip_step(const TileId& id, int step) {
  return kCoarseD == 4 && !id.fine && step >= 3 && step <= 6;
}

template <int kCluster, class G, int kOp, bool kSave = false, bool kPair = false, int kCoarseD = 8>
__global__ void __launch_bounds__(kBfThreads, 1) snerf_bf16_render_kernel(const RenderParams p, const int T) {
  static_assert(kCoarseD == 8 || (kCoarseD == 4 && !kSave && !kPair), "4-layer coarse network: inference kernels only");
  static_assert(!(kSave && kOp == OP_F16X3), "the activation store holds single 16-bit operands");
  static_assert(!kPair || (kCluster == 2 && kOp != OP_F16X3), "cta_group::2 variant: 2-CTA clusters, single-pass operands");
  constexpr bool kF16 = kOp != OP_BF16;
  constexpr bool kSplit = kOp == OP_F16X3;
  // the bias enters the accumulator as one extra K=16 MMA per accumulator half (constant-ones A operand x bias tile)
  // instead of 64 shared-memory loads + adds per epilogue thread: the epilogue is what paces the layer chain
  constexpr bool kBiasMma = !kSplit && !kPair;
  constexpr uint32_t kPkBytes = kSplit ? kBfPacketHeadBytes : kBfPacketBytes;
  constexpr int kPkBufs = PkRing<kSplit>::kBufs, kPkShift = PkRing<kSplit>::kShift;
  using Img = BfImage<kSplit>;
  constexpr int kRing = RingFor<G, kSplit>::value;
  using Smem = BfSmemT<G, kRing, kSplit>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);  // stays in the shared address space (LDS/STS, not generic LD/ST)
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();     // 128B-swizzled UMMA tiles need 1024-byte alignment
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned char* img[2] = {p.img_coarse, p.img_fine ? p.img_fine : p.img_coarse};
  const int n_tiles = G::n_tiles(T);

  if (tid == 0) {
    // the packed images must have been built for this operand type (snerf_pack_weights(mode))
    constexpr uint32_t kMagic = kSplit ? kF16x3Magic : (kF16 ? kF16Magic : kBf16Magic);
    if (reinterpret_cast<const Bf16Header*>(img[0])->magic != kMagic ||
        reinterpret_cast<const Bf16Header*>(img[1])->magic != kMagic) __trap();
    // ... and for the network depths this instantiation walks (with a fine pass: coarse kCoarseD, fine 8)
    if (reinterpret_cast<const Bf16Header*>(img[0])->depth != kCoarseD ||
        (G::Nf > 0 && reinterpret_cast<const Bf16Header*>(img[1])->depth != 8)) __trap();
    for (int s = 0; s < kRing; ++s) {
      mbar_init(&sm.w_full[s], 1); mbar_init(&sm.w_empty[s], kPair ? 1 : kCluster); mbar_init(&sm.w_peer[s], 1);
    }
    for (int i = 0; i < kPkBufs; ++i) { mbar_init(&sm.pk_full[i], 1); mbar_init(&sm.pk_empty[i], 2 * kGroup); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.enc_full[i], kGroup);
      mbar_init(&sm.acc_ready[i], 1);
      mbar_init(&sm.raw_full[i], 2 * kGroup);
      mbar_init(&sm.raw_free[i], kGroup);
      mbar_init(&sm.enc_peer[i], 1);
    }
    // a_ready[kb]: both epilogue groups contribute half of k-block kb (split mode: one group per k-block);
    // kPair: one arrival per epilogue warp, both CTAs
    for (int i = 0; i < 4; ++i) mbar_init(&sm.a_ready[i], kSplit ? kGroup : (kPair ? 16 : 2 * kGroup));
    mbar_init(&sm.tile_started, 1);
    mbar_fence_init();
  }
  if (kBiasMma && tid >= 64 && tid < 64 + 128) {   // 4 cores x 8 rows x 4 words
    const int i = tid - 64, core = i >> 5, w = i & 3;          // word w of a row = operand values 2w, 2w + 1
    const uint32_t one = kF16 ? 0x3C00u : 0x3F80u;
    uint32_t v = 0;
    if (core == 0) v = w == 0 ? (one | (one << 16)) : (w == 1 ? one : 0u);            // [1,1,1,0,0,0,0,0]
    if (core == 2) v = w == 1 ? (one << 16) : (w == 2 ? (one | (one << 16)) : 0u);    // [0,0,0,1,1,1,0,0]
    reinterpret_cast<uint32_t*>(sm.ones)[i] = v;
    fence_proxy_async();
  }
  if (warp == 1) {  // all 512 TMEM columns: accumulator + two A-operand buffers
    if (kPair) {    // the same warp of both CTAs allocates the pair's tensor memory
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&sm.tmem_base))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&sm.tmem_base))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;
  uint32_t cta_rank = 0;
  if (kCluster > 1) {  // barriers of every CTA initialised before any multicast copy / commit can reach them
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }

  if (warp == 0) {
    // ================================ weight producer ================================
    // With clusters every weight chunk is fetched from L2 once per cluster: CTA r issues chunks r, r+kCluster, ...
    // as a multicast bulk copy into the same ring slot of every CTA (each CTA posts its own expect_tx).
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, g = 0;  // g = global step counter
      uint32_t nchunk = 0;
      for (int n = 0; n < n_tiles; ++n) {
        const TileId id = tile_info<G>(n);
        const unsigned char* im = img[id.fine];
        for (int step = 0; step < kBfSteps; ++step) {
          if (tc_skip_step<kCoarseD>(id, step)) continue;
          // (OP_F16X3: every chunk is followed by its lo part in the image -> twice the chunks, same order)
          const int first = (kSplit ? 2 : 1) * bf_step_first_chunk(step), cnt = (kSplit ? 2 : 1) * bf_step_chunks(step);
          {  // the step's parameter packet (4 buffers; wait until the epilogue of step g-4 is done with this one)
            const int pb = g & (kPkBufs - 1);
            mbar_wait(&sm.pk_empty[pb], ((g >> kPkShift) & 1) ^ 1);
            mbar_arrive_expect_tx(&sm.pk_full[pb], kPkBytes);
            bulk_g2s(sm.packet[pb], im + Img::kPacketsOffset + step * kBfPacketBytes, kPkBytes, &sm.pk_full[pb]);
          }
          for (int i = 0; i < cnt; ++i) {
            mbar_wait(&sm.w_empty[stage], phase ^ 1);
            const unsigned char* src = im + kBfChunksOffset + (size_t)(first + i) * kBfChunkBytes;
            if (kPair) {   // this CTA's half of the chunk: output channels [64 rank, 64 rank + 64) = whole 8-row swizzle atoms
              mbar_arrive_expect_tx(&sm.w_full[stage], kBfChunkBytes / 2);
              bulk_g2s(sm.ring[stage], src + cta_rank * (kBfChunkBytes / 2), kBfChunkBytes / 2, &sm.w_full[stage]);
              if (++stage == kRing) { stage = 0; phase ^= 1; }
              continue;
            }
            mbar_arrive_expect_tx(&sm.w_full[stage], kBfChunkBytes);
            if (kCluster == 1) bulk_g2s(sm.ring[stage], src, kBfChunkBytes, &sm.w_full[stage]);
            else if (nchunk % kCluster == cta_rank)
              bulk_g2s_multicast(sm.ring[stage], src, kBfChunkBytes, &sm.w_full[stage], (uint16_t)((1 << kCluster) - 1));
            ++nchunk;
            if (++stage == kRing) { stage = 0; phase ^= 1; }
          }
          ++g;
        }
      }
    }
  } else if (warp == 1) {
    // ================================== MMA issuer ==================================
    // The whole warp runs the (warp-uniform, straight-line per step) control flow so descriptors and barrier
    // addresses live in uniform registers; the elected lane issues each k-block (4 x tcgen05.mma + the commit
    // that frees its weight stage) as one predicated block.
    if (kPair && cta_rank != 0) {
      // peer CTA of a pair: no MMAs to issue.  Tell the leader when this CTA's operands are in place: the encoding of
      // each tile and this CTA's half of every weight chunk, in the order the leader consumes them.
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int n = 0; n < n_tiles; ++n) {
          mbar_wait(&sm.enc_full[n & 1], (n >> 1) & 1);
          mbar_arrive_remote(&sm.enc_peer[n & 1], 0);
          for (int step = 0; step < kBfSteps; ++step) {
            const int cnt = bf_step_chunks(step);
            for (int i = 0; i < cnt; ++i) {
              mbar_wait(&sm.w_full[stage], phase);
              mbar_arrive_remote(&sm.w_peer[stage], 0);
              if (++stage == kRing) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    } else {
    const uint32_t leader = elect_one() ? 1u : 0u;
    constexpr uint32_t idesc = kPair ? (kF16 ? umma_idesc_f16(256, 128) : umma_idesc_bf16(256, 128))
                                     : (kF16 ? umma_idesc_f16(128, 128) : umma_idesc_bf16(128, 128));
    constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);  // SBO | version | SWIZZLE_128B
    const uint32_t enc_lo[2] = {((smem_u32(sm.enc[0][0]) & 0x3FFFFu) >> 4) | (1u << 16),
                                ((smem_u32(sm.enc[1][0]) & 0x3FFFFu) >> 4) | (1u << 16)};
    const uint32_t ring_lo0 = ((smem_u32(sm.ring[0]) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t full0 = smem_u32(&sm.w_full[0]), empty0 = smem_u32(&sm.w_empty[0]);
    const uint32_t accr0 = smem_u32(&sm.acc_ready[0]), accr1 = smem_u32(&sm.acc_ready[1]);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t aphase = 1;  // a_ready parity to wait for; a fresh barrier reports the "previous" phase complete
    const uint32_t acc_h0 = tmem_base + kAccCol, acc_h1 = tmem_base + kAccCol + 128;
    // bias MMA operands (no-swizzle K-major cores): A = constant ones rows (every 8-row group aliases one core: SBO = 0;
    // K values 8..15 = the zero core 128 B further: LBO = 128), B = the packet's bias tile (row n at 16 n: SBO = 128;
    // K values 8..15 meet zeros of A, so they alias the same core: LBO = 0)
    const uint32_t ones_lo[2] = {((smem_u32(sm.ones[0]) & 0x3FFFFu) >> 4) | ((128u >> 4) << 16),
                                 ((smem_u32(sm.ones[2]) & 0x3FFFFu) >> 4) | ((128u >> 4) << 16)};
    constexpr uint32_t kOnesHi = (1u << 14), kTileHi = (128u >> 4) | (1u << 14);
    uint32_t g = 0;   // global step counter (packet buffer g & 3)
#define SNERF_BIAS_MMA(D_TMEM, HALF)                                                                     \
    do {                                                                                                  \
      if (kBiasMma) {                                                                                     \
        const uint32_t tile_lo = (smem_u32(&sm.packet[g & (kPkBufs - 1)][kSplit ? 0 : kBfPacketHeadFloats]) & 0x3FFFFu) >> 4; \
        issue1_ss(leader, (D_TMEM), ones_lo[HALF], kOnesHi, tile_lo, kTileHi, idesc);                     \
      }                                                                                                   \
    } while (0)

    // one 64-wide k-block: wait for its weight chunk, issue, advance the ring
#define SNERF_KBLOCK_TS(D_TMEM, A_TMEM, ACCUM)                                                          \
    do {                                                                                                  \
      mbar_wait(&sm.w_full[stage], phase);                                                                \
      if (kPair) mbar_wait(&sm.w_peer[stage], phase);                                                     \
      issue_kblock_ts<kCluster, kPair>(leader, (D_TMEM), (A_TMEM), ring_lo0 + (uint32_t)stage * (kBfChunkBytes >> 4), kDescHi, idesc, \
                      (ACCUM), empty0 + (uint32_t)stage * 8);                                             \
      if (++stage == kRing) { stage = 0; phase ^= 1; }                                                    \
    } while (0)
#define SNERF_KBLOCK_SS(D_TMEM, A_LO, ACCUM)                                                            \
    do {                                                                                                  \
      mbar_wait(&sm.w_full[stage], phase);                                                                \
      if (kPair) mbar_wait(&sm.w_peer[stage], phase);                                                     \
      issue_kblock_ss<kCluster, kPair>(leader, (D_TMEM), (A_LO), ring_lo0 + (uint32_t)stage * (kBfChunkBytes >> 4), kDescHi, idesc,   \
                      (ACCUM), empty0 + (uint32_t)stage * 8);                                             \
      if (++stage == kRing) { stage = 0; phase ^= 1; }                                                    \
    } while (0)

    // OP_F16X3 k-block: hi weights against the hi and the lo operand halves, then the lo weights against the hi halves
    // (two weight chunks, three groups of four MMAs).  The lo halves of the A operand sit 128 TMEM columns / one
    // 16 KiB shared-memory buffer above the hi halves.
#define SNERF_KBLOCK_TS3(D_TMEM, A_TMEM, ACCUM)                                                         \
    do {                                                                                                  \
      mbar_wait(&sm.w_full[stage], phase);                                                                \
      issue4_ts(leader, (D_TMEM), (A_TMEM), ring_lo0 + (uint32_t)stage * (kBfChunkBytes >> 4), kDescHi, idesc, (ACCUM)); \
      issue_kblock_ts<kCluster>(leader, (D_TMEM), (A_TMEM) + (kAloCol - kAhiCol),                         \
                      ring_lo0 + (uint32_t)stage * (kBfChunkBytes >> 4), kDescHi, idesc, 1u, empty0 + (uint32_t)stage * 8); \
      if (++stage == kRing) { stage = 0; phase ^= 1; }                                                    \
      SNERF_KBLOCK_TS((D_TMEM), (A_TMEM), 1u);                                                            \
    } while (0)
#define SNERF_KBLOCK_SS3(D_TMEM, A_LO, ACCUM)                                                           \
    do {                                                                                                  \
      mbar_wait(&sm.w_full[stage], phase);                                                                \
      issue4_ss(leader, (D_TMEM), (A_LO), ring_lo0 + (uint32_t)stage * (kBfChunkBytes >> 4), kDescHi, idesc, (ACCUM)); \
      issue_kblock_ss<kCluster>(leader, (D_TMEM), (A_LO) + (kBfChunkBytes >> 4),                          \
                      ring_lo0 + (uint32_t)stage * (kBfChunkBytes >> 4), kDescHi, idesc, 1u, empty0 + (uint32_t)stage * 8); \
      if (++stage == kRing) { stage = 0; phase ^= 1; }                                                    \
      SNERF_KBLOCK_SS((D_TMEM), (A_LO), 1u);                                                              \
    } while (0)
    // one k-block in the arithmetic of this instantiation
#define SNERF_KB_TS(D_TMEM, A_TMEM, ACCUM)                                                              \
    do { if (kSplit) SNERF_KBLOCK_TS3((D_TMEM), (A_TMEM), (ACCUM)); else SNERF_KBLOCK_TS((D_TMEM), (A_TMEM), (ACCUM)); } while (0)
#define SNERF_KB_SS(D_TMEM, A_LO, ACCUM)                                                                \
    do { if (kSplit) SNERF_KBLOCK_SS3((D_TMEM), (A_LO), (ACCUM)); else SNERF_KBLOCK_SS((D_TMEM), (A_LO), (ACCUM)); } while (0)

    for (int n = 0; n < n_tiles; ++n) {
      mbar_wait(&sm.enc_full[n & 1], (n >> 1) & 1);
      if (kPair) mbar_wait(&sm.enc_peer[n & 1], (n >> 1) & 1);
      const uint32_t a_enc = enc_lo[n & 1];
      const TileId id = tile_info<G>(n);
      for (int step = 0; step < kBfSteps; ++step) {
        if (tc_skip_step<kCoarseD>(id, step)) continue;
        const uint32_t biased = (kBiasMma && step != 9) ? 1u : 0u;   // the accumulator halves start from the bias tile
        if (biased) mbar_wait(&sm.pk_full[g & (kPkBufs - 1)], (g >> kPkShift) & 1);
        // A operand of the hidden k-blocks: the TMEM buffer the previous epilogue wrote
        // (epilogue(s) writes ping for even s, pong for odd s; OP_F16X3: the single hi/lo buffer)
        const uint32_t a_tmem = tmem_base + (kSplit ? kAhiCol : (((step - 1) & 1) ? kAbufCol1 : kAbufCol0));
        // ---- accumulator half 0.  a_ready[kb] = k-block kb of A is in TMEM; a_ready[0] / [2] also mean that
        // accumulator half 0 / 1 has been drained (split mode: [0] and [1] / [2] and [3] together)
        mbar_wait(&sm.a_ready[0], aphase);
        if (kSplit || step == 0) mbar_wait(&sm.a_ready[1], aphase);
        tc_fence_after();
        if (biased) SNERF_BIAS_MMA(acc_h0, 0);
        if (step == 0) {
          SNERF_KB_SS(acc_h0, a_enc, biased);
          commit_if<kPair>(leader, smem_u32(&sm.tile_started));
        } else {
          uint32_t first = biased;
          if (step == 5) { SNERF_KB_SS(acc_h0, a_enc, biased); first = 1u; }
          SNERF_KB_TS(acc_h0, a_tmem, first);
          if (!kSplit) { mbar_wait(&sm.a_ready[1], aphase); tc_fence_after(); }
          SNERF_KB_TS(acc_h0, a_tmem + 32, 1u);
          mbar_wait(&sm.a_ready[2], aphase);
          tc_fence_after();
          SNERF_KB_TS(acc_h0, a_tmem + 64, 1u);
          mbar_wait(&sm.a_ready[3], aphase);
          tc_fence_after();
          SNERF_KB_TS(acc_h0, a_tmem + 96, 1u);
        }
        commit_if<kPair>(leader, accr0);
        // ---- accumulator half 1 (not for the N=128 views step); a_ready[2], [3] also mean half 1 is drained
        if (step == 0) {
          mbar_wait(&sm.a_ready[2], aphase);
          mbar_wait(&sm.a_ready[3], aphase);
          tc_fence_after();
          if (biased) SNERF_BIAS_MMA(acc_h1, 1);
          SNERF_KB_SS(acc_h1, a_enc, biased);
        } else if (step != 9) {
          uint32_t first = biased;
          if (biased) SNERF_BIAS_MMA(acc_h1, 1);
          if (step == 5) { SNERF_KB_SS(acc_h1, a_enc, biased); first = 1u; }
          SNERF_KB_TS(acc_h1, a_tmem, first);
          SNERF_KB_TS(acc_h1, a_tmem + 32, 1u);
          SNERF_KB_TS(acc_h1, a_tmem + 64, 1u);
          SNERF_KB_TS(acc_h1, a_tmem + 96, 1u);
        }
        commit_if<kPair>(leader, accr1);  // (step 9: completes together with half 0; keeps phase counts uniform)
        aphase ^= 1;
        ++g;
      }
    }
#undef SNERF_KB_TS
#undef SNERF_KB_SS
#undef SNERF_KBLOCK_TS3
#undef SNERF_KBLOCK_SS3
#undef SNERF_KBLOCK_TS
#undef SNERF_KBLOCK_SS
#undef SNERF_BIAS_MMA
    (void)full0;
    }
  } else if (warp < 10) {
    // ============================= epilogue warpgroups (2) ============================
    const int e = (warp - 2) >> 2;   // epilogue group: which chunks of each accumulator half it drains
    const int wq = warp & 3;         // TMEM lane quarter this warp may access
    const int row = wq * 32 + lane;  // tile row owned by this thread
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const uint32_t acc_addr = tmem_base + lane_base + kAccCol;
    uint32_t acc_phase = 0, g = 0;
    for (int n = 0; n < n_tiles; ++n) {
      const TileId id = tile_info<G>(n);
      const PairData<G>& pd = sm.pair[id.q % 3];
      int ray, s;
      row_to_sample<G>(id, row, ray, s);
      float sigma = 0.f;
      long long pair_index = -1;
      for (int step = 0; step < kBfSteps; ++step) {
        if (tc_skip_step<kCoarseD>(id, step)) continue;
        const int pb = g & (kPkBufs - 1);
        mbar_wait(&sm.pk_full[pb], (g >> kPkShift) & 1);
        const float* pk = sm.packet[pb];
        const uint32_t anext = tmem_base + lane_base + ((step & 1) ? kAbufCol1 : kAbufCol0);
        const uint32_t ahi = tmem_base + lane_base + kAhiCol, alo = tmem_base + lane_base + kAloCol;  // OP_F16X3
        float h0 = 0.f, h1 = 0.f, h2 = 0.f;
        // training: where this row's output of the step goes (slot 1 + step for the trunk, 9 for the feature layer,
        // the upper half of slot 0 for the views layer)
        uint4* save = nullptr;
        unsigned long long* bits = nullptr;
        if (kSave) {
          if (step == 0) {  // the pair record was written by the front-end before it released this tile's encoding
            mbar_wait(&sm.acc_ready[0], acc_phase);
            pair_index = pd.pair_index;
          }
          save = act_row<G>(p, id, pair_index, step < 8 ? 1 + step : (step == 8 ? 9 : 0), row);
          if (save && step == 9) save += 2 * 256;                       // views: column blocks 2, 3 of slot 0
          if (step != 8) bits = mask_row<G>(p, id, pair_index, step < 8 ? step : 8, row);
        }
        if (step < 7) {
          if (kSplit) epilogue_x3<EPI_RELU>(sm.acc_ready, sm.a_ready, acc_addr, ahi, alo, acc_phase, pk, pk, e, h0, h1, h2);
          else epilogue<EPI_RELU, kF16, kSave, kPair, kBiasMma>(sm.acc_ready, sm.a_ready, acc_addr, anext, acc_phase, pk, pk, e, h0, h1, h2, save, bits);
        } else if (step == 7) {
          if (kSplit) epilogue_x3<EPI_ALPHA>(sm.acc_ready, sm.a_ready, acc_addr, ahi, alo, acc_phase, pk, pk + 256, e, h0, h1, h2);
          else epilogue<EPI_ALPHA, kF16, kSave, kPair, kBiasMma>(sm.acc_ready, sm.a_ready, acc_addr, anext, acc_phase, pk, pk + 256, e, h0, h1, h2, save, bits);
          sigma = h0 + (e == 0 ? pk[512] : 0.f);
        } else if (step == 8) {
          if (kSplit) epilogue_x3<EPI_LINEAR>(sm.acc_ready, sm.a_ready, acc_addr, ahi, alo, acc_phase, pk, pk, e, h0, h1, h2);
          else epilogue<EPI_LINEAR, kF16, kSave, kPair, kBiasMma>(sm.acc_ready, sm.a_ready, acc_addr, anext, acc_phase, pk, pk, e, h0, h1, h2, save, bits);
        } else {
          if (kSplit) epilogue_x3<EPI_RGB>(sm.acc_ready, sm.a_ready, acc_addr, ahi, alo, acc_phase, pd.dirbias[id.fine][ray], pk + 128, e, h0, h1, h2);
          else epilogue<EPI_RGB, kF16, kSave, kPair, kBiasMma>(sm.acc_ready, sm.a_ready, acc_addr, anext, acc_phase, pd.dirbias[id.fine][ray], pk + 128, e, h0, h1, h2, save, bits);
          mbar_wait(&sm.raw_free[n & 1], ((n >> 1) & 1) ^ 1);  // front-end is done with this buffer (tile n-2)
          const float br = e == 0 ? pk[512] : 0.f, bg = e == 0 ? pk[513] : 0.f, bb = e == 0 ? pk[514] : 0.f;
          sm.raw[n & 1][e][row] = make_float4(h0 + br, h1 + bg, h2 + bb, sigma);
          mbar_arrive(&sm.raw_full[n & 1]);
        }
        mbar_arrive(&sm.pk_empty[pb]);
        acc_phase ^= 1;
        ++g;
      }
    }
  } else {
    // ============================== front-end warpgroup =============================
    const int wt = tid - 320;  // 0..127
    const int wl = wt >> 5;
    const int bar_id = 2;
    const long long n_pairs = (p.n_rays + 1) >> 1;
    {  // prologue: pair 0 and the encoding of tile 0
      const long long gp = blockIdx.x;
      frontend_load_pair<G, kSplit>(p, img, sm.pair[0], gp, gp < n_pairs, wt, bar_id);
      named_bar_sync(bar_id, kGroup);
      frontend_encode<G, kOp>(sm.pair[0], tile_info<G>(0), sm.enc[0][0], wt,
                              kSave ? act_row<G>(p, tile_info<G>(0), sm.pair[0].pair_index, 0, wt) : nullptr);
      fence_proxy_async();
      mbar_arrive(&sm.enc_full[0]);
    }
    for (int n = 0; n <= n_tiles; ++n) {
      // (a) composite the tile that just finished (tile n-1): its raw is in raw[(n-1)&1]
      if (n >= 1) {
        const int m = n - 1;
        const TileId id = tile_info<G>(m);
        mbar_wait_relaxed(&sm.raw_full[m & 1], (m >> 1) & 1);
        {  // the two epilogue groups each hold the head sums over their half of the columns
          const float4 a = sm.raw[m & 1][0][wt], b = sm.raw[m & 1][1][wt];
          sm.raw[m & 1][0][wt] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
        }
        named_bar_sync(bar_id, kGroup);
        frontend_composite<G>(sm, p, sm.pair[id.q % 3], id, sm.raw[m & 1][0], wl, lane);
        mbar_arrive(&sm.raw_free[m & 1]);
        named_bar_sync(bar_id, kGroup);  // zf / carry of the pair visible to the whole warpgroup
      }
      // (b) prepare tile n+1 while tile n runs on the tensor core
      if (n + 1 < n_tiles) {
        const TileId id = tile_info<G>(n + 1);
        PairData<G>& pd = sm.pair[id.q % 3];
        if (!id.fine && id.t == 0) {  // first tile of a new pair
          const long long gp = (long long)id.q * gridDim.x + blockIdx.x;
          frontend_load_pair<G, kSplit>(p, img, pd, gp, id.q < T && gp < n_pairs, wt, bar_id);
          named_bar_sync(bar_id, kGroup);
        }
        mbar_wait_relaxed(&sm.tile_started, n & 1);  // tile n has started => tile n-1 no longer reads enc[(n+1)&1]
        frontend_encode<G, kOp>(pd, id, sm.enc[(n + 1) & 1][0], wt, kSave ? act_row<G>(p, id, pd.pair_index, 0, wt) : nullptr);
        fence_proxy_async();
        mbar_arrive(&sm.enc_full[(n + 1) & 1]);
      }
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1)  // no CTA may exit while a peer can still multicast into its ring / arrive on its barriers
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------------------------
// host launchers (shared by the translation units that instantiate the kernel)
// ------------------------------------------------------------------------------------
template <int kCluster, class G, int kOp, bool kSave = false, bool kPair = false, int kCoarseD = 8>
static int launch_bf16_render_t(const RenderParams& p, long long grid, int T, cudaStream_t stream) {
  using Smem = BfSmemT<G, RingFor<G, kOp == OP_F16X3>::value, kOp == OP_F16X3>;
  const size_t smem = sizeof(Smem);
  auto kern = snerf_bf16_render_kernel<kCluster, G, kOp, kSave, kPair, kCoarseD>;
  if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                 "cudaFuncSetAttribute(bf16 kernel smem)"))
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
  return check_cuda(cudaLaunchKernelEx(&cfg, kern, p, T), "launch snerf_bf16_render_kernel");
}

template <class G, int kOp, bool kSave = false, int kCoarseD = 8>
static int launch_bf16_render_gf(const RenderParams& p, cudaStream_t stream) {
  const long long n_pairs = (p.n_rays + 1) / 2;
  long long grid = n_pairs < (long long)sm_count() ? n_pairs : (long long)sm_count();
  // 2-CTA clusters share every weight chunk through a multicast bulk copy (one L2 read per cluster)
  static const int cluster_env = [] { const char* e = getenv("SNERF_B200_CLUSTER"); return e ? atoi(e) : 2; }();
  const bool use_cluster = cluster_env == 2 && grid >= 2;
  if (use_cluster) grid &= ~1ll;
  const int T = (int)((n_pairs + grid - 1) / grid);
  // SNERF_B200_PAIR=1: the two CTAs of a cluster as one tcgen05 cta_group::2 unit (inference, single-pass operands)
  static const int pair_env = [] { const char* e = getenv("SNERF_B200_PAIR"); return e ? atoi(e) : 0; }();
  if constexpr (kOp != OP_F16X3 && !kSave && kCoarseD == 8) {
    if (use_cluster && pair_env == 1) return launch_bf16_render_t<2, G, kOp, kSave, true>(p, grid, T, stream);
  }
  return use_cluster ? launch_bf16_render_t<2, G, kOp, kSave, false, kCoarseD>(p, grid, T, stream)
                     : launch_bf16_render_t<1, G, kOp, kSave, false, kCoarseD>(p, grid, T, stream);
}

// every sample geometry the tensor-core kernel is instantiated for
template <int kOp, bool kSave = false, int kCoarseD = 8>
static int launch_tc_render_op(const RenderParams& p, cudaStream_t stream) {
  if (p.Nc == 64 && p.Nf == 128) return launch_bf16_render_gf<Geo<64, 128>, kOp, kSave, kCoarseD>(p, stream);
  if (p.Nc == 64 && p.Nf == 0) return launch_bf16_render_gf<Geo<64, 0>, kOp, kSave, kCoarseD>(p, stream);
  if (p.Nc == 64 && p.Nf == 64) return launch_bf16_render_gf<Geo<64, 64>, kOp, kSave, kCoarseD>(p, stream);
  if (p.Nc == 64 && p.Nf == 192) return launch_bf16_render_gf<Geo<64, 192>, kOp, kSave, kCoarseD>(p, stream);
  if (p.Nc == 128 && p.Nf == 0) return launch_bf16_render_gf<Geo<128, 0>, kOp, kSave, kCoarseD>(p, stream);
  if (p.Nc == 128 && p.Nf == 128) return launch_bf16_render_gf<Geo<128, 128>, kOp, kSave, kCoarseD>(p, stream);
  set_error("tensor-core modes: (N_samples, N_importance) = (%d, %d) is not instantiated", p.Nc, p.Nf);
  return SNERF_ERR_UNSUPPORTED;
}

}  // namespace snerf
