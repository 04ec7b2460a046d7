// snerf_fp32_core.cuh -- the FFMA tile GEMM shared by the fp32 forward kernel (snerf_fp32.cu) and the
// training kernels (snerf_train.cu).
//
// One tile = 64 rows.  Activations live transposed in shared memory ([channel][row], row pitch kLd);
// each of the 8 compute warps owns 8 rows x all output columns, lane l owns columns l + 32 j.
// Weights arrive as [K][n_out] fp32 (K = contraction index) in chunks of kFp32ChunkRows K-rows through a
// kStages-deep ring filled by one producer thread with bulk async copies and mbarriers.
#pragma once
#include "snerf_common.cuh"
#include "snerf_packed.h"

namespace snerf {

constexpr int kTileRows = 64;
constexpr int kLd = 68;  // row pitch (floats) of the transposed activation buffers
constexpr int kStages = 3;
constexpr int kComputeThreads = 256;
constexpr int kFp32Threads = kComputeThreads + 32;

struct Fp32Ring {
  float* wstage;      // kStages x stage_floats
  uint64_t* full;     // [kStages], count 1 (+ tx bytes)
  uint64_t* empty;    // [kStages], count kComputeThreads / 32
  int stage_floats;
};

__device__ __forceinline__ void ring_init(const Fp32Ring& rg) {
  for (int s = 0; s < kStages; ++s) { mbar_init(&rg.full[s], 1); mbar_init(&rg.empty[s], kComputeThreads / 32); }
  mbar_fence_init();
}

// producer (one thread): stream a [rows][n_out] fp32 matrix, kFp32ChunkRows rows per stage
__device__ __forceinline__ void ring_stream(const Fp32Ring& rg, const float* src, int rows, int n_out, int& stage,
                                            uint32_t& phase) {
  const uint32_t bytes = kFp32ChunkRows * n_out * 4;
  for (int kc = 0; kc < rows; kc += kFp32ChunkRows) {
    mbar_wait(&rg.empty[stage], phase ^ 1);
    mbar_arrive_expect_tx(&rg.full[stage], bytes);
    bulk_g2s(rg.wstage + (size_t)stage * rg.stage_floats, src, bytes, &rg.full[stage]);
    src += kFp32ChunkRows * n_out;
    if (++stage == kStages) { stage = 0; phase ^= 1; }
  }
}

// consumer: acc[i][j] += sum_k a[k][r0 + i] * Wt[k][lane + 32 j] over `rows` K-rows of one segment.
// `abase` points at row r0 of channel 0 of the segment's activation buffer.
template <int NJ>
__device__ __forceinline__ void ring_gemm(const Fp32Ring& rg, const float* abase, int rows, float (&acc)[8][NJ],
                                          int& stage, uint32_t& phase, int lane) {
  constexpr int n_out = NJ * 32;
  for (int kc = 0; kc < rows; kc += kFp32ChunkRows) {
    mbar_wait(&rg.full[stage], phase);
    const float* ws = rg.wstage + (size_t)stage * rg.stage_floats + lane;
#pragma unroll
    for (int kk = 0; kk < kFp32ChunkRows; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(abase + (kc + kk) * kLd);
      const float4 a1 = *reinterpret_cast<const float4*>(abase + (kc + kk) * kLd + 4);
      float w[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) w[j] = ws[kk * n_out + 32 * j];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        acc[0][j] = fmaf(a0.x, w[j], acc[0][j]);
        acc[1][j] = fmaf(a0.y, w[j], acc[1][j]);
        acc[2][j] = fmaf(a0.z, w[j], acc[2][j]);
        acc[3][j] = fmaf(a0.w, w[j], acc[3][j]);
        acc[4][j] = fmaf(a1.x, w[j], acc[4][j]);
        acc[5][j] = fmaf(a1.y, w[j], acc[5][j]);
        acc[6][j] = fmaf(a1.z, w[j], acc[6][j]);
        acc[7][j] = fmaf(a1.w, w[j], acc[7][j]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&rg.empty[stage]);
    if (++stage == kStages) { stage = 0; phase ^= 1; }
  }
}

}  // namespace snerf
