// snerf_bf16.cu -- throughput renderer: render_rays as ONE persistent sm_100a kernel with the MLP on tcgen05
// tensor cores.  This translation unit instantiates the kernel template of snerf_tc_kernel.cuh for the single-pass
// operand types (bf16, fp16); snerf_x3.cu instantiates the three-pass fp32-class variant.
#include "snerf_tc_kernel.cuh"

namespace snerf {

// ------------------------------------------------------------------------------------
// bring-up self test: 128x128x64 through the production descriptors / swizzle / TMEM path.
// variant 0: A from shared memory (SS);  variant 1: A staged into TMEM with tcgen05.st, then the TS form.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) snerf_selftest_umma_kernel(const float* __restrict__ a,
                                                                     const float* __restrict__ b,
                                                                     float* __restrict__ d, int variant) {
  __shared__ alignas(1024) uint8_t sa[kBfChunkBytes];
  __shared__ alignas(1024) uint8_t sb[kBfChunkBytes];
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&done, 1); mbar_fence_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  // row `tid` of A and of B -> bf16
  uint32_t arow[32];
#pragma unroll
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
    arow[q8 * 4 + 0] = va.x; arow[q8 * 4 + 1] = va.y; arow[q8 * 4 + 2] = va.z; arow[q8 * 4 + 3] = va.w;
  }
  if (variant == 1) {  // A operand into TMEM columns 128..159 (64 bf16 = 32 columns), as the epilogue does
    uint32_t lo[16], hi[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { lo[i] = arow[i]; hi[i] = arow[16 + i]; }
    tmem_st16(tmem_base + lane_base + 128, lo);
    tmem_st16(tmem_base + lane_base + 144, hi);
    tmem_st_wait();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (variant == 0)
        tc_mma_ss(tmem_base, umma_desc_sw128(smem_u32(sa) + kk * 32), umma_desc_sw128(smem_u32(sb) + kk * 32), idesc,
                  kk != 0 ? 1u : 0u);
      else
        tc_mma_ts(tmem_base, tmem_base + 128 + kk * 8, umma_desc_sw128(smem_u32(sb) + kk * 32), idesc, kk != 0 ? 1u : 0u);
    }
    tc_commit(&done);
  }
  mbar_wait(&done, 0);
  tc_fence_after();
  for (int j = 0; j < 4; ++j) {
    uint32_t v[32];
    tmem_ld32(tmem_base + lane_base + j * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) d[(warp * 32 + lane) * 128 + j * 32 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
  }
}

// sample counts the tensor-core kernel is instantiated for (N_samples, N_importance)
bool bf16_geometry_supported(int nc, int nf) {
  return (nc == 64 && (nf == 0 || nf == 64 || nf == 128 || nf == 192)) || (nc == 128 && (nf == 0 || nf == 128));
}

int launch_bf16_render(const RenderParams& p, cudaStream_t stream) {
  if (p.n_rays <= 0) return SNERF_OK;
  if (p.coarse_depth == 4) return p.tc_op == OP_F16X3 ? launch_x3_render_d4(p, stream) : launch_bf16_render_d4(p, stream);
  if (p.tc_op == OP_F16X3) return launch_x3_render(p, stream);
  return p.tc_op == OP_F16 ? launch_tc_render_op<OP_F16>(p, stream) : launch_tc_render_op<OP_BF16>(p, stream);
}

int launch_bf16_query(const RenderParams&, cudaStream_t) {
  set_error("network_query_fn in the tensor-core modes is not available on its own; use mode fp32 or the fused renderer");
  return SNERF_ERR_UNSUPPORTED;
}

int launch_selftest_umma(const float* a, const float* b, float* d, int variant, cudaStream_t stream) {
  snerf_selftest_umma_kernel<<<1, 128, 0, stream>>>(a, b, d, variant);
  return check_cuda(cudaGetLastError(), "launch snerf_selftest_umma_kernel");
}

}  // namespace snerf
