// snerf_x3.cu -- fp32-class tensor-core renderer (SNERF_MODE_FP16X3): the kernel template of snerf_tc_kernel.cuh with
// every MLP operand split into fp16 hi + fp16 lo parts and three tcgen05.mma passes per k-block
// (hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM; the dropped lo*lo term is 2^-22 relative).  Same pipeline,
// same fp32 sampling / compositing code as the bf16 build; meets the 1e-4 parity bar of the FFMA mode on the
// tensor cores.
#include "snerf_tc_kernel.cuh"

namespace snerf {

int launch_x3_render(const RenderParams& p, cudaStream_t stream) {
  if (p.n_rays <= 0) return SNERF_OK;
  return launch_tc_render_op<OP_F16X3>(p, stream);
}

}  // namespace snerf
