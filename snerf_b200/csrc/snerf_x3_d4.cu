// snerf_x3_d4.cu -- SNERF_MODE_FP16X3 (fp32-class, three fp16 passes) for a 4-layer coarse network; see snerf_bf16_d4.cu.
#include "snerf_tc_kernel.cuh"

namespace snerf {

int launch_x3_render_d4(const RenderParams& p, cudaStream_t stream) {
  if (p.n_rays <= 0) return SNERF_OK;
  return launch_tc_render_op<OP_F16X3, false, 4>(p, stream);
}

}  // namespace snerf
