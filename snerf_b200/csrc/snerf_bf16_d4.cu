// snerf_bf16_d4.cu -- the tensor-core renderer for a 4-layer coarse network (create_nerf with netdepth = 4 against
// netdepth_fine = 8, as the shipped configs set them; render.py:176-201): the kernel template of snerf_tc_kernel.cuh with
// kCoarseD = 4, bf16 and fp16 operands.  Its own translation unit so that it compiles beside the 8-layer instantiations.
#include "snerf_tc_kernel.cuh"

namespace snerf {

int launch_bf16_render_d4(const RenderParams& p, cudaStream_t stream) {
  if (p.n_rays <= 0) return SNERF_OK;
  return p.tc_op == OP_F16 ? launch_tc_render_op<OP_F16, false, 4>(p, stream) : launch_tc_render_op<OP_BF16, false, 4>(p, stream);
}

}  // namespace snerf
