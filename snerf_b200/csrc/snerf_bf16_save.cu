// snerf_bf16_save.cu -- training forward of the tensor-core modes: the fused renderer of snerf_tc_kernel.cuh instantiated
// with kSave = true (every layer's 16-bit output also goes to the activation store the backward kernels of
// snerf_train_tc.cu read).  Its own translation unit so the inference instantiations do not pay for it at build time.
#include "snerf_tc_kernel.cuh"
#include "snerf_train_tc.h"

namespace snerf {

int launch_tc_render_save(const RenderParams& p, cudaStream_t stream) {
  if (p.n_rays <= 0) return SNERF_OK;
  if (p.tc_op != OP_BF16) { set_error("the training forward stores bf16 activations (the weight-gradient GEMM multiplies them with bf16 gradients)"); return SNERF_ERR_UNSUPPORTED; }
  return launch_tc_render_op<OP_BF16, true>(p, stream);
}

}  // namespace snerf
