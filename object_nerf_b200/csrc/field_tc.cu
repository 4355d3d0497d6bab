// placeholder until the tcgen05 kernel lands (next commit)
#include "field_common.cuh"
int onerf_launch_field_bf16(onerf_ctx* ctx, const FieldParams& p, cudaStream_t stream) {
  (void)ctx; (void)p; (void)stream;
  onerf_set_error("onerf_field_fwd: ONERF_PREC_BF16 not built yet");
  return ONERF_ERR_UNSUPPORTED;
}
