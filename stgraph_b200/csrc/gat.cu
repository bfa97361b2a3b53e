// placeholder until the fused edge-softmax kernels land (replaced below in this round)
#include "common.cuh"
using namespace stg;
STG_API int stg_gat_softmax_fwd_f32(const StgCsrView*, const float*, const float*, const float*, int32_t, int32_t, float,
                                    float*, float*, float*, void*) {
  set_error("stg_gat_softmax_fwd_f32 not built yet");
  return STG_ERR_UNSUPPORTED;
}
STG_API int stg_gat_softmax_bwd_f32(const StgCsrView*, const StgCsrView*, const float*, const float*, const float*,
                                    const float*, const float*, const float*, const float*, int32_t, int32_t, float,
                                    float*, float*, float*, float*, void*) {
  set_error("stg_gat_softmax_bwd_f32 not built yet");
  return STG_ERR_UNSUPPORTED;
}
