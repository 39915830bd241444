#include "lemon_common.cuh"
extern "C" int lemon_knn_candidates(lemon_ctx* ctx, const void* q16, const void* db16, int64_t nq, int64_t m,
                                    int d16, int nseg, int cta_group, float* cand_val, int32_t* cand_idx,
                                    void* stream) {
  return lemon_set_error(ctx, LEMON_ERR_UNSUPPORTED, "tensor-core kernel not built yet");
}
