// tcgen05 / TMEM / TMA pointwise GEMM (placeholder until the tensor-core path lands).
#pragma once
#include "common.cuh"
#include "layers.h"

namespace mc {
struct PwTcPlan {};
inline int pw_tc_build(PwTcPlan** out, const NetCfg&, const float*, int, int, int) {
  *out = nullptr;
  return MC_OK;
}
inline void pw_tc_free(PwTcPlan*) {}
inline bool pw_tc_has(const PwTcPlan*, int) { return false; }
inline int pw_tc_run(PwTcPlan*, int, const void*, const float*, const void*, void*, int64_t, int, cudaStream_t) {
  return fail(MC_ERR_UNSUPPORTED, "tcgen05 path not built");
}
}  // namespace mc
