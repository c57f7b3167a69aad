// K1 / K3, TF32 path: tcgen05 tensor-core contractions with TMA-staged operands (sm_100a).
#include "common.cuh"

namespace pdephi {

bool tc_supported(const Plan*) { return false; }

int analysis_tc(const Plan*, const float*, float2*, long long, int, float2*, cudaStream_t) {
  return fail(PDEPHI_ERR_UNSUPPORTED, "analysis_tc: not built");
}
int synthesis_tc(const Plan*, const float2*, float*, const float*, const float*, const float*, const float*, long long,
                 long long, int, float2*, cudaStream_t) {
  return fail(PDEPHI_ERR_UNSUPPORTED, "synthesis_tc: not built");
}

}  // namespace pdephi
