// tcgen05 (5th-gen tensor core) implicit-GEMM path -- placeholder until the kernel lands.
#include "common.cuh"
namespace b2s {
bool conv_tc_supported(int, int, int) { return false; }
int conv_table_tc(const float*, const float*, const int32_t*, float*, int64_t, int, int, int, int, int, int, cudaStream_t) {
  set_error("tcgen05 path not built");
  return B2S_E_INVALID;
}
}  // namespace b2s
