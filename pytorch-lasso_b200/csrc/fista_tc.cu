// tcgen05 / TMEM FISTA step (placeholder until the tensor-core kernel lands).
#include "common.cuh"

namespace lasso {

bool fista_tc_supported(int64_t, int, int) { return false; }

int fista_tc_run(const FistaArgs&, float*, cudaStream_t) {
  set_error("tcgen05 path not built");
  return LASSO_B200_ERR_UNSUPPORTED;
}

}  // namespace lasso
