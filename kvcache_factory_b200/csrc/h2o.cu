// H2O all-query scoring: placeholder (fails loudly) until the kernel lands.
#include "common.cuh"
namespace kvc {
size_t h2o_scores_workspace_bytes(int, int, int, int, int, int) { return 256; }
int h2o_scores(const void*, int64_t, int64_t, int64_t, const void*, int64_t, int64_t, int64_t, int, int, int, int,
               int, int, int, void*, void*, size_t, cudaStream_t) {
  return fail(KVC_ERR_UNSUPPORTED, "H2O scoring kernel not built");
}
}  // namespace kvc
