// tcgen05/TMA window-scoring kernel: placeholder until the kernel lands (reports "unsupported" so
// KVC_IMPL_AUTO resolves to the mma.sync implementation).
#include "common.cuh"
namespace kvc {
size_t window_scores_tc_workspace_bytes(int, int, int, int, int, int) { return 0; }
bool window_scores_tc_supported(int, int, int, int, int, int) { return false; }
int window_scores_tc(const void*, int64_t, int64_t, int64_t, const void*, int64_t, int64_t, int64_t, int, int, int,
                     int, int, int, int, int, void*, void*, size_t, cudaStream_t) {
  return fail(KVC_ERR_UNSUPPORTED, "tcgen05 scoring kernel not built");
}
}  // namespace kvc
