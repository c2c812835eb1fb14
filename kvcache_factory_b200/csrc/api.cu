// extern "C" surface of libkvc_b200.so (declared in include/kvc_b200.h): argument validation,
// device check, dispatch to the sm_100a kernels.  No torch types cross this boundary.
#include <algorithm>
#include <atomic>
#include <vector>

#include "common.cuh"

namespace kvc {

// kernels implemented in the other translation units
int gather_dense(const void*, int64_t, int64_t, int64_t, const void*, int64_t, int64_t, int64_t, const int32_t*, int,
                 int, int, int, int, int, int, int, void*, void*, cudaStream_t);
int gather_varlen(const void*, int64_t, int64_t, const void*, int64_t, int64_t, const int32_t*, int64_t,
                  const int32_t*, const int32_t*, int, int, int, int, int, int, int, void*, void*, cudaStream_t);
int key_norm_scores(const void*, int64_t, int64_t, int64_t, int, int, int, int, int, int, void*, cudaStream_t);
int append_rows(void*, void*, const void*, const void*, const int32_t*, const int32_t*, int, int, int, cudaStream_t);
int flatten_regrow(void*, const void*, const int32_t*, const int32_t*, const int32_t*, int, int, int, int64_t, cudaStream_t);
int update_flatten_view(void*, const void*, const void*, const int32_t*, const int32_t*, int, int, int, int64_t,
                        cudaStream_t);
size_t pool_topk_workspace_bytes(int rows, int n, int k);
int pool_topk(const void*, int, int, int, int, int, int, const int32_t*, int32_t*, int64_t, void*, float*, void*,
              size_t, cudaStream_t);
size_t window_scores_mma_workspace_bytes(int B, int Hq, int Hkv, int L, int d, int w);
int window_scores_mma(const void*, int64_t, int64_t, int64_t, const void*, int64_t, int64_t, int64_t, int, int, int,
                      int, int, int, int, int, void*, void*, size_t, cudaStream_t);
size_t window_scores_tc_workspace_bytes(int B, int Hq, int Hkv, int L, int d, int w);
bool window_scores_tc_supported(int B, int Hq, int Hkv, int L, int d, int w);
bool window_scores_dyn_supported(int B, int Hq, int Hkv, int L, int d, int w);
size_t window_scores_dyn_workspace_bytes(int B, int Hq, int Hkv, int L, int d, int w);
int window_scores_dyn(const void*, int64_t, int64_t, int64_t, const void*, int64_t, int64_t, int64_t, int, int, int,
                      int, int, int, int, int, void*, void*, size_t, const float*, cudaStream_t);
bool window_scores_fused_supported(int B, int Hq, int Hkv, int L, int d, int w);
bool window_scores_split_preferred(int B, int Hq, int Hkv, int L, int d, int w);
size_t window_scores_fused_workspace_bytes(int B, int Hq, int Hkv, int L, int d, int w);
int window_scores_fused(const void*, int64_t, int64_t, int64_t, const void*, int64_t, int64_t, int64_t, int, int, int,
                        int, int, int, int, int, void*, void*, size_t, int, cudaStream_t);
int window_scores_tc(const void*, int64_t, int64_t, int64_t, const void*, int64_t, int64_t, int64_t, int, int, int,
                     int, int, int, int, int, void*, void*, size_t, cudaStream_t);
size_t h2o_scores_workspace_bytes(int B, int Hq, int Hkv, int L, int d, int w);
int h2o_scores(const void*, int64_t, int64_t, int64_t, const void*, int64_t, int64_t, int64_t, int, int, int, int,
               int, int, int, void*, void*, size_t, cudaStream_t);
size_t adakv_allocate_workspace_bytes(int H, int n);
int adakv_key_histogram(const void*, const float*, int, int, int, int, int32_t*, cudaStream_t);
size_t think_prune_workspace_bytes(int BH, int rows, int d);
int think_prune(const void*, const void*, int64_t, int64_t, int64_t, int, int, int, int, int, int, int, int, void*, uint8_t*, void*,
                size_t, cudaStream_t);
int adakv_hist_coarse(const void*, const float*, int, int, int, int, uint16_t*, int32_t*, cudaStream_t);
int adakv_hist_fine(const uint16_t*, const int32_t*, int, int, const int32_t*, int32_t*, int32_t*, cudaStream_t);
int adakv_allocate(const void*, const float*, int, int, int, int, float, int, int, int, int32_t*, int32_t*, int32_t*,
                   void*, size_t, cudaStream_t);

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

#ifdef KVC_TRACE
static unsigned long long* g_tl_dev = nullptr;
static int g_tl_kind[4096];
static std::atomic<int> g_tl_next{0};
unsigned long long* timeline_slot(int kind) {
  if (!g_tl_dev) {
    cudaMalloc(&g_tl_dev, 4096 * 16);
    cudaMemset(g_tl_dev, 0xff, 4096 * 16);
  }
  const int i = g_tl_next.fetch_add(1) % 4096;
  g_tl_kind[i] = kind;
  return g_tl_dev + 2 * i;
}
#endif

static thread_local char g_err[512] = "";
char* last_error_buffer() { return g_err; }

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// The library carries sm_100a SASS only: refuse anything else loudly instead of failing at launch.
static int check_device() {
  static int ok[64] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(KVC_ERR_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  if (dev >= 0 && dev < 64 && ok[dev]) return 0;
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) return fail(KVC_ERR_DEVICE, "device %d is sm_%d%d; libkvc_b200 is built for sm_100a only", dev, major, minor);
  if (dev >= 0 && dev < 64) ok[dev] = 1;
  return 0;
}

static int check_shape(int B, int Hq, int Hkv, int L, int d, int w, int dtype) {
  KVC_REQUIRE(B > 0 && Hq > 0 && Hkv > 0 && L > 0 && d > 0 && w > 0, KVC_ERR_ARG, "non-positive dimension");
  KVC_REQUIRE(Hq % Hkv == 0, KVC_ERR_ARG, "Hq=%d is not a multiple of Hkv=%d", Hq, Hkv);
  KVC_REQUIRE(w < L, KVC_ERR_ARG, "window %d must be smaller than the prompt length %d", w, L);
  KVC_REQUIRE(dtype == KVC_BF16 || dtype == KVC_FP16, KVC_ERR_UNSUPPORTED,
              "dtype %d: only bf16 and fp16 are implemented (as in the reference kernel)", dtype);
  return 0;
}

}  // namespace kvc

using namespace kvc;

extern "C" {

#ifdef KVC_TRACE
// out[3*i] = kind (0 scoring, 1 select, 2 gather), out[3*i+1] = start ns, out[3*i+2] = end ns; returns #entries
int kvc_debug_timeline(long long* out, int max_entries, int reset) {
  const int n = std::min(std::min(g_tl_next.load(), 4096), max_entries);
  if (n > 0 && g_tl_dev) {
    std::vector<unsigned long long> h(2 * n);
    cudaMemcpy(h.data(), g_tl_dev, 16 * (size_t)n, cudaMemcpyDeviceToHost);
    for (int i = 0; i < n; ++i) { out[3 * i] = g_tl_kind[i]; out[3 * i + 1] = (long long)h[2 * i]; out[3 * i + 2] = (long long)~h[2 * i + 1]; }
  }
  if (reset && g_tl_dev) { cudaMemset(g_tl_dev, 0xff, 4096 * 16); }
  return n;
}
#endif

int kvc_version(void) { return KVC_ABI_VERSION; }
const char* kvc_last_error(void) { return last_error_buffer(); }
uint64_t kvc_launch_count(void) { return (uint64_t)launch_count(); }

int kvc_sm_count(int* out_sms) {
  KVC_REQUIRE(out_sms != nullptr, KVC_ERR_ARG, "out_sms is null");
  if (int rc = check_device()) return rc;
  *out_sms = sm_count();
  return 0;
}

size_t kvc_window_scores_workspace_bytes(int B, int Hq, int Hkv, int L, int d, int w) {
  if (B <= 0 || Hq <= 0 || Hkv <= 0 || L <= 0 || w <= 0 || Hq % Hkv) return 0;
  size_t a = window_scores_mma_workspace_bytes(B, Hq, Hkv, L, d, w);
  size_t b = window_scores_tc_workspace_bytes(B, Hq, Hkv, L, d, w);
  size_t c = window_scores_fused_workspace_bytes(B, Hq, Hkv, L, d, w);
  size_t e = window_scores_dyn_workspace_bytes(B, Hq, Hkv, L, d, w);
  if (b > a) a = b;
  if (e > a) a = e;
  return a > c ? a : c;
}

int kvc_window_scores(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const void* k, int64_t k_sb,
                      int64_t k_sh, int64_t k_sl, int B, int Hq, int Hkv, int L, int d, int w, int dtype, int reduce,
                      void* out_scores, void* workspace, size_t workspace_bytes, int impl, void* stream) {
  if (int rc = check_device()) return rc;
  if (int rc = check_shape(B, Hq, Hkv, L, d, w, dtype)) return rc;
  KVC_REQUIRE(q && k && out_scores && workspace, KVC_ERR_ARG, "null pointer argument");
  KVC_REQUIRE(reduce == KVC_REDUCE_SUM || reduce == KVC_REDUCE_MEAN, KVC_ERR_ARG, "bad reduce %d", reduce);
  KVC_REQUIRE(impl >= KVC_IMPL_AUTO && impl <= KVC_IMPL_TCGEN05_DYN, KVC_ERR_ARG, "bad impl %d", impl);
  const bool tc_ok = window_scores_tc_supported(B, Hq, Hkv, L, d, w);
  const bool fused_ok = tc_ok && window_scores_fused_supported(B, Hq, Hkv, L, d, w);
  const bool dyn_ok = tc_ok && window_scores_dyn_supported(B, Hq, Hkv, L, d, w);
  // query/key rows that are not 16-byte aligned: only the mma.sync kernel takes them (IMPL_AUTO falls through)
  const bool aligned = ((q_sb | q_sh | q_sl | k_sb | k_sh | k_sl) & 7) == 0 &&
                       ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k)) & 15) == 0;
  if (impl == KVC_IMPL_TCGEN05_DYN && !dyn_ok)
    return fail(KVC_ERR_UNSUPPORTED, "tcgen05 scoring kernel does not cover this shape (Hq/Hkv=%d, w=%d, d=%d, units=%d)",
                Hq / Hkv, w, d, B * Hkv);
  if ((impl == KVC_IMPL_TCGEN05 && !tc_ok) || ((impl == KVC_IMPL_TCGEN05_FUSED || impl == KVC_IMPL_TCGEN05_HALF) && !fused_ok))
    return fail(KVC_ERR_UNSUPPORTED, "tcgen05 scoring kernel does not cover this shape (Hq/Hkv=%d, w=%d, d=%d)",
                Hq / Hkv, w, d);
  // IMPL_AUTO: the chunked single-launch kernel where its grid fits the SMs (fastest alone: 35 vs 38 us per cfg-2
  // layer), the dynamically scheduled kernel for more units than SMs, then the two-launch tcgen05 kernel, then mma.sync
  if (impl == KVC_IMPL_TCGEN05_FUSED || impl == KVC_IMPL_TCGEN05_HALF || (impl == KVC_IMPL_AUTO && fused_ok && aligned))
    return window_scores_fused(q, q_sb, q_sh, q_sl, k, k_sb, k_sh, k_sl, B, Hq, Hkv, L, d, w, dtype, reduce, out_scores,
                               workspace, workspace_bytes,
                               impl == KVC_IMPL_TCGEN05_HALF ? 1 : (impl == KVC_IMPL_AUTO && window_scores_split_preferred(B, Hq, Hkv, L, d, w) ? 2 : 0),
                               (cudaStream_t)stream);
  if (impl == KVC_IMPL_TCGEN05_DYN || (impl == KVC_IMPL_AUTO && dyn_ok && aligned))
    return window_scores_dyn(q, q_sb, q_sh, q_sl, k, k_sb, k_sh, k_sl, B, Hq, Hkv, L, d, w, dtype, reduce, out_scores,
                             workspace, workspace_bytes, nullptr, (cudaStream_t)stream);
  if (impl == KVC_IMPL_TCGEN05 || (impl == KVC_IMPL_AUTO && tc_ok && aligned))
    return window_scores_tc(q, q_sb, q_sh, q_sl, k, k_sb, k_sh, k_sl, B, Hq, Hkv, L, d, w, dtype, reduce, out_scores,
                            workspace, workspace_bytes, (cudaStream_t)stream);
  return window_scores_mma(q, q_sb, q_sh, q_sl, k, k_sb, k_sh, k_sl, B, Hq, Hkv, L, d, w, dtype, reduce, out_scores,
                           workspace, workspace_bytes, (cudaStream_t)stream);
}

int kvc_window_scores_lse(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const void* k, int64_t k_sb,
                          int64_t k_sh, int64_t k_sl, const float* lse, int B, int Hq, int Hkv, int L, int d, int w, int dtype,
                          int reduce, void* out_scores, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_device()) return rc;
  if (int rc = check_shape(B, Hq, Hkv, L, d, w, dtype)) return rc;
  KVC_REQUIRE(q && k && lse && out_scores && workspace, KVC_ERR_ARG, "null pointer argument");
  KVC_REQUIRE(reduce == KVC_REDUCE_SUM || reduce == KVC_REDUCE_MEAN, KVC_ERR_ARG, "bad reduce %d", reduce);
  KVC_REQUIRE(window_scores_dyn_supported(B, Hq, Hkv, L, d, w), KVC_ERR_UNSUPPORTED,
              "scores-from-LSE is implemented by the dynamic tcgen05 kernel only (d = 128, w in {8,16,32}, Hq/Hkv*w %% 32 == 0)");
  return window_scores_dyn(q, q_sb, q_sh, q_sl, k, k_sb, k_sh, k_sl, B, Hq, Hkv, L, d, w, dtype, reduce, out_scores,
                           workspace, workspace_bytes, lse, (cudaStream_t)stream);
}

size_t kvc_h2o_scores_workspace_bytes(int B, int Hq, int Hkv, int L, int d, int w) {
  if (B <= 0 || Hq <= 0 || Hkv <= 0 || L <= 0 || w <= 0 || Hq % Hkv) return 0;
  return h2o_scores_workspace_bytes(B, Hq, Hkv, L, d, w);
}

int kvc_h2o_scores(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const void* k, int64_t k_sb, int64_t k_sh,
                   int64_t k_sl, int B, int Hq, int Hkv, int L, int d, int w, int dtype, void* out_scores,
                   void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_device()) return rc;
  if (int rc = check_shape(B, Hq, Hkv, L, d, w, dtype)) return rc;
  KVC_REQUIRE(q && k && out_scores && workspace, KVC_ERR_ARG, "null pointer argument");
  return h2o_scores(q, q_sb, q_sh, q_sl, k, k_sb, k_sh, k_sl, B, Hq, Hkv, L, d, w, dtype, out_scores, workspace,
                    workspace_bytes, (cudaStream_t)stream);
}

size_t kvc_pool_topk_workspace_bytes(int rows, int n, int k) {
  if (rows <= 0 || n <= 0 || k < 0) return 0;
  return pool_topk_workspace_bytes(rows, n, k);
}

int kvc_pool_topk(const void* scores, int rows, int n, int dtype, int kernel_size, int pool, int k,
                  const int32_t* k_per_row, int32_t* out_idx, int64_t idx_stride, void* out_pooled, float* out_sums,
                  void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_device()) return rc;
  KVC_REQUIRE(scores && workspace && (out_idx || k == 0 || out_sums), KVC_ERR_ARG, "null pointer argument");
  KVC_REQUIRE(rows > 0 && n > 0, KVC_ERR_ARG, "rows=%d n=%d must be positive", rows, n);
  KVC_REQUIRE(k >= 0 && k <= n, KVC_ERR_ARG, "k=%d outside [0, n=%d] (torch.topk raises for k > n)", k, n);
  KVC_REQUIRE(idx_stride >= k, KVC_ERR_ARG, "idx_stride %lld < k %d", (long long)idx_stride, k);
  KVC_REQUIRE(dtype == KVC_BF16 || dtype == KVC_FP16, KVC_ERR_UNSUPPORTED, "dtype %d", dtype);
  KVC_REQUIRE(pool == KVC_POOL_NONE || pool == KVC_POOL_AVG || pool == KVC_POOL_MAX, KVC_ERR_ARG,
              "Pooling method not supported (pool=%d)", pool);
  KVC_REQUIRE(pool == KVC_POOL_NONE || (kernel_size >= 1 && (kernel_size & 1)), KVC_ERR_UNSUPPORTED,
              "kernel_size %d: only odd kernels keep the sequence length (see DESIGN.md)", kernel_size);
  return pool_topk(scores, rows, n, dtype, kernel_size, pool, k, k_per_row, out_idx, idx_stride, out_pooled, out_sums,
                   workspace, workspace_bytes, (cudaStream_t)stream);
}

int kvc_gather_dense(const void* k, int64_t k_sb, int64_t k_sh, int64_t k_sl, const void* v, int64_t v_sb,
                     int64_t v_sh, int64_t v_sl, const int32_t* idx, int B, int Hq, int Hkv, int L, int d, int n_sel,
                     int w, int elem_bytes, void* out_k, void* out_v, void* stream) {
  if (int rc = check_device()) return rc;
  KVC_REQUIRE(k && v && out_k && out_v, KVC_ERR_ARG, "null pointer argument");
  KVC_REQUIRE(B > 0 && Hq > 0 && Hkv > 0 && Hq % Hkv == 0 && L > 0 && d > 0, KVC_ERR_ARG, "bad head geometry");
  KVC_REQUIRE(w >= 0 && n_sel >= 0 && n_sel + w > 0 && n_sel <= L - w, KVC_ERR_ARG,
              "n_sel=%d w=%d do not fit L=%d", n_sel, w, L);
  KVC_REQUIRE(elem_bytes == 2 || elem_bytes == 4, KVC_ERR_UNSUPPORTED, "elem_bytes %d", elem_bytes);
  return gather_dense(k, k_sb, k_sh, k_sl, v, v_sb, v_sh, v_sl, idx, B, Hq, Hkv, L, d, n_sel, w, elem_bytes, out_k,
                      out_v, (cudaStream_t)stream);
}

int kvc_key_norm_scores(const void* k, int64_t k_sb, int64_t k_sh, int64_t k_sl, int B, int Hq, int Hkv, int L, int d,
                        int dtype, void* out_scores, void* stream) {
  if (int rc = check_device()) return rc;
  KVC_REQUIRE(k && out_scores, KVC_ERR_ARG, "null pointer argument");
  KVC_REQUIRE(B > 0 && Hq > 0 && Hkv > 0 && Hq % Hkv == 0 && L > 0 && d > 0, KVC_ERR_ARG, "bad head geometry");
  KVC_REQUIRE(dtype == KVC_BF16 || dtype == KVC_FP16, KVC_ERR_UNSUPPORTED, "dtype %d", dtype);
  return key_norm_scores(k, k_sb, k_sh, k_sl, B, Hq, Hkv, L, d, dtype, out_scores, (cudaStream_t)stream);
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t kvc_compress_dense_workspace_bytes(int B, int Hq, int Hkv, int L, int d, int w, int n_keep) {
  if (B <= 0 || Hq <= 0 || Hkv <= 0 || L <= w || w <= 0 || Hq % Hkv || n_keep < 0) return 0;
  const size_t scores = align256((size_t)B * Hq * (L - w) * 2);
  const size_t idx = align256((size_t)B * Hq * (n_keep > 0 ? n_keep : 1) * 4);
  return scores + idx + align256(kvc_window_scores_workspace_bytes(B, Hq, Hkv, L, d, w)) +
         align256(pool_topk_workspace_bytes(B * Hq, L - w, n_keep)) + 256;
}

int kvc_compress_dense(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const void* k, int64_t k_sb,
                       int64_t k_sh, int64_t k_sl, const void* v, int64_t v_sb, int64_t v_sh, int64_t v_sl, int B, int Hq,
                       int Hkv, int L, int d, int w, int dtype, int kernel_size, int pool, int n_keep, int score_impl,
                       void* out_k, void* out_v, void* workspace, size_t workspace_bytes, size_t* idx_offset,
                       size_t* scores_offset, void* stream) {
  if (int rc = check_device()) return rc;
  if (int rc = check_shape(B, Hq, Hkv, L, d, w, dtype)) return rc;
  KVC_REQUIRE(k && v && out_k && out_v && workspace, KVC_ERR_ARG, "null pointer argument");
  KVC_REQUIRE(n_keep >= 0 && n_keep <= L - w, KVC_ERR_ARG, "n_keep=%d does not fit L-w=%d", n_keep, L - w);
  const bool scored = !(pool == KVC_POOL_NONE && kernel_size == 0);
  const int e = 2;
  if (!scored)   // StreamingLLM: first n_keep rows + window, no index tensor at all
    return kvc_gather_dense(k, k_sb, k_sh, k_sl, v, v_sb, v_sh, v_sl, nullptr, B, Hq, Hkv, L, d, n_keep, w, e, out_k, out_v, stream);
  KVC_REQUIRE(q != nullptr, KVC_ERR_ARG, "null query pointer");
  const size_t need = kvc_compress_dense_workspace_bytes(B, Hq, Hkv, L, d, w, n_keep);
  KVC_REQUIRE(workspace_bytes >= need, KVC_ERR_WORKSPACE, "compress_dense workspace: have %zu need %zu", workspace_bytes, need);
  char* ws = (char*)workspace;
  const size_t scores_bytes = align256((size_t)B * Hq * (L - w) * 2);
  const size_t idx_bytes = align256((size_t)B * Hq * (n_keep > 0 ? n_keep : 1) * 4);
  const size_t sc_ws = align256(kvc_window_scores_workspace_bytes(B, Hq, Hkv, L, d, w));
  const size_t sel_ws = align256(pool_topk_workspace_bytes(B * Hq, L - w, n_keep));
  char* scores = ws;
  int32_t* idx = (int32_t*)(ws + scores_bytes);
  char* ws_scores = ws + scores_bytes + idx_bytes;
  char* ws_select = ws_scores + sc_ws;
  if (idx_offset) *idx_offset = scores_bytes;
  if (scores_offset) *scores_offset = 0;
  if (int rc = kvc_window_scores(q, q_sb, q_sh, q_sl, k, k_sb, k_sh, k_sl, B, Hq, Hkv, L, d, w, dtype, KVC_REDUCE_SUM, scores,
                                 ws_scores, sc_ws > 256 ? sc_ws : 256, score_impl, stream))
    return rc;
  if (int rc = kvc_pool_topk(scores, B * Hq, L - w, dtype, kernel_size, pool, n_keep, nullptr, idx, n_keep, nullptr, nullptr,
                             ws_select, sel_ws + 256, stream))
    return rc;
  return kvc_gather_dense(k, k_sb, k_sh, k_sl, v, v_sb, v_sh, v_sl, idx, B, Hq, Hkv, L, d, n_keep, w, e, out_k, out_v, stream);
}

int kvc_gather_varlen(const void* k, int64_t k_sh, int64_t k_sl, const void* v, int64_t v_sh, int64_t v_sl,
                      const int32_t* idx, int64_t idx_stride, const int32_t* cap, const int32_t* cu_klen, int Hq,
                      int Hkv, int L, int d, int w, int max_cap, int elem_bytes, void* out_k, void* out_v,
                      void* stream) {
  if (int rc = check_device()) return rc;
  KVC_REQUIRE(k && v && idx && cap && cu_klen && out_k && out_v, KVC_ERR_ARG, "null pointer argument");
  KVC_REQUIRE(Hq > 0 && Hkv > 0 && Hq % Hkv == 0 && L > 0 && d > 0 && w >= 0 && max_cap >= 0, KVC_ERR_ARG,
              "bad head geometry");
  KVC_REQUIRE(elem_bytes == 2 || elem_bytes == 4, KVC_ERR_UNSUPPORTED, "elem_bytes %d", elem_bytes);
  return gather_varlen(k, k_sh, k_sl, v, v_sh, v_sl, idx, idx_stride, cap, cu_klen, Hq, Hkv, L, d, w, max_cap,
                       elem_bytes, out_k, out_v, (cudaStream_t)stream);
}

size_t kvc_adakv_allocate_workspace_bytes(int H, int n) {
  if (H <= 0 || n <= 0) return 0;
  return adakv_allocate_workspace_bytes(H, n);
}

int kvc_adakv_allocate(const void* pooled, const float* sums, int H, int n, int dtype, int base_capacity,
                       float one_minus_floor, int floor_capacity, int normalize, int w, int32_t* out_cap,
                       int32_t* out_head_lens, int32_t* out_cu_klen, void* workspace, size_t workspace_bytes,
                       void* stream) {
  if (int rc = check_device()) return rc;
  KVC_REQUIRE(pooled && sums && out_cap && out_head_lens && out_cu_klen && workspace, KVC_ERR_ARG, "null pointer");
  KVC_REQUIRE(H > 0 && n > 0 && base_capacity > 0 && base_capacity <= n, KVC_ERR_ARG,
              "H=%d n=%d base_capacity=%d", H, n, base_capacity);
  KVC_REQUIRE(dtype == KVC_BF16 || dtype == KVC_FP16, KVC_ERR_UNSUPPORTED, "dtype %d", dtype);
  return adakv_allocate(pooled, sums, H, n, dtype, base_capacity, one_minus_floor, floor_capacity, normalize, w,
                        out_cap, out_head_lens, out_cu_klen, workspace, workspace_bytes, (cudaStream_t)stream);
}

// ---- AdaKV prefill in one call -------------------------------------------------------------------------------------
// Workspace layout: scores | pooled | sums | indices | scoring ws | select ws | allocation ws (each 256-aligned).
struct AdaLayout { size_t scores, pooled, sums, idx, ws_scores, ws_select, ws_alloc, total, sc_ws, sel_ws, al_ws; };
static AdaLayout ada_layout(int Hq, int Hkv, int L, int d, int w, int max_cap_bound) {
  AdaLayout a{};
  const size_t n = (size_t)(L - w);
  size_t off = 0;
  a.scores = off; off += align256((size_t)Hq * n * 2);
  a.pooled = off; off += align256((size_t)Hq * n * 2);
  a.sums = off; off += align256((size_t)Hq * 2 * sizeof(float));
  a.idx = off; off += align256((size_t)Hq * (size_t)max_cap_bound * 4);
  a.sc_ws = align256(kvc_window_scores_workspace_bytes(1, Hq, Hkv, L, d, w)) + 256;
  a.sel_ws = align256(pool_topk_workspace_bytes(Hq, (int)n, max_cap_bound)) + 256;
  a.al_ws = align256(adakv_allocate_workspace_bytes(Hq, (int)n)) + 256;
  a.ws_scores = off; off += a.sc_ws;
  a.ws_select = off; off += a.sel_ws;
  a.ws_alloc = off; off += a.al_ws;
  a.total = off;
  return a;
}

size_t kvc_adakv_prefill_workspace_bytes(int Hq, int Hkv, int L, int d, int w, int max_cap_bound) {
  if (Hq <= 0 || Hkv <= 0 || Hq % Hkv || L <= w || w <= 0 || max_cap_bound <= 0) return 0;
  return ada_layout(Hq, Hkv, L, d, w, max_cap_bound).total;
}

int kvc_adakv_prefill(const void* q, int64_t q_sh, int64_t q_sl, const void* k, int64_t k_sh, int64_t k_sl, const void* v,
                      int64_t v_sh, int64_t v_sl, int Hq, int Hkv, int L, int d, int w, int dtype, int kernel_size,
                      int pool, int base_capacity, float one_minus_floor, int floor_capacity, int normalize,
                      int score_impl, void* out_k, void* out_v, int64_t out_rows, int32_t* out_cap,
                      int32_t* out_head_lens, int32_t* out_cu_klen, int32_t* host_cap, int max_cap_bound,
                      void* workspace, size_t workspace_bytes, int* out_max_cap, int64_t* out_klen_sum,
                      size_t* offsets, void* stream) {
  if (int rc = check_device()) return rc;
  if (int rc = check_shape(1, Hq, Hkv, L, d, w, dtype)) return rc;
  KVC_REQUIRE(q && k && v && out_k && out_v && out_cap && out_head_lens && out_cu_klen && host_cap && workspace &&
                  out_max_cap && out_klen_sum,
              KVC_ERR_ARG, "null pointer argument");
  const int n = L - w;
  KVC_REQUIRE(base_capacity > 0 && base_capacity <= n, KVC_ERR_ARG, "base_capacity=%d outside (0, L-w=%d]", base_capacity, n);
  KVC_REQUIRE(max_cap_bound >= base_capacity && max_cap_bound <= n, KVC_ERR_ARG, "max_cap_bound=%d outside [base_capacity, L-w]",
              max_cap_bound);
  KVC_REQUIRE(pool == KVC_POOL_AVG || pool == KVC_POOL_MAX, KVC_ERR_ARG, "Pooling method not supported (pool=%d)", pool);
  const AdaLayout a = ada_layout(Hq, Hkv, L, d, w, max_cap_bound);
  KVC_REQUIRE(workspace_bytes >= a.total, KVC_ERR_WORKSPACE, "adakv_prefill workspace: have %zu need %zu", workspace_bytes, a.total);
  char* ws = (char*)workspace;
  if (offsets) { offsets[0] = a.scores; offsets[1] = a.pooled; offsets[2] = a.idx; }
  float* sums = (float*)(ws + a.sums);
  int32_t* idx = (int32_t*)(ws + a.idx);
  cudaStream_t st = (cudaStream_t)stream;
  // window-mean scores -> pooled scores + (top-base, all) sums -> capacities (pyramidkv_utils.py:647-672, 706-719)
  if (int rc = kvc_window_scores(q, 0, q_sh, q_sl, k, 0, k_sh, k_sl, 1, Hq, Hkv, L, d, w, dtype, KVC_REDUCE_MEAN, ws + a.scores,
                                 ws + a.ws_scores, a.sc_ws, score_impl, stream))
    return rc;
  if (int rc = kvc_pool_topk(ws + a.scores, Hq, n, dtype, kernel_size, pool, base_capacity, nullptr, nullptr, base_capacity,
                             ws + a.pooled, sums, ws + a.ws_select, a.sel_ws, stream))
    return rc;
  if (int rc = kvc_adakv_allocate(ws + a.pooled, sums, Hq, n, dtype, base_capacity, one_minus_floor, floor_capacity, normalize, w,
                                  out_cap, out_head_lens, out_cu_klen, ws + a.ws_alloc, a.al_ws, stream))
    return rc;
  // the reference reads the capacities back here as well (:718); everything after it is sized by them.  (Queuing the
  // per-head select and the gather BEFORE the read, on the device-side capacities with 8192-wide index rows, was measured
  // twice -- from Python and from this function -- and is slower: 254 vs 226 us per cfg-4 layer.)
  KVC_CUDA_TRY(cudaMemcpyAsync(host_cap, out_cap, (size_t)Hq * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  KVC_CUDA_TRY(cudaStreamSynchronize(st));
  int max_cap = 0;
  int64_t klen_sum = (int64_t)Hq * w;
  for (int h = 0; h < Hq; ++h) { max_cap = host_cap[h] > max_cap ? host_cap[h] : max_cap; klen_sum += host_cap[h]; }
  *out_max_cap = max_cap;
  *out_klen_sum = klen_sum;
  KVC_REQUIRE(max_cap <= max_cap_bound, KVC_ERR_WORKSPACE, "largest head capacity %d exceeds max_cap_bound %d", max_cap, max_cap_bound);
  KVC_REQUIRE(klen_sum <= out_rows, KVC_ERR_WORKSPACE, "flattened cache needs %lld rows, output holds %lld", (long long)klen_sum,
              (long long)out_rows);
  if (max_cap > 0)
    if (int rc = kvc_pool_topk(ws + a.scores, Hq, n, dtype, kernel_size, pool, max_cap, out_cap, idx, max_cap, nullptr, nullptr,
                               ws + a.ws_select, a.sel_ws, stream))
      return rc;
  return kvc_gather_varlen(k, k_sh, k_sl, v, v_sh, v_sl, idx, max_cap, out_cap, out_cu_klen, Hq, Hkv, L, d, w, max_cap, 2, out_k,
                           out_v, stream);
}

int kvc_adakv_hist_coarse(const void* pooled, const float* sums, int H, int n, int dtype, int normalize, void* akeys,
                          int32_t* hist1, void* stream) {
  if (int rc = check_device()) return rc;
  KVC_REQUIRE(pooled && akeys && hist1 && (sums || !normalize), KVC_ERR_ARG, "null pointer");
  KVC_REQUIRE(H > 0 && n > 0, KVC_ERR_ARG, "H=%d n=%d", H, n);
  KVC_REQUIRE(dtype == KVC_BF16 || dtype == KVC_FP16, KVC_ERR_UNSUPPORTED, "dtype %d", dtype);
  return adakv_hist_coarse(pooled, sums, H, n, dtype, normalize, (uint16_t*)akeys, hist1, (cudaStream_t)stream);
}

int kvc_adakv_hist_fine(const void* akeys, const int32_t* hist1, int H, int n, const int32_t* thr_bin, int32_t* hist2,
                        int32_t* above, void* stream) {
  if (int rc = check_device()) return rc;
  KVC_REQUIRE(akeys && hist1 && thr_bin && hist2 && above, KVC_ERR_ARG, "null pointer");
  KVC_REQUIRE(H > 0 && n > 0, KVC_ERR_ARG, "H=%d n=%d", H, n);
  return adakv_hist_fine((const uint16_t*)akeys, hist1, H, n, thr_bin, hist2, above, (cudaStream_t)stream);
}

int kvc_adakv_key_histogram(const void* pooled, const float* sums, int H, int n, int dtype, int normalize,
                            int32_t* out_hist, void* stream) {
  if (int rc = check_device()) return rc;
  KVC_REQUIRE(pooled && out_hist && (sums || !normalize), KVC_ERR_ARG, "null pointer");
  KVC_REQUIRE(H > 0 && n > 0, KVC_ERR_ARG, "H=%d n=%d", H, n);
  KVC_REQUIRE(dtype == KVC_BF16 || dtype == KVC_FP16, KVC_ERR_UNSUPPORTED, "dtype %d", dtype);
  return adakv_key_histogram(pooled, sums, H, n, dtype, normalize, out_hist, (cudaStream_t)stream);
}

size_t kvc_think_prune_workspace_bytes(int B, int H, int rows, int d) {
  if (B <= 0 || H <= 0 || rows <= 0 || d <= 0) return 0;
  return think_prune_workspace_bytes(B * H, rows, d);
}

int kvc_think_prune(const void* k_compressed, const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, int B, int H, int L,
                    int rows, int d, int recent, int n_drop, int dtype, void* out_pruned, uint8_t* out_mask, void* workspace,
                    size_t workspace_bytes, void* stream) {
  if (int rc = check_device()) return rc;
  KVC_REQUIRE(k_compressed && q && out_mask && workspace && (out_pruned || rows <= recent || n_drop == d), KVC_ERR_ARG, "null pointer argument");
  KVC_REQUIRE(B > 0 && H > 0 && L > 0 && rows > 0 && d > 0 && d <= 1024, KVC_ERR_ARG, "B=%d H=%d L=%d rows=%d d=%d", B, H, L, rows, d);
  KVC_REQUIRE(recent >= 0 && n_drop >= 0 && n_drop <= d, KVC_ERR_ARG, "recent=%d n_drop=%d", recent, n_drop);
  KVC_REQUIRE(dtype == KVC_BF16 || dtype == KVC_FP16, KVC_ERR_UNSUPPORTED, "dtype %d", dtype);
  return think_prune(k_compressed, q, q_sb, q_sh, q_sl, B, H, L, rows, d, recent > rows ? rows : recent, n_drop, dtype, out_pruned,
                     out_mask, workspace, workspace_bytes, (cudaStream_t)stream);
}

int kvc_update_flatten_view(void* dst, const void* src, const void* state, const int32_t* headlens,
                            const int32_t* cu_headlens, int H, int d, int elem_bytes, int64_t total_rows,
                            void* stream) {
  if (int rc = check_device()) return rc;
  KVC_REQUIRE(dst && src && state && headlens && cu_headlens, KVC_ERR_ARG, "null pointer argument");
  KVC_REQUIRE(H > 0 && d > 0 && total_rows >= 0, KVC_ERR_ARG, "H=%d d=%d total_rows=%lld", H, d, (long long)total_rows);
  KVC_REQUIRE(elem_bytes == 2, KVC_ERR_UNSUPPORTED,
              "elem_bytes %d: the reference kernel is fp16/bf16 only (static_switch.h:3-12)", elem_bytes);
  return update_flatten_view(dst, src, state, headlens, cu_headlens, H, d, elem_bytes, total_rows,
                             (cudaStream_t)stream);
}

int kvc_append_rows(void* k_buf, void* v_buf, const void* new_k, const void* new_v, const int32_t* cu_cap,
                    const int32_t* head_lens, int H, int d, int elem_bytes, void* stream) {
  if (int rc = check_device()) return rc;
  KVC_REQUIRE(k_buf && v_buf && new_k && new_v && cu_cap && head_lens, KVC_ERR_ARG, "null pointer argument");
  KVC_REQUIRE(H > 0 && d > 0, KVC_ERR_ARG, "H=%d d=%d", H, d);
  KVC_REQUIRE(elem_bytes == 2, KVC_ERR_UNSUPPORTED, "elem_bytes %d: the flattened cache is fp16/bf16", elem_bytes);
  return append_rows(k_buf, v_buf, new_k, new_v, cu_cap, head_lens, H, d, elem_bytes, (cudaStream_t)stream);
}

int kvc_flatten_regrow(void* dst, const void* src, const int32_t* head_lens, const int32_t* cu_src,
                       const int32_t* cu_dst, int H, int d, int elem_bytes, int64_t total_rows, void* stream) {
  if (int rc = check_device()) return rc;
  KVC_REQUIRE(dst && src && head_lens && cu_src && cu_dst, KVC_ERR_ARG, "null pointer argument");
  KVC_REQUIRE(H > 0 && d > 0 && total_rows >= 0, KVC_ERR_ARG, "H=%d d=%d total_rows=%lld", H, d, (long long)total_rows);
  KVC_REQUIRE(elem_bytes == 2, KVC_ERR_UNSUPPORTED, "elem_bytes %d: the flattened cache is fp16/bf16", elem_bytes);
  return flatten_regrow(dst, src, head_lens, cu_src, cu_dst, H, d, elem_bytes, total_rows, (cudaStream_t)stream);
}

}  // extern "C"
