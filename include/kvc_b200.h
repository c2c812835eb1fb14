/*
 * kvc_b200.h — C ABI of libkvc_b200.so: the B200-native (sm_100a) prefill KV-compression hot path.
 *
 * This is the drop-in boundary for the path that the reference (assassin808/KVCache-Factory)
 * implements as a chain of PyTorch ops in pyramidkv/pyramidkv_utils.py plus one pybind11 CUDA
 * function in csrc/csrc/cuda_api.cu.  Every entry point takes plain pointers, sizes and element
 * strides (no torch types), returns an int status and launches on the caller's stream.
 *
 *   status == 0            ok (work is enqueued on `stream`; nothing is synchronised)
 *   status  < 0            KVC_ERR_* argument / capability error, message via kvc_last_error()
 *   status  > 0            a cudaError_t raised by a launch or an API call
 *
 * Conventions
 *   - All tensor pointers are DEVICE pointers unless the name ends in `_host`.
 *   - Q is logical [B,Hq,L,d], K and V are logical [B,Hkv,L,d] (un-repeated GQA layout; pass
 *     Hkv == Hq for the reference's repeat_kv-expanded layout).  The last dimension must be
 *     contiguous; the other three are described by ELEMENT strides (sb, sh, sl), because HF hands
 *     query_states over as a transposed view.
 *   - dtype: KVC_BF16 or KVC_FP16 (the reference kernel's FP16_SWITCH, csrc/static_switch.h:3-12).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - Workspaces are caller-allocated device memory; size them with the *_workspace_bytes calls.
 *
 * Reference interface each function replaces is cited as path:line under /root/reference.
 */
#ifndef KVC_B200_H_
#define KVC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KVC_ABI_VERSION 1

enum { KVC_BF16 = 0, KVC_FP16 = 1 };
enum { KVC_POOL_NONE = 0, KVC_POOL_AVG = 1, KVC_POOL_MAX = 2 };
enum { KVC_REDUCE_SUM = 0, KVC_REDUCE_MEAN = 1 };
/* scoring implementation: AUTO picks the tcgen05/TMA kernel when the shape allows it */
enum { KVC_IMPL_AUTO = 0, KVC_IMPL_MMA_SYNC = 1, KVC_IMPL_TCGEN05 = 2 /* two launches */,
       KVC_IMPL_TCGEN05_FUSED = 3 /* one cooperative launch with a grid barrier between the passes */,
       KVC_IMPL_TCGEN05_HALF = 4  /* same, half-footprint CTAs: two launches (two streams) share every SM */,
       KVC_IMPL_TCGEN05_DYN = 5   /* one launch, work claimed dynamically per (batch, KV head, row block) unit: no grid
                                     barrier and no co-residency requirement; AUTO uses it for more units than SMs */ };

enum {
  KVC_OK = 0,
  KVC_ERR_ARG = -1,         /* malformed argument (null pointer, non-positive size, bad enum)      */
  KVC_ERR_UNSUPPORTED = -2, /* shape/dtype outside what the sm_100a kernels implement              */
  KVC_ERR_WORKSPACE = -3,   /* workspace too small                                                 */
  KVC_ERR_DEVICE = -4       /* not an sm_100 device / no CUDA device                               */
};

/* ABI version of the loaded library (== KVC_ABI_VERSION it was built with). */
int kvc_version(void);
/* Thread-local description of the last non-zero status returned on this thread. */
const char* kvc_last_error(void);
/* Number of kernels this library has launched in this process (monotonic; used by bench.py). */
uint64_t kvc_launch_count(void);
/* Number of SMs of the current device (grids are sized in multiples of it). */
int kvc_sm_count(int* out_sms);

/* ------------------------------------------------------------------------------------------------
 * Window scoring — replaces the matmul / div / mask / softmax / sum chain of
 *   SnapKVCluster.update_kv      pyramidkv/pyramidkv_utils.py:317-327
 *   PyramidKVCluster.update_kv   pyramidkv/pyramidkv_utils.py:221-232, 253-260
 *   AdaKVCluster.calcul_attn_sore pyramidkv/pyramidkv_utils.py:649-661 (reduce = MEAN)
 * and folds repeat_kv (pyramidkv_utils.py:108-117) into the kernel: each KV head is read once for
 * its Hq/Hkv query heads.
 *
 * out_scores[b,h,j] (dtype, contiguous [B,Hq,L-w]) = model-dtype rounding of
 *   sum|mean over the last w query rows i of  round( softmax_j( round(round(q_i.k_j)/sqrt(d)) + causal_ww ) )
 * with the reference's rounding points (GEMM result, scaled logits, probabilities and the
 * reduction are each rounded to `dtype`; softmax itself is fp32).
 * ---------------------------------------------------------------------------------------------- */
size_t kvc_window_scores_workspace_bytes(int B, int Hq, int Hkv, int L, int d, int w);
int kvc_window_scores(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sl,
                      const void* k, int64_t k_sb, int64_t k_sh, int64_t k_sl,
                      int B, int Hq, int Hkv, int L, int d, int w,
                      int dtype, int reduce, void* out_scores,
                      void* workspace, size_t workspace_bytes, int impl, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Window scores from statistics the caller already has: `lse` is float32
 * [B,Hq,w], the log-sum-exp (natural log, of the 1/sqrt(d)-scaled logits) of each observation-window query
 * row over ALL keys with the causal mask inside the window -- exactly the `softmax_lse` rows the prefill's
 * flash attention returns for the last w queries.  Only the second softmax pass runs (about half a scoring
 * launch).  NOT bit-parity with the reference: flash attention's logits are not rounded to the model dtype
 * before the softmax (pyramidkv_utils.py:317 rounds twice), so scores differ in the last bits; opt-in.
 * Same workspace as kvc_window_scores.
 * ---------------------------------------------------------------------------------------------- */
int kvc_window_scores_lse(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sl,
                          const void* k, int64_t k_sb, int64_t k_sh, int64_t k_sl, const float* lse,
                          int B, int Hq, int Hkv, int L, int d, int w, int dtype, int reduce,
                          void* out_scores, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * H2O scoring — replaces H2OKVCluster.update_kv pyramidkv/pyramidkv_utils.py:544-554: ALL L query
 * rows score every key (non-causal except the bottom-right w x w block), column-summed.  The
 * [B,H,L,L] matrix is never materialised.
 * ---------------------------------------------------------------------------------------------- */
size_t kvc_h2o_scores_workspace_bytes(int B, int Hq, int Hkv, int L, int d, int w);
int kvc_h2o_scores(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sl,
                   const void* k, int64_t k_sb, int64_t k_sh, int64_t k_sl,
                   int B, int Hq, int Hkv, int L, int d, int w,
                   int dtype, void* out_scores,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Pool + top-k — replaces F.avg_pool1d / F.max_pool1d (stride 1, padding kernel/2) followed by
 * Tensor.topk(k).indices: pyramidkv/pyramidkv_utils.py:328-334 (SnapKV), :233-239 / :261-270
 * (PyramidKV), :562-563 (H2O, pool = NONE), and the per-head sort prefix of AdaKV (:706, :734).
 *
 * scores: [rows, n] dtype contiguous.  For every row the k largest pooled values are selected;
 * out_idx[row, 0..k) lists their positions ordered by DESCENDING pooled value (torch.topk
 * sorted=True), equal values by ascending position (torch leaves tie order unspecified).
 *   k_per_row == NULL : every row selects `k` entries, out_idx row stride = idx_stride (>= k)
 *   k_per_row != NULL : device int32[rows]; row r selects k_per_row[r] <= k_max = k entries
 * out_pooled (optional, may be NULL): [rows, n] dtype, the pooled scores the selection was made on.
 * out_sums (optional, may be NULL): float[rows,2] = {fp32 sum of the selected values, fp32 sum of
 * all pooled values} (AdaKV ratio normalisation, pyramidkv_utils.py:709-711).
 *   out_idx == NULL with k > 0 and out_sums != NULL: "sums only" -- the kernel stops after the
 *   threshold search (no compaction, no sort, no indices); KVC_ERR_UNSUPPORTED for shapes outside the
 *   cluster kernel (k > 8192, n > 131072, kernel_size > 7), where the caller passes an index buffer.
 * ---------------------------------------------------------------------------------------------- */
size_t kvc_pool_topk_workspace_bytes(int rows, int n, int k);
int kvc_pool_topk(const void* scores, int rows, int n, int dtype, int kernel_size, int pool,
                  int k, const int32_t* k_per_row, int32_t* out_idx, int64_t idx_stride,
                  void* out_pooled, float* out_sums,
                  void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dense gather/compact — replaces gather(dim=2, idx) + cat([.., window]) for K and V:
 * pyramidkv/pyramidkv_utils.py:341-346 (SnapKV), :245-250 / :277-282 (PyramidKV), :569-574 (H2O),
 * :607-619 (StreamingLLM: pass idx == NULL for idx = arange(n_sel), no index tensor is built).
 *
 * out_k/out_v: contiguous [B,Hq,n_sel+w,d].  Row r < n_sel of head h copies K[b, h/(Hq/Hkv),
 * idx[b,h,r], :]; rows n_sel.. copy the last w rows.  idx: int32 [B,Hq,n_sel], values in [0,L-w).
 * Bytes move untouched (bit-exact); elem_bytes in {2,4}.
 * v (and only v) may point into PINNED host memory: V is read at the selected rows only, so a caller
 * whose KV lives in host memory need not move the rest of V to the device (unified addressing).
 * ---------------------------------------------------------------------------------------------- */
int kvc_gather_dense(const void* k, int64_t k_sb, int64_t k_sh, int64_t k_sl,
                     const void* v, int64_t v_sb, int64_t v_sh, int64_t v_sl,
                     const int32_t* idx, int B, int Hq, int Hkv, int L, int d, int n_sel, int w,
                     int elem_bytes, void* out_k, void* out_v, void* stream);

/* ------------------------------------------------------------------------------------------------
 * L2Norm policy (pyramidkv/pyramidkv_utils.py:394-429): selection scores from the
 * key norms alone.  out[b, hq, l] = -||K[b, hq / (Hq/Hkv), l, :]||_2 in the model dtype [B,Hq,L] (negated:
 * kvc_pool_topk(pool = NONE) then returns the `cap` smallest norms in ascending order, the reference's
 * argsort()[:cap]); follow with kvc_gather_dense(w = 0).
 * ---------------------------------------------------------------------------------------------- */
int kvc_key_norm_scores(const void* k, int64_t k_sb, int64_t k_sh, int64_t k_sl,
                        int B, int Hq, int Hkv, int L, int d, int dtype, void* out_scores, void* stream);

/* ------------------------------------------------------------------------------------------------
 * One call for the whole dense update_kv of SnapKV / PyramidKV / StreamingLLM
 * (pyramidkv/pyramidkv_utils.py:197-283, 306-347, 595-620): window scores -> pool + top-k -> gather +
 * window, three launches on `stream` from ONE host call.  The eager drop-in path is launch bound
 * (a layer of the reference's CPU-runnable config costs 22 us of GPU time and ~95 us of host time when
 * every kernel is a separate Python -> ctypes round trip with its own allocations); this entry point
 * carves scores, indices and the kernels' scratch out of one caller-provided workspace.
 *
 * pool == KVC_POOL_NONE with kernel_size == 0 means "no scoring at all": the first n_keep rows are kept
 * (StreamingLLM).  out_idx / out_scores are written inside the workspace; their byte offsets are returned
 * through idx_offset / scores_offset (may be NULL) for callers that want to inspect them.
 * ---------------------------------------------------------------------------------------------- */
size_t kvc_compress_dense_workspace_bytes(int B, int Hq, int Hkv, int L, int d, int w, int n_keep);
int kvc_compress_dense(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sl,
                       const void* k, int64_t k_sb, int64_t k_sh, int64_t k_sl,
                       const void* v, int64_t v_sb, int64_t v_sh, int64_t v_sl,
                       int B, int Hq, int Hkv, int L, int d, int w, int dtype,
                       int kernel_size, int pool, int n_keep, int score_impl,
                       void* out_k, void* out_v, void* workspace, size_t workspace_bytes,
                       size_t* idx_offset, size_t* scores_offset, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Varlen (flattened) gather — replaces the per-head Python loop of AdaKVCluster.update_kv
 * pyramidkv/pyramidkv_utils.py:733-755: head h writes cap[h] selected rows then the w window rows
 * at row offset cu_klen[h] of the flattened [sum_h(cap[h]+w), d] cache.  Batch is 1 (:724).
 * idx: int32 [Hq, idx_stride]; cap, cu_klen: device int32 [Hq], [Hq+1].
 * ---------------------------------------------------------------------------------------------- */
int kvc_gather_varlen(const void* k, int64_t k_sh, int64_t k_sl,
                      const void* v, int64_t v_sh, int64_t v_sl,
                      const int32_t* idx, int64_t idx_stride, const int32_t* cap,
                      const int32_t* cu_klen, int Hq, int Hkv, int L, int d, int w, int max_cap,
                      int elem_bytes, void* out_k, void* out_v, void* stream);

/* ------------------------------------------------------------------------------------------------
 * AdaKV head-adaptive allocation — replaces pyramidkv/pyramidkv_utils.py:706-719 + init_metadata
 * (:682-698): ratio-normalise each head's pooled scores, take the global top-(H*base) over all
 * heads, count winners per head and apply the floor:
 *   cap[h] = round_half_even(count[h] * (1 - floor_ratio) + int(base * floor_ratio))
 * pooled: [H, n] dtype; sums: float[H,2] from kvc_pool_topk(k = base).  Outputs (device int32):
 * cap[H], head_lens[H] = cap + w, cu_klen[H+1] (exclusive prefix + total).
 * counts_inout (optional): int32[H] raw winner counts; with n_ranks > 1 the caller all-reduces the
 * two histograms between the phases (see kvcache_factory_b200/sharding.py), the single-GPU entry
 * does everything in one call.
 * ---------------------------------------------------------------------------------------------- */
size_t kvc_adakv_allocate_workspace_bytes(int H, int n);
int kvc_adakv_allocate(const void* pooled, const float* sums, int H, int n, int dtype,
                       int base_capacity, float one_minus_floor, int floor_capacity, int normalize,
                       int w, int32_t* out_cap, int32_t* out_head_lens, int32_t* out_cu_klen,
                       void* workspace, size_t workspace_bytes, void* stream);

/* The whole AdaKV prefill update_kv (AdaKVCluster.update_kv, pyramidkv/pyramidkv_utils.py:700-757) in ONE
 * call: window-mean scores -> pooled scores + ratio sums -> global allocation -> read of the capacities (the
 * reference synchronises at the same point, :718) -> per-head top-cap_h -> varlen gather.  B = 1 (:724).
 * out_k / out_v: flattened [out_rows, d] buffers sized by the caller with an upper bound
 * (H*(base_capacity+w) + H rows always suffices); the rows actually used are returned in *out_klen_sum.
 * host_cap: H ints of (preferably pinned) host memory; max_cap_bound: widest index row the workspace is sized
 * for (KVC_ERR_WORKSPACE if a head's capacity exceeds it: retry with a larger bound).  offsets (optional,
 * 3 entries): byte offsets of the scores [H,L-w], pooled scores [H,L-w] and indices [H,*out_max_cap] in the
 * workspace.  Synchronises `stream` once. */
size_t kvc_adakv_prefill_workspace_bytes(int Hq, int Hkv, int L, int d, int w, int max_cap_bound);
int kvc_adakv_prefill(const void* q, int64_t q_sh, int64_t q_sl, const void* k, int64_t k_sh, int64_t k_sl,
                      const void* v, int64_t v_sh, int64_t v_sl, int Hq, int Hkv, int L, int d, int w,
                      int dtype, int kernel_size, int pool, int base_capacity, float one_minus_floor,
                      int floor_capacity, int normalize, int score_impl, void* out_k, void* out_v,
                      int64_t out_rows, int32_t* out_cap, int32_t* out_head_lens, int32_t* out_cu_klen,
                      int32_t* host_cap, int max_cap_bound, void* workspace, size_t workspace_bytes,
                      int* out_max_cap, int64_t* out_klen_sum, size_t* offsets, void* stream);

/* Multi-GPU building block of the same allocation (one KV-head shard per GPU):
 * out_hist int32[H,65536] = histogram of each local head's ratio-normalised scores over an
 * order-preserving 16-bit key (key of +0.0 is 0x8000; larger value -> larger key).  The host sums it
 * over heads, all-reduces it over the ranks (one 256 KiB collective) and derives the global
 * threshold and the per-head winner counts (kvcache_factory_b200/sharding.py). */
int kvc_adakv_key_histogram(const void* pooled, const float* sums, int H, int n, int dtype,
                            int normalize, int32_t* out_hist, void* stream);

/* The same exchange on the two-level histograms the single-GPU allocation uses (16 KiB + 64 B all-reduces instead
 * of 256 KiB): coarse = per-head 4096-bin histogram of key >> 4 (out_hist1 int32[H,4096]) and the normalised keys
 * themselves (akeys uint16[H,n], scratch for the second phase); fine = per-head 16-bin histogram of the keys
 * inside the threshold bin (*thr_bin: DEVICE int, so no host read sits between the phases) and the per-head count
 * above that bin.  Reference anchor: the global top-(H*base) of pyramidkv/pyramidkv_utils.py:712-717. */
int kvc_adakv_hist_coarse(const void* pooled, const float* sums, int H, int n, int dtype, int normalize,
                          void* akeys, int32_t* out_hist1, void* stream);
int kvc_adakv_hist_fine(const void* akeys, const int32_t* hist1, int H, int n, const int32_t* thr_bin,
                        int32_t* out_hist2, int32_t* out_above, void* stream);

/* ------------------------------------------------------------------------------------------------
 * ThinK query-driven key-channel pruning of the compressed keys — replaces key_pruner_query_driven
 * pyramidkv/pyramidkv_utils.py:13-26 (called from SnapKVCluster.update_think :349-392).
 *
 * k_compressed: contiguous [B,H,rows,d] (what update_kv returned); q: [B,H,L,d] with element strides, the last
 * min(32, L) rows are read (:17).  The n_drop = int(d*ratio) channels with the smallest
 * mean(q^2) * mean(k^2) (every step rounded to the model dtype) are removed from rows [0, rows-recent):
 * out_pruned contiguous [B*H, rows-recent, d-n_drop], kept channels in channel order; out_mask uint8 [B,H,d],
 * 1 = kept (the reference's ~mask).  The last `recent` rows stay as they are (a slice of k_compressed).
 * ---------------------------------------------------------------------------------------------- */
size_t kvc_think_prune_workspace_bytes(int B, int H, int rows, int d);
int kvc_think_prune(const void* k_compressed, const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sl,
                    int B, int H, int L, int rows, int d, int recent, int n_drop, int dtype,
                    void* out_pruned, uint8_t* out_mask, void* workspace, size_t workspace_bytes,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * Flattened-cache append — replaces tiny_api_cuda.update_flatten_view
 * csrc/csrc/cuda_api.cu:55-85 (kernel :11-53), called by DynamicCacheSplitHeadFlatten.update
 * pyramidkv/pyramidkv_utils.py:66-67.
 *
 * dst: [total_rows + H, d]; src: [total_rows, d]; state: [H, d]; headlens int32[H];
 * cu_headlens int32[H+1].  For head h: dst rows [cu[h]+h, cu[h]+h+headlens[h]) = src rows
 * [cu[h], cu[h]+headlens[h]) and dst row cu[h+1]+h = state[h].  d*elem_bytes <= 256 as in the
 * reference (cuda_api.cu:74 requires head_dim <= 128 of a 16-bit type).
 * ---------------------------------------------------------------------------------------------- */
int kvc_update_flatten_view(void* dst, const void* src, const void* state,
                            const int32_t* headlens, const int32_t* cu_headlens,
                            int H, int d, int elem_bytes, int64_t total_rows, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Slack layout of the flattened cache: O(H*d) decode append.  No reference
 * counterpart: the reference rebuilds the whole layer cache on every decoded token
 * (pyramidkv/pyramidkv_utils.py:52-75 -> cuda_api.cu:55-85); these two entry points replace that with
 * an in-place write of H rows plus a re-layout every `slack` tokens.
 *
 * Buffers are [cu_cap[H], d]; head h owns rows [cu_cap[h], cu_cap[h+1]) and uses the first
 * head_lens[h].  kvc_append_rows writes new_k[h] / new_v[h] ([H, d]) to row cu_cap[h]+head_lens[h]
 * of k_buf / v_buf (heads without slack left are skipped; head_lens is NOT modified: the caller bumps
 * it, as llama_model.py:2375 does).  kvc_flatten_regrow moves head h's head_lens[h] rows from src rows
 * [cu_src[h], ...) to dst rows [cu_dst[h], ...) (src/dst must not overlap).
 * ---------------------------------------------------------------------------------------------- */
int kvc_append_rows(void* k_buf, void* v_buf, const void* new_k, const void* new_v, const int32_t* cu_cap,
                    const int32_t* head_lens, int H, int d, int elem_bytes, void* stream);
int kvc_flatten_regrow(void* dst, const void* src, const int32_t* head_lens, const int32_t* cu_src,
                       const int32_t* cu_dst, int H, int d, int elem_bytes, int64_t total_rows, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KVC_B200_H_ */
