/*
 * CPU ORACLE (test infrastructure only; never on the product path).
 *
 * Plain-C restatement of the reference's only native kernel, update_flatten_view_kernel
 * (/root/reference/csrc/csrc/cuda_api.cu:11-53) and of the output sizing in its host wrapper
 * (:59-65).  The CUDA kernel walks every element of every head with a grid-stride loop; this file
 * states the same index arithmetic element by element so the byte layout it defines can be checked
 * against kvc_update_flatten_view.  Elements are 16-bit patterns (fp16/bf16 are never interpreted).
 *
 * Build: gcc -O2 -shared -fPIC oracle/flatten_oracle.c -o oracle/_build/libflatten_oracle.so
 */
#include <stdint.h>

/* dst: [total_rows + H, dim], src: [total_rows, dim], state: [H, dim] */
void flatten_oracle_update(uint16_t* dst, const uint16_t* src, const uint16_t* state, const int32_t* headlens,
                           const int32_t* cu_headlens, int32_t H, int32_t dim) {
  for (int32_t head = 0; head < H; ++head) {
    const int64_t headlen = headlens[head];
    const int64_t src_off = (int64_t)cu_headlens[head] * dim;             /* cuda_api.cu:26 */
    const int64_t dst_off = src_off + (int64_t)head * dim;                /* :27 */
    const int64_t ins_off = ((int64_t)cu_headlens[head + 1] + head) * dim; /* :28 */
    for (int64_t e = 0; e < headlen * dim; ++e) dst[dst_off + e] = src[src_off + e]; /* :34-45 */
    for (int32_t t = 0; t < dim; ++t) dst[ins_off + t] = state[(int64_t)head * dim + t]; /* :48-52 */
  }
}
