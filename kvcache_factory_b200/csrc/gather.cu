// Row movers: dense/varlen gather-compact of the selected K,V rows and the flattened-cache append.
// Pure HBM traffic: every row is moved as 16-byte vectors (a d=128 bf16 row is 256 B = 16 lanes),
// several rows in flight per thread, streaming (no-L1-allocate) loads and stores.
#include <cstdlib>

#include "common.cuh"

namespace kvc {

constexpr int kGatherThreads = 256;
constexpr int kGatherUnroll = 4;

struct RowSrc {
  const char* base;  // tensor base
  int64_t sb, sh, sl;  // byte strides
};

// One thread moves one 16-B chunk of kGatherUnroll rows of K and of V.
// rows are numbered over [units, rows_per_unit]; unit = b*Hq + h.
template <bool kVarlen>
__global__ void __launch_bounds__(kGatherThreads)
gather_rows_kernel(RowSrc ksrc, RowSrc vsrc, const int32_t* __restrict__ idx, int64_t idx_stride,
                   const int32_t* __restrict__ cap, const int32_t* __restrict__ cu_klen,
                   int Hq, int groups, int L, int w, int n_sel, int chunks_per_row,
                   char* __restrict__ out_k, char* __restrict__ out_v, unsigned long long* tl) {
  asm volatile("griddepcontrol.wait;" ::: "memory");   // programmatic dependent launch (KVC_PDL_TAIL): see gather_dense
  tl_start(tl);
  const int unit = blockIdx.y;
  const int b = unit / Hq, h = unit % Hq;
  const int nsel = kVarlen ? min(cap[h], n_sel) : n_sel;   // varlen: n_sel = width of an index row (memory safety)
  const int rows = nsel + w;
  const int64_t out_row0 = kVarlen ? (int64_t)cu_klen[h] : (int64_t)unit * rows;
  const int rows_per_pass = kGatherThreads / chunks_per_row;
  const int c = threadIdx.x % chunks_per_row;
  const int rg = threadIdx.x / chunks_per_row;
  if (rg >= rows_per_pass) return;
  const int block_rows = rows_per_pass * kGatherUnroll;

  const int64_t rowbytes = (int64_t)chunks_per_row * 16;
  const char* kbase = ksrc.base + b * ksrc.sb + (h / groups) * ksrc.sh + c * 16;
  const char* vbase = vsrc.base + b * vsrc.sb + (h / groups) * vsrc.sh + c * 16;
  const int32_t* my_idx = idx ? idx + (int64_t)unit * idx_stride : nullptr;

  // (grid-stride so that a caller may launch fewer CTAs than row blocks; the shipped launches cover them all.
  // A one-CTA-per-SM "background" grid next to the next layer's select was measured: both kernels slow down
  // by more than the overlap gains -- they share the LSU/MIO path -- 540 vs 513 us for 8 cfg-2 layers.)
  for (int blk = blockIdx.x; blk * block_rows < rows; blk += gridDim.x) {
    const int r0 = blk * block_rows + rg;
    int srow[kGatherUnroll];
#pragma unroll
    for (int u = 0; u < kGatherUnroll; ++u) {
      const int r = r0 + u * rows_per_pass;
      int s = -1;
      if (r < rows) s = (r < nsel) ? (my_idx ? my_idx[r] : r) : (L - w + (r - nsel));
      srow[u] = s;
    }
    int4 kv[kGatherUnroll], vv[kGatherUnroll];
#pragma unroll
    for (int u = 0; u < kGatherUnroll; ++u) {
      if (srow[u] >= 0) {
        kv[u] = ld_stream16(kbase + (int64_t)srow[u] * ksrc.sl);
        vv[u] = ld_stream16(vbase + (int64_t)srow[u] * vsrc.sl);
      }
    }
#pragma unroll
    for (int u = 0; u < kGatherUnroll; ++u) {
      if (srow[u] >= 0) {
        const int64_t off = (out_row0 + r0 + u * rows_per_pass) * rowbytes + c * 16;
        st_stream16(out_k + off, kv[u]);
        st_stream16(out_v + off, vv[u]);
      }
    }
  }
  tl_end(tl);
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int check_rows(const void* k, const void* v, int64_t rowbytes, int64_t ksb, int64_t ksh, int64_t ksl,
                      int64_t vsb, int64_t vsh, int64_t vsl, int e, const void* ok, const void* ov) {
  KVC_REQUIRE(rowbytes % 16 == 0 && rowbytes <= 16 * kGatherThreads, KVC_ERR_UNSUPPORTED,
              "row of %lld bytes: need a multiple of 16 B and <= %d B", (long long)rowbytes, 16 * kGatherThreads);
  KVC_REQUIRE(aligned16(k) && aligned16(v) && aligned16(ok) && aligned16(ov), KVC_ERR_UNSUPPORTED,
              "K/V/out pointers must be 16-byte aligned");
  KVC_REQUIRE((ksb * e) % 16 == 0 && (ksh * e) % 16 == 0 && (ksl * e) % 16 == 0 && (vsb * e) % 16 == 0 &&
                  (vsh * e) % 16 == 0 && (vsl * e) % 16 == 0,
              KVC_ERR_UNSUPPORTED, "K/V strides must be multiples of 16 bytes");
  return 0;
}

int gather_dense(const void* k, int64_t k_sb, int64_t k_sh, int64_t k_sl, const void* v, int64_t v_sb,
                 int64_t v_sh, int64_t v_sl, const int32_t* idx, int B, int Hq, int Hkv, int L, int d, int n_sel,
                 int w, int e, void* out_k, void* out_v, cudaStream_t stream) {
  const int64_t rowbytes = (int64_t)d * e;
  if (int rc = check_rows(k, v, rowbytes, k_sb, k_sh, k_sl, v_sb, v_sh, v_sl, e, out_k, out_v)) return rc;
  const int cpr = (int)(rowbytes / 16);
  const int rows = n_sel + w;
  const int block_rows = (kGatherThreads / cpr) * kGatherUnroll;
  dim3 grid((rows + block_rows - 1) / block_rows, B * Hq);
  RowSrc ks{(const char*)k, k_sb * e, k_sh * e, k_sl * e}, vs{(const char*)v, v_sb * e, v_sh * e, v_sl * e};
  static const bool pdl = [] { const char* e = getenv("KVC_PDL_TAIL"); return e && atoi(e) == 1; }();
  if (pdl) {   // let the grid be scheduled while the select that produces the indices is still draining
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kGatherThreads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int32_t* nocap = nullptr;
    KVC_CUDA_TRY(cudaLaunchKernelEx(&cfg, gather_rows_kernel<false>, ks, vs, idx, (int64_t)n_sel, nocap, nocap, Hq, Hq / Hkv, L, w,
                                    n_sel, cpr, (char*)out_k, (char*)out_v, KVC_TL_SLOT(2)));
    return launch_status("gather_rows_kernel<dense>");
  }
  gather_rows_kernel<false><<<grid, kGatherThreads, 0, stream>>>(ks, vs, idx, n_sel, nullptr, nullptr, Hq, Hq / Hkv,
                                                                  L, w, n_sel, cpr, (char*)out_k, (char*)out_v, KVC_TL_SLOT(2));
  return launch_status("gather_rows_kernel<dense>");
}

int gather_varlen(const void* k, int64_t k_sh, int64_t k_sl, const void* v, int64_t v_sh, int64_t v_sl,
                  const int32_t* idx, int64_t idx_stride, const int32_t* cap, const int32_t* cu_klen, int Hq,
                  int Hkv, int L, int d, int w, int max_cap, int e, void* out_k, void* out_v,
                  cudaStream_t stream) {
  const int64_t rowbytes = (int64_t)d * e;
  if (int rc = check_rows(k, v, rowbytes, 0, k_sh, k_sl, 0, v_sh, v_sl, e, out_k, out_v)) return rc;
  const int cpr = (int)(rowbytes / 16);
  const int rows = max_cap + w;
  const int block_rows = (kGatherThreads / cpr) * kGatherUnroll;
  dim3 grid((rows + block_rows - 1) / block_rows, Hq);
  RowSrc ks{(const char*)k, 0, k_sh * e, k_sl * e}, vs{(const char*)v, 0, v_sh * e, v_sl * e};
  gather_rows_kernel<true><<<grid, kGatherThreads, 0, stream>>>(ks, vs, idx, idx_stride, cap, cu_klen, Hq, Hq / Hkv,
                                                                 L, w, max_cap, cpr, (char*)out_k, (char*)out_v, KVC_TL_SLOT(2));
  return launch_status("gather_rows_kernel<varlen>");
}

// ---- flattened-cache append (cuda_api.cu:11-53 restated as a 16-byte segmented shift copy) ------
// Head h's old rows move down by h rows; its new row lands right after them.  grid = (x, H); the
// x-blocks of a head stride over that head's 16-byte chunks, kFlatUnroll chunks in flight.
constexpr int kFlatThreads = 256;
constexpr int kFlatUnroll = 4;

__global__ void __launch_bounds__(kFlatThreads)
flatten_append_kernel(char* __restrict__ dst, const char* __restrict__ src, const char* __restrict__ state,
                      const int32_t* __restrict__ headlens, const int32_t* __restrict__ cu, int rowbytes) {
  const int h = blockIdx.y;
  const int64_t begin = cu[h];
  const int64_t nchunks = (int64_t)headlens[h] * rowbytes / 16;
  const char* s = src + begin * rowbytes;
  char* d = dst + (begin + h) * rowbytes;
  const int64_t stride = (int64_t)gridDim.x * kFlatThreads;
  int64_t i = (int64_t)blockIdx.x * kFlatThreads + threadIdx.x;
  for (; i + (kFlatUnroll - 1) * stride < nchunks; i += kFlatUnroll * stride) {
    int4 t[kFlatUnroll];
#pragma unroll
    for (int u = 0; u < kFlatUnroll; ++u) t[u] = ld_stream16(s + (i + u * stride) * 16);
#pragma unroll
    for (int u = 0; u < kFlatUnroll; ++u) st_stream16(d + (i + u * stride) * 16, t[u]);
  }
  for (; i < nchunks; i += stride) st_stream16(d + i * 16, ld_stream16(s + i * 16));
  if (blockIdx.x == 0 && threadIdx.x < rowbytes / 16) {
    const int64_t ins = ((int64_t)cu[h + 1] + h) * rowbytes;
    st_stream16(dst + ins + threadIdx.x * 16, ld_stream16(state + (int64_t)h * rowbytes + threadIdx.x * 16));
  }
}

int update_flatten_view(void* dst, const void* src, const void* state, const int32_t* headlens,
                        const int32_t* cu_headlens, int H, int d, int e, int64_t total_rows, cudaStream_t stream) {
  const int rowbytes = d * e;
  KVC_REQUIRE(rowbytes % 16 == 0, KVC_ERR_UNSUPPORTED, "row of %d bytes is not a multiple of 16", rowbytes);
  KVC_REQUIRE(rowbytes <= 256, KVC_ERR_UNSUPPORTED,
              "head_dim*elem_bytes = %d > 256 (the reference requires head_dim <= 128, cuda_api.cu:74)", rowbytes);
  KVC_REQUIRE(aligned16(dst) && aligned16(src) && aligned16(state), KVC_ERR_UNSUPPORTED,
              "cache/state pointers must be 16-byte aligned");
  const int64_t chunks_per_head = (total_rows * rowbytes / 16 + H - 1) / H;
  int64_t gx = (chunks_per_head + (int64_t)kFlatThreads * kFlatUnroll - 1) / ((int64_t)kFlatThreads * kFlatUnroll);
  const int64_t cap_x = (int64_t)(8 * sm_count() + H - 1) / H;  // ~8 CTAs per SM over all heads
  if (gx > cap_x) gx = cap_x;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, H);
  flatten_append_kernel<<<grid, kFlatThreads, 0, stream>>>((char*)dst, (const char*)src, (const char*)state, headlens,
                                                           cu_headlens, rowbytes);
  return launch_status("flatten_append_kernel");
}

// ---- key L2 norms as selection scores (L2Norm policy, pyramidkv_utils.py:394-429) -------
// out[b, hq, l] = -||K[b, hq / groups, l, :]||_2 in the model dtype (fp32 sum of squares, sqrt, one rounding, as
// torch.norm on a 16-bit tensor), NEGATED so that the top-k select (largest first) yields the smallest norms in
// ascending order -- the reference's argsort()[:cap].  One row per `lanes` = rowbytes / 16 threads.
template <typename T>
__global__ void __launch_bounds__(256)
key_norm_kernel(const char* __restrict__ k, int64_t sb, int64_t sh, int64_t sl, int Hq, int groups, int L, int lanes,
                uint16_t* __restrict__ out) {
  const int rows_per_cta = 256 / lanes;
  const int c = threadIdx.x % lanes, rg = threadIdx.x / lanes;
  const int unit = blockIdx.y, b = unit / Hq, h = unit % Hq;
  const int l = blockIdx.x * rows_per_cta + rg;
  float ss = 0.f;
  if (l < L && rg < rows_per_cta) {
    const int4 q = ld_stream16(k + b * sb + (int64_t)(h / groups) * sh + (int64_t)l * sl + c * 16);
    const uint32_t w4[4] = {(uint32_t)q.x, (uint32_t)q.y, (uint32_t)q.z, (uint32_t)q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = T::to_float((uint16_t)(w4[i] & 0xffffu)), bb = T::to_float((uint16_t)(w4[i] >> 16));
      ss = fmaf(a, a, ss);
      ss = fmaf(bb, bb, ss);
    }
  }
  for (int o = lanes >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (c == 0 && l < L && rg < rows_per_cta) out[(int64_t)unit * L + l] = T::round1(-sqrtf(ss));
}

int key_norm_scores(const void* k, int64_t k_sb, int64_t k_sh, int64_t k_sl, int B, int Hq, int Hkv, int L, int d,
                    int dtype, void* out, cudaStream_t stream) {
  const int rowbytes = d * 2;
  const int lanes = rowbytes / 16;
  KVC_REQUIRE(rowbytes % 16 == 0 && lanes >= 1 && lanes <= 32 && (lanes & (lanes - 1)) == 0, KVC_ERR_UNSUPPORTED,
              "head_dim %d: a row must be a power-of-two number (<= 32) of 16-byte chunks", d);
  KVC_REQUIRE(aligned16(k) && ((k_sb | k_sh | k_sl) * 2) % 16 == 0, KVC_ERR_UNSUPPORTED, "key rows must be 16-byte aligned");
  const int rows_per_cta = 256 / lanes;
  dim3 grid((L + rows_per_cta - 1) / rows_per_cta, B * Hq);
  if (dtype == KVC_BF16)
    key_norm_kernel<Bf16><<<grid, 256, 0, stream>>>((const char*)k, k_sb * 2, k_sh * 2, k_sl * 2, Hq, Hq / Hkv, L, lanes, (uint16_t*)out);
  else
    key_norm_kernel<Fp16><<<grid, 256, 0, stream>>>((const char*)k, k_sb * 2, k_sh * 2, k_sl * 2, Hq, Hq / Hkv, L, lanes, (uint16_t*)out);
  return launch_status("key_norm_kernel");
}

// ---- slack layout of the flattened cache: O(H*d) decode append ----------------------
// Head h owns rows [cu_cap[h], cu_cap[h+1]) of the buffer and uses the first head_lens[h] of them.
// append: the new row of head h goes to row cu_cap[h] + head_lens[h] of the K and of the V buffer (the caller bumps
// head_lens afterwards, as the reference's forward does, llama_model.py:2372-2375).  One CTA per head and tensor.
__global__ void __launch_bounds__(32)
append_rows_kernel(char* __restrict__ kbuf, char* __restrict__ vbuf, const char* __restrict__ knew,
                   const char* __restrict__ vnew, const int32_t* __restrict__ cu_cap,
                   const int32_t* __restrict__ head_lens, int rowbytes) {
  const int h = blockIdx.x;
  const int len = head_lens[h];
  if (len >= cu_cap[h + 1] - cu_cap[h]) return;   // no slack left: the host regrows before this can happen
  char* dst = (blockIdx.y ? vbuf : kbuf) + ((int64_t)cu_cap[h] + len) * rowbytes;
  const char* src = (blockIdx.y ? vnew : knew) + (int64_t)h * rowbytes;
  if (threadIdx.x < rowbytes / 16) st_stream16(dst + threadIdx.x * 16, ld_stream16(src + threadIdx.x * 16));
}

// regrow / re-layout: head h's head_lens[h] rows move from src rows [cu_src[h], ...) to dst rows [cu_dst[h], ...).
__global__ void __launch_bounds__(kFlatThreads)
flatten_regrow_kernel(char* __restrict__ dst, const char* __restrict__ src, const int32_t* __restrict__ head_lens,
                      const int32_t* __restrict__ cu_src, const int32_t* __restrict__ cu_dst, int rowbytes) {
  const int h = blockIdx.y;
  const int64_t nchunks = (int64_t)head_lens[h] * rowbytes / 16;
  const char* s = src + (int64_t)cu_src[h] * rowbytes;
  char* d = dst + (int64_t)cu_dst[h] * rowbytes;
  const int64_t stride = (int64_t)gridDim.x * kFlatThreads;
  int64_t i = (int64_t)blockIdx.x * kFlatThreads + threadIdx.x;
  for (; i + (kFlatUnroll - 1) * stride < nchunks; i += kFlatUnroll * stride) {
    int4 t[kFlatUnroll];
#pragma unroll
    for (int u = 0; u < kFlatUnroll; ++u) t[u] = ld_stream16(s + (i + u * stride) * 16);
#pragma unroll
    for (int u = 0; u < kFlatUnroll; ++u) st_stream16(d + (i + u * stride) * 16, t[u]);
  }
  for (; i < nchunks; i += stride) st_stream16(d + i * 16, ld_stream16(s + i * 16));
}

int append_rows(void* kbuf, void* vbuf, const void* knew, const void* vnew, const int32_t* cu_cap,
                const int32_t* head_lens, int H, int d, int e, cudaStream_t stream) {
  const int rowbytes = d * e;
  KVC_REQUIRE(rowbytes % 16 == 0 && rowbytes <= 512, KVC_ERR_UNSUPPORTED, "row of %d bytes: need a multiple of 16, <= 512", rowbytes);
  KVC_REQUIRE(aligned16(kbuf) && aligned16(vbuf) && aligned16(knew) && aligned16(vnew), KVC_ERR_UNSUPPORTED,
              "cache/state pointers must be 16-byte aligned");
  append_rows_kernel<<<dim3(H, 2), 32, 0, stream>>>((char*)kbuf, (char*)vbuf, (const char*)knew, (const char*)vnew, cu_cap,
                                                   head_lens, rowbytes);
  return launch_status("append_rows_kernel");
}

int flatten_regrow(void* dst, const void* src, const int32_t* head_lens, const int32_t* cu_src, const int32_t* cu_dst,
                   int H, int d, int e, int64_t total_rows, cudaStream_t stream) {
  const int rowbytes = d * e;
  KVC_REQUIRE(rowbytes % 16 == 0, KVC_ERR_UNSUPPORTED, "row of %d bytes is not a multiple of 16", rowbytes);
  KVC_REQUIRE(aligned16(dst) && aligned16(src), KVC_ERR_UNSUPPORTED, "cache pointers must be 16-byte aligned");
  const int64_t chunks_per_head = (total_rows * rowbytes / 16 + H - 1) / H;
  int64_t gx = (chunks_per_head + (int64_t)kFlatThreads * kFlatUnroll - 1) / ((int64_t)kFlatThreads * kFlatUnroll);
  const int64_t cap_x = (int64_t)(8 * sm_count() + H - 1) / H;
  if (gx > cap_x) gx = cap_x;
  if (gx < 1) gx = 1;
  flatten_regrow_kernel<<<dim3((unsigned)gx, H), kFlatThreads, 0, stream>>>((char*)dst, (const char*)src, head_lens, cu_src,
                                                                           cu_dst, rowbytes);
  return launch_status("flatten_regrow_kernel");
}

}  // namespace kvc
