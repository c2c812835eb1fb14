// Window scoring, legacy tensor-core path (mma.sync m16n8k16): the always-available implementation
// that the tcgen05/TMA kernel (scores_tc.cu) is validated against.
//
// For one (batch, KV head) unit the g = Hq/Hkv query heads that share it are stacked into one row
// block (row = head_in_group * wp + window_row, wp = w rounded up to 16), so K is streamed ONCE for
// the whole group (repeat_kv folded away, pyramidkv_utils.py:108-117).
//   pass 1  (stats):  per row, running max m and Z = sum_j exp(l_ij - m) over a chunk of keys
//   pass 2  (scores): p_ij = exp(l_ij - M_i) / Z_i, rounded to the model dtype, summed over the
//                     window rows of each head, rounded again, written for j < L - w
// l_ij reproduces the reference's rounding points (pyramidkv_utils.py:317-326): the GEMM result
// is rounded to the model dtype, multiplied by 1/sqrt(d) in fp32 and rounded again; the w x w
// causal block is masked to -inf; softmax is fp32.
#include "common.cuh"

namespace kvc {

constexpr int kTileN = 64;  // keys per shared-memory stage
constexpr int kStages = 3;

struct ScoreParams {
  const uint16_t* q; int64_t q_sb, q_sh, q_sl;
  const uint16_t* k; int64_t k_sb, k_sh, k_sl;
  int B, Hq, Hkv, L, w, wp, groups;
  int heads_per_blk, nblk;  // query heads per row block, row blocks per unit
  int chunk_keys, nchunks;
  float scale;
  int reduce_mean;
  float2* stats;   // [B*Hkv*nblk][nchunks][rows_pad]
  uint16_t* out;   // [B,Hq,L-w]
};

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gptr), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

template <typename T>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (T::kDtype == KVC_BF16) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}

template <int SLABS>
struct Tiling {
  static constexpr int WN = SLABS >= 4 ? 1 : 4 / SLABS;  // warps along the key axis
  static constexpr int NW = SLABS * WN;
  static constexpr int NT = NW * 32;
  static constexpr int NB = kTileN / WN / 8;  // 8-key blocks per warp per tile
  static constexpr int ROWS = SLABS * 32;
};

template <int D>
constexpr size_t score_smem_bytes(int slabs) {
  return (size_t)kStages * kTileN * D * 2 + 2 * (size_t)(slabs * 2) * kTileN * 4 + (size_t)slabs * 32 * 4 +
         (size_t)slabs * 32 * 8 * 4;
}

template <typename T, int SLABS, int D, bool kPass2>
__global__ void __launch_bounds__(Tiling<SLABS>::NT)
window_score_mma_kernel(const ScoreParams p) {
  using TL = Tiling<SLABS>;
  constexpr int KS = D / 16;        // k-steps
  constexpr int KP = D / 32;        // k-step pairs (one ldmatrix.x4 each)
  constexpr int CPR = D * 2 / 16;   // 16-byte chunks per key row
  constexpr int ROWB = D * 2;

  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* ktile = smem;                                               // [kStages][kTileN][D] swizzled
  float* colsum = reinterpret_cast<float*>(smem + kStages * kTileN * ROWB);  // [2][SLABS*2][kTileN]
  float* crow = colsum + 2 * (SLABS * 2) * kTileN;                           // [ROWS]
  float2* wstat = reinterpret_cast<float2*>(crow + TL::ROWS);               // [WN][ROWS] (pass 1)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slab = warp / TL::WN, wn = warp % TL::WN;
  const int chunk = blockIdx.x;
  const int ub = blockIdx.y;  // unit * nblk + rb
  const int unit = ub / p.nblk, rb = ub % p.nblk;
  const int b = unit / p.Hkv, kvh = unit % p.Hkv;
  const int nscore = p.L - p.w;

  const int key_begin = chunk * p.chunk_keys;
  int key_end = min(key_begin + p.chunk_keys, kPass2 ? nscore : p.L);
  if (key_begin >= key_end) return;
  const int ntiles = (key_end - key_begin + kTileN - 1) / kTileN;

  // ---- this thread's four rows: (mt, half) -> row = slab*32 + mt*16 + half*8 + lane/4 ---------
  int row_i[4];       // window row index, or -1 for a padding row
  int row_head[4];    // query head (global), valid only when row_i >= 0
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int r = slab * 32 + (q >> 1) * 16 + (q & 1) * 8 + (lane >> 2);
    const int hl = r / p.wp, i = r % p.wp;
    const int head = rb * p.heads_per_blk + hl;
    const bool ok = hl < p.heads_per_blk && head < p.groups && i < p.w;
    row_i[q] = ok ? i : -1;
    row_head[q] = kvh * p.groups + head;
  }

  // ---- Q fragments straight from global (strided view, read once per CTA) ----------------------
  uint32_t qf[2][KS][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int q = mt * 2 + hf;
      const uint16_t* qrow = nullptr;
      if (row_i[q] >= 0)
        qrow = p.q + b * p.q_sb + (int64_t)row_head[q] * p.q_sh + (int64_t)(p.L - p.w + row_i[q]) * p.q_sl;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int col = ks * 16 + (lane & 3) * 2;
        qf[mt][ks][hf] = qrow ? *reinterpret_cast<const uint32_t*>(qrow + col) : 0u;
        qf[mt][ks][hf + 2] = qrow ? *reinterpret_cast<const uint32_t*>(qrow + col + 8) : 0u;
      }
    }
  }

  // ---- pass 2: fold all chunks' (m, Z) into c_row = -M*log2e - log2 Z ------------------------
  float c[4] = {0.f, 0.f, 0.f, 0.f};
  if constexpr (kPass2) {
    for (int r = tid; r < TL::ROWS; r += TL::NT) {
      const float2* st = p.stats + (int64_t)ub * p.nchunks * TL::ROWS + r;
      float M = -1e30f;
      for (int cidx = 0; cidx < p.nchunks; ++cidx) M = fmaxf(M, st[(int64_t)cidx * TL::ROWS].x);
      float Z = 0.f;
      for (int cidx = 0; cidx < p.nchunks; ++cidx) {
        const float2 s = st[(int64_t)cidx * TL::ROWS];
        Z += s.y * fast_exp2((s.x - M) * kLog2e);
      }
      const int hl = r / p.wp, i = r % p.wp;
      const bool ok = hl < p.heads_per_blk && rb * p.heads_per_blk + hl < p.groups && i < p.w;
      crow[r] = ok ? (-M * kLog2e - log2f(Z)) : -INFINITY;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) c[q] = crow[slab * 32 + (q >> 1) * 16 + (q & 1) * 8 + (lane >> 2)];
  }

  // ---- K tile pipeline (cp.async, XOR-swizzled 16-byte chunks) -------------------------------
  const uint16_t* kbase = p.k + b * p.k_sb + (int64_t)kvh * p.k_sh;
  const uint32_t ktile_s = (uint32_t)__cvta_generic_to_shared(ktile);
  auto load_tile = [&](int t, int stage) {
    const int key0 = key_begin + t * kTileN;
#pragma unroll
    for (int it = 0; it < kTileN * CPR / TL::NT; ++it) {
      const int cid = tid + it * TL::NT;
      const int r = cid / CPR, ch = cid % CPR;
      const int key = key0 + r;
      const bool ok = key < p.L;
      const uint16_t* g = kbase + (int64_t)(ok ? key : 0) * p.k_sl + ch * 8;
      cp_async16(ktile_s + stage * (kTileN * ROWB) + r * ROWB + ((ch ^ (r & 7)) << 4), g, ok);
    }
  };
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < ntiles) load_tile(s, s);
    cp_async_commit();
  }

  float m_run[4], z_run[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) { m_run[q] = -1e30f; z_run[q] = 0.f; }

  for (int t = 0; t < ntiles; ++t) {
    cp_async_wait<kStages - 2>();
    __syncthreads();
    if (t + kStages - 1 < ntiles) load_tile(t + kStages - 1, (t + kStages - 1) % kStages);
    cp_async_commit();

    const int stage = t % kStages;
    const int key0 = key_begin + t * kTileN;
    const uint32_t st_base = ktile_s + stage * (kTileN * ROWB);
    const bool needs_mask = key0 + kTileN > nscore;  // tile touches the window block or runs past L

    uint32_t r2p[TL::NB][4];
#pragma unroll
    for (int nb = 0; nb < TL::NB; ++nb) {
      const int kb = wn * TL::NB + nb;
      float acc[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[mt][v] = 0.f;
      const int lrow = kb * 8 + (lane & 7);
#pragma unroll
      for (int kp = 0; kp < KP; ++kp) {
        const int ch = 4 * kp + (lane >> 3);
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(st_base + lrow * ROWB + ((ch ^ (lrow & 7)) << 4), b0, b1, b2, b3);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma16816<T>(acc[mt], qf[mt][2 * kp], b0, b1);
          mma16816<T>(acc[mt], qf[mt][2 * kp + 1], b2, b3);
        }
      }
      // reference rounding points: GEMM output -> dtype, * 1/sqrt(d) in fp32 -> dtype
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int mt = q >> 1, hf = q & 1;
        const uint32_t p1 = T::pack(acc[mt][hf * 2], acc[mt][hf * 2 + 1]);
        uint32_t p2 = T::pack(T::lo(p1) * p.scale, T::hi(p1) * p.scale);
        if (!kPass2 && needs_mask) {
          const int key = key0 + kb * 8 + (lane & 3) * 2;
          const int lim = row_i[q] >= 0 ? nscore + row_i[q] : p.L - 1;  // last visible key of the row
          const int last = min(lim, p.L - 1);
          if (key > last) p2 = (p2 & 0xffff0000u) | (T::kNegInfPair & 0xffffu);
          if (key + 1 > last) p2 = (p2 & 0x0000ffffu) | (T::kNegInfPair & 0xffff0000u);
        }
        r2p[nb][q] = p2;
      }

      if constexpr (kPass2) {
        // probabilities -> dtype, summed over the 16 rows of each m-tile
        float* cs = colsum + (t & 1) * (SLABS * 2) * kTileN;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int q = mt * 2 + hf;
            const float e0 = fast_exp2(fmaf(T::lo(r2p[nb][q]), kLog2e, c[q]));
            const float e1 = fast_exp2(fmaf(T::hi(r2p[nb][q]), kLog2e, c[q]));
            const uint32_t pp = T::pack(e0, e1);
            s0 += T::lo(pp);
            s1 += T::hi(pp);
          }
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          }
          if (lane < 4) {
            float2* dst = reinterpret_cast<float2*>(cs + (slab * 2 + mt) * kTileN + kb * 8 + lane * 2);
            *dst = make_float2(s0, s1);
          }
        }
      }
    }

    if constexpr (!kPass2) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t pm = r2p[0][q];
#pragma unroll
        for (int nb = 1; nb < TL::NB; ++nb) pm = T::max2(pm, r2p[nb][q]);
        const float m_new = fmaxf(m_run[q], fmaxf(T::lo(pm), T::hi(pm)));
        float z = z_run[q] * fast_exp2((m_run[q] - m_new) * kLog2e);
        const float ml = -m_new * kLog2e;
#pragma unroll
        for (int nb = 0; nb < TL::NB; ++nb) {
          z += fast_exp2(fmaf(T::lo(r2p[nb][q]), kLog2e, ml));
          z += fast_exp2(fmaf(T::hi(r2p[nb][q]), kLog2e, ml));
        }
        m_run[q] = m_new;
        z_run[q] = z;
      }
    } else {
      __syncthreads();
      const float* cs = colsum + (t & 1) * (SLABS * 2) * kTileN;
      const int mt_per_head = p.wp / 16;
      for (int e = tid; e < p.heads_per_blk * kTileN; e += TL::NT) {
        const int hl = e / kTileN, kk = e % kTileN;
        const int head = rb * p.heads_per_blk + hl;
        const int key = key0 + kk;
        if (head < p.groups && key < nscore) {
          float s = 0.f;
          for (int m = 0; m < mt_per_head; ++m) s += cs[(hl * mt_per_head + m) * kTileN + kk];
          if (p.reduce_mean) s = div_exact(s, (float)p.w, 1.0f / (float)p.w);
          p.out[((int64_t)b * p.Hq + kvh * p.groups + head) * nscore + key] = T::round1(s);
        }
      }
    }
  }

  if constexpr (!kPass2) {
    // combine the four lanes that share a row, then the WN warps that split the key axis
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int o = 1; o < 4; o <<= 1) {
        const float mo = __shfl_xor_sync(0xffffffffu, m_run[q], o);
        const float zo = __shfl_xor_sync(0xffffffffu, z_run[q], o);
        const float mn = fmaxf(m_run[q], mo);
        z_run[q] = z_run[q] * fast_exp2((m_run[q] - mn) * kLog2e) + zo * fast_exp2((mo - mn) * kLog2e);
        m_run[q] = mn;
      }
    }
    if ((lane & 3) == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = slab * 32 + (q >> 1) * 16 + (q & 1) * 8 + (lane >> 2);
        wstat[wn * TL::ROWS + r] = make_float2(m_run[q], z_run[q]);
      }
    }
    __syncthreads();
    for (int r = tid; r < TL::ROWS; r += TL::NT) {
      float2 a = wstat[r];
#pragma unroll
      for (int x = 1; x < TL::WN; ++x) {
        const float2 o = wstat[x * TL::ROWS + r];
        const float mn = fmaxf(a.x, o.x);
        a.y = a.y * fast_exp2((a.x - mn) * kLog2e) + o.y * fast_exp2((o.x - mn) * kLog2e);
        a.x = mn;
      }
      p.stats[((int64_t)ub * p.nchunks + chunk) * TL::ROWS + r] = a;
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------
struct ScorePlan {
  int wp, heads_per_blk, nblk, slabs, rows_pad, chunk_keys, nchunks;
  size_t stats_bytes;
};

static ScorePlan make_plan(int B, int Hq, int Hkv, int L, int w) {
  ScorePlan pl{};
  const int g = Hq / Hkv;
  pl.wp = (w + 15) / 16 * 16;
  pl.heads_per_blk = 256 / pl.wp < g ? 256 / pl.wp : g;
  if (pl.heads_per_blk < 1) pl.heads_per_blk = 1;
  pl.nblk = (g + pl.heads_per_blk - 1) / pl.heads_per_blk;
  const int rows = pl.heads_per_blk * pl.wp;
  pl.slabs = rows <= 32 ? 1 : rows <= 64 ? 2 : rows <= 128 ? 4 : 8;
  pl.rows_pad = pl.slabs * 32;
  const int units = B * Hkv * pl.nblk;
  int want = (4 * sm_count() + units - 1) / units;  // ~4 CTAs per SM in flight
  if (want < 1) want = 1;
  int ck = (L + want - 1) / want;
  ck = (ck + kTileN - 1) / kTileN * kTileN;
  if (ck < 4 * kTileN) ck = 4 * kTileN;
  pl.chunk_keys = ck;
  pl.nchunks = (L + ck - 1) / ck;
  pl.stats_bytes = (size_t)units * pl.nchunks * pl.rows_pad * sizeof(float2);
  return pl;
}

size_t window_scores_mma_workspace_bytes(int B, int Hq, int Hkv, int L, int d, int w) {
  (void)d;
  return (make_plan(B, Hq, Hkv, L, w).stats_bytes + 255) & ~(size_t)255;
}

template <typename T, int SLABS, int D>
static int launch_both(const ScoreParams& sp, int units_blk, cudaStream_t stream) {
  const size_t smem = score_smem_bytes<D>(SLABS);
  auto k1 = window_score_mma_kernel<T, SLABS, D, false>;
  auto k2 = window_score_mma_kernel<T, SLABS, D, true>;
  KVC_SET_MAX_SMEM(k1, (int)smem);
  KVC_SET_MAX_SMEM(k2, (int)smem);
  dim3 grid(sp.nchunks, units_blk);
  k1<<<grid, Tiling<SLABS>::NT, smem, stream>>>(sp);
  if (int rc = launch_status("window_score_mma_kernel<stats>")) return rc;
  k2<<<grid, Tiling<SLABS>::NT, smem, stream>>>(sp);
  return launch_status("window_score_mma_kernel<scores>");
}

template <typename T, int D>
static int dispatch_slabs(const ScoreParams& sp, int slabs, int units_blk, cudaStream_t stream) {
  switch (slabs) {
    case 1: return launch_both<T, 1, D>(sp, units_blk, stream);
    case 2: return launch_both<T, 2, D>(sp, units_blk, stream);
    case 4: return launch_both<T, 4, D>(sp, units_blk, stream);
    default: return launch_both<T, 8, D>(sp, units_blk, stream);
  }
}

int window_scores_mma(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const void* k, int64_t k_sb,
                      int64_t k_sh, int64_t k_sl, int B, int Hq, int Hkv, int L, int d, int w, int dtype, int reduce,
                      void* out_scores, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  KVC_REQUIRE(d == 64 || d == 128, KVC_ERR_UNSUPPORTED, "head_dim %d: the scoring kernels implement 64 and 128", d);
  KVC_REQUIRE(w >= 1 && w <= 256, KVC_ERR_UNSUPPORTED, "window_size %d outside [1,256]", w);
  KVC_REQUIRE(((q_sb | q_sh | q_sl) & 1) == 0 && (reinterpret_cast<uintptr_t>(q) & 3) == 0, KVC_ERR_UNSUPPORTED,
              "query strides must be even and the pointer 4-byte aligned");
  KVC_REQUIRE(((k_sb | k_sh | k_sl) & 7) == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0, KVC_ERR_UNSUPPORTED,
              "key strides must be multiples of 8 elements and the pointer 16-byte aligned");
  const ScorePlan pl = make_plan(B, Hq, Hkv, L, w);
  KVC_REQUIRE(workspace_bytes >= pl.stats_bytes, KVC_ERR_WORKSPACE, "window_scores workspace: have %zu need %zu",
              workspace_bytes, pl.stats_bytes);
  ScoreParams sp{};
  sp.q = (const uint16_t*)q; sp.q_sb = q_sb; sp.q_sh = q_sh; sp.q_sl = q_sl;
  sp.k = (const uint16_t*)k; sp.k_sb = k_sb; sp.k_sh = k_sh; sp.k_sl = k_sl;
  sp.B = B; sp.Hq = Hq; sp.Hkv = Hkv; sp.L = L; sp.w = w; sp.wp = pl.wp; sp.groups = Hq / Hkv;
  sp.heads_per_blk = pl.heads_per_blk; sp.nblk = pl.nblk;
  sp.chunk_keys = pl.chunk_keys; sp.nchunks = pl.nchunks;
  sp.scale = 1.0f / sqrtf((float)d);
  sp.reduce_mean = reduce == KVC_REDUCE_MEAN;
  sp.stats = (float2*)workspace;
  sp.out = (uint16_t*)out_scores;
  const int units_blk = B * Hkv * pl.nblk;
  if (dtype == KVC_BF16)
    return d == 128 ? dispatch_slabs<Bf16, 128>(sp, pl.slabs, units_blk, stream)
                    : dispatch_slabs<Bf16, 64>(sp, pl.slabs, units_blk, stream);
  return d == 128 ? dispatch_slabs<Fp16, 128>(sp, pl.slabs, units_blk, stream)
                  : dispatch_slabs<Fp16, 64>(sp, pl.slabs, units_blk, stream);
}

}  // namespace kvc
