// Weight-gradient GEMM of the TGCN cell for sm_100a:  C[K, Nc] = A[M, K]^T * B[M, Nc]  (+ column sums of B),
// fp32, M (the number of vertices) in the millions, K and Nc a few dozen to a few hundred.
//
// Reference: the three Linear(2H, H) layers and the GCNConv weights of stgraph/nn/pytorch/temporal/tgcn.py:16-47 get
// their gradients from torch autograd as cuBLAS GEMMs with a [K, Nc] result and a reduction over all M vertices.  For
// that shape cuBLAS picks a SIMT kernel with a handful of CTAs plus a split-K reduce (config 4, N = 10^6, H = 64:
// 0.59 ms per call = 14 % of the HBM bandwidth the two operands need, profiles/r02_config4_profile.json), eight calls
// per time step.  Here the M rows are split into slabs over twice the SM count; a CTA streams its slab through shared
// memory (16 rows per stage, double-buffered through registers) and keeps a 64 x 64 ... 32 x 192 tile of C in
// registers, a 4 x 4 ... 2 x 12 micro-tile per thread; slab partials go to a workspace and a second kernel sums them in
// slab order: deterministic, no atomics.  The bias gradient (column sums of B) rides along in the threads that own
// the first micro-tile row.  Operands may be column blocks of wider matrices (leading dimensions lda / ldb).
// Bound by fp32 FMA issue (16 FFMA per two shared-memory loads), not by HBM; exact fp32 (no TF32), sums in a
// different order than cuBLAS.
#include "common.cuh"

#include <algorithm>

namespace stg {
namespace {

constexpr int kTnThreads = 256;
constexpr int kTnRows = 16;      // rows of A / B per shared-memory stage

struct TnParams {
  const float* __restrict__ A;
  const float* __restrict__ B;
  int64_t lda, ldb, M;
  int K, Nc;
  float* __restrict__ part;      // [slabs, K, Nc] (or C itself when there is one slab)
  float* __restrict__ part_cs;   // [slabs, Nc] or NULL
  int64_t rows_per_slab;
  int vec_a, vec_b;              // 128-bit global loads allowed
};

// one 4-float piece of a stage: row `r` (global), columns [c, c+4) of a matrix with `cols` valid columns
__device__ __forceinline__ float4 load4(const float* __restrict__ base, int64_t ld, int64_t r, int64_t rows, int c, int cols,
                                        int vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r >= rows || c >= cols) return v;
  const float* p = base + r * ld + c;
  if (vec && c + 4 <= cols) return __ldcs(reinterpret_cast<const float4*>(p));
  v.x = __ldcs(p);
  if (c + 1 < cols) v.y = __ldcs(p + 1);
  if (c + 2 < cols) v.z = __ldcs(p + 2);
  if (c + 3 < cols) v.w = __ldcs(p + 3);
  return v;
}

// MT x NT micro-tile per thread, 16 x 16 threads: a CTA owns a (16*MT) x (16*NT) tile of C.
template <int MT, int NT>
__global__ void __launch_bounds__(kTnThreads) gemm_tn_kernel(const TnParams p) {
  constexpr int TK = 16 * MT, TN = 16 * NT;
  constexpr int A_VEC_PER_ROW = TK / 4;                     // float4 pieces per stage row of A
  constexpr int A_LOADERS = kTnRows * A_VEC_PER_ROW;        // threads that load a piece of A (256 or 128)
  constexpr int B_PIECES = NT / 4;                          // float4 pieces of a B stage per thread
  static_assert(NT % 4 == 0 && (MT == 2 || MT == 4), "micro-tile shape");
  __shared__ __align__(16) float As[2][kTnRows][TK];
  __shared__ __align__(16) float Bs[2][kTnRows][TN];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int k0 = blockIdx.y * TK, n0 = blockIdx.z * TN;
  const int64_t m_beg = static_cast<int64_t>(blockIdx.x) * p.rows_per_slab;
  const int64_t m_end = min(p.M, m_beg + p.rows_per_slab);
  const int ar = tid / A_VEC_PER_ROW, ac = (tid % A_VEC_PER_ROW) * 4;       // this thread's piece of an A stage
  const int br = tid >> 4, bc = (tid & 15) * 4;                              // ... of a B stage: pieces bc + 64 * q
  const float* Ab = p.A + k0;
  const float* Bb = p.B + n0;
  const int a_cols = p.K - k0, b_cols = p.Nc - n0;

  float acc[MT][NT];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j] = 0.f;
  float cs[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) cs[j] = 0.f;

  float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb[B_PIECES];
  auto fetch = [&](int64_t m) {
    if (tid < A_LOADERS) ra = load4(Ab, p.lda, m + ar, m_end, ac, a_cols, p.vec_a);
#pragma unroll
    for (int q = 0; q < B_PIECES; ++q) rb[q] = load4(Bb, p.ldb, m + br, m_end, bc + 64 * q, b_cols, p.vec_b);
  };
  fetch(m_beg);
  int buf = 0;
  for (int64_t m = m_beg; m < m_end; m += kTnRows) {
    if (tid < A_LOADERS) *reinterpret_cast<float4*>(&As[buf][ar][ac]) = ra;
#pragma unroll
    for (int q = 0; q < B_PIECES; ++q) *reinterpret_cast<float4*>(&Bs[buf][br][bc + 64 * q]) = rb[q];
    __syncthreads();                                        // one barrier per stage: the other buffer was last read a stage ago
    if (m + kTnRows < m_end) fetch(m + kTnRows);            // next stage in flight while this one is consumed
#pragma unroll
    for (int r = 0; r < kTnRows; ++r) {
      float a[MT];
      if constexpr (MT == 4) {
        const float4 t = *reinterpret_cast<const float4*>(&As[buf][r][ty * 4]);
        a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w;
      } else {
        const float2 t = *reinterpret_cast<const float2*>(&As[buf][r][ty * 2]);
        a[0] = t.x; a[1] = t.y;
      }
      // thread tx owns columns 64*q + 4*tx .. +3 of the tile: the 16 threads of a row read 256 contiguous bytes per piece
#pragma unroll
      for (int q = 0; q < B_PIECES; ++q) {
        const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][r][64 * q + tx * 4]);
#pragma unroll
        for (int i = 0; i < MT; ++i) {
          acc[i][4 * q + 0] = fmaf(a[i], b.x, acc[i][4 * q + 0]);
          acc[i][4 * q + 1] = fmaf(a[i], b.y, acc[i][4 * q + 1]);
          acc[i][4 * q + 2] = fmaf(a[i], b.z, acc[i][4 * q + 2]);
          acc[i][4 * q + 3] = fmaf(a[i], b.w, acc[i][4 * q + 3]);
        }
        if (ty == 0) {                                      // rows past m_end were zero-filled
          cs[4 * q + 0] += b.x; cs[4 * q + 1] += b.y; cs[4 * q + 2] += b.z; cs[4 * q + 3] += b.w;
        }
      }
    }
    buf ^= 1;
  }
  float* out = p.part + static_cast<size_t>(blockIdx.x) * p.K * p.Nc;
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int k = k0 + ty * MT + i;
    if (k >= p.K) continue;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int n = n0 + 64 * (j / 4) + tx * 4 + (j % 4);
      if (n < p.Nc) out[static_cast<size_t>(k) * p.Nc + n] = acc[i][j];
    }
  }
  if (p.part_cs != nullptr && ty == 0 && blockIdx.y == 0) {
    float* oc = p.part_cs + static_cast<size_t>(blockIdx.x) * p.Nc;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int n = n0 + 64 * (j / 4) + tx * 4 + (j % 4);
      if (n < p.Nc) oc[n] = cs[j];
    }
  }
}

// C[i] = sum over slabs (in slab order) of part[s][i]; the same for the column sums
__global__ void __launch_bounds__(256) gemm_tn_reduce_kernel(const float* __restrict__ part, const float* __restrict__ part_cs,
                                                             int slabs, int kn, int nc, float* __restrict__ C,
                                                             float* __restrict__ colsum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kn) {
    float s = 0.f;
    for (int t = 0; t < slabs; ++t) s += part[static_cast<size_t>(t) * kn + i];
    C[i] = s;
  } else if (colsum != nullptr && i < kn + nc) {
    const int j = i - kn;
    float s = 0.f;
    for (int t = 0; t < slabs; ++t) s += part_cs[static_cast<size_t>(t) * nc + j];
    colsum[j] = s;
  }
}

struct TnPlan {
  int mt, nt, k_tiles, n_tiles, slabs;
  int64_t rows_per_slab;
};

// Tile shapes: 64 x 64 (4 x 4 per thread), 64 x 128 (4 x 8) for wide results, 32 x 64 / 128 / 192 for K <= 32 (2 x 4 / 8 / 12):
// a tile as wide as the result reads A once.
TnPlan plan_tn(int64_t M, int K, int Nc) {
  TnPlan pl;
  pl.mt = K <= 32 ? 2 : 4;
  pl.nt = Nc <= 64 ? 4 : (Nc <= 128 || pl.mt == 4 ? 8 : 12);
  const int tk = 16 * pl.mt, tn = 16 * pl.nt;
  pl.k_tiles = (K + tk - 1) / tk;
  pl.n_tiles = (Nc + tn - 1) / tn;
  const int tiles = pl.k_tiles * pl.n_tiles;
  const int64_t want = (2LL * sm_count() + tiles - 1) / tiles;              // about two CTAs per SM in total
  const int64_t most = std::max<int64_t>(1, M / (8 * kTnRows));             // at least 128 rows per slab
  pl.slabs = static_cast<int>(std::max<int64_t>(1, std::min(want, most)));
  pl.rows_per_slab = ((M + pl.slabs - 1) / pl.slabs + kTnRows - 1) / kTnRows * kTnRows;
  pl.slabs = static_cast<int>((M + pl.rows_per_slab - 1) / std::max<int64_t>(pl.rows_per_slab, 1));
  if (pl.slabs < 1) pl.slabs = 1;
  return pl;
}

}  // namespace
}  // namespace stg

using namespace stg;

STG_API size_t stg_gemm_tn_workspace_bytes(int64_t M, int32_t K, int32_t Nc) {
  if (M <= 0 || K <= 0 || Nc <= 0) return 0;
  const TnPlan pl = plan_tn(M, K, Nc);
  if (pl.slabs <= 1) return 0;
  return static_cast<size_t>(pl.slabs) * (static_cast<size_t>(K) * Nc + Nc) * sizeof(float);
}

STG_API int stg_gemm_tn_f32(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int32_t K, int32_t Nc,
                            float* C, float* colsum_b, void* workspace, size_t workspace_bytes, void* stream) {
  STG_CHECK_ARG(M >= 0 && K > 0 && Nc > 0, "bad shape (M=%lld, K=%d, Nc=%d)", static_cast<long long>(M), K, Nc);
  STG_CHECK_ARG(lda >= K && ldb >= Nc, "leading dimensions (%lld, %lld) smaller than the row lengths (%d, %d)",
                static_cast<long long>(lda), static_cast<long long>(ldb), K, Nc);
  STG_CHECK_ARG(C != nullptr, "C is NULL");
  STG_CHECK_ARG(static_cast<int64_t>(K) * Nc + Nc < (1LL << 31), "result too large");
  cudaStream_t s = as_stream(stream);
  if (M == 0) {
    STG_CUDA(cudaMemsetAsync(C, 0, static_cast<size_t>(K) * Nc * sizeof(float), s));
    if (colsum_b) STG_CUDA(cudaMemsetAsync(colsum_b, 0, static_cast<size_t>(Nc) * sizeof(float), s));
    return STG_OK;
  }
  STG_CHECK_ARG(A && B, "NULL operand");
  const TnPlan pl = plan_tn(M, K, Nc);
  const size_t need = stg_gemm_tn_workspace_bytes(M, K, Nc);
  STG_CHECK_ARG(need == 0 || (workspace != nullptr && workspace_bytes >= need),
                "workspace of %zu bytes needed (stg_gemm_tn_workspace_bytes), got %zu", need, workspace_bytes);
  TnParams p{};
  p.A = A;
  p.B = B;
  p.lda = lda;
  p.ldb = ldb;
  p.M = M;
  p.K = K;
  p.Nc = Nc;
  p.rows_per_slab = pl.rows_per_slab;
  p.vec_a = aligned16(A) && lda % 4 == 0;
  p.vec_b = aligned16(B) && ldb % 4 == 0;
  const size_t kn = static_cast<size_t>(K) * Nc;
  if (pl.slabs == 1) {
    p.part = C;
    p.part_cs = colsum_b;
  } else {
    p.part = static_cast<float*>(workspace);
    p.part_cs = colsum_b ? p.part + static_cast<size_t>(pl.slabs) * kn : nullptr;
  }
  const dim3 grid(pl.slabs, pl.k_tiles, pl.n_tiles);
  if (pl.mt == 2 && pl.nt == 4) gemm_tn_kernel<2, 4><<<grid, kTnThreads, 0, s>>>(p);
  else if (pl.mt == 2 && pl.nt == 8) gemm_tn_kernel<2, 8><<<grid, kTnThreads, 0, s>>>(p);
  else if (pl.mt == 2) gemm_tn_kernel<2, 12><<<grid, kTnThreads, 0, s>>>(p);
  else if (pl.nt == 4) gemm_tn_kernel<4, 4><<<grid, kTnThreads, 0, s>>>(p);
  else gemm_tn_kernel<4, 8><<<grid, kTnThreads, 0, s>>>(p);
  STG_LAUNCH_CHECK("gemm_tn_kernel");
  if (pl.slabs > 1) {
    const int total = static_cast<int>(kn) + (colsum_b ? Nc : 0);
    gemm_tn_reduce_kernel<<<(total + 255) / 256, 256, 0, s>>>(p.part, p.part_cs, pl.slabs, static_cast<int>(kn), Nc, C,
                                                               colsum_b);
    STG_LAUNCH_CHECK("gemm_tn_reduce_kernel");
  }
  return STG_OK;
}
