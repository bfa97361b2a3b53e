// Fused edge-softmax attention aggregation (GAT) for sm_100a.
//
// Reference: GATConv's vertex program (stgraph/nn/pytorch/static/gat_conv.py:48-56) compiles to two
// forward kernels with a materialised [E,H,1] score tensor (K0: scores + row sums, K1: weighted sum)
// and one backward kernel that accumulates d_el / d_er with per-lane atomicAdd, E*H*D of them inside
// the edge loop.  As shipped the program even degenerates to a mean (trap
// T2); the stock program therefore runs through the generic VM kernel, bit-faithful to the trace.
// These kernels are the genuine edge softmax a fixed program would compute:
//   forward : ONE pass per destination row with an online softmax (running max / sum per head),
//             nothing of size E is written; row max and row sum ([N,H]) are kept for backward.
//   backward: alpha is recomputed from (el, er, max, sum).  Pass A walks the in-edge CSR
//             (destination-parallel): dot[v,h] = <dout[v,h,:], out[v,h,:]>, d_er.  Pass B walks the
//             out-edge CSR (source-parallel): d_feat, d_el.  Every output row is owned by one lane
//             group: no atomics, deterministic.
// A lane group owns a row; a lane owns VEC consecutive floats of the flattened [H*D] row per chunk,
// so one neighbour row is one coalesced (128-bit when D % 4 == 0) load.  HBM-roofline kernels;
// algorithmic bytes: fwd 4*(2*N*H*D + 4*N*H + E + N+1), bwd 4*(4*N*H*D + 6*N*H + 2*(E+N+1)).
#include "common.cuh"

namespace stg {
namespace {

constexpr int kGatThreads = 256;

template <int VEC>
__device__ __forceinline__ typename VecT<VEC>::type ldv(const float* p) {
  using T = typename VecT<VEC>::type;
  return __ldg(reinterpret_cast<const T*>(p));
}
template <int VEC>
__device__ __forceinline__ void stv(float* p, typename VecT<VEC>::type v) {
  using T = typename VecT<VEC>::type;
  *reinterpret_cast<T*>(p) = v;
}
__device__ __forceinline__ float dotv(float a, float b) { return a * b; }
__device__ __forceinline__ float dotv(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float dotv(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float lrelu(float x, float slope) { return x > 0.f ? x : slope * x; }

struct GatParams {
  const int32_t* __restrict__ row_off;
  const int32_t* __restrict__ col;
  int num_rows;
  int heads, dim, hd;        // hd = heads*dim
  int lph;                   // lanes per head inside a chunk (dim / VEC), power of two
  float slope;
  const float* __restrict__ el;     // [N,H]
  const float* __restrict__ er;     // [N,H]
  const float* __restrict__ feat;   // [N,H,D]
  const float* __restrict__ out;    // [N,H,D]  (backward)
  const float* __restrict__ gout;   // [N,H,D]  (backward)
  float* __restrict__ row_max;      // [N,H]
  float* __restrict__ row_sum;      // [N,H]
  float* __restrict__ dot;          // [N,H]    (backward scratch)
  float* __restrict__ o_vec;        // forward: out; backward pass B: d_feat
  float* __restrict__ o_head;       // backward: d_er (pass A) / d_el (pass B)
};

#define GAT_PROLOGUE()                                                                              \
  using T = typename VecT<VEC>::type;                                                               \
  constexpr int GROUPS_PER_WARP = 32 / GROUP;                                                       \
  const int lane = threadIdx.x & 31;                                                                \
  const int gl = lane & (GROUP - 1);                                                                \
  const unsigned gmask = (GROUP == 32) ? 0xffffffffu : (((1u << GROUP) - 1u) << (lane & ~(GROUP - 1))); \
  const int warp = blockIdx.x * (kGatThreads / 32) + (threadIdx.x >> 5);                            \
  const int row = warp * GROUPS_PER_WARP + lane / GROUP;                                            \
  if (row >= p.num_rows) return;                                                                    \
  const int beg = __ldg(p.row_off + row), end = __ldg(p.row_off + row + 1);                         \
  int off[NACC], hk[NACC];                                                                          \
  bool act[NACC];                                                                                   \
  _Pragma("unroll") for (int k = 0; k < NACC; ++k) {                                                \
    off[k] = (gl + k * GROUP) * VEC;                                                                \
    act[k] = off[k] < p.hd;                                                                         \
    hk[k] = act[k] ? off[k] / p.dim : 0;                                                            \
  }

// sum over the lanes of one head (lph consecutive lanes, power of two)
__device__ __forceinline__ float head_sum(float v, int lph, unsigned gmask, int width) {
  for (int o = 1; o < lph; o <<= 1) v += __shfl_xor_sync(gmask, v, o, width);
  return v;
}

template <int VEC, int GROUP, int NACC>
__global__ void __launch_bounds__(kGatThreads) gat_fwd_kernel(const GatParams p) {
  GAT_PROLOGUE();
  float m[NACC], s[NACC], erk[NACC];
  T acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    m[k] = -INFINITY;
    s[k] = 0.f;
    zero_vec(acc[k]);
    erk[k] = act[k] ? __ldg(p.er + static_cast<size_t>(row) * p.heads + hk[k]) : 0.f;
  }
  for (int base = beg; base < end; base += GROUP) {
    const int n = min(GROUP, end - base);
    const int my_c = (gl < n) ? ld_stream(p.col + base + gl) : 0;
#pragma unroll 2
    for (int j = 0; j < n; ++j) {
      const int c = __shfl_sync(gmask, my_c, j, GROUP);
#pragma unroll
      for (int k = 0; k < NACC; ++k) {
        if (!act[k]) continue;
        const float sc = lrelu(__ldg(p.el + static_cast<size_t>(c) * p.heads + hk[k]) + erk[k], p.slope);
        const T v = ldv<VEC>(p.feat + static_cast<size_t>(c) * p.hd + off[k]);
        if (sc > m[k]) {                       // new running max: rescale what has been accumulated
          const float r = __expf(m[k] - sc);   // exp(-inf) = 0 on the first edge
          s[k] *= r;
          scale_vec(acc[k], r);
          m[k] = sc;
        }
        const float pe = expf(sc - m[k]);
        s[k] += pe;
        fma_vec(acc[k], pe, v);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    if (!act[k]) continue;
    const float inv = s[k] > 0.f ? 1.f / s[k] : 0.f;
    scale_vec(acc[k], inv);
    stv<VEC>(p.o_vec + static_cast<size_t>(row) * p.hd + off[k], acc[k]);
    if (off[k] % p.dim == 0) {
      p.row_max[static_cast<size_t>(row) * p.heads + hk[k]] = (end > beg) ? m[k] : 0.f;
      p.row_sum[static_cast<size_t>(row) * p.heads + hk[k]] = s[k];
    }
  }
}

// SRC_PARALLEL = false: pass A (rows = destinations): dot, d_er.
// SRC_PARALLEL = true : pass B (rows = sources): d_feat, d_el.
template <int VEC, int GROUP, int NACC, bool SRC_PARALLEL>
__global__ void __launch_bounds__(kGatThreads) gat_bwd_kernel(const GatParams p) {
  GAT_PROLOGUE();
  const size_t rh = static_cast<size_t>(row) * p.heads;
  T cen[NACC], accv[NACC];
  float acch[NACC], c_a[NACC], c_m[NACC], c_inv[NACC], c_dot[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    acch[k] = 0.f;
    zero_vec(accv[k]);
    zero_vec(cen[k]);
    c_a[k] = c_m[k] = c_inv[k] = c_dot[k] = 0.f;
    float part = 0.f;
    const size_t ro = static_cast<size_t>(row) * p.hd + off[k];
    if (act[k]) {
      if (SRC_PARALLEL) {
        cen[k] = ldv<VEC>(p.feat + ro);
        c_a[k] = __ldg(p.el + rh + hk[k]);
      } else {
        cen[k] = ldv<VEC>(p.gout + ro);
        c_a[k] = __ldg(p.er + rh + hk[k]);
        c_m[k] = __ldg(p.row_max + rh + hk[k]);
        const float sv = __ldg(p.row_sum + rh + hk[k]);
        c_inv[k] = sv > 0.f ? 1.f / sv : 0.f;
        part = dotv(cen[k], ldv<VEC>(p.out + ro));
      }
    }
    if (!SRC_PARALLEL) {                       // every lane of the group takes part in the shuffles
      c_dot[k] = head_sum(part, p.lph, gmask, GROUP);
      if (act[k] && off[k] % p.dim == 0) p.dot[rh + hk[k]] = c_dot[k];
    }
  }
  for (int base = beg; base < end; base += GROUP) {
    const int n = min(GROUP, end - base);
    const int my_c = (gl < n) ? ld_stream(p.col + base + gl) : 0;
#pragma unroll 2
    for (int j = 0; j < n; ++j) {
      const int c = __shfl_sync(gmask, my_c, j, GROUP);
      const size_t ch = static_cast<size_t>(c) * p.heads;
#pragma unroll
      for (int k = 0; k < NACC; ++k) {
        float part = 0.f, pre = 0.f, mm = 0.f, inv = 0.f, dt = 0.f;
        T nb;
        zero_vec(nb);
        if (act[k]) {
          if (SRC_PARALLEL) {      // neighbour = destination: its er, max, sum, dot, dout row
            nb = ldv<VEC>(p.gout + static_cast<size_t>(c) * p.hd + off[k]);
            pre = c_a[k] + __ldg(p.er + ch + hk[k]);
            mm = __ldg(p.row_max + ch + hk[k]);
            const float sv = __ldg(p.row_sum + ch + hk[k]);
            inv = sv > 0.f ? 1.f / sv : 0.f;
            dt = __ldg(p.dot + ch + hk[k]);
          } else {                 // neighbour = source: its el and feature row
            nb = ldv<VEC>(p.feat + static_cast<size_t>(c) * p.hd + off[k]);
            pre = __ldg(p.el + ch + hk[k]) + c_a[k];
            mm = c_m[k];
            inv = c_inv[k];
            dt = c_dot[k];
          }
          part = dotv(cen[k], nb);
        }
        const float dalpha = head_sum(part, p.lph, gmask, GROUP);      // <dout[v,h,:], feat[u,h,:]>
        if (act[k]) {
          const float alpha = expf(lrelu(pre, p.slope) - mm) * inv;
          const float g = alpha * (dalpha - dt) * (pre > 0.f ? 1.f : p.slope);
          acch[k] += g;
          if (SRC_PARALLEL) fma_vec(accv[k], alpha, nb);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    if (!act[k]) continue;
    if (SRC_PARALLEL) stv<VEC>(p.o_vec + static_cast<size_t>(row) * p.hd + off[k], accv[k]);
    if (off[k] % p.dim == 0) p.o_head[rh + hk[k]] = acch[k];
  }
}

template <int VEC, int GROUP, int NACC>
int launch_gat(const GatParams& p, int which, cudaStream_t s) {
  const int rows_per_block = (kGatThreads / 32) * (32 / GROUP);
  const int blocks = (p.num_rows + rows_per_block - 1) / rows_per_block;
  if (blocks <= 0) return STG_OK;
  if (which == 0) gat_fwd_kernel<VEC, GROUP, NACC><<<blocks, kGatThreads, 0, s>>>(p);
  else if (which == 1) gat_bwd_kernel<VEC, GROUP, NACC, false><<<blocks, kGatThreads, 0, s>>>(p);
  else gat_bwd_kernel<VEC, GROUP, NACC, true><<<blocks, kGatThreads, 0, s>>>(p);
  STG_LAUNCH_CHECK("gat kernel");
  return STG_OK;
}

template <int VEC>
int dispatch_gat(const GatParams& p, int which, cudaStream_t s) {
  const int nvec = p.hd / VEC;
  if (nvec <= 1) return launch_gat<VEC, 1, 1>(p, which, s);
  if (nvec <= 2) return launch_gat<VEC, 2, 1>(p, which, s);
  if (nvec <= 4) return launch_gat<VEC, 4, 1>(p, which, s);
  if (nvec <= 8) return launch_gat<VEC, 8, 1>(p, which, s);
  if (nvec <= 16) return launch_gat<VEC, 16, 1>(p, which, s);
  if (nvec <= 32) return launch_gat<VEC, 32, 1>(p, which, s);
  if (nvec <= 64) return launch_gat<VEC, 32, 2>(p, which, s);
  if (nvec <= 128) return launch_gat<VEC, 32, 4>(p, which, s);
  set_error("heads*dim = %d is wider than the fused GAT kernels support (512 floats with dim %% 4 == 0)", p.hd);
  return STG_ERR_UNSUPPORTED;
}

int run_gat(GatParams p, int which, bool aligned, cudaStream_t s) {
  const int vec = (p.dim % 4 == 0 && aligned) ? 4 : 1;
  p.lph = p.dim / vec;
  if ((p.lph & (p.lph - 1)) != 0 || p.lph > 32) {
    set_error("fused GAT kernels need dim/%d = %d lanes per head to be a power of two <= 32 (dim = %d)", vec, p.lph,
              p.dim);
    return STG_ERR_UNSUPPORTED;
  }
  return vec == 4 ? dispatch_gat<4>(p, which, s) : dispatch_gat<1>(p, which, s);
}

int check_view(const StgCsrView* g) {
  STG_CHECK_ARG(g && g->row_offset, "graph view / row_offset is NULL");
  STG_CHECK_ARG(g->num_edges == 0 || g->column_indices, "column_indices is NULL");
  return STG_OK;
}

}  // namespace
}  // namespace stg

using namespace stg;

STG_API int stg_gat_softmax_fwd_f32(const StgCsrView* g_in, const float* el, const float* er, const float* feat,
                                    int32_t heads, int32_t dim, float slope, float* out, float* row_max,
                                    float* row_sum, void* stream) {
  int rc = check_view(g_in);
  if (rc != STG_OK) return rc;
  STG_CHECK_ARG(heads > 0 && dim > 0, "heads and dim must be positive");
  if (g_in->num_nodes == 0) return STG_OK;
  STG_CHECK_ARG(el && er && feat && out && row_max && row_sum, "NULL tensor");
  GatParams p{};
  p.row_off = g_in->row_offset;
  p.col = g_in->column_indices;
  p.num_rows = g_in->num_nodes;
  p.heads = heads;
  p.dim = dim;
  p.hd = heads * dim;
  p.slope = slope;
  p.el = el;
  p.er = er;
  p.feat = feat;
  p.row_max = row_max;
  p.row_sum = row_sum;
  p.o_vec = out;
  return run_gat(p, 0, aligned16(feat) && aligned16(out), as_stream(stream));
}

STG_API int stg_gat_softmax_bwd_f32(const StgCsrView* g_in, const StgCsrView* g_out, const float* el, const float* er,
                                    const float* feat, const float* out, const float* grad_out, const float* row_max,
                                    const float* row_sum, int32_t heads, int32_t dim, float slope, float* d_feat,
                                    float* d_el, float* d_er, float* dot_scratch, void* stream) {
  int rc = check_view(g_in);
  if (rc != STG_OK) return rc;
  rc = check_view(g_out);
  if (rc != STG_OK) return rc;
  STG_CHECK_ARG(heads > 0 && dim > 0, "heads and dim must be positive");
  STG_CHECK_ARG(g_in->num_nodes == g_out->num_nodes, "forward / backward views disagree on the node count");
  if (g_in->num_nodes == 0) return STG_OK;
  STG_CHECK_ARG(el && er && feat && out && grad_out && row_max && row_sum && d_feat && d_el && d_er && dot_scratch,
                "NULL tensor");
  GatParams p{};
  p.num_rows = g_in->num_nodes;
  p.heads = heads;
  p.dim = dim;
  p.hd = heads * dim;
  p.slope = slope;
  p.el = el;
  p.er = er;
  p.feat = feat;
  p.out = out;
  p.gout = grad_out;
  p.row_max = const_cast<float*>(row_max);
  p.row_sum = const_cast<float*>(row_sum);
  p.dot = dot_scratch;
  const bool al = aligned16(feat) && aligned16(out) && aligned16(grad_out) && aligned16(d_feat);
  cudaStream_t s = as_stream(stream);
  // pass A: destination-parallel on the in-edge view -> dot, d_er
  p.row_off = g_in->row_offset;
  p.col = g_in->column_indices;
  p.o_vec = nullptr;
  p.o_head = d_er;
  rc = run_gat(p, 1, al, s);
  if (rc != STG_OK) return rc;
  // pass B: source-parallel on the out-edge view -> d_feat, d_el
  p.row_off = g_out->row_offset;
  p.col = g_out->column_indices;
  p.o_vec = d_feat;
  p.o_head = d_el;
  return run_gat(p, 2, al, s);
}
