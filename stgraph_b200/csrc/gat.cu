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
//             group (or one block for hub rows): no atomics, deterministic.
// A lane group owns a row; a lane owns VEC consecutive floats of the flattened [H*D] row per chunk,
// so one neighbour row is one coalesced (128-bit when D % 4 == 0) load; UNROLL neighbour rows are
// loaded before any of them is consumed (the online-softmax update is applied per batch, so the
// running max does not serialise the loads).  Rows longer than the view's hub threshold go to a
// block-per-row kernel: 16 warps take strided batches and their (max, sum, acc) partials are merged
// in a fixed order through shuffles + shared memory.
// HBM-roofline kernels; algorithmic bytes: fwd 4*(2*N*H*D + 4*N*H + E + N+1),
// bwd 4*(4*N*H*D + 6*N*H + 2*(E+N+1)).
#include "common.cuh"

namespace stg {
namespace {

constexpr int kGatThreads = 256;
constexpr int kGatHubThreads = 512;

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

__device__ __forceinline__ float shfl_xor_t(unsigned m, float v, int o) { return __shfl_xor_sync(m, v, o); }
__device__ __forceinline__ float2 shfl_xor_t(unsigned m, float2 v, int o) {
  return make_float2(__shfl_xor_sync(m, v.x, o), __shfl_xor_sync(m, v.y, o));
}
__device__ __forceinline__ float4 shfl_xor_t(unsigned m, float4 v, int o) {
  return make_float4(__shfl_xor_sync(m, v.x, o), __shfl_xor_sync(m, v.y, o), __shfl_xor_sync(m, v.z, o),
                     __shfl_xor_sync(m, v.w, o));
}

struct GatParams {
  const int32_t* __restrict__ row_off;
  const int32_t* __restrict__ col;
  const int32_t* __restrict__ hub_rows;
  const int32_t* __restrict__ hub_count;
  int hub_threshold, hub_capacity;
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

template <int VEC, int GROUP, int NACC>
struct Lanes {
  int gl;
  unsigned gmask;
  int off[NACC], hk[NACC];
  bool act[NACC];
  __device__ __forceinline__ void init(const GatParams& p) {
    const int lane = threadIdx.x & 31;
    gl = lane & (GROUP - 1);
    gmask = (GROUP == 32) ? 0xffffffffu : (((1u << GROUP) - 1u) << (lane & ~(GROUP - 1)));
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
      off[k] = (gl + k * GROUP) * VEC;
      act[k] = off[k] < p.hd;
      hk[k] = act[k] ? off[k] / p.dim : 0;
    }
  }
};

// sum over the lanes of one head (lph consecutive lanes, power of two)
__device__ __forceinline__ float head_sum(float v, int lph, unsigned gmask, int width) {
  for (int o = 1; o < lph; o <<= 1) v += __shfl_xor_sync(gmask, v, o, width);
  return v;
}

template <int NACC> struct UnrollFor { static constexpr int value = NACC >= 4 ? 2 : 4; };

// ---------------------------------------------------------------------------------- forward
template <int VEC, int GROUP, int NACC>
struct FwdState {
  using T = typename VecT<VEC>::type;
  float m[NACC], s[NACC];
  T acc[NACC];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
      m[k] = -INFINITY;
      s[k] = 0.f;
      zero_vec(acc[k]);
    }
  }
  // merge another partial (online-softmax combine); safe when either side is still empty (m = -inf)
  __device__ __forceinline__ void merge(int k, float m2, float s2, T a2) {
    const float mm = fmaxf(m[k], m2);
    const float f1 = (m[k] == -INFINITY) ? 0.f : __expf(m[k] - mm);
    const float f2 = (m2 == -INFINITY) ? 0.f : __expf(m2 - mm);
    s[k] = s[k] * f1 + s2 * f2;
    scale_vec(acc[k], f1);
    fma_vec(acc[k], f2, a2);
    m[k] = mm;
  }
};

template <int VEC, int GROUP, int NACC>
__device__ __forceinline__ void gat_fwd_edges(const GatParams& p, const Lanes<VEC, GROUP, NACC>& L, int row, int beg,
                                              int end, int first_batch, int batch_step,
                                              FwdState<VEC, GROUP, NACC>& st) {
  using T = typename VecT<VEC>::type;
  constexpr int U = UnrollFor<NACC>::value;
  float erk[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) erk[k] = L.act[k] ? __ldg(p.er + static_cast<size_t>(row) * p.heads + L.hk[k]) : 0.f;
  for (int base = beg + first_batch * GROUP; base < end; base += batch_step * GROUP) {
    const int n = min(GROUP, end - base);
    const int my_c = (L.gl < n) ? ld_stream(p.col + base + L.gl) : 0;
    for (int j = 0; j < n; j += U) {
      float sc[U][NACC];
      T v[U][NACC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = __shfl_sync(L.gmask, my_c, min(j + u, GROUP - 1), GROUP);
        const bool valid = (j + u) < n;
#pragma unroll
        for (int k = 0; k < NACC; ++k) {
          sc[u][k] = -INFINITY;
          zero_vec(v[u][k]);
          if (valid && L.act[k]) {
            sc[u][k] = lrelu(__ldg(p.el + static_cast<size_t>(c) * p.heads + L.hk[k]) + erk[k], p.slope);
            v[u][k] = ldv<VEC>(p.feat + static_cast<size_t>(c) * p.hd + L.off[k]);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < NACC; ++k) {
        if (!L.act[k]) continue;
        float mb = sc[0][k];
#pragma unroll
        for (int u = 1; u < U; ++u) mb = fmaxf(mb, sc[u][k]);
        if (mb > st.m[k]) {                       // new running max: rescale what has been accumulated
          const float r = __expf(st.m[k] - mb);   // exp(-inf) = 0 on the first batch
          st.s[k] *= r;
          scale_vec(st.acc[k], r);
          st.m[k] = mb;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float pe = expf(sc[u][k] - st.m[k]);   // exp(-inf) = 0 for padding edges
          st.s[k] += pe;
          fma_vec(st.acc[k], pe, v[u][k]);
        }
      }
    }
  }
}

template <int VEC, int GROUP, int NACC>
__device__ __forceinline__ void gat_fwd_store(const GatParams& p, const Lanes<VEC, GROUP, NACC>& L, int row,
                                              bool nonempty, FwdState<VEC, GROUP, NACC>& st) {
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    if (!L.act[k]) continue;
    const float inv = st.s[k] > 0.f ? 1.f / st.s[k] : 0.f;
    scale_vec(st.acc[k], inv);
    stv<VEC>(p.o_vec + static_cast<size_t>(row) * p.hd + L.off[k], st.acc[k]);
    if (L.off[k] % p.dim == 0) {
      p.row_max[static_cast<size_t>(row) * p.heads + L.hk[k]] = nonempty ? st.m[k] : 0.f;
      p.row_sum[static_cast<size_t>(row) * p.heads + L.hk[k]] = st.s[k];
    }
  }
}

template <int VEC, int GROUP, int NACC>
__global__ void __launch_bounds__(kGatThreads) gat_fwd_kernel(const GatParams p) {
  Lanes<VEC, GROUP, NACC> L;
  L.init(p);
  const int warp = blockIdx.x * (kGatThreads / 32) + (threadIdx.x >> 5);
  const int row = warp * (32 / GROUP) + (threadIdx.x & 31) / GROUP;
  if (row >= p.num_rows) return;
  const int beg = __ldg(p.row_off + row), end = __ldg(p.row_off + row + 1);
  if (p.hub_threshold > 0 && (end - beg) > p.hub_threshold) return;   // hub kernel owns this row
  FwdState<VEC, GROUP, NACC> st;
  st.init();
  gat_fwd_edges<VEC, GROUP, NACC>(p, L, row, beg, end, 0, 1, st);
  gat_fwd_store<VEC, GROUP, NACC>(p, L, row, end > beg, st);
}

template <int VEC, int GROUP, int NACC>
__global__ void __launch_bounds__(kGatHubThreads) gat_fwd_hub_kernel(const GatParams p) {
  using T = typename VecT<VEC>::type;
  constexpr int WARPS = kGatHubThreads / 32;
  constexpr int GPW = 32 / GROUP;
  __shared__ T s_acc[WARPS][GROUP * NACC];
  __shared__ float s_m[WARPS][GROUP * NACC], s_s[WARPS][GROUP * NACC];
  Lanes<VEC, GROUP, NACC> L;
  L.init(p);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n_hub = min(__ldg(p.hub_count), p.hub_capacity);
  for (int i = blockIdx.x; i < n_hub; i += gridDim.x) {
    const int row = __ldg(p.hub_rows + i);
    const int beg = __ldg(p.row_off + row), end = __ldg(p.row_off + row + 1);
    FwdState<VEC, GROUP, NACC> st;
    st.init();
    gat_fwd_edges<VEC, GROUP, NACC>(p, L, row, beg, end, wid * GPW + lane / GROUP, WARPS * GPW, st);
    // groups of one warp -> lanes [0, GROUP)
#pragma unroll
    for (int o = GROUP; o < 32; o <<= 1) {
#pragma unroll
      for (int k = 0; k < NACC; ++k) {
        const float m2 = __shfl_xor_sync(0xffffffffu, st.m[k], o);
        const float s2 = __shfl_xor_sync(0xffffffffu, st.s[k], o);
        const T a2 = shfl_xor_t(0xffffffffu, st.acc[k], o);
        st.merge(k, m2, s2, a2);
      }
    }
    if (lane < GROUP) {
#pragma unroll
      for (int k = 0; k < NACC; ++k) {
        s_acc[wid][k * GROUP + lane] = st.acc[k];
        s_m[wid][k * GROUP + lane] = st.m[k];
        s_s[wid][k * GROUP + lane] = st.s[k];
      }
    }
    __syncthreads();
    if (wid == 0 && lane < GROUP) {
#pragma unroll
      for (int k = 0; k < NACC; ++k)
        for (int w = 1; w < WARPS; ++w)
          st.merge(k, s_m[w][k * GROUP + lane], s_s[w][k * GROUP + lane], s_acc[w][k * GROUP + lane]);
      gat_fwd_store<VEC, GROUP, NACC>(p, L, row, end > beg, st);
    }
    __syncthreads();
  }
}

// --------------------------------------------------------------------------------- backward
// SRC_PARALLEL = false: pass A (rows = destinations): dot, d_er.
// SRC_PARALLEL = true : pass B (rows = sources): d_feat, d_el.
template <int VEC, int GROUP, int NACC>
struct BwdState {
  using T = typename VecT<VEC>::type;
  T cen[NACC], accv[NACC];
  float acch[NACC], c_a[NACC], c_m[NACC], c_inv[NACC], c_dot[NACC];
};

template <int VEC, int GROUP, int NACC, bool SRC_PARALLEL>
__device__ __forceinline__ void gat_bwd_prologue(const GatParams& p, const Lanes<VEC, GROUP, NACC>& L, int row,
                                                 BwdState<VEC, GROUP, NACC>& st, bool write_dot) {
  const size_t rh = static_cast<size_t>(row) * p.heads;
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    st.acch[k] = 0.f;
    zero_vec(st.accv[k]);
    zero_vec(st.cen[k]);
    st.c_a[k] = st.c_m[k] = st.c_inv[k] = st.c_dot[k] = 0.f;
    float part = 0.f;
    const size_t ro = static_cast<size_t>(row) * p.hd + L.off[k];
    if (L.act[k]) {
      if (SRC_PARALLEL) {
        st.cen[k] = ldv<VEC>(p.feat + ro);
        st.c_a[k] = __ldg(p.el + rh + L.hk[k]);
      } else {
        st.cen[k] = ldv<VEC>(p.gout + ro);
        st.c_a[k] = __ldg(p.er + rh + L.hk[k]);
        st.c_m[k] = __ldg(p.row_max + rh + L.hk[k]);
        const float sv = __ldg(p.row_sum + rh + L.hk[k]);
        st.c_inv[k] = sv > 0.f ? 1.f / sv : 0.f;
        part = dotv(st.cen[k], ldv<VEC>(p.out + ro));
      }
    }
    if (!SRC_PARALLEL) {                       // every lane of the group takes part in the shuffles
      st.c_dot[k] = head_sum(part, p.lph, L.gmask, GROUP);
      if (write_dot && L.act[k] && L.off[k] % p.dim == 0) p.dot[rh + L.hk[k]] = st.c_dot[k];
    }
  }
}

template <int VEC, int GROUP, int NACC, bool SRC_PARALLEL>
__device__ __forceinline__ void gat_bwd_edges(const GatParams& p, const Lanes<VEC, GROUP, NACC>& L, int beg, int end,
                                              int first_batch, int batch_step, BwdState<VEC, GROUP, NACC>& st) {
  using T = typename VecT<VEC>::type;
  constexpr int U = UnrollFor<NACC>::value;
  for (int base = beg + first_batch * GROUP; base < end; base += batch_step * GROUP) {
    const int n = min(GROUP, end - base);
    const int my_c = (L.gl < n) ? ld_stream(p.col + base + L.gl) : 0;
    for (int j = 0; j < n; j += U) {
      T nb[U][NACC];
      float pre[U][NACC], mm[U][NACC], inv[U][NACC], dt[U][NACC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = __shfl_sync(L.gmask, my_c, min(j + u, GROUP - 1), GROUP);
        const size_t ch = static_cast<size_t>(c) * p.heads;
        const bool valid = (j + u) < n;
#pragma unroll
        for (int k = 0; k < NACC; ++k) {
          zero_vec(nb[u][k]);
          pre[u][k] = mm[u][k] = inv[u][k] = dt[u][k] = 0.f;
          if (valid && L.act[k]) {
            if (SRC_PARALLEL) {      // neighbour = destination: its er, max, sum, dot, dout row
              nb[u][k] = ldv<VEC>(p.gout + static_cast<size_t>(c) * p.hd + L.off[k]);
              pre[u][k] = st.c_a[k] + __ldg(p.er + ch + L.hk[k]);
              mm[u][k] = __ldg(p.row_max + ch + L.hk[k]);
              const float sv = __ldg(p.row_sum + ch + L.hk[k]);
              inv[u][k] = sv > 0.f ? 1.f / sv : 0.f;
              dt[u][k] = __ldg(p.dot + ch + L.hk[k]);
            } else {                 // neighbour = source: its el and feature row
              nb[u][k] = ldv<VEC>(p.feat + static_cast<size_t>(c) * p.hd + L.off[k]);
              pre[u][k] = __ldg(p.el + ch + L.hk[k]) + st.c_a[k];
              mm[u][k] = st.c_m[k];
              inv[u][k] = st.c_inv[k];
              dt[u][k] = st.c_dot[k];
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool valid = (j + u) < n;
#pragma unroll
        for (int k = 0; k < NACC; ++k) {
          const float part = L.act[k] ? dotv(st.cen[k], nb[u][k]) : 0.f;
          const float dalpha = head_sum(part, p.lph, L.gmask, GROUP);      // <dout[v,h,:], feat[u,h,:]>
          if (valid && L.act[k]) {
            const float alpha = expf(lrelu(pre[u][k], p.slope) - mm[u][k]) * inv[u][k];
            const float g = alpha * (dalpha - dt[u][k]) * (pre[u][k] > 0.f ? 1.f : p.slope);
            st.acch[k] += g;
            if (SRC_PARALLEL) fma_vec(st.accv[k], alpha, nb[u][k]);
          }
        }
      }
    }
  }
}

template <int VEC, int GROUP, int NACC, bool SRC_PARALLEL>
__device__ __forceinline__ void gat_bwd_store(const GatParams& p, const Lanes<VEC, GROUP, NACC>& L, int row,
                                              BwdState<VEC, GROUP, NACC>& st) {
  const size_t rh = static_cast<size_t>(row) * p.heads;
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    if (!L.act[k]) continue;
    if (SRC_PARALLEL) stv<VEC>(p.o_vec + static_cast<size_t>(row) * p.hd + L.off[k], st.accv[k]);
    if (L.off[k] % p.dim == 0) p.o_head[rh + L.hk[k]] = st.acch[k];
  }
}

template <int VEC, int GROUP, int NACC, bool SRC_PARALLEL>
__global__ void __launch_bounds__(kGatThreads) gat_bwd_kernel(const GatParams p) {
  Lanes<VEC, GROUP, NACC> L;
  L.init(p);
  const int warp = blockIdx.x * (kGatThreads / 32) + (threadIdx.x >> 5);
  const int row = warp * (32 / GROUP) + (threadIdx.x & 31) / GROUP;
  if (row >= p.num_rows) return;
  const int beg = __ldg(p.row_off + row), end = __ldg(p.row_off + row + 1);
  if (p.hub_threshold > 0 && (end - beg) > p.hub_threshold) return;
  BwdState<VEC, GROUP, NACC> st;
  gat_bwd_prologue<VEC, GROUP, NACC, SRC_PARALLEL>(p, L, row, st, true);
  gat_bwd_edges<VEC, GROUP, NACC, SRC_PARALLEL>(p, L, beg, end, 0, 1, st);
  gat_bwd_store<VEC, GROUP, NACC, SRC_PARALLEL>(p, L, row, st);
}

template <int VEC, int GROUP, int NACC, bool SRC_PARALLEL>
__global__ void __launch_bounds__(kGatHubThreads) gat_bwd_hub_kernel(const GatParams p) {
  using T = typename VecT<VEC>::type;
  constexpr int WARPS = kGatHubThreads / 32;
  constexpr int GPW = 32 / GROUP;
  __shared__ T s_v[WARPS][GROUP * NACC];
  __shared__ float s_h[WARPS][GROUP * NACC];
  Lanes<VEC, GROUP, NACC> L;
  L.init(p);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n_hub = min(__ldg(p.hub_count), p.hub_capacity);
  for (int i = blockIdx.x; i < n_hub; i += gridDim.x) {
    const int row = __ldg(p.hub_rows + i);
    const int beg = __ldg(p.row_off + row), end = __ldg(p.row_off + row + 1);
    BwdState<VEC, GROUP, NACC> st;
    gat_bwd_prologue<VEC, GROUP, NACC, SRC_PARALLEL>(p, L, row, st, wid == 0 && lane < GROUP);
    gat_bwd_edges<VEC, GROUP, NACC, SRC_PARALLEL>(p, L, beg, end, wid * GPW + lane / GROUP, WARPS * GPW, st);
#pragma unroll
    for (int o = GROUP; o < 32; o <<= 1) {
#pragma unroll
      for (int k = 0; k < NACC; ++k) {
        st.acch[k] += __shfl_xor_sync(0xffffffffu, st.acch[k], o);
        add_vec(st.accv[k], shfl_xor_t(0xffffffffu, st.accv[k], o));
      }
    }
    if (lane < GROUP) {
#pragma unroll
      for (int k = 0; k < NACC; ++k) {
        s_v[wid][k * GROUP + lane] = st.accv[k];
        s_h[wid][k * GROUP + lane] = st.acch[k];
      }
    }
    __syncthreads();
    if (wid == 0 && lane < GROUP) {
#pragma unroll
      for (int k = 0; k < NACC; ++k)
        for (int w = 1; w < WARPS; ++w) {
          add_vec(st.accv[k], s_v[w][k * GROUP + lane]);
          st.acch[k] += s_h[w][k * GROUP + lane];
        }
      gat_bwd_store<VEC, GROUP, NACC, SRC_PARALLEL>(p, L, row, st);
    }
    __syncthreads();
  }
}

template <int VEC, int GROUP, int NACC>
int launch_gat(const GatParams& p, int which, cudaStream_t s) {
  const int rows_per_block = (kGatThreads / 32) * (32 / GROUP);
  const int blocks = (p.num_rows + rows_per_block - 1) / rows_per_block;
  if (blocks <= 0) return STG_OK;
  const bool hubs = p.hub_threshold > 0 && p.hub_rows != nullptr;
  const int hub_grid = 2 * sm_count();
  if (which == 0) {
    gat_fwd_kernel<VEC, GROUP, NACC><<<blocks, kGatThreads, 0, s>>>(p);
    if (hubs) gat_fwd_hub_kernel<VEC, GROUP, NACC><<<hub_grid, kGatHubThreads, 0, s>>>(p);
  } else if (which == 1) {
    gat_bwd_kernel<VEC, GROUP, NACC, false><<<blocks, kGatThreads, 0, s>>>(p);
    if (hubs) gat_bwd_hub_kernel<VEC, GROUP, NACC, false><<<hub_grid, kGatHubThreads, 0, s>>>(p);
  } else {
    gat_bwd_kernel<VEC, GROUP, NACC, true><<<blocks, kGatThreads, 0, s>>>(p);
    if (hubs) gat_bwd_hub_kernel<VEC, GROUP, NACC, true><<<hub_grid, kGatHubThreads, 0, s>>>(p);
  }
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

void bind_view(GatParams& p, const StgCsrView* g) {
  p.row_off = g->row_offset;
  p.col = g->column_indices;
  const bool hubs = g->hub_rows && g->hub_count && g->hub_threshold > 0;
  p.hub_rows = hubs ? g->hub_rows : nullptr;
  p.hub_count = hubs ? g->hub_count : nullptr;
  p.hub_threshold = hubs ? g->hub_threshold : 0;
  p.hub_capacity = hubs ? g->hub_capacity : 0;
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
  bind_view(p, g_in);
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
  bind_view(p, g_in);
  p.o_vec = nullptr;
  p.o_head = d_er;
  rc = run_gat(p, 1, al, s);
  if (rc != STG_OK) return rc;
  // pass B: source-parallel on the out-edge view -> d_feat, d_el
  bind_view(p, g_out);
  p.o_vec = d_feat;
  p.o_head = d_el;
  return run_gat(p, 2, al, s);
}
