// Static graph structure on the GPU: CSR (in-edge) / CSC (out-edge) build,
// degrees, degree-sorted row order, degree norm, hub-row schedule.
//
// Replaces the host-side construction of the reference:
//   stgraph/graph/static/static_graph.py:65-78  (Python tuple sorts, eid = rank in (dst,src))
//   stgraph/graph/static/csr.cu:68-170          (host loop + std::sort + 4 cudaMemcpy)
// Pipeline (all stream-ordered, no host sync, workspace supplied by the caller):
//   pack (dst<<32|src) -> radix sort (stable) -> unpack cols / boundary-fill row offsets
//   pack (src<<32|dst) of the forward-sorted edges -> radix sort carrying the forward eid
//   radix sort of (maxdeg - degree) carrying the row id -> node_ids
// Every step is HBM-bound integer work; the sorts are cub::DeviceRadixSort restricted to
// the significant key bits.
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace stg {
namespace {

constexpr int kThreads = 256;

inline int blocks_for(int64_t n, int per_block = kThreads) {
  int64_t b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

__global__ void pack_fwd_keys(const int32_t* __restrict__ src, const int32_t* __restrict__ dst, int64_t n,
                              uint64_t* __restrict__ keys, int32_t* __restrict__ vals) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    keys[i] = (static_cast<uint64_t>(static_cast<uint32_t>(dst[i])) << 32) | static_cast<uint32_t>(src[i]);
    vals[i] = static_cast<int32_t>(i);
  }
}

// From keys sorted by (row<<32|col): write col, optional identity eids, and fill
// row_offset[r] for every r in (row[i-1], row[i]] with i (boundary fill, no atomics).
__global__ void unpack_sorted(const uint64_t* __restrict__ keys, int64_t n, int32_t num_nodes,
                              int32_t* __restrict__ col, int32_t* __restrict__ eids_identity,
                              int32_t* __restrict__ row_offset, int32_t* __restrict__ num_unique) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t k = keys[i];
    const int32_t r = static_cast<int32_t>(k >> 32);
    col[i] = static_cast<int32_t>(k & 0xffffffffu);
    if (eids_identity) eids_identity[i] = static_cast<int32_t>(i);
    const uint64_t kprev = (i == 0) ? ~0ull : keys[i - 1];
    const int32_t prev = (i == 0) ? -1 : static_cast<int32_t>(kprev >> 32);
    if (num_unique && (i == 0 || kprev != k)) atomicAdd(num_unique, 1);   // integer atomic: exact
    for (int32_t q = prev + 1; q <= r; ++q) row_offset[q] = static_cast<int32_t>(i);
    if (i == n - 1) {
      for (int32_t q = r + 1; q <= num_nodes; ++q) row_offset[q] = static_cast<int32_t>(n);
    }
  }
}

__global__ void fill_i32(int32_t* __restrict__ p, int64_t n, int32_t v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// keys for the backward sort, built from the forward-sorted keys: swap halves, value = forward eid
__global__ void pack_bwd_keys(const uint64_t* __restrict__ fwd_keys, int64_t n, uint64_t* __restrict__ keys,
                              int32_t* __restrict__ vals) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t k = fwd_keys[i];
    keys[i] = (k << 32) | (k >> 32);
    vals[i] = static_cast<int32_t>(i);
  }
}

__global__ void degrees_from_offsets(const int32_t* __restrict__ row_offset, int32_t n, int32_t* __restrict__ deg,
                                     uint32_t* __restrict__ sort_key, int32_t* __restrict__ ids) {
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int32_t d = row_offset[i + 1] - row_offset[i];
    if (deg) deg[i] = d;
    if (sort_key) sort_key[i] = 0x7fffffffu - static_cast<uint32_t>(d);  // ascending key = descending degree
    if (ids) ids[i] = i;
  }
}

__global__ void degree_norm_kernel(const int32_t* __restrict__ deg, int32_t n, float* __restrict__ norm) {
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int32_t d = deg[i];
    norm[i] = d > 0 ? 1.0f / sqrtf(static_cast<float>(d)) : 0.f;
  }
}

// thread per row, sequential fp32 sum in row order: the same addition order as the
// host loop of csr.cu:96-128, so the result is bit-identical to the reference's.
__global__ void weighted_row_degree_kernel(const int32_t* __restrict__ row_offset, const int32_t* __restrict__ eids,
                                           int eids_identity, int eid_base, int32_t n,
                                           const float* __restrict__ w, float* __restrict__ out) {
  for (int32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    float acc = 0.f;
    const int32_t e0 = row_offset[r], e1 = row_offset[r + 1];
    for (int32_t e = e0; e < e1; ++e) {
      const int32_t eid = eids_identity ? e : eids[e] - eid_base;
      acc = __fadd_rn(acc, w[eid]);
    }
    out[r] = acc;
  }
}

__global__ void hub_rows_kernel(const int32_t* __restrict__ row_offset, int32_t n, int32_t threshold,
                                int32_t* __restrict__ hub_rows, int32_t capacity, int32_t* __restrict__ hub_count) {
  for (int32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    if (row_offset[r + 1] - row_offset[r] > threshold) {
      const int32_t slot = atomicAdd(hub_count, 1);
      if (slot < capacity) hub_rows[slot] = r;
    }
  }
}

int bits_for(uint32_t max_value) {
  int b = 1;
  while (b < 32 && (max_value >> b) != 0) ++b;
  return b;
}

struct BuildWorkspace {
  uint64_t *keys_a, *keys_b;
  int32_t *vals_a, *vals_b;
  uint32_t *nk_a, *nk_b;
  int32_t *ni_a;
  void* cub_temp;
  size_t cub_bytes;
  size_t total;
};

size_t cub_temp_bytes(int64_t e, int32_t n) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, e > 0 ? e : 1, 0, 64, (cudaStream_t)0);
  cub::DeviceRadixSort::SortPairs(nullptr, b, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int64_t)(n > 0 ? n : 1), 0, 32, (cudaStream_t)0);
  return a > b ? a : b;
}

BuildWorkspace carve(void* base, int64_t e, int32_t n) {
  BuildWorkspace w;
  const size_t e_ = static_cast<size_t>(e > 0 ? e : 1), n_ = static_cast<size_t>(n > 0 ? n : 1);
  char* p = static_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* q = p ? p + off : nullptr;
    off += align_up(bytes, 256);
    return q;
  };
  w.keys_a = reinterpret_cast<uint64_t*>(take(8 * e_));
  w.keys_b = reinterpret_cast<uint64_t*>(take(8 * e_));
  w.vals_a = reinterpret_cast<int32_t*>(take(4 * e_));
  w.vals_b = reinterpret_cast<int32_t*>(take(4 * e_));
  w.nk_a = reinterpret_cast<uint32_t*>(take(4 * n_));
  w.nk_b = reinterpret_cast<uint32_t*>(take(4 * n_));
  w.ni_a = reinterpret_cast<int32_t*>(take(4 * n_));
  w.cub_bytes = cub_temp_bytes(e, n);
  w.cub_temp = take(w.cub_bytes);
  w.total = off;
  return w;
}

}  // namespace
}  // namespace stg

using namespace stg;

STG_API size_t stg_csr_build_workspace_bytes(int64_t num_edges, int32_t num_nodes) {
  return carve(nullptr, num_edges, num_nodes).total;
}

STG_API int stg_csr_build(const int32_t* src, const int32_t* dst, int64_t num_edges, int32_t num_nodes,
                             int32_t* fwd_row_offset, int32_t* fwd_col, int32_t* fwd_eids, int32_t* fwd_node_ids,
                             int32_t* bwd_row_offset, int32_t* bwd_col, int32_t* bwd_eids, int32_t* bwd_node_ids,
                             int32_t* in_degree, int32_t* out_degree, int32_t* edge_perm, int32_t* num_unique,
                             void* workspace, size_t workspace_bytes, void* stream) {
  STG_CHECK_ARG(num_edges >= 0 && num_nodes >= 0, "negative sizes");
  STG_CHECK_ARG(num_edges <= 0x7fffffffLL, "edge count %lld exceeds the int32 index type of the reference ABI",
                (long long)num_edges);
  STG_CHECK_ARG(fwd_row_offset && bwd_row_offset, "row offset outputs must not be NULL");
  STG_CHECK_ARG(num_edges == 0 || (src && dst && fwd_col && fwd_eids && bwd_col && bwd_eids),
                "edge arrays must not be NULL");
  BuildWorkspace w = carve(workspace, num_edges, num_nodes);
  if (workspace == nullptr || workspace_bytes < w.total) {
    set_error("workspace too small: %zu bytes given, %zu needed", workspace_bytes, w.total);
    return STG_ERR_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t s = as_stream(stream);
  if (num_unique) STG_CUDA(cudaMemsetAsync(num_unique, 0, sizeof(int32_t), s));
  const int node_bits = bits_for(num_nodes > 0 ? static_cast<uint32_t>(num_nodes - 1) : 0u);
  const int grid_e = min(blocks_for(num_edges), 148 * 16);
  const int grid_n = min(blocks_for(num_nodes), 148 * 16);

  if (num_edges == 0) {
    fill_i32<<<grid_n, kThreads, 0, s>>>(fwd_row_offset, (int64_t)num_nodes + 1, 0);
    fill_i32<<<grid_n, kThreads, 0, s>>>(bwd_row_offset, (int64_t)num_nodes + 1, 0);
    STG_LAUNCH_CHECK("fill_i32");
  } else {
    // ---- forward: sort by (dst, src); stable, so duplicate edges keep list order
    pack_fwd_keys<<<grid_e, kThreads, 0, s>>>(src, dst, num_edges, w.keys_a, w.vals_a);
    STG_LAUNCH_CHECK("pack_fwd_keys");
    size_t tb = w.cub_bytes;
    STG_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_temp, tb, (const uint64_t*)w.keys_a, w.keys_b,
                                             (const int32_t*)w.vals_a, edge_perm ? edge_perm : w.vals_b,
                                             num_edges, 0, 32 + node_bits, s));
    unpack_sorted<<<grid_e, kThreads, 0, s>>>(w.keys_b, num_edges, num_nodes, fwd_col, fwd_eids, fwd_row_offset,
                                                   num_unique);
    STG_LAUNCH_CHECK("unpack_sorted(fwd)");
    // ---- backward: sort the forward-ordered edges by (src, dst); value = forward eid
    pack_bwd_keys<<<grid_e, kThreads, 0, s>>>(w.keys_b, num_edges, w.keys_a, w.vals_a);
    STG_LAUNCH_CHECK("pack_bwd_keys");
    tb = w.cub_bytes;
    STG_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_temp, tb, (const uint64_t*)w.keys_a, w.keys_b,
                                             (const int32_t*)w.vals_a, bwd_eids, num_edges, 0, 32 + node_bits, s));
    unpack_sorted<<<grid_e, kThreads, 0, s>>>(w.keys_b, num_edges, num_nodes, bwd_col, nullptr, bwd_row_offset,
                                                   nullptr);
    STG_LAUNCH_CHECK("unpack_sorted(bwd)");
  }
  if (num_nodes == 0) return STG_OK;
  // ---- degrees + degree-descending row order (stable => ascending id inside a tie)
  const int deg_bits = 31;
  for (int dir = 0; dir < 2; ++dir) {
    const int32_t* ro = dir == 0 ? fwd_row_offset : bwd_row_offset;
    int32_t* deg = dir == 0 ? in_degree : out_degree;
    int32_t* ids = dir == 0 ? fwd_node_ids : bwd_node_ids;
    degrees_from_offsets<<<grid_n, kThreads, 0, s>>>(ro, num_nodes, deg, ids ? w.nk_a : nullptr,
                                                     ids ? w.ni_a : nullptr);
    STG_LAUNCH_CHECK("degrees_from_offsets");
    if (ids) {
      size_t tb = w.cub_bytes;
      STG_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_temp, tb, (const uint32_t*)w.nk_a, w.nk_b,
                                               (const int32_t*)w.ni_a, ids, (int64_t)num_nodes, 0, deg_bits, s));
    }
  }
  return STG_OK;
}

STG_API int stg_degree_norm_f32(const int32_t* degree, int32_t num_nodes, float* norm, void* stream) {
  STG_CHECK_ARG(num_nodes >= 0, "negative node count");
  if (num_nodes == 0) return STG_OK;
  STG_CHECK_ARG(degree && norm, "NULL pointer");
  degree_norm_kernel<<<min(blocks_for(num_nodes), 148 * 8), kThreads, 0, as_stream(stream)>>>(degree, num_nodes, norm);
  STG_LAUNCH_CHECK("degree_norm_kernel");
  return STG_OK;
}

STG_API int stg_weighted_row_degree_f32(const StgCsrView* g, const float* edge_weight, float* out, void* stream) {
  STG_CHECK_ARG(g && g->row_offset, "graph view / row_offset is NULL");
  if (g->num_nodes == 0) return STG_OK;
  STG_CHECK_ARG(out != nullptr, "out is NULL");
  STG_CHECK_ARG(g->num_edges == 0 || edge_weight != nullptr, "edge_weight is NULL");
  STG_CHECK_ARG(g->eids_identity || g->num_edges == 0 || g->eids, "eids is NULL");
  weighted_row_degree_kernel<<<min(blocks_for(g->num_nodes), 148 * 8), kThreads, 0, as_stream(stream)>>>(
      g->row_offset, g->eids, g->eids_identity, g->eid_base, g->num_nodes, edge_weight, out);
  STG_LAUNCH_CHECK("weighted_row_degree_kernel");
  return STG_OK;
}

STG_API int stg_csr_hub_rows(const int32_t* row_offset, int32_t num_nodes, int32_t threshold,
                                int32_t* hub_rows, int32_t capacity, int32_t* hub_count, void* stream) {
  STG_CHECK_ARG(num_nodes >= 0 && capacity >= 0 && threshold >= 0, "negative argument");
  STG_CHECK_ARG(hub_count != nullptr, "hub_count is NULL");
  cudaStream_t s = as_stream(stream);
  STG_CUDA(cudaMemsetAsync(hub_count, 0, sizeof(int32_t), s));
  if (num_nodes == 0) return STG_OK;
  STG_CHECK_ARG(row_offset && (hub_rows || capacity == 0), "NULL pointer");
  hub_rows_kernel<<<min(blocks_for(num_nodes), 148 * 8), kThreads, 0, s>>>(row_offset, num_nodes, threshold,
                                                                          hub_rows, capacity, hub_count);
  STG_LAUNCH_CHECK("hub_rows_kernel");
  return STG_OK;
}
