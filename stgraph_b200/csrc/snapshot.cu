// Dynamic-graph snapshots on the GPU: edge-set keys, snapshot diffs, batched insert/delete,
// labelled forward / backward views.
//
// Replaces, for NaiveGraph / PCSRGraph / GPMAGraph:
//   stgraph/graph/dynamic/dynamic_graph.py:56-79    Python set() differences per timestamp
//   stgraph/graph/dynamic/gpma/gpma.cu:838-911      update_gpma (locate leaf, per-level rebalance; CDP1)
//   stgraph/graph/dynamic/gpma/gpma.cu:1064-1119    edge_update_t (apply / revert a timestamp)
//   stgraph/graph/dynamic/gpma/gpma.cu:1121-1163    label_edges (thread per row)
//   stgraph/graph/dynamic/gpma/gpma.cu:1165-1231    build_backward_csr (atomic counting sort, nondeterministic)
//   stgraph/graph/dynamic/pcsr/pcsr.cu:404-883      host PMA + build_csr / build_reverse_csr + 4 cudaMemcpy
//
// Representation: a snapshot is the sorted array of its live keys (dst<<32 | src) -- a packed
// memory array with zero gaps.  The reference keeps gaps so that a batch of U updates costs
// O(U log^2 E) on Pascal-class bandwidth; at B200's 6.5 TB/s rewriting all 10^7 keys of config 4
// is ~25 us, cheaper than the 20+ dependent launches of a level-by-level rebalance, fully
// deterministic and sync-free.  The contract is the COMPACTED VIEW: per
// row the sorted live columns, label = 1 + rank among live keys, row_offset, degrees -- all
// bit-exact; "gap placement is free".  Batch update = two binary-search rank kernels + two scans +
// a scatter (merge path); views = boundary fill + one radix sort for the transpose.
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace stg {
namespace {

constexpr int kT = 256;
inline int grid_for(int64_t n) {
  int64_t b = (n + kT - 1) / kT;
  if (b < 1) b = 1;
  if (b > 148 * 32) b = 148 * 32;
  return static_cast<int>(b);
}
#define GRID_STRIDE(i, n) \
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)

__device__ __forceinline__ int64_t lower_bound_u64(const uint64_t* a, int64_t n, uint64_t key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__global__ void pack_keys(const int32_t* __restrict__ src, const int32_t* __restrict__ dst, int64_t n,
                          uint64_t* __restrict__ keys) {
  GRID_STRIDE(i, n) keys[i] = (static_cast<uint64_t>(static_cast<uint32_t>(dst[i])) << 32) | static_cast<uint32_t>(src[i]);
}

// flag[i] = 1 for the first occurrence of each key in a sorted array
__global__ void flag_first(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ flag) {
  GRID_STRIDE(i, n) flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// flag[i] = 1 if a[i] is NOT contained in sorted b
__global__ void flag_absent(const uint64_t* __restrict__ a, int64_t na, const uint64_t* __restrict__ b, int64_t nb,
                            int32_t* __restrict__ flag) {
  GRID_STRIDE(i, na) {
    const uint64_t k = a[i];
    const int64_t p = lower_bound_u64(b, nb, k);
    flag[i] = (p < nb && b[p] == k) ? 0 : 1;
  }
}

// out[rank[i]] = in[i] for flagged items; the total is written by the last thread
__global__ void compact_flagged(const uint64_t* __restrict__ in, const int32_t* __restrict__ flag,
                                const int32_t* __restrict__ rank, int64_t n, uint64_t* __restrict__ out,
                                int64_t* __restrict__ count) {
  GRID_STRIDE(i, n) {
    if (flag[i]) out[rank[i]] = in[i];
    if (i == n - 1 && count) *count = static_cast<int64_t>(rank[i]) + flag[i];
  }
}

__global__ void set_count(int64_t* count, int64_t v) { *count = v; }

// merge-path scatter of the surviving old keys and the kept new keys
__global__ void merge_scatter_old(const uint64_t* __restrict__ keys, const int32_t* __restrict__ keep,
                                  const int32_t* __restrict__ rank_old, int64_t n, const uint64_t* __restrict__ add,
                                  const int32_t* __restrict__ keep_add, const int32_t* __restrict__ rank_add, int64_t na,
                                  uint64_t* __restrict__ out) {
  GRID_STRIDE(i, n) {
    if (!keep[i]) continue;
    const uint64_t k = keys[i];
    const int64_t p = lower_bound_u64(add, na, k);                 // adds strictly smaller than k
    const int64_t smaller_adds = (p < na) ? rank_add[p] : (na > 0 ? rank_add[na - 1] + keep_add[na - 1] : 0);
    out[rank_old[i] + smaller_adds] = k;
  }
}
__global__ void merge_scatter_add(const uint64_t* __restrict__ keys, const int32_t* __restrict__ keep,
                                  const int32_t* __restrict__ rank_old, int64_t n, const uint64_t* __restrict__ add,
                                  const int32_t* __restrict__ keep_add, const int32_t* __restrict__ rank_add, int64_t na,
                                  uint64_t* __restrict__ out, int64_t* __restrict__ count) {
  GRID_STRIDE(j, na) {
    if (keep_add[j]) {
      const uint64_t k = add[j];
      const int64_t p = lower_bound_u64(keys, n, k);               // old keys strictly smaller than k
      const int64_t smaller_old = (p < n) ? rank_old[p] : (n > 0 ? rank_old[n - 1] + keep[n - 1] : 0);
      out[rank_add[j] + smaller_old] = k;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && count) {
    const int64_t a = n > 0 ? rank_old[n - 1] + keep[n - 1] : 0;
    const int64_t b = na > 0 ? rank_add[na - 1] + keep_add[na - 1] : 0;
    *count = a + b;
  }
}

// keep_add[j] = 1 unless add[j] is already live (present in keys and not deleted) or repeats add[j-1]
__global__ void flag_new_adds(const uint64_t* __restrict__ add, int64_t na, const uint64_t* __restrict__ keys,
                              const int32_t* __restrict__ keep, int64_t n, int32_t* __restrict__ keep_add) {
  GRID_STRIDE(j, na) {
    const uint64_t k = add[j];
    int f = (j == 0 || add[j - 1] != k) ? 1 : 0;
    const int64_t p = lower_bound_u64(keys, n, k);
    if (p < n && keys[p] == k && keep[p]) f = 0;
    keep_add[j] = f;
  }
}

// forward view from sorted keys: row_offset by boundary fill, col = low half, label = rank + 1.
// descending: each row is emitted back to front (PCSR build_csr order, pcsr.cu:842-855).
__global__ void view_fill_offsets(const uint64_t* __restrict__ keys, int64_t n, int32_t num_nodes,
                                  int32_t* __restrict__ row_offset) {
  GRID_STRIDE(i, n) {
    const int32_t r = static_cast<int32_t>(keys[i] >> 32);
    const int32_t prev = (i == 0) ? -1 : static_cast<int32_t>(keys[i - 1] >> 32);
    for (int32_t q = prev + 1; q <= r; ++q) row_offset[q] = static_cast<int32_t>(i);
    if (i == n - 1)
      for (int32_t q = r + 1; q <= num_nodes; ++q) row_offset[q] = static_cast<int32_t>(n);
  }
}
__global__ void fill_offsets_empty(int32_t* __restrict__ row_offset, int32_t num_nodes) {
  GRID_STRIDE(i, (int64_t)num_nodes + 1) row_offset[i] = 0;
}
__global__ void view_emit(const uint64_t* __restrict__ keys, const int32_t* __restrict__ labels_in, int64_t n,
                          const int32_t* __restrict__ row_offset, int descending, int label_base,
                          int32_t* __restrict__ col, int32_t* __restrict__ labels) {
  GRID_STRIDE(i, n) {
    const uint64_t k = keys[i];
    int64_t pos = i;
    if (descending) {
      const int32_t r = static_cast<int32_t>(k >> 32);
      pos = static_cast<int64_t>(row_offset[r]) + (row_offset[r + 1] - 1 - i);
    }
    col[pos] = static_cast<int32_t>(k & 0xffffffffu);
    labels[pos] = labels_in ? labels_in[i] : static_cast<int32_t>(i) + label_base;
  }
}
__global__ void swap_halves(const uint64_t* __restrict__ keys, int64_t n, int label_base, uint64_t* __restrict__ out,
                            int32_t* __restrict__ labels) {
  GRID_STRIDE(i, n) {
    const uint64_t k = keys[i];
    out[i] = (k << 32) | (k >> 32);
    labels[i] = static_cast<int32_t>(i) + label_base;
  }
}
__global__ void degrees_and_sortkeys(const int32_t* __restrict__ row_offset, int32_t n, int32_t* __restrict__ deg,
                                     uint32_t* __restrict__ sort_key, int32_t* __restrict__ ids) {
  GRID_STRIDE(i, n) {
    const int32_t d = row_offset[i + 1] - row_offset[i];
    if (deg) deg[i] = d;
    if (sort_key) sort_key[i] = 0x7fffffffu - static_cast<uint32_t>(d);
    if (ids) ids[i] = static_cast<int32_t>(i);
  }
}

struct Ws {
  uint64_t *k0, *k1;
  int32_t *f0, *f1, *r0, *r1, *l0;
  void* cub;
  size_t cub_bytes, total;
};

size_t cub_bytes_for(int64_t m) {
  size_t a = 0, b = 0, c = 0;
  const int64_t mm = m > 0 ? m : 1;
  cub::DeviceRadixSort::SortKeys(nullptr, a, (const uint64_t*)nullptr, (uint64_t*)nullptr, mm, 0, 64, (cudaStream_t)0);
  cub::DeviceRadixSort::SortPairs(nullptr, b, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, mm, 0, 64, (cudaStream_t)0);
  cub::DeviceScan::ExclusiveSum(nullptr, c, (const int32_t*)nullptr, (int32_t*)nullptr, mm, (cudaStream_t)0);
  size_t r = a > b ? a : b;
  return r > c ? r : c;
}

Ws carve_ws(void* base, int64_t m) {
  Ws w;
  const size_t mm = static_cast<size_t>(m > 0 ? m : 1);
  char* p = static_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* q = p ? p + off : nullptr;
    off += align_up(bytes, 256);
    return q;
  };
  w.k0 = reinterpret_cast<uint64_t*>(take(8 * mm));
  w.k1 = reinterpret_cast<uint64_t*>(take(8 * mm));
  w.f0 = reinterpret_cast<int32_t*>(take(4 * mm));
  w.f1 = reinterpret_cast<int32_t*>(take(4 * mm));
  w.r0 = reinterpret_cast<int32_t*>(take(4 * mm));
  w.r1 = reinterpret_cast<int32_t*>(take(4 * mm));
  w.l0 = reinterpret_cast<int32_t*>(take(4 * mm));
  w.cub_bytes = cub_bytes_for(m);
  w.cub = take(w.cub_bytes);
  w.total = off;
  return w;
}

int bits_for_nodes(int32_t num_nodes) {
  uint32_t v = num_nodes > 0 ? static_cast<uint32_t>(num_nodes - 1) : 0u;
  int b = 1;
  while (b < 32 && (v >> b) != 0) ++b;
  return b;
}

#define CHECK_WS(w, ws, ws_bytes)                                                              \
  if ((ws) == nullptr || (ws_bytes) < (w).total) {                                             \
    set_error("workspace too small: %zu bytes given, %zu needed", (size_t)(ws_bytes), (w).total); \
    return STG_ERR_WORKSPACE_TOO_SMALL;                                                        \
  }

}  // namespace
}  // namespace stg

using namespace stg;

STG_API size_t stg_snapshot_workspace_bytes(int64_t max_items) { return carve_ws(nullptr, max_items).total; }

STG_API int stg_snapshot_keys_from_edges(const int32_t* src, const int32_t* dst, int64_t n, int32_t num_nodes,
                                         uint64_t* keys_out, int64_t* count_out, void* ws, size_t ws_bytes,
                                         void* stream) {
  STG_CHECK_ARG(n >= 0 && count_out != nullptr, "bad arguments");
  cudaStream_t s = as_stream(stream);
  if (n == 0) {
    set_count<<<1, 1, 0, s>>>(count_out, 0);
    return STG_OK;
  }
  STG_CHECK_ARG(src && dst && keys_out, "NULL pointer");
  Ws w = carve_ws(ws, n);
  CHECK_WS(w, ws, ws_bytes);
  pack_keys<<<grid_for(n), kT, 0, s>>>(src, dst, n, w.k0);
  size_t tb = w.cub_bytes;
  STG_CUDA(cub::DeviceRadixSort::SortKeys(w.cub, tb, (const uint64_t*)w.k0, w.k1, n, 0, 32 + bits_for_nodes(num_nodes), s));
  flag_first<<<grid_for(n), kT, 0, s>>>(w.k1, n, w.f0);
  tb = w.cub_bytes;
  STG_CUDA(cub::DeviceScan::ExclusiveSum(w.cub, tb, (const int32_t*)w.f0, w.r0, n, s));
  compact_flagged<<<grid_for(n), kT, 0, s>>>(w.k1, w.f0, w.r0, n, keys_out, count_out);
  STG_LAUNCH_CHECK("keys_from_edges");
  return STG_OK;
}

STG_API int stg_snapshot_diff(const uint64_t* a, int64_t na, const uint64_t* b, int64_t nb, uint64_t* out,
                              int64_t* count_out, void* ws, size_t ws_bytes, void* stream) {
  STG_CHECK_ARG(na >= 0 && nb >= 0 && count_out != nullptr, "bad arguments");
  cudaStream_t s = as_stream(stream);
  if (na == 0) {
    set_count<<<1, 1, 0, s>>>(count_out, 0);
    return STG_OK;
  }
  STG_CHECK_ARG(a && out && (b || nb == 0), "NULL pointer");
  Ws w = carve_ws(ws, na);
  CHECK_WS(w, ws, ws_bytes);
  flag_absent<<<grid_for(na), kT, 0, s>>>(a, na, b, nb, w.f0);
  size_t tb = w.cub_bytes;
  STG_CUDA(cub::DeviceScan::ExclusiveSum(w.cub, tb, (const int32_t*)w.f0, w.r0, na, s));
  compact_flagged<<<grid_for(na), kT, 0, s>>>(a, w.f0, w.r0, na, out, count_out);
  STG_LAUNCH_CHECK("snapshot_diff");
  return STG_OK;
}

STG_API int stg_snapshot_apply(const uint64_t* keys, int64_t n, const uint64_t* add, int64_t na, const uint64_t* del,
                               int64_t nd, uint64_t* out, int64_t* count_out, void* ws, size_t ws_bytes, void* stream) {
  STG_CHECK_ARG(n >= 0 && na >= 0 && nd >= 0, "negative sizes");
  STG_CHECK_ARG((keys || n == 0) && (add || na == 0) && (del || nd == 0) && (out || n + na == 0), "NULL pointer");
  STG_CHECK_ARG(out != keys, "stg_snapshot_apply is out of place");
  cudaStream_t s = as_stream(stream);
  const int64_t m = (n > na ? n : na);
  Ws w = carve_ws(ws, m);
  CHECK_WS(w, ws, ws_bytes);
  size_t tb;
  if (n > 0) {
    flag_absent<<<grid_for(n), kT, 0, s>>>(keys, n, del, nd, w.f0);          // keep[i]
    tb = w.cub_bytes;
    STG_CUDA(cub::DeviceScan::ExclusiveSum(w.cub, tb, (const int32_t*)w.f0, w.r0, n, s));
  }
  if (na > 0) {
    flag_new_adds<<<grid_for(na), kT, 0, s>>>(add, na, keys, w.f0, n, w.f1);  // keep_add[j]
    tb = w.cub_bytes;
    STG_CUDA(cub::DeviceScan::ExclusiveSum(w.cub, tb, (const int32_t*)w.f1, w.r1, na, s));
  }
  if (n > 0) merge_scatter_old<<<grid_for(n), kT, 0, s>>>(keys, w.f0, w.r0, n, add, w.f1, w.r1, na, out);
  merge_scatter_add<<<grid_for(na), kT, 0, s>>>(keys, w.f0, w.r0, n, add, w.f1, w.r1, na, out, count_out);
  STG_LAUNCH_CHECK("snapshot_apply");
  return STG_OK;
}

STG_API int stg_snapshot_views(const uint64_t* keys, int64_t n, int32_t num_nodes, int32_t descending_rows,
                               int32_t label_base,
                               int32_t* fwd_row_offset, int32_t* fwd_col, int32_t* fwd_labels, int32_t* fwd_node_ids,
                               int32_t* bwd_row_offset, int32_t* bwd_col, int32_t* bwd_labels, int32_t* bwd_node_ids,
                               int32_t* in_degree, int32_t* out_degree, void* ws, size_t ws_bytes, void* stream) {
  STG_CHECK_ARG(n >= 0 && num_nodes >= 0, "negative sizes");
  STG_CHECK_ARG(n <= 0x7fffffffLL, "snapshot has more edges than the int32 index type allows");
  STG_CHECK_ARG(label_base == 0 || label_base == 1, "label_base must be 0 or 1");
  STG_CHECK_ARG(fwd_row_offset != nullptr, "fwd_row_offset is NULL");
  STG_CHECK_ARG(n == 0 || (keys && fwd_col && fwd_labels), "NULL pointer");
  cudaStream_t s = as_stream(stream);
  const int64_t m = n > num_nodes ? n : num_nodes;
  Ws w = carve_ws(ws, m);
  CHECK_WS(w, ws, ws_bytes);
  const bool want_bwd = bwd_row_offset != nullptr;
  if (n == 0) {
    fill_offsets_empty<<<grid_for(num_nodes + 1), kT, 0, s>>>(fwd_row_offset, num_nodes);
    if (want_bwd) fill_offsets_empty<<<grid_for(num_nodes + 1), kT, 0, s>>>(bwd_row_offset, num_nodes);
  } else {
    view_fill_offsets<<<grid_for(n), kT, 0, s>>>(keys, n, num_nodes, fwd_row_offset);
    // labels: label_base + rank among live keys -- 1-based for PCSR/GPMA (gpma.cu:1121-1146,
    // pcsr.cu:748-760), 0-based for NaiveGraph's per-snapshot CSR (static_graph.py:65-72)
    view_emit<<<grid_for(n), kT, 0, s>>>(keys, nullptr, n, fwd_row_offset, descending_rows, label_base, fwd_col,
                                         fwd_labels);
    if (want_bwd) {
      swap_halves<<<grid_for(n), kT, 0, s>>>(keys, n, label_base, w.k0, w.l0);
      STG_CHECK_ARG(bwd_col && bwd_labels, "NULL backward arrays");
      size_t tb = w.cub_bytes;
      STG_CUDA(cub::DeviceRadixSort::SortPairs(w.cub, tb, (const uint64_t*)w.k0, w.k1, (const int32_t*)w.l0, w.f0, n, 0,
                                               32 + bits_for_nodes(num_nodes), s));
      view_fill_offsets<<<grid_for(n), kT, 0, s>>>(w.k1, n, num_nodes, bwd_row_offset);
      view_emit<<<grid_for(n), kT, 0, s>>>(w.k1, w.f0, n, bwd_row_offset, descending_rows, label_base, bwd_col, bwd_labels);
    }
  }
  if (num_nodes > 0) {
    for (int dir = 0; dir < 2; ++dir) {
      const int32_t* ro = dir == 0 ? fwd_row_offset : bwd_row_offset;
      int32_t* deg = dir == 0 ? in_degree : out_degree;
      int32_t* ids = dir == 0 ? fwd_node_ids : bwd_node_ids;
      if (ro == nullptr || (deg == nullptr && ids == nullptr)) continue;
      degrees_and_sortkeys<<<grid_for(num_nodes), kT, 0, s>>>(ro, num_nodes, deg, ids ? (uint32_t*)w.r0 : nullptr,
                                                              ids ? w.f1 : nullptr);
      if (ids) {
        size_t tb = w.cub_bytes;
        STG_CUDA(cub::DeviceRadixSort::SortPairs(w.cub, tb, (const uint32_t*)w.r0, (uint32_t*)w.r1, (const int32_t*)w.f1,
                                                 ids, (int64_t)num_nodes, 0, 31, s));
      }
    }
  }
  STG_LAUNCH_CHECK("snapshot_views");
  return STG_OK;
}
