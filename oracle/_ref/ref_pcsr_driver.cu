
// Driver around the reference's host-side PCSR: includes it from where it lies, keeps its "pinned" and
// "device" CSR buffers on the host heap (there is no GPU in the build container) and exposes the arrays
// build_csr() / build_reverse_csr() fill through a C ABI.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#include <thrust/device_vector.h>
#include <thrust/host_vector.h>
#include <thrust/remove.h>
#include <thrust/sort.h>
#include <cub/cub.cuh>
#include <cstdlib>
#include <cstring>
static inline cudaError_t ref_host_alloc(void** p, size_t s) { *p = calloc(s ? s : 1, 1); return cudaSuccess; }
#define cudaMallocHost(p, s) ref_host_alloc((void**)(p), (s))
#define cudaMalloc(p, s) ref_host_alloc((void**)(p), (s))
#define cudaMemcpy(d, s, n, k) (memcpy((d), (s), (n)), cudaSuccess)
#include "/root/reference/stgraph/graph/dynamic/pcsr/pcsr.cu"
extern "C" void* ref_pcsr_new(int n, int max_edges) { return new PCSR(n, max_edges); }
extern "C" void ref_pcsr_update(void* h, const uint32_t* a, const uint32_t* b, int cnt, int is_delete, int is_reverse) {
  std::vector<std::tuple<uint32_t, uint32_t>> el;
  for (int i = 0; i < cnt; ++i) el.emplace_back(a[i], b[i]);
  static_cast<PCSR*>(h)->edge_update_list(el, is_delete != 0, is_reverse != 0);
}
extern "C" void ref_pcsr_label(void* h) { static_cast<PCSR*>(h)->label_edges(); }
extern "C" int ref_pcsr_build(void* h, int reverse, uint32_t* ro, uint32_t* col, uint32_t* eids, uint32_t* nid,
                              uint32_t* in_deg, uint32_t* out_deg) {
  PCSR* p = static_cast<PCSR*>(h);
  if (reverse) p->build_reverse_csr(); else p->build_csr();
  const size_t n = p->get_n();
  std::copy(p->row_offset_pinned, p->row_offset_pinned + n + 1, ro);
  std::copy(p->column_indices_pinned, p->column_indices_pinned + p->edge_count, col);
  std::copy(p->eids_pinned, p->eids_pinned + p->edge_count, eids);
  std::copy(p->node_ids_pinned, p->node_ids_pinned + n, nid);
  std::copy(p->in_degrees.begin(), p->in_degrees.end(), in_deg);
  std::copy(p->out_degrees.begin(), p->out_degrees.end(), out_deg);
  return (int)p->edge_count;
}
extern "C" void ref_pcsr_free(void* h) { delete static_cast<PCSR*>(h); }
