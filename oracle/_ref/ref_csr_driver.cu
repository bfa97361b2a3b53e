
// Driver around the reference's CSR builder: includes it from where it lies and exposes the host
// vectors (row_offset, column_indices, eids, node_ids, degrees) through a C ABI.
#include "/root/reference/stgraph/graph/static/csr.cu"
extern "C" void* ref_csr_new(const int* a, const int* b, const int* eid, int n_edges, const float* w, int num_nodes,
                             int is_edge_reverse) {
  std::vector<std::tuple<int, int, int>> el;
  for (int i = 0; i < n_edges; ++i) el.emplace_back(a[i], b[i], eid[i]);
  std::vector<float> ew(w, w + n_edges);
  return new CSR(el, ew, num_nodes, is_edge_reverse != 0);
}
extern "C" void ref_csr_get(void* h, int* row_offset, int* column_indices, int* eids, int* node_ids, int* in_deg,
                            int* out_deg, float* wdeg) {
  CSR* c = static_cast<CSR*>(h);
  std::copy(c->row_offset.begin(), c->row_offset.end(), row_offset);
  std::copy(c->column_indices.begin(), c->column_indices.end(), column_indices);
  std::copy(c->eids.begin(), c->eids.end(), eids);
  std::copy(c->node_ids.begin(), c->node_ids.end(), node_ids);
  std::copy(c->in_degrees.begin(), c->in_degrees.end(), in_deg);
  std::copy(c->out_degrees.begin(), c->out_degrees.end(), out_deg);
  std::copy(c->weighted_out_degrees.begin(), c->weighted_out_degrees.end(), wdeg);
}
extern "C" void ref_csr_free(void* h) { delete static_cast<CSR*>(h); }
