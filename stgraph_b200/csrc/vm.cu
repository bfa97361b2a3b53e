// Generic fused vertex-program kernel: interprets a lowered execution unit.
//
// The reference turns every fused unit into CUDA text (Jinja templates
// stgraph/compiler/code_gen/templates/fa/tpl_fa_csr*.jinja:1-57, per-op C snippets in
// stgraph/compiler/registry.py:195-406) and compiles it with nvcc at run time
// (code_gen/compiler.py:14-44).  Here the unit is lowered once (compiler/lowering.py) to a small
// register program and ONE pre-compiled sm_100a kernel evaluates it with the same loop
// structure as the reference template:
//     per (row, feature lane):  PRE instrs;  for e in row: LOOP instrs;  POST instrs
// Differences that matter: lanes of a row sit in one warp (a GROUP of 2^k lanes per row when the
// unit is narrower than a warp, several rows per warp), cross-lane reductions ([H,D] -> [H,1]
// outputs, the reference's per-lane atomicAdd, kernel_context.py:126-149) use segmented warp
// shuffles when the segment is a power of two, and aggregations onto the *other* side are never
// written with atomics from inside the loop: lowering runs a second launch on the transposed CSR.
// Virtual registers live in shared memory ([reg][thread], conflict-free).
// This is the generality path; hot shapes are matched to agg.cu / gat.cu first.
#include "common.cuh"

namespace stg {
namespace {

constexpr int kVmThreads = 128;
constexpr int kVmHubThreads = 1024;   // one block per long row: 32 warps share its edges

// Rows longer than this are split over a whole block (interpreting an instruction costs a few hundred
// cycles, so long rows dominate otherwise).  The hub launch finds them by scanning the row lengths itself.
constexpr int kVmHubThreshold = 96;

struct VmArgs {
  StgCsrView g;
  void* tensors[STG_VM_MAX_TENSORS];
};

__device__ __forceinline__ int tensor_elem(const StgVmTensor& t, int i0, int i1, int dim1) {
  return (t.bc0 ? i0 : 0) * (t.bc1 ? dim1 : 1) + (t.bc1 ? i1 : 0);
}
__device__ __forceinline__ int tensor_size(const StgVmTensor& t, int dim0, int dim1) {
  return (t.bc0 ? dim0 : 1) * (t.bc1 ? dim1 : 1);
}

// HUB = false: one lane group per row (rows longer than the view's hub threshold are skipped).
// HUB = true : one BLOCK per hub row; every (warp, group) slot takes a strided share of the row's edges,
//              the accumulators are merged through shared memory by kind (sum / max / min) in a fixed
//              order, and slot 0 alone runs the POST phase.
template <int GROUP, bool HUB>
__global__ void __launch_bounds__(HUB ? kVmHubThreads : kVmThreads)
    vm_kernel(const __grid_constant__ VmArgs a, const __grid_constant__ StgVmProgram prog) {
  constexpr int NT = HUB ? kVmHubThreads : kVmThreads;     // threads per block
  extern __shared__ float smem[];
  float* regs = smem;                                      // [n_regs][NT]
  float* accs = smem + prog.n_regs * NT;                   // [n_acc][NT]
  float* scratch = accs + prog.n_acc * NT;                 // [NT] lane-reduction staging
  constexpr int GROUPS_PER_WARP = 32 / GROUP;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int gl = lane & (GROUP - 1);
  const unsigned gmask = (GROUP == 32) ? 0xffffffffu : (((1u << GROUP) - 1u) << (lane & ~(GROUP - 1)));
  constexpr int NSLOT = HUB ? (NT / 32) * GROUPS_PER_WARP : 1;
  const int slot = HUB ? (tid >> 5) * GROUPS_PER_WARP + lane / GROUP : 0;
  const int lanes = prog.dim0 * prog.dim1;
  const int dim1 = prog.dim1;
  const int n_iter = HUB ? a.g.num_nodes : 1;
  for (int it = HUB ? blockIdx.x : 0; it < n_iter; it += HUB ? gridDim.x : 1) {
  int row;
  if (HUB) {
    row = it;
  } else {
    const int warp = blockIdx.x * (kVmThreads / 32) + (tid >> 5);
    row = warp * GROUPS_PER_WARP + lane / GROUP;
    if (row >= a.g.num_nodes) return;
  }
  const int beg = __ldg(a.g.row_offset + row);
  const int end = __ldg(a.g.row_offset + row + 1);
  if (HUB && (end - beg) <= kVmHubThreshold) continue;    // block-uniform: every thread sees the same row
  if (!HUB && (end - beg) > kVmHubThreshold) return;      // the hub launch owns this row
  const bool seg_pow2 = (dim1 & (dim1 - 1)) == 0 && dim1 <= GROUP;
  // lanes per chunk: whole dim1-segments only, so a segment never straddles two chunks
  const int chunk = (dim1 <= GROUP) ? (GROUP / dim1) * dim1 : GROUP;
  // sum over the dim1 consecutive lanes of this lane's segment, result in every lane of the segment
  auto segment_sum = [&](float v, int i1) -> float {
    if (seg_pow2) {
      for (int o = 1; o < dim1; o <<= 1) v += __shfl_xor_sync(gmask, v, o, GROUP);
      return v;
    }
    scratch[tid] = v;
    __syncwarp(gmask);
    float s = 0.f;
    for (int k = 0; k < dim1; ++k) s += scratch[tid - i1 + k];
    __syncwarp(gmask);
    return s;
  };

#define R(i) regs[(i) * NT + tid]
#define ACC(i) accs[(i) * NT + tid]

  for (int tx0 = 0; tx0 < lanes; tx0 += chunk) {
    const int tx = tx0 + gl;
    const bool active = gl < chunk && tx < lanes;
    const int txc = active ? tx : lanes - 1;
    const int i0 = txc / dim1, i1 = txc - i0 * dim1;
    for (int k = 0; k < prog.n_acc; ++k) ACC(k) = prog.acc_init[k];

    int nbr = 0, eid = 0;
    auto exec = [&](const StgVmInstr& in) {
      switch (in.op) {
        case STG_OP_LOAD: {
          const StgVmTensor& t = prog.tensors[in.a];
          const float* base = static_cast<const float*>(a.tensors[in.a]);
          long long id = 0;
          if (t.side == STG_VM_CENTER) id = row;
          else if (t.side == STG_VM_NBR) id = nbr;
          else if (t.side == STG_VM_EDGE) id = eid;
          R(in.dst) = __ldg(base + id * tensor_size(t, prog.dim0, dim1) + tensor_elem(t, i0, i1, dim1));
          break;
        }
        case STG_OP_CONST: R(in.dst) = in.imm; break;
        case STG_OP_ADD: R(in.dst) = R(in.a) + R(in.b); break;
        case STG_OP_SUB: R(in.dst) = R(in.a) - R(in.b); break;
        case STG_OP_MUL: R(in.dst) = R(in.a) * R(in.b); break;
        case STG_OP_DIV: R(in.dst) = R(in.a) / R(in.b); break;
        case STG_OP_EXP: R(in.dst) = expf(R(in.a)); break;
        case STG_OP_LRELU: { const float v = R(in.a); R(in.dst) = v > 0.f ? v : in.imm * v; break; }
        case STG_OP_LRELU_BWD: R(in.dst) = R(in.a) > 0.f ? 1.f : in.imm; break;
        case STG_OP_RELU: { const float v = R(in.a); R(in.dst) = v > 0.f ? v : 0.f; break; }
        case STG_OP_RELU_BWD: R(in.dst) = R(in.a) > 0.f ? R(in.b) : 0.f; break;
        case STG_OP_AMAX_BWD: R(in.dst) = R(in.a) == R(in.b) ? 1.f : 0.f; break;
        case STG_OP_ACC_SUM: ACC(in.dst) += R(in.a); break;
        case STG_OP_ACC_MAX: ACC(in.dst) = fmaxf(ACC(in.dst), R(in.a)); break;
        case STG_OP_ACC_MIN: ACC(in.dst) = fminf(ACC(in.dst), R(in.a)); break;
        case STG_OP_ACC_READ: {
          float v = ACC(in.a);
          if (in.b == 1) v = (end > beg) ? v / static_cast<float>(end - beg) : 0.f;
          R(in.dst) = v;
          break;
        }
        case STG_OP_GSUM: {
          // sum over dim1 inside each dim0 slice (segments of dim1 consecutive lanes), broadcast back
          R(in.dst) = segment_sum(active ? R(in.a) : 0.f, i1);
          break;
        }
        case STG_OP_STORE: {
          const StgVmTensor& t = prog.tensors[in.a];
          float* base = static_cast<float*>(a.tensors[in.a]);
          long long id = 0;
          if (t.side == STG_VM_CENTER) id = row;
          else if (t.side == STG_VM_NBR) id = nbr;
          else if (t.side == STG_VM_EDGE) id = eid;
          float* dst = base + id * tensor_size(t, prog.dim0, dim1) + tensor_elem(t, i0, i1, dim1);
          float v = R(in.b);
          const bool full = t.bc0 && t.bc1;
          if (in.imm == 0.f) {
            // value already has the tensor's shape: one lane per distinct element writes
            const bool leader = (t.bc0 || i0 == 0) && (t.bc1 || i1 == 0);
            if (active && leader) *dst = v;
          } else if (full) {
            if (active) *dst = v;
          } else if (t.bc0 && !t.bc1 && dim1 <= GROUP) {
            // [dim0,dim1] -> [dim0,1]: segmented reduction over dim1 consecutive lanes
            v = segment_sum(active ? v : 0.f, i1);
            if (active && i1 == 0) *dst = v;
          } else {
            if (active) atomicAdd(dst, v);   // caller zero-fills; generic cross-lane reduction
          }
          break;
        }
        default: break;
      }
    };

    int pc = 0;
    for (; pc < prog.n_pre; ++pc) exec(prog.instr[pc]);
    const int loop_end = prog.n_pre + prog.n_loop;
    if (prog.n_loop > 0) {
      for (int e = beg + slot; e < end; e += NSLOT) {
        nbr = __ldg(a.g.column_indices + e);
        eid = a.g.eids_identity ? e : (__ldg(a.g.eids + e) - a.g.eid_base);
        for (int q = prog.n_pre; q < loop_end; ++q) exec(prog.instr[q]);
      }
    }
    if (HUB) {
      __syncthreads();
      if (slot == 0) {
        for (int k = 0; k < prog.n_acc; ++k) {
          float v = ACC(k);
          for (int s2 = 1; s2 < NSLOT; ++s2) {
            const int other = (s2 / GROUPS_PER_WARP) * 32 + (s2 % GROUPS_PER_WARP) * GROUP + gl;
            const float o = accs[k * NT + other];
            v = prog.acc_kind[k] == 1 ? fmaxf(v, o) : prog.acc_kind[k] == 2 ? fminf(v, o) : v + o;
          }
          ACC(k) = v;
        }
        for (int q = loop_end; q < prog.n_instr; ++q) exec(prog.instr[q]);
      }
      __syncthreads();
    } else {
      for (int q = loop_end; q < prog.n_instr; ++q) exec(prog.instr[q]);
    }
  }
  }
#undef R
#undef ACC
}

template <int GROUP>
int launch_vm(const VmArgs& a, const StgVmProgram& prog, cudaStream_t stream) {
  const int rows_per_block = (kVmThreads / 32) * (32 / GROUP);
  const int blocks = (a.g.num_nodes + rows_per_block - 1) / rows_per_block;
  const size_t smem = static_cast<size_t>(prog.n_regs + prog.n_acc + 1) * kVmThreads * sizeof(float);
  vm_kernel<GROUP, false><<<blocks, kVmThreads, smem, stream>>>(a, prog);
  STG_LAUNCH_CHECK("vm_kernel");
  if (a.g.num_edges > kVmHubThreshold) {     // a row longer than the threshold can only exist then
    const size_t hub_smem = static_cast<size_t>(prog.n_regs + prog.n_acc + 1) * kVmHubThreads * sizeof(float);
    static size_t configured = 0;
    if (hub_smem > configured) {
      STG_CUDA(cudaFuncSetAttribute(vm_kernel<GROUP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(hub_smem)));
      configured = hub_smem;
    }
    vm_kernel<GROUP, true><<<2 * sm_count(), kVmHubThreads, hub_smem, stream>>>(a, prog);
    STG_LAUNCH_CHECK("vm_kernel(hub)");
  }
  return STG_OK;
}

}  // namespace
}  // namespace stg

using namespace stg;

STG_API int stg_vm_run_f32(const StgCsrView* g, const StgVmProgram* prog, void* const* tensors, void* stream) {
  STG_CHECK_ARG(g && prog && tensors, "NULL argument");
  STG_CHECK_ARG(g->row_offset != nullptr, "row_offset is NULL");
  STG_CHECK_ARG(prog->dim0 > 0 && prog->dim1 > 0, "program dims must be positive");
  STG_CHECK_ARG(prog->n_tensors >= 0 && prog->n_tensors <= STG_VM_MAX_TENSORS, "too many tensors (%d)", prog->n_tensors);
  STG_CHECK_ARG(prog->n_instr >= 0 && prog->n_instr <= STG_VM_MAX_INSTR, "too many instructions (%d)", prog->n_instr);
  STG_CHECK_ARG(prog->n_regs >= 0 && prog->n_regs <= STG_VM_MAX_REGS, "too many registers (%d)", prog->n_regs);
  STG_CHECK_ARG(prog->n_acc >= 0 && prog->n_acc <= STG_VM_MAX_ACC, "too many accumulators (%d)", prog->n_acc);
  STG_CHECK_ARG(prog->n_pre >= 0 && prog->n_loop >= 0 && prog->n_pre + prog->n_loop <= prog->n_instr,
                "inconsistent phase counts");
  bool needs_eids = false;
  for (int i = 0; i < prog->n_instr; ++i) {
    const StgVmInstr& in = prog->instr[i];
    STG_CHECK_ARG(in.op >= 0 && in.op < STG_OP_COUNT_, "bad opcode %d at %d", in.op, i);
    if (in.op == STG_OP_LOAD || in.op == STG_OP_STORE) {
      STG_CHECK_ARG(in.a >= 0 && in.a < prog->n_tensors, "tensor index out of range at instr %d", i);
      STG_CHECK_ARG(tensors[in.a] != nullptr, "tensor %d is NULL", in.a);
      if (prog->tensors[in.a].side == STG_VM_EDGE) needs_eids = true;
    }
    if (in.op == STG_OP_GSUM) {
      const int d1 = prog->dim1;
      STG_CHECK_ARG(d1 <= 32, "GSUM needs dim1 <= 32 (got %d)", d1);
    }
  }
  STG_CHECK_ARG(!needs_eids || g->eids_identity || g->eids || g->num_edges == 0, "eids is NULL");
  if (g->num_nodes == 0) return STG_OK;
  STG_CHECK_ARG(g->num_edges == 0 || g->column_indices, "column_indices is NULL");
  VmArgs a;
  a.g = *g;
  for (int i = 0; i < STG_VM_MAX_TENSORS; ++i) a.tensors[i] = i < prog->n_tensors ? tensors[i] : nullptr;
  const int lanes = prog->dim0 * prog->dim1;
  cudaStream_t s = as_stream(stream);
  if (lanes <= 1) return launch_vm<1>(a, *prog, s);
  if (lanes <= 2) return launch_vm<2>(a, *prog, s);
  if (lanes <= 4) return launch_vm<4>(a, *prog, s);
  if (lanes <= 8) return launch_vm<8>(a, *prog, s);
  if (lanes <= 16) return launch_vm<16>(a, *prog, s);
  return launch_vm<32>(a, *prog, s);
}
