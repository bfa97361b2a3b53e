// Halo pull over peer memory: copy the remote feature rows this rank's edges reference from the
// owners' blocks (mapped through CUDA IPC / symmetric memory, read with NVLink loads) into a local,
// compact halo buffer.  Runs on a side stream with a SMALL grid so it overlaps the aggregation pass over
// locally owned sources (dist/halo.py); NVLink needs ~1.5 MB in flight (775 GB/s x 2 us), which a few
// dozen CTAs with several independent 128-bit loads per lane provide.  Replaces an NCCL all-to-all
// whose send/recv kernels could not start while the aggregation kernel filled every SM.
// New functionality: the reference is single-GPU.
#include "common.cuh"

namespace stg {
namespace {

constexpr int kPullThreads = 256;
constexpr int kRowsInFlight = 8;

struct PullParams {
  const int64_t* __restrict__ ids;   // sorted global row ids
  int64_t n_ids;
  int feat;
  float* __restrict__ out;
  int nparts;
  int bounds[STG_MAX_PARTS + 1];
  const float* xs[STG_MAX_PARTS];
};

__device__ __forceinline__ const float* owner_row(const PullParams& p, int64_t c) {
  int o = 0;
#pragma unroll
  for (int q = 1; q < STG_MAX_PARTS; ++q) o += (q < p.nparts && c >= p.bounds[q]) ? 1 : 0;
  return p.xs[o] + static_cast<size_t>(c - p.bounds[o]) * p.feat;
}

template <int VEC>
__global__ void __launch_bounds__(kPullThreads) halo_pull_kernel(const PullParams p) {
  using T = typename VecT<VEC>::type;
  const int lane = threadIdx.x & 31;
  const int64_t warp = blockIdx.x * (kPullThreads / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * (kPullThreads / 32);
  const int nvec = p.feat / VEC;
  for (int64_t i0 = warp * kRowsInFlight; i0 < p.n_ids; i0 += nwarps * kRowsInFlight) {
    const float* src[kRowsInFlight];
#pragma unroll
    for (int r = 0; r < kRowsInFlight; ++r) {
      const int64_t i = i0 + r;
      src[r] = i < p.n_ids ? owner_row(p, p.ids[i]) : nullptr;
    }
    for (int v0 = 0; v0 < nvec; v0 += 32) {
      const int v = v0 + lane;
      T val[kRowsInFlight];
#pragma unroll
      for (int r = 0; r < kRowsInFlight; ++r)
        if (src[r] != nullptr && v < nvec) val[r] = *reinterpret_cast<const T*>(src[r] + v * VEC);
#pragma unroll
      for (int r = 0; r < kRowsInFlight; ++r)
        if (src[r] != nullptr && v < nvec)
          *reinterpret_cast<T*>(p.out + static_cast<size_t>(i0 + r) * p.feat + v * VEC) = val[r];
    }
  }
}

// Halo push: the OWNER writes the rows its peers need straight into their halo buffers (peer memory,
// NVLink stores).  Stores are posted -- no round trip to hide -- so a small grid reaches link rate, and
// the local reads come from this GPU's own L2/HBM.  Item j copies local row send_rows[j] to row
// send_slot[j] of peer send_peer[j]'s halo buffer.
struct PushParams {
  const float* __restrict__ own;
  const int64_t* __restrict__ send_rows;
  const int32_t* __restrict__ send_peer;
  const int64_t* __restrict__ send_slot;
  int64_t n_items;
  int feat;
  float* halo[STG_MAX_PARTS];
};

template <int VEC>
__global__ void __launch_bounds__(kPullThreads) halo_push_kernel(const PushParams p) {
  using T = typename VecT<VEC>::type;
  const int lane = threadIdx.x & 31;
  const int64_t warp = blockIdx.x * (kPullThreads / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * (kPullThreads / 32);
  const int nvec = p.feat / VEC;
  for (int64_t i0 = warp * kRowsInFlight; i0 < p.n_items; i0 += nwarps * kRowsInFlight) {
    const float* src[kRowsInFlight];
    float* dst[kRowsInFlight];
#pragma unroll
    for (int r = 0; r < kRowsInFlight; ++r) {
      const int64_t i = i0 + r;
      src[r] = nullptr;
      dst[r] = nullptr;
      if (i < p.n_items) {
        src[r] = p.own + static_cast<size_t>(p.send_rows[i]) * p.feat;
        dst[r] = p.halo[p.send_peer[i]] + static_cast<size_t>(p.send_slot[i]) * p.feat;
      }
    }
    for (int v0 = 0; v0 < nvec; v0 += 32) {
      const int v = v0 + lane;
      T val[kRowsInFlight];
#pragma unroll
      for (int r = 0; r < kRowsInFlight; ++r)
        if (src[r] != nullptr && v < nvec) val[r] = __ldg(reinterpret_cast<const T*>(src[r] + v * VEC));
#pragma unroll
      for (int r = 0; r < kRowsInFlight; ++r)
        if (src[r] != nullptr && v < nvec) *reinterpret_cast<T*>(dst[r] + v * VEC) = val[r];
    }
  }
}

}  // namespace
}  // namespace stg

using namespace stg;

STG_API int stg_halo_push_f32(const float* own, int32_t feat, const int64_t* send_rows, const int32_t* send_peer,
                              const int64_t* send_slot, int64_t n_items, float* const* peer_halo, int32_t num_parts,
                              int32_t max_blocks, void* stream) {
  STG_CHECK_ARG(num_parts >= 1 && num_parts <= STG_MAX_PARTS, "num_parts must be in [1, %d]", STG_MAX_PARTS);
  STG_CHECK_ARG(feat > 0 && n_items >= 0, "bad sizes");
  if (n_items == 0) return STG_OK;
  STG_CHECK_ARG(own && send_rows && send_peer && send_slot && peer_halo, "NULL argument");
  PushParams p;
  p.own = own;
  p.send_rows = send_rows;
  p.send_peer = send_peer;
  p.send_slot = send_slot;
  p.n_items = n_items;
  p.feat = feat;
  bool al16 = aligned16(own);
  for (int q = 0; q < STG_MAX_PARTS; ++q) {
    p.halo[q] = q < num_parts ? peer_halo[q] : nullptr;
    if (q < num_parts && peer_halo[q]) al16 = al16 && aligned16(peer_halo[q]);
  }
  int blocks = max_blocks > 0 ? max_blocks : 32;
  const int64_t need = (n_items + (kPullThreads / 32) * kRowsInFlight - 1) / ((kPullThreads / 32) * kRowsInFlight);
  if (need < blocks) blocks = static_cast<int>(need);
  if (feat % 4 == 0 && al16) halo_push_kernel<4><<<blocks, kPullThreads, 0, as_stream(stream)>>>(p);
  else halo_push_kernel<1><<<blocks, kPullThreads, 0, as_stream(stream)>>>(p);
  STG_LAUNCH_CHECK("halo_push_kernel");
  return STG_OK;
}

STG_API int stg_halo_pull_f32(const float* const* x_parts, const int32_t* part_bounds, int32_t num_parts,
                              const int64_t* ids, int64_t n_ids, int32_t feat, float* out, int32_t max_blocks,
                              void* stream) {
  STG_CHECK_ARG(num_parts >= 1 && num_parts <= STG_MAX_PARTS, "num_parts must be in [1, %d]", STG_MAX_PARTS);
  STG_CHECK_ARG(feat > 0 && n_ids >= 0, "bad sizes");
  if (n_ids == 0) return STG_OK;
  STG_CHECK_ARG(x_parts && part_bounds && ids && out, "NULL argument");
  PullParams p;
  p.ids = ids;
  p.n_ids = n_ids;
  p.feat = feat;
  p.out = out;
  p.nparts = num_parts;
  bool al16 = aligned16(out);
  for (int q = 0; q <= STG_MAX_PARTS; ++q) p.bounds[q] = q <= num_parts ? part_bounds[q] : 0;
  for (int q = 0; q < STG_MAX_PARTS; ++q) {
    p.xs[q] = q < num_parts ? x_parts[q] : nullptr;
    if (q < num_parts && x_parts[q]) al16 = al16 && aligned16(x_parts[q]);
  }
  int blocks = max_blocks > 0 ? max_blocks : 32;
  const int64_t need = (n_ids + (kPullThreads / 32) * kRowsInFlight - 1) / ((kPullThreads / 32) * kRowsInFlight);
  if (need < blocks) blocks = static_cast<int>(need);
  if (feat % 4 == 0 && al16) halo_pull_kernel<4><<<blocks, kPullThreads, 0, as_stream(stream)>>>(p);
  else halo_pull_kernel<1><<<blocks, kPullThreads, 0, as_stream(stream)>>>(p);
  STG_LAUNCH_CHECK("halo_pull_kernel");
  return STG_OK;
}
