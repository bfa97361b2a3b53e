// Halo pull over peer memory: copy the remote feature rows this rank's edges reference from the
// owners' blocks (mapped through CUDA IPC / symmetric memory, read with NVLink loads) into a local,
// compact halo buffer.  Runs on a side stream with a SMALL grid so it overlaps the aggregation pass over
// locally owned sources (dist/halo.py); NVLink needs ~1.5 MB in flight (775 GB/s x 2 us), which a few
// dozen CTAs with several independent 128-bit loads per lane provide.  Replaces an NCCL all-to-all
// whose send/recv kernels could not start while the aggregation kernel filled every SM.
// New functionality: the reference is single-GPU.
#include "common.cuh"

#include <algorithm>

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

// ---- copy-engine halo exchange: pack -> P-1 peer copies -> flags ------------------------------------------
// The SM push above shares the SMs (and the L1TEX queues) with the own-source aggregation pass it is meant to
// hide behind: measured at 8 GPUs it ran at 300 GB/s and slowed that pass from 0.32 to 0.6 ms.  Here the
// rows a peer needs are first packed into one contiguous send buffer (local gather, 2 x 181 MB of HBM traffic
// on config 5 at P=8: ~0.06 ms), then shipped by the copy engines with one cudaMemcpyAsync per peer straight
// into that peer's halo buffer (symmetric memory), and a one-warp kernel posts an arrival flag on every peer.
template <int VEC>
__global__ void __launch_bounds__(256) rows_gather_kernel(const float* __restrict__ own, int feat,
                                                          const int64_t* __restrict__ rows, int64_t n,
                                                          float* __restrict__ buf) {
  using T = typename VecT<VEC>::type;
  const int nvec = feat / VEC;
  const int64_t total = n * nvec;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  constexpr int U = 4;          // independent row-id -> row loads in flight per thread (the kernel shares the chip with
                                // the aggregation pass: it is latency-, not occupancy-, bound)
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < total; i += U * stride) {
    int64_t j[U];
    int v[U];
    const float* src[U];
    T val[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t ii = i + u * stride;
      j[u] = ii / nvec;
      v[u] = static_cast<int>(ii - j[u] * nvec);
      src[u] = own + static_cast<size_t>(__ldg(rows + j[u])) * feat;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) val[u] = __ldg(reinterpret_cast<const T*>(src[u]) + v[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) __stcs(reinterpret_cast<T*>(buf + static_cast<size_t>(j[u]) * feat) + v[u], val[u]);
  }
  for (; i < total; i += stride) {
    const int64_t j = i / nvec;
    const int v = static_cast<int>(i - j * nvec);
    const T val = __ldg(reinterpret_cast<const T*>(own + static_cast<size_t>(__ldg(rows + j)) * feat) + v);
    __stcs(reinterpret_cast<T*>(buf + static_cast<size_t>(j) * feat) + v, val);
  }
}

struct SignalParams {
  int32_t* flag[STG_MAX_PARTS];   // flag[q] = address of MY slot in peer q's flag array (peer memory)
  int nparts, rank, value;
};

// Everything enqueued before this kernel on its stream (the peer copies) has completed; publish that.
__global__ void peer_signal_kernel(const SignalParams p) {
  const int q = threadIdx.x;
  if (q < p.nparts && q != p.rank && p.flag[q] != nullptr) {
    __threadfence_system();
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p.flag[q]), "r"(p.value) : "memory");
  }
}

// Spin until every peer's flag has reached `value` (flags only grow).  Gives up after `timeout_cycles` SM
// clocks and raises *status so that a lost peer cannot hang the device (the host checks status later).
__global__ void peer_wait_kernel(const int32_t* flags, int nparts, int rank, int value, long long timeout_cycles,
                                 int32_t* status) {
  const int q = threadIdx.x;
  if (q < nparts && q != rank) {
    const long long t0 = clock64();
    int v;
    for (;;) {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + q) : "memory");
      if (((v - value) & 0xFFFF) < 0x8000) break;        // sequence numbers live modulo 2^16 (stg_halo_exchange_f32)
      if (clock64() - t0 > timeout_cycles) {
        if (status) atomicExch(status, 1 + q);
        break;
      }
      __nanosleep(200);
    }
  }
}

}  // namespace
}  // namespace stg

using namespace stg;

STG_API int stg_rows_gather_f32(const float* own, int32_t feat, const int64_t* rows, int64_t n, float* buf,
                                int32_t max_blocks, void* stream) {
  STG_CHECK_ARG(feat > 0 && n >= 0, "bad sizes");
  if (n == 0) return STG_OK;
  STG_CHECK_ARG(own && rows && buf, "NULL argument");
  const bool v4 = feat % 4 == 0 && aligned16(own) && aligned16(buf);
  const int64_t items = n * (v4 ? feat / 4 : feat);
  int blocks = static_cast<int>(std::min<int64_t>((items + 255) / 256, max_blocks > 0 ? max_blocks : 4 * sm_count()));
  if (v4) rows_gather_kernel<4><<<blocks, 256, 0, as_stream(stream)>>>(own, feat, rows, n, buf);
  else rows_gather_kernel<1><<<blocks, 256, 0, as_stream(stream)>>>(own, feat, rows, n, buf);
  STG_LAUNCH_CHECK("rows_gather_kernel");
  return STG_OK;
}

STG_API int stg_halo_send_f32(const float* send_buf, int32_t feat, int32_t num_parts, int32_t my_rank,
                              const int64_t* send_off, float* const* peer_dst, void* stream) {
  STG_CHECK_ARG(num_parts >= 1 && num_parts <= STG_MAX_PARTS, "num_parts must be in [1, %d]", STG_MAX_PARTS);
  STG_CHECK_ARG(my_rank >= 0 && my_rank < num_parts && feat > 0, "bad rank / feat");
  STG_CHECK_ARG(send_off && peer_dst, "NULL argument");
  for (int i = 1; i < num_parts; ++i) {        // start with the next rank: the P copies of a step fan out over the switch
    const int q = (my_rank + i) % num_parts;
    const int64_t rows = send_off[q + 1] - send_off[q];
    STG_CHECK_ARG(rows >= 0, "send_off must be non-decreasing");
    if (rows == 0) continue;
    STG_CHECK_ARG(send_buf && peer_dst[q], "NULL buffer for a non-empty segment (peer %d)", q);
    STG_CUDA(cudaMemcpyAsync(peer_dst[q], send_buf + static_cast<size_t>(send_off[q]) * feat,
                             static_cast<size_t>(rows) * feat * sizeof(float), cudaMemcpyDeviceToDevice, as_stream(stream)));
  }
  return STG_OK;
}

STG_API int stg_peer_signal(int32_t* const* peer_flags, int32_t num_parts, int32_t my_rank, int32_t value, void* stream) {
  STG_CHECK_ARG(num_parts >= 1 && num_parts <= STG_MAX_PARTS, "num_parts must be in [1, %d]", STG_MAX_PARTS);
  STG_CHECK_ARG(peer_flags && my_rank >= 0 && my_rank < num_parts, "bad arguments");
  SignalParams p;
  for (int q = 0; q < STG_MAX_PARTS; ++q) p.flag[q] = q < num_parts ? peer_flags[q] : nullptr;
  p.nparts = num_parts;
  p.rank = my_rank;
  p.value = value;
  peer_signal_kernel<<<1, 32, 0, as_stream(stream)>>>(p);
  STG_LAUNCH_CHECK("peer_signal_kernel");
  return STG_OK;
}

STG_API int stg_peer_wait(const int32_t* flags, int32_t num_parts, int32_t my_rank, int32_t value, int64_t timeout_cycles,
                          int32_t* status, void* stream) {
  STG_CHECK_ARG(num_parts >= 1 && num_parts <= STG_MAX_PARTS, "num_parts must be in [1, %d]", STG_MAX_PARTS);
  STG_CHECK_ARG(flags && my_rank >= 0 && my_rank < num_parts, "bad arguments");
  peer_wait_kernel<<<1, 32, 0, as_stream(stream)>>>(flags, num_parts, my_rank, value,
                                                     timeout_cycles > 0 ? timeout_cycles : (4LL << 30), status);
  STG_LAUNCH_CHECK("peer_wait_kernel");
  return STG_OK;
}

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

STG_API int stg_halo_exchange_f32(const float* own, int32_t feat, const int64_t* send_rows, const int64_t* send_off,
                                  float* send_buf, float* const* peer_dst, int32_t* const* peer_flags,
                                  const int32_t* seq_values, int32_t value, int32_t num_parts, int32_t my_rank,
                                  int32_t gather_blocks, void* stream) {
  STG_CHECK_ARG(num_parts >= 1 && num_parts <= STG_MAX_PARTS, "num_parts must be in [1, %d]", STG_MAX_PARTS);
  STG_CHECK_ARG(my_rank >= 0 && my_rank < num_parts && feat > 0, "bad rank / feat");
  STG_CHECK_ARG(send_off && peer_dst && peer_flags && seq_values, "NULL argument");
  STG_CHECK_ARG(value >= 0 && value < 65536, "sequence value must be in [0, 65536)");
  const int64_t n_send = send_off[num_parts];
  int rc = stg_rows_gather_f32(own, feat, send_rows, n_send, send_buf, gather_blocks, stream);
  if (rc != STG_OK) return rc;
  rc = stg_halo_send_f32(send_buf, feat, num_parts, my_rank, send_off, peer_dst, stream);
  if (rc != STG_OK) return rc;
  // arrival flags by the copy engines too: 4 bytes of a device-resident table of sequence numbers, stream-ordered
  // behind the data copies -- a signalling KERNEL would have to wait for a free SM slot behind the persistent
  // aggregation grid (measured: 0.47 ms for a one-warp kernel).
  for (int i = 1; i < num_parts; ++i) {
    const int q = (my_rank + i) % num_parts;
    if (peer_flags[q] == nullptr) continue;
    STG_CUDA(cudaMemcpyAsync(peer_flags[q], seq_values + value, sizeof(int32_t), cudaMemcpyDeviceToDevice, as_stream(stream)));
  }
  return STG_OK;
}
