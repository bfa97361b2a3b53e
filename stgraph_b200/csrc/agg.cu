// Fused neighbour aggregation (gather - scale - sum) for sm_100a.
//
//   out[r,:] = row_scale[r] * sum_{e in row r} nbr_scale[col[e]] * edge_scale[eid(e)] * x[col[e],:]
//
// This is the one kernel shape on STGraph's hot path: the reference emits it
// from stgraph/compiler/code_gen/templates/fa/tpl_fa_csr*.jinja:1-57 with one
// thread per (row, feature) and a serial edge loop of scalar loads.  Here:
//   * a GROUP of lanes (1..32, power of two) owns a row and each lane owns VEC
//     (1/2/4) consecutive floats x NACC chunks, so one neighbour row is fetched
//     with 128-bit coalesced loads;
//   * column indices / scales of the next GROUP edges are loaded cooperatively
//     (one coalesced load per group) and prefetched one batch ahead, then
//     broadcast with warp shuffles;
//   * UNROLL independent neighbour-row loads are in flight per lane;
//   * rows longer than hub_threshold go to a block-per-row kernel whose warps
//     split the row and reduce through shared memory in a fixed order
//     (deterministic: no atomics anywhere).
// Memory bound: HBM roofline; algorithmic bytes per launch
//   4*(2*N*F + E + (N+1) + 2*N [+ E for edge_scale]).
#include "common.cuh"

#include <stdlib.h>

namespace stg {
namespace {

constexpr int kBlockThreads = 256;
constexpr int kHubThreads = 512;

template <int VEC>
__device__ __forceinline__ typename VecT<VEC>::type ld_row(const float* p) {
  using T = typename VecT<VEC>::type;
  return __ldg(reinterpret_cast<const T*>(p));
}
template <int VEC>
__device__ __forceinline__ void st_row(float* p, typename VecT<VEC>::type v) {
  using T = typename VecT<VEC>::type;
  __stcs(reinterpret_cast<T*>(p), v);
}

struct AggParams {
  const int32_t* __restrict__ row_off;
  const int32_t* __restrict__ col;
  const int32_t* __restrict__ eids;
  const int32_t* __restrict__ hub_rows;
  const int32_t* __restrict__ hub_count;
  int num_rows;
  int eid_base;
  int eids_identity;
  int hub_threshold;
  int hub_capacity;
  const float* __restrict__ x;    // already offset to the chunk's first column
  float* __restrict__ out;        // same
  int ld;                         // floats between consecutive rows (= feat)
  int width;                      // floats of this chunk handled by the launch
  const float* __restrict__ ns;
  const float* __restrict__ es;
  const float* __restrict__ rs;
};

// Accumulate edges [beg,end) visited with stride `step` batches of GROUP edges,
// starting at batch `first`.  All lanes of a group execute this together.
template <int VEC, int GROUP, int NACC>
__device__ __forceinline__ void accumulate_edges(const AggParams& p, int beg, int end, int first_batch,
                                                 int batch_step, int gl, unsigned gmask,
                                                 typename VecT<VEC>::type (&acc)[NACC]) {
  using T = typename VecT<VEC>::type;
  constexpr int UNROLL = (GROUP >= 8 ? 8 : GROUP) / (NACC > 2 ? 2 : 1);
  bool act[NACC];
  int off[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    off[k] = (gl + k * GROUP) * VEC;
    act[k] = off[k] < p.width;
  }

  auto load_meta = [&](int base, int& c, float& s) {
    c = 0;
    s = 0.f;
    const int e = base + gl;
    if (e < end) {
      c = ld_stream(p.col + e);
      float sc = 1.f;
      if (p.ns) sc = __ldg(p.ns + c);
      if (p.es) {
        const int eid = p.eids_identity ? e : (ld_stream(p.eids + e) - p.eid_base);
        sc *= __ldg(p.es + eid);
      }
      s = sc;
    }
  };

  int base = beg + first_batch * GROUP;
  int my_c, nx_c;
  float my_s, nx_s;
  load_meta(base, my_c, my_s);
  for (; base < end; base += batch_step * GROUP) {
    load_meta(base + batch_step * GROUP, nx_c, nx_s);   // prefetch next batch
    const int n = min(GROUP, end - base);
    for (int j = 0; j < n; j += UNROLL) {
      int c[UNROLL];
      float s[UNROLL];
      T v[UNROLL][NACC];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        c[u] = __shfl_sync(gmask, my_c, j + u, GROUP);
        s[u] = __shfl_sync(gmask, my_s, j + u, GROUP);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const float* src = p.x + static_cast<size_t>(c[u]) * p.ld;
#pragma unroll
        for (int k = 0; k < NACC; ++k) {
          if (act[k] && (j + u) < n) v[u][k] = ld_row<VEC>(src + off[k]);
          else zero_vec(v[u][k]);
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) fma_vec(acc[k], s[u], v[u][k]);
      }
    }
    my_c = nx_c;
    my_s = nx_s;
  }
}

// One group per row (baseline variant, kept for A/B measurements: STG_AGG_VARIANT=rows).
template <int VEC, int GROUP, int NACC>
__global__ void __launch_bounds__(kBlockThreads) agg_rows_kernel(const AggParams p) {
  using T = typename VecT<VEC>::type;
  constexpr int GROUPS_PER_WARP = 32 / GROUP;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (GROUP - 1);
  const unsigned gmask = (GROUP == 32) ? 0xffffffffu : (((1u << GROUP) - 1u) << (lane & ~(GROUP - 1)));
  const int warp = blockIdx.x * (kBlockThreads / 32) + (threadIdx.x >> 5);
  const int row = warp * GROUPS_PER_WARP + lane / GROUP;
  if (row >= p.num_rows) return;
  const int beg = __ldg(p.row_off + row);
  const int end = __ldg(p.row_off + row + 1);
  if (p.hub_threshold > 0 && (end - beg) > p.hub_threshold) return;  // hub kernel owns this row

  T acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) zero_vec(acc[k]);
  accumulate_edges<VEC, GROUP, NACC>(p, beg, end, 0, 1, gl, gmask, acc);

  const float r = p.rs ? __ldg(p.rs + row) : 1.f;
  float* dst = p.out + static_cast<size_t>(row) * p.ld;
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    const int o = (gl + k * GROUP) * VEC;
    if (o < p.width) {
      scale_vec(acc[k], r);
      st_row<VEC>(dst + o, acc[k]);
    }
  }
}

// Rows-per-group of the streaming kernel: a group owns kRowsPerGroup consecutive rows and
// walks their (contiguous) CSR edge range as ONE stream, so the load pipeline (index prefetch
// one batch ahead, UNROLL neighbour rows in flight) never drains at a row boundary -- most
// rows of a power-law graph are shorter than one batch.
constexpr int kRowsPerGroup = 8;

template <int VEC, int GROUP, int NACC>
__global__ void __launch_bounds__(kBlockThreads) agg_stream_kernel(const AggParams p) {
  using T = typename VecT<VEC>::type;
  constexpr int GROUPS_PER_WARP = 32 / GROUP;
  constexpr int UNROLL = (GROUP >= 8 ? 8 : GROUP) / (NACC > 2 ? 2 : 1);
  const int lane = threadIdx.x & 31;
  const int gl = lane & (GROUP - 1);
  const unsigned gmask = (GROUP == 32) ? 0xffffffffu : (((1u << GROUP) - 1u) << (lane & ~(GROUP - 1)));
  const int warp = blockIdx.x * (kBlockThreads / 32) + (threadIdx.x >> 5);
  const int group = warp * GROUPS_PER_WARP + lane / GROUP;
  const int r0 = group * kRowsPerGroup;
  if (r0 >= p.num_rows) return;
  const int r1 = min(r0 + kRowsPerGroup, p.num_rows);
  const int thr = p.hub_threshold;

  bool act[NACC];
  int off[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    off[k] = (gl + k * GROUP) * VEC;
    act[k] = off[k] < p.width;
  }
  T acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) zero_vec(acc[k]);

  auto flush = [&](int row) {
    const float r = p.rs ? __ldg(p.rs + row) : 1.f;
    float* dst = p.out + static_cast<size_t>(row) * p.ld;
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
      if (act[k]) {
        scale_vec(acc[k], r);
        st_row<VEC>(dst + off[k], acc[k]);
      }
      zero_vec(acc[k]);
    }
  };

  int cur = r0;
  while (cur < r1) {
    // ---- next run of consecutive non-hub rows [cur, rr) with edges [e_lo, e_hi)
    const int e_lo = __ldg(p.row_off + cur);
    int e_hi = __ldg(p.row_off + cur + 1);
    if (thr > 0 && (e_hi - e_lo) > thr) { ++cur; continue; }   // hub kernel owns this row
    int rr = cur + 1;
    if (thr > 0) {
      while (rr < r1) {
        const int ne = __ldg(p.row_off + rr + 1);
        if (ne - e_hi > thr) break;
        e_hi = ne;
        ++rr;
      }
    } else {
      rr = r1;
      e_hi = __ldg(p.row_off + r1);
    }
    int cur_end = __ldg(p.row_off + cur + 1);

    auto load_meta = [&](int base, int& c, float& s) {
      c = 0;
      s = 0.f;
      const int e = base + gl;
      if (e < e_hi) {
        c = ld_stream(p.col + e);
        float sc = 1.f;
        if (p.ns) sc = __ldg(p.ns + c);
        if (p.es) {
          const int eid = p.eids_identity ? e : (ld_stream(p.eids + e) - p.eid_base);
          sc *= __ldg(p.es + eid);
        }
        s = sc;
      }
    };

    int my_c, nx_c;
    float my_s, nx_s;
    load_meta(e_lo, my_c, my_s);
    for (int base = e_lo; base < e_hi; base += GROUP) {
      load_meta(base + GROUP, nx_c, nx_s);
      const int n = min(GROUP, e_hi - base);
      for (int j = 0; j < n; j += UNROLL) {
        float s[UNROLL];
        T v[UNROLL][NACC];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const int c = __shfl_sync(gmask, my_c, j + u, GROUP);
          s[u] = __shfl_sync(gmask, my_s, j + u, GROUP);
          const float* src = p.x + static_cast<size_t>(c) * p.ld;
#pragma unroll
          for (int k = 0; k < NACC; ++k) {
            if (act[k] && (j + u) < n) v[u][k] = ld_row<VEC>(src + off[k]);
            else zero_vec(v[u][k]);
          }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          if (j + u < n) {
            const int e = base + j + u;
            while (e >= cur_end) {          // crossed into the next row (empty rows flush zeros)
              flush(cur);
              ++cur;
              cur_end = __ldg(p.row_off + cur + 1);
            }
#pragma unroll
            for (int k = 0; k < NACC; ++k) fma_vec(acc[k], s[u], v[u][k]);
          }
        }
      }
      my_c = nx_c;
      my_s = nx_s;
    }
    while (cur < rr) {   // last row of the run and trailing empty rows
      flush(cur);
      ++cur;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Asynchronous-copy gather pipeline (cp.async -> SASS LDGSTS).
// The LDG kernels above are latency bound: ncu shows long_scoreboard as the only stall that
// matters, ~20 warps x <=8 neighbour rows in flight per SM, DRAM and L2 both below 30 %
// (profiles/r01_agg_ncu.md).  Registers limit the bytes in flight, so this variant copies each
// neighbour row with ONE warp-wide cp.async.cg (16 B per lane, L2 -> shared memory, no register
// staging) into a per-warp ring of K*G rows; commit groups of G rows give a K-deep pipeline and
// consumption is a conflict-free LDS.128 per row.  A warp walks kRowsPerGroup rows as one
// contiguous edge stream and flushes the accumulator whenever the stream crosses a row end.
// (A cp.async.bulk/UBLKCP version was measured first: one bulk copy per 400-512 B row runs at
// ~40-70 cycles per copy per SM and was 2-4x SLOWER; see profiles/r01_agg_ncu.md.)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int G, int K>
__global__ void __launch_bounds__(kBlockThreads) agg_async_kernel(const AggParams p) {
  extern __shared__ __align__(128) unsigned char async_smem[];
  constexpr int RR = G * K;        // rows in the data ring
  constexpr int RS = 64;           // scales ring (>= 32 + RR is required; RR <= 32)
  static_assert(RR <= 32 && 32 % G == 0, "ring geometry");
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int rowb = p.width * 4;                                   // bytes per neighbour row (multiple of 16)
  unsigned char* ring = async_smem + static_cast<size_t>(wid) * RR * rowb;
  float* scales = reinterpret_cast<float*>(async_smem + static_cast<size_t>(nwarps) * RR * rowb) + wid * RS;
  const uint32_t ring_s = smem_u32(ring) + lane * 16;

  const int warp = blockIdx.x * nwarps + wid;
  const int r0 = warp * kRowsPerGroup;
  if (r0 >= p.num_rows) return;
  const int r1 = min(r0 + kRowsPerGroup, p.num_rows);
  const int thr = p.hub_threshold;
  const bool act = lane * 4 < p.width;
  const float* xl = p.x + lane * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);

  auto flush = [&](int row) {
    const float r = p.rs ? __ldg(p.rs + row) : 1.f;
    if (act) {
      scale_vec(acc, r);
      st_row<4>(p.out + static_cast<size_t>(row) * p.ld + lane * 4, acc);
    }
    acc = make_float4(0.f, 0.f, 0.f, 0.f);
  };

  int cur = r0;
  while (cur < r1) {
    const int e_lo = __ldg(p.row_off + cur);
    int e_hi = __ldg(p.row_off + cur + 1);
    if (thr > 0 && (e_hi - e_lo) > thr) { ++cur; continue; }   // hub kernel owns this row
    int rr = cur + 1;
    if (thr > 0) {
      while (rr < r1) {
        const int ne = __ldg(p.row_off + rr + 1);
        if (ne - e_hi > thr) break;
        e_hi = ne;
        ++rr;
      }
    } else {
      rr = r1;
      e_hi = __ldg(p.row_off + r1);
    }
    int cur_end = __ldg(p.row_off + cur + 1);
    const int total = e_hi - e_lo;
    const int ngroups = (total + G - 1) / G;
    int my_c = 0;

    for (int gi = 0; gi < ngroups + K - 1; ++gi) {
      if (gi < ngroups) {
        const int idx0 = gi * G;                     // stream index of the group's first edge
        if ((idx0 & 31) == 0) {                      // new batch of 32 edges: cooperative index/scale load
          const int e = e_lo + idx0 + lane;
          my_c = 0;
          float sc = 0.f;
          if (e < e_hi) {
            my_c = ld_stream(p.col + e);
            sc = 1.f;
            if (p.ns) sc = __ldg(p.ns + my_c);
            if (p.es) {
              const int eid = p.eids_identity ? e : (ld_stream(p.eids + e) - p.eid_base);
              sc *= __ldg(p.es + eid);
            }
          }
          scales[(idx0 + lane) & (RS - 1)] = sc;
        }
        const int n = min(G, total - idx0);
#pragma unroll
        for (int u = 0; u < G; ++u) {
          const int c = __shfl_sync(0xffffffffu, my_c, (idx0 + u) & 31);
          if (u < n && act)
            cp_async16(ring_s + ((idx0 + u) % RR) * rowb, xl + static_cast<size_t>(c) * p.ld);
        }
      }
      cp_async_commit();
      if (gi >= K - 1) {
        cp_async_wait<K - 1>();
        __syncwarp();
        const int idx0 = (gi - (K - 1)) * G;
        const int n = min(G, total - idx0);
#pragma unroll
        for (int u = 0; u < G; ++u) {
          if (u < n) {
            const int e = e_lo + idx0 + u;
            while (e >= cur_end) {
              flush(cur);
              ++cur;
              cur_end = __ldg(p.row_off + cur + 1);
            }
            const float s = scales[(idx0 + u) & (RS - 1)];
            if (act) {
              const float4 v = *reinterpret_cast<const float4*>(ring + ((idx0 + u) % RR) * rowb + lane * 16);
              fma_vec(acc, s, v);
            }
          }
        }
        __syncwarp();          // slots of this group are refilled by the next iteration's copies
      }
    }
    cp_async_wait<0>();
    __syncwarp();
    while (cur < rr) {
      flush(cur);
      ++cur;
    }
  }
}

// Block per hub row: every (warp, group) pair takes a strided share of the row's
// edge batches; partials are reduced group->warp by shuffles and warp->block
// through shared memory in a fixed order.
template <int VEC, int GROUP, int NACC>
__global__ void __launch_bounds__(kHubThreads) agg_hub_kernel(const AggParams p) {
  using T = typename VecT<VEC>::type;
  constexpr int GROUPS_PER_WARP = 32 / GROUP;
  constexpr int WARPS = kHubThreads / 32;
  __shared__ T partial[WARPS][GROUP * NACC];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int gl = lane & (GROUP - 1);
  const int gidx = lane / GROUP;
  const unsigned gmask = (GROUP == 32) ? 0xffffffffu : (((1u << GROUP) - 1u) << (lane & ~(GROUP - 1)));
  const int n_hub = min(__ldg(p.hub_count), p.hub_capacity);
  for (int i = blockIdx.x; i < n_hub; i += gridDim.x) {
    const int row = __ldg(p.hub_rows + i);
    const int beg = __ldg(p.row_off + row);
    const int end = __ldg(p.row_off + row + 1);
    T acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) zero_vec(acc[k]);
    accumulate_edges<VEC, GROUP, NACC>(p, beg, end, wid * GROUPS_PER_WARP + gidx,
                                       WARPS * GROUPS_PER_WARP, gl, gmask, acc);
    // groups of one warp -> lanes [0, GROUP)
#pragma unroll
    for (int o = GROUP; o < 32; o <<= 1) {
#pragma unroll
      for (int k = 0; k < NACC; ++k) {
        T other;
        if constexpr (VEC == 1) {
          other = __shfl_xor_sync(0xffffffffu, acc[k], o);
        } else if constexpr (VEC == 2) {
          other.x = __shfl_xor_sync(0xffffffffu, acc[k].x, o);
          other.y = __shfl_xor_sync(0xffffffffu, acc[k].y, o);
        } else {
          other.x = __shfl_xor_sync(0xffffffffu, acc[k].x, o);
          other.y = __shfl_xor_sync(0xffffffffu, acc[k].y, o);
          other.z = __shfl_xor_sync(0xffffffffu, acc[k].z, o);
          other.w = __shfl_xor_sync(0xffffffffu, acc[k].w, o);
        }
        add_vec(acc[k], other);
      }
    }
    if (lane < GROUP) {
#pragma unroll
      for (int k = 0; k < NACC; ++k) partial[wid][k * GROUP + lane] = acc[k];
    }
    __syncthreads();
    if (wid == 0 && lane < GROUP) {
      const float r = p.rs ? __ldg(p.rs + row) : 1.f;
      float* dst = p.out + static_cast<size_t>(row) * p.ld;
#pragma unroll
      for (int k = 0; k < NACC; ++k) {
        T sum = partial[0][k * GROUP + lane];
        for (int w = 1; w < WARPS; ++w) add_vec(sum, partial[w][k * GROUP + lane]);
        const int o = (lane + k * GROUP) * VEC;
        if (o < p.width) {
          scale_vec(sum, r);
          st_row<VEC>(dst + o, sum);
        }
      }
    }
    __syncthreads();
  }
}

// 0 = group per row, 1 = streaming multi-row groups, 2 = cp.async ring pipeline
// (env STG_AGG_VARIANT=rows|stream|async)
inline int agg_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("STG_AGG_VARIANT");
    v = (e && e[0] == 's') ? 1 : (e && e[0] == 'a') ? 2 : 0;
  }
  return v;
}

inline int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

template <int G, int K>
int launch_async_cfg(const AggParams& p, int warps, cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(warps) * (static_cast<size_t>(G * K) * p.width * 4 + 64 * 4) + 128;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(agg_async_kernel<G, K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) {
      set_error("cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
      return STG_ERR_CUDA;
    }
    configured = smem;
  }
  const int rows_per_block = warps * kRowsPerGroup;
  const int blocks = (p.num_rows + rows_per_block - 1) / rows_per_block;
  agg_async_kernel<G, K><<<blocks, warps * 32, smem, stream>>>(p);
  STG_LAUNCH_CHECK("agg_async_kernel");
  return STG_OK;
}

int launch_hub_only(const AggParams& p, cudaStream_t stream);

inline int launch_async(const AggParams& p, cudaStream_t stream) {
  static const int g = env_int("STG_ASYNC_G", 8), k = env_int("STG_ASYNC_K", 4), warps = env_int("STG_ASYNC_WARPS", 8);
  int rc;
  if (g == 8 && k == 2) rc = launch_async_cfg<8, 2>(p, warps, stream);
  else if (g == 8 && k == 3) rc = launch_async_cfg<8, 3>(p, warps, stream);
  else if (g == 4 && k == 4) rc = launch_async_cfg<4, 4>(p, warps, stream);
  else if (g == 4 && k == 8) rc = launch_async_cfg<4, 8>(p, warps, stream);
  else if (g == 16 && k == 2) rc = launch_async_cfg<16, 2>(p, warps, stream);
  else if (g == 2 && k == 8) rc = launch_async_cfg<2, 8>(p, warps, stream);
  else rc = launch_async_cfg<8, 4>(p, warps, stream);
  if (rc != STG_OK) return rc;
  return launch_hub_only(p, stream);
}

template <int VEC, int GROUP, int NACC>
int launch_agg(const AggParams& p, cudaStream_t stream) {
  if (VEC == 4 && GROUP == 32 && NACC == 1 && agg_variant() == 2) {
    return launch_async(p, stream);
  } else if (agg_variant() == 1) {
    constexpr int rows_per_block = (kBlockThreads / 32) * (32 / GROUP) * kRowsPerGroup;
    const int blocks = (p.num_rows + rows_per_block - 1) / rows_per_block;
    if (blocks > 0) {
      agg_stream_kernel<VEC, GROUP, NACC><<<blocks, kBlockThreads, 0, stream>>>(p);
      STG_LAUNCH_CHECK("agg_stream_kernel");
    }
  } else {
    constexpr int rows_per_block = (kBlockThreads / 32) * (32 / GROUP);
    const int blocks = (p.num_rows + rows_per_block - 1) / rows_per_block;
    if (blocks > 0) {
      agg_rows_kernel<VEC, GROUP, NACC><<<blocks, kBlockThreads, 0, stream>>>(p);
      STG_LAUNCH_CHECK("agg_rows_kernel");
    }
  }
  if (p.hub_threshold > 0 && p.hub_rows != nullptr) {
    agg_hub_kernel<VEC, GROUP, NACC><<<2 * sm_count(), kHubThreads, 0, stream>>>(p);
    STG_LAUNCH_CHECK("agg_hub_kernel");
  }
  return STG_OK;
}

int launch_hub_only(const AggParams& p, cudaStream_t stream) {
  if (p.hub_threshold > 0 && p.hub_rows != nullptr) {
    agg_hub_kernel<4, 32, 1><<<2 * sm_count(), kHubThreads, 0, stream>>>(p);
    STG_LAUNCH_CHECK("agg_hub_kernel");
  }
  return STG_OK;
}

template <int VEC>
int dispatch_group(const AggParams& p, cudaStream_t stream) {
  const int nvec = p.width / VEC;
  if (nvec <= 1) return launch_agg<VEC, 1, 1>(p, stream);
  if (nvec <= 2) return launch_agg<VEC, 2, 1>(p, stream);
  if (nvec <= 4) return launch_agg<VEC, 4, 1>(p, stream);
  if (nvec <= 8) return launch_agg<VEC, 8, 1>(p, stream);
  if (nvec <= 16) return launch_agg<VEC, 16, 1>(p, stream);
  if (nvec <= 32) return launch_agg<VEC, 32, 1>(p, stream);
  if (nvec <= 64) return launch_agg<VEC, 32, 2>(p, stream);
  return launch_agg<VEC, 32, 4>(p, stream);
}

}  // namespace

int agg_scaled_sum_device(const StgCsrView* g, const float* x, int32_t feat, const float* ns,
                          const float* es, const float* rs, float* out, cudaStream_t stream) {
  AggParams p;
  p.row_off = g->row_offset;
  p.col = g->column_indices;
  p.eids = g->eids;
  p.hub_rows = g->hub_rows;
  p.hub_count = g->hub_count;
  p.num_rows = g->num_nodes;
  p.eid_base = g->eid_base;
  p.eids_identity = g->eids_identity;
  p.hub_threshold = (g->hub_rows && g->hub_count) ? g->hub_threshold : 0;
  p.hub_capacity = g->hub_capacity;
  p.ld = feat;
  p.ns = ns;
  p.es = es;
  p.rs = rs;
  int vec = 1;
  if (feat % 4 == 0 && aligned16(x) && aligned16(out)) vec = 4;
  else if (feat % 2 == 0 && aligned8(x) && aligned8(out)) vec = 2;
  const int chunk = 32 * 4 * vec;  // widest tile one launch covers
  for (int f0 = 0; f0 < feat; f0 += chunk) {
    p.x = x + f0;
    p.out = out + f0;
    p.width = min(chunk, feat - f0);
    int rc;
    if (vec == 4) rc = dispatch_group<4>(p, stream);
    else if (vec == 2) rc = dispatch_group<2>(p, stream);
    else rc = dispatch_group<1>(p, stream);
    if (rc != STG_OK) return rc;
  }
  return STG_OK;
}

}  // namespace stg

using namespace stg;

static int validate_view(const StgCsrView* g, bool need_eids) {
  STG_CHECK_ARG(g != nullptr, "graph view is NULL");
  STG_CHECK_ARG(g->num_nodes >= 0 && g->num_edges >= 0, "negative graph size (%d nodes, %d edges)",
                g->num_nodes, g->num_edges);
  STG_CHECK_ARG(g->row_offset != nullptr, "row_offset is NULL");
  STG_CHECK_ARG(g->num_edges == 0 || g->column_indices != nullptr, "column_indices is NULL");
  STG_CHECK_ARG(!need_eids || g->eids_identity || g->num_edges == 0 || g->eids != nullptr,
                "eids is NULL but an edge tensor is indexed by edge id");
  return STG_OK;
}

STG_API int stg_agg_scaled_sum_f32(const StgCsrView* g, const float* x, int32_t feat,
                                      const float* nbr_scale, const float* edge_scale,
                                      const float* row_scale, float* out, void* stream) {
  int rc = validate_view(g, edge_scale != nullptr);
  if (rc != STG_OK) return rc;
  STG_CHECK_ARG(feat > 0, "feat must be positive (got %d)", feat);
  if (g->num_nodes == 0) return STG_OK;
  STG_CHECK_ARG(x != nullptr && out != nullptr, "x / out is NULL");
  STG_CHECK_ARG(x != out, "x and out must not alias");
  return agg_scaled_sum_device(g, x, feat, nbr_scale, edge_scale, row_scale, out, as_stream(stream));
}

STG_API int stg_agg_scaled_sum_f32_host(const StgCsrView* g, const float* x_host, int32_t feat,
                                           const float* nbr_scale_host, const float* edge_scale_host,
                                           const float* row_scale_host, float* out_host,
                                           void* dev_scratch, size_t dev_scratch_bytes, void* stream) {
  int rc = validate_view(g, edge_scale_host != nullptr);
  if (rc != STG_OK) return rc;
  STG_CHECK_ARG(feat > 0, "feat must be positive (got %d)", feat);
  if (g->num_nodes == 0) return STG_OK;
  STG_CHECK_ARG(x_host && out_host && dev_scratch, "host buffers / scratch must not be NULL");
  const size_t n = static_cast<size_t>(g->num_nodes), e = static_cast<size_t>(g->num_edges);
  const size_t nf = align_up(n * feat, 4), n4 = align_up(n, 4), e4 = align_up(e, 4);
  const size_t need = (2 * nf + 2 * n4 + e4) * sizeof(float);
  if (dev_scratch_bytes < need) {
    set_error("device scratch too small: %zu bytes given, %zu needed", dev_scratch_bytes, need);
    return STG_ERR_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t s = as_stream(stream);
  float* dx = static_cast<float*>(dev_scratch);
  float* dout = dx + nf;
  float* dns = dout + nf;
  float* drs = dns + n4;
  float* des = drs + n4;
  STG_CUDA(cudaMemcpyAsync(dx, x_host, n * feat * sizeof(float), cudaMemcpyHostToDevice, s));
  if (nbr_scale_host) STG_CUDA(cudaMemcpyAsync(dns, nbr_scale_host, n * sizeof(float), cudaMemcpyHostToDevice, s));
  if (row_scale_host) STG_CUDA(cudaMemcpyAsync(drs, row_scale_host, n * sizeof(float), cudaMemcpyHostToDevice, s));
  if (edge_scale_host) STG_CUDA(cudaMemcpyAsync(des, edge_scale_host, e * sizeof(float), cudaMemcpyHostToDevice, s));
  rc = agg_scaled_sum_device(g, dx, feat, nbr_scale_host ? dns : nullptr, edge_scale_host ? des : nullptr,
                             row_scale_host ? drs : nullptr, dout, s);
  if (rc != STG_OK) return rc;
  STG_CUDA(cudaMemcpyAsync(out_host, dout, n * feat * sizeof(float), cudaMemcpyDeviceToHost, s));
  STG_CUDA(cudaStreamSynchronize(s));
  return STG_OK;
}
