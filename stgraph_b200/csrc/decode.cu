// Link-prediction decode of the dynamic-temporal benchmark: score[p] = <z[a[p]], z[b[p]]> for a list of vertex
// pairs (benchmarking/dynamic-temporal-tgcn/seastar/model.py:18-21: (z[idx[0]] * z[idx[1]]).sum(-1), which torch runs
// as two gathers that materialise [P, F] twice, a multiply and a row reduction), and its backward
//   d_z[a[p]] += g[p] * z[b[p]],   d_z[b[p]] += g[p] * z[a[p]].
// One lane group (power of two >= F/VEC, <= 32 lanes) per pair, 128-bit row loads, shuffle reduction; the backward
// accumulates with vector red.global.add (a vertex appears in many pairs).  HBM roofline: 4*(2*P*F + 3*P) bytes
// forward when no row is shared, 4*(4*P*F + 3*P) backward.
#include "common.cuh"

namespace stg {
namespace {

constexpr int kThreads = 256;

template <int VEC, int GROUP>
__global__ void __launch_bounds__(kThreads) edge_dot_fwd_kernel(const float* __restrict__ z, int feat,
                                                                const int64_t* __restrict__ a, const int64_t* __restrict__ b,
                                                                int64_t n_pairs, float* __restrict__ out) {
  using T = typename VecT<VEC>::type;
  const int gl = threadIdx.x & (GROUP - 1);
  const int64_t group = (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) / GROUP;
  const int64_t n_groups = static_cast<int64_t>(gridDim.x) * kThreads / GROUP;
  const int nvec = feat / VEC;
  for (int64_t p = group; p < n_pairs; p += n_groups) {
    const float* za = z + static_cast<size_t>(__ldg(a + p)) * feat;
    const float* zb = z + static_cast<size_t>(__ldg(b + p)) * feat;
    float acc = 0.f;
    for (int v = gl; v < nvec; v += GROUP) {
      const T x = __ldg(reinterpret_cast<const T*>(za) + v);
      const T y = __ldg(reinterpret_cast<const T*>(zb) + v);
      if constexpr (VEC == 4) acc += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
      else acc += x * y;
    }
#pragma unroll
    for (int o = GROUP / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (gl == 0) out[p] = acc;
  }
}

template <int VEC, int GROUP>
__global__ void __launch_bounds__(kThreads) edge_dot_bwd_kernel(const float* __restrict__ z, int feat,
                                                                const int64_t* __restrict__ a, const int64_t* __restrict__ b,
                                                                int64_t n_pairs, const float* __restrict__ g,
                                                                float* __restrict__ d_z) {
  using T = typename VecT<VEC>::type;
  const int gl = threadIdx.x & (GROUP - 1);
  const int64_t group = (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) / GROUP;
  const int64_t n_groups = static_cast<int64_t>(gridDim.x) * kThreads / GROUP;
  const int nvec = feat / VEC;
  for (int64_t p = group; p < n_pairs; p += n_groups) {
    const size_t ia = static_cast<size_t>(__ldg(a + p)) * feat, ib = static_cast<size_t>(__ldg(b + p)) * feat;
    const float gp = __ldg(g + p);
    for (int v = gl; v < nvec; v += GROUP) {
      T x = __ldg(reinterpret_cast<const T*>(z + ia) + v);
      T y = __ldg(reinterpret_cast<const T*>(z + ib) + v);
      if constexpr (VEC == 4) {
        atomicAdd(reinterpret_cast<float4*>(d_z + ia) + v, make_float4(gp * y.x, gp * y.y, gp * y.z, gp * y.w));
        atomicAdd(reinterpret_cast<float4*>(d_z + ib) + v, make_float4(gp * x.x, gp * x.y, gp * x.z, gp * x.w));
      } else {
        atomicAdd(d_z + ia + v, gp * y);
        atomicAdd(d_z + ib + v, gp * x);
      }
    }
  }
}

template <int VEC, int GROUP>
int launch_dot(bool bwd, const float* z, int feat, const int64_t* a, const int64_t* b, int64_t n_pairs, const float* g,
               float* out, cudaStream_t stream) {
  const int64_t need = (n_pairs * GROUP + kThreads - 1) / kThreads;
  const int blocks = static_cast<int>(need < 8LL * sm_count() ? (need > 0 ? need : 1) : 8LL * sm_count());
  if (bwd) edge_dot_bwd_kernel<VEC, GROUP><<<blocks, kThreads, 0, stream>>>(z, feat, a, b, n_pairs, g, out);
  else edge_dot_fwd_kernel<VEC, GROUP><<<blocks, kThreads, 0, stream>>>(z, feat, a, b, n_pairs, out);
  STG_LAUNCH_CHECK("edge_dot kernel");
  return STG_OK;
}

int dispatch_dot(bool bwd, const float* z, int feat, const int64_t* a, const int64_t* b, int64_t n_pairs, const float* g,
                 float* out, cudaStream_t stream) {
  const bool v4 = feat % 4 == 0 && aligned16(z) && (!bwd || aligned16(out));
  if (v4) {
    const int nvec = feat / 4;
    if (nvec <= 2) return launch_dot<4, 2>(bwd, z, feat, a, b, n_pairs, g, out, stream);
    if (nvec <= 4) return launch_dot<4, 4>(bwd, z, feat, a, b, n_pairs, g, out, stream);
    if (nvec <= 8) return launch_dot<4, 8>(bwd, z, feat, a, b, n_pairs, g, out, stream);
    if (nvec <= 16) return launch_dot<4, 16>(bwd, z, feat, a, b, n_pairs, g, out, stream);
    return launch_dot<4, 32>(bwd, z, feat, a, b, n_pairs, g, out, stream);
  }
  if (feat <= 4) return launch_dot<1, 4>(bwd, z, feat, a, b, n_pairs, g, out, stream);
  if (feat <= 16) return launch_dot<1, 16>(bwd, z, feat, a, b, n_pairs, g, out, stream);
  return launch_dot<1, 32>(bwd, z, feat, a, b, n_pairs, g, out, stream);
}

}  // namespace
}  // namespace stg

using namespace stg;

STG_API int stg_edge_dot_f32(const float* z, int32_t feat, const int64_t* a, const int64_t* b, int64_t n_pairs,
                             float* out, void* stream) {
  STG_CHECK_ARG(feat > 0 && n_pairs >= 0, "bad sizes (feat %d, pairs %lld)", feat, static_cast<long long>(n_pairs));
  if (n_pairs == 0) return STG_OK;
  STG_CHECK_ARG(z && a && b && out, "NULL argument");
  return dispatch_dot(false, z, feat, a, b, n_pairs, nullptr, out, as_stream(stream));
}

STG_API int stg_edge_dot_bwd_f32(const float* z, int32_t feat, const int64_t* a, const int64_t* b, int64_t n_pairs,
                                 const float* grad_out, float* d_z, void* stream) {
  STG_CHECK_ARG(feat > 0 && n_pairs >= 0, "bad sizes (feat %d, pairs %lld)", feat, static_cast<long long>(n_pairs));
  if (n_pairs == 0) return STG_OK;
  STG_CHECK_ARG(z && a && b && grad_out && d_z, "NULL argument");
  STG_CHECK_ARG(z != d_z, "z and d_z must not alias");
  return dispatch_dot(true, z, feat, a, b, n_pairs, grad_out, d_z, as_stream(stream));
}
