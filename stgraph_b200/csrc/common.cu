// Error reporting + device queries shared by every entry point.
#include "common.cuh"

#include <mutex>

namespace stg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace stg

STG_API int stg_abi_version(void) { return STG_ABI_VERSION; }

STG_API const char* stg_last_error(void) { return stg::g_err; }

STG_API int stg_device_info(int device, int32_t* sm_count, int64_t* l2_bytes, int32_t* cc_major,
                               int32_t* cc_minor) {
  int v = 0;
  if (sm_count) {
    STG_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
    *sm_count = v;
  }
  if (l2_bytes) {
    STG_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, device));
    *l2_bytes = v;
  }
  if (cc_major) {
    STG_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, device));
    *cc_major = v;
  }
  if (cc_minor) {
    STG_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, device));
    *cc_minor = v;
  }
  return STG_OK;
}

STG_API int stg_get_array_i32(const int32_t* dev_ptr, int64_t count, int32_t* host_out, void* stream) {
  STG_CHECK_ARG(count >= 0, "negative count");
  if (count == 0) return STG_OK;
  STG_CHECK_ARG(dev_ptr && host_out, "NULL pointer");
  cudaStream_t s = stg::as_stream(stream);
  STG_CUDA(cudaMemcpyAsync(host_out, dev_ptr, count * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  STG_CUDA(cudaStreamSynchronize(s));
  return STG_OK;
}
