// C-ABI plumbing: thread-local error message, ABI version.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace phifem {
namespace {
thread_local char g_error[512] = "";
}
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
}  // namespace phifem

extern "C" const char* phifem_last_error(void) { return phifem::g_error; }
extern "C" int phifem_abi_version(void) { return PHIFEM_B200_ABI_VERSION; }

namespace phifem {
namespace {
__global__ void k_post_words(const int64_t* __restrict__ src, volatile int64_t* dst, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
}
}  // namespace
}  // namespace phifem

// A few words of device memory -> page-locked host memory by a KERNEL (stores from an SM over PCIe), not by a copy
// engine: the counter block of a classification (128 bytes) reaches the host without queueing behind a download of
// hundreds of MB that another stream has put on the device -> host engine (an end-to-end pipeline with two steps in
// flight stalls exactly there).  The words are visible to the host once an event recorded on `stream` after this call
// has completed.
extern "C" int phifem_post_to_host(const int64_t* device_words, int64_t* pinned_host_words, int32_t n_words,
                                   void* stream) {
  PHIFEM_CHECK_ARG(device_words && pinned_host_words, "null pointer");
  PHIFEM_CHECK_ARG(n_words > 0 && n_words <= 4096, "between 1 and 4096 words");
  void* mapped = nullptr;
  cudaError_t err = cudaHostGetDevicePointer(&mapped, pinned_host_words, 0);
  if (err != cudaSuccess) {
    cudaGetLastError();
    phifem::set_error("%s: the destination is not page-locked, device-mapped host memory (%s)", __func__,
                      cudaGetErrorString(err));
    return PHIFEM_ERR_ARGUMENT;
  }
  phifem::k_post_words<<<1, 128, 0, (cudaStream_t)stream>>>(device_words, (volatile int64_t*)mapped, n_words);
  PHIFEM_CHECK_LAUNCH();
  return PHIFEM_OK;
}
