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

namespace phifem {
namespace {
// what an assembly plan sees of a facet tag: 1 = ghost-penalty facet (tags 2, 3), 2 = Gamma_h (tag 4), 0 otherwise
__device__ __forceinline__ unsigned facet_class(unsigned t) { return (t == 2u || t == 3u) ? 1u : (t == 4u ? 2u : 0u); }

__global__ void __launch_bounds__(256) k_tags_match(const uint8_t* __restrict__ ct, const uint8_t* __restrict__ ct_sig,
                                                    int64_t n_cells, const uint8_t* __restrict__ ft,
                                                    const uint8_t* __restrict__ ft_sig, int64_t n_facets,
                                                    unsigned long long* __restrict__ mismatches) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned bad = 0;
  // 16 bytes per thread and step where the arrays allow it (torch allocations are 256-byte aligned), bytes otherwise
  const bool vec = ((reinterpret_cast<uintptr_t>(ct) | reinterpret_cast<uintptr_t>(ct_sig) |
                     reinterpret_cast<uintptr_t>(ft) | reinterpret_cast<uintptr_t>(ft_sig)) & 15u) == 0;
  const int64_t nc16 = vec ? n_cells / 16 : 0, nf16 = vec ? n_facets / 16 : 0;
  for (int64_t i = t0; i < nc16; i += stride) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(ct) + i), b = __ldg(reinterpret_cast<const uint4*>(ct_sig) + i);
    bad |= (a.x ^ b.x) | (a.y ^ b.y) | (a.z ^ b.z) | (a.w ^ b.w);
  }
  for (int64_t i = nc16 * 16 + t0; i < n_cells; i += stride) bad |= ct[i] ^ ct_sig[i];
  for (int64_t i = t0; i < nf16; i += stride) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(ft) + i), b = __ldg(reinterpret_cast<const uint4*>(ft_sig) + i);
    const unsigned wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int k = 0; k < 4; ++k) bad |= facet_class((wa[q] >> (8 * k)) & 0xffu) ^ ((wb[q] >> (8 * k)) & 0xffu);
  }
  for (int64_t i = nf16 * 16 + t0; i < n_facets; i += stride) bad |= facet_class(ft[i]) ^ ft_sig[i];
  if (__any_sync(0xffffffffu, bad != 0) && (threadIdx.x & 31) == 0) atomicAdd(mismatches, 1ull);
}
}  // namespace
}  // namespace phifem

// Do these tags still belong to the plan that kept `cell_signature` (a copy of the one-byte cell tags it was built
// from) and `facet_signature` (the facet classes it depends on: 1 = tag 2 or 3, 2 = tag 4, 0 otherwise)?  One pass over
// the four arrays; *mismatches (device, zeroed here on the stream) ends up non-zero iff a cell tag or a facet class
// differs.  The question a moving-interface loop asks after every classification before it reuses its plan (the
// reference rebuilds forms and matrices at every step, demo/strong-dirichlet/flower/main.py:59-66,121-123).
extern "C" int phifem_tags_match(const int8_t* cell_tags8, const int8_t* cell_signature, int64_t n_cells,
                                 const int8_t* facet_tags8, const int8_t* facet_signature, int64_t n_facets,
                                 int64_t* mismatches, void* stream) {
  PHIFEM_CHECK_ARG(mismatches != nullptr, "null counter");
  PHIFEM_CHECK_ARG(n_cells >= 0 && n_facets >= 0, "negative size");
  PHIFEM_CHECK_ARG((n_cells == 0 || (cell_tags8 && cell_signature)) && (n_facets == 0 || (facet_tags8 && facet_signature)),
                   "null tag array");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(mismatches, 0, sizeof(int64_t), st) != cudaSuccess) {
    phifem::set_error("%s: CUDA error: %s", __func__, cudaGetErrorString(cudaGetLastError()));
    return PHIFEM_ERR_CUDA;
  }
  const int64_t work = (n_cells + n_facets) / 16 + 1;
  const int grid = phifem::grid_for(work, 256, 8);
  phifem::k_tags_match<<<grid, 256, 0, st>>>(
      reinterpret_cast<const uint8_t*>(cell_tags8), reinterpret_cast<const uint8_t*>(cell_signature), n_cells,
      reinterpret_cast<const uint8_t*>(facet_tags8), reinterpret_cast<const uint8_t*>(facet_signature), n_facets,
      reinterpret_cast<unsigned long long*>(mismatches));
  PHIFEM_CHECK_LAUNCH();
  return PHIFEM_OK;
}
