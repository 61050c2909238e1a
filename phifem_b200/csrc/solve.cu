// CSR sparse matrix-vector product for the Krylov solve that follows the assembly.
//
// The reference hands the assembled system to PETSc KSP "preonly" + MUMPS LU with null-pivot detection
// (demo/strong-dirichlet/flower/main.py:138-157, demo/weak-dirichlet/flower/main.py:161-184).  On the GPU the
// system stays where the assembly left it and is solved by Jacobi-preconditioned BiCGStab
// (phifem_b200/solve.py); this file provides its one bandwidth-bound building block.  phi-FEM rows are short
// (15 entries for P1 tetrahedra, 7 for P1 triangles), so a row is reduced by a group of 8 lanes: 4 rows per
// warp, coalesced 64-byte segments of `data` / `indices`, x gathered through the read-only path.
#include "common.cuh"

namespace phifem {
namespace {

constexpr int kSpmvBlock = 256;
constexpr int kLanesPerRow = 8;

__global__ void __launch_bounds__(kSpmvBlock) k_csr_spmv(int64_t n_rows, const int32_t* __restrict__ indptr,
                                                         const int32_t* __restrict__ indices,
                                                         const double* __restrict__ data,
                                                         const double* __restrict__ x, double* __restrict__ y) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = t / kLanesPerRow;
  const int sub = (int)(t % kLanesPerRow);
  double acc = 0.0;
  if (row < n_rows) {
    const int lo = __ldg(indptr + row), hi = __ldg(indptr + row + 1);
    for (int k = lo + sub; k < hi; k += kLanesPerRow) acc += __ldg(data + k) * __ldg(x + __ldg(indices + k));
  }
#pragma unroll
  for (int off = kLanesPerRow / 2; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off, kLanesPerRow);
  if (row < n_rows && sub == 0) y[row] = acc;
}

}  // namespace
}  // namespace phifem

using namespace phifem;

extern "C" int phifem_csr_spmv(int64_t n_rows, const int32_t* indptr, const int32_t* indices,
                               const double* data, const double* x, double* y, void* stream) {
  PHIFEM_CHECK_ARG(n_rows >= 0 && indptr && x && y, "null pointer");
  if (n_rows == 0) return PHIFEM_OK;
  PHIFEM_CHECK_ARG(indices && data, "null CSR arrays");
  const int64_t threads = n_rows * kLanesPerRow;
  k_csr_spmv<<<(unsigned)((threads + kSpmvBlock - 1) / kSpmvBlock), kSpmvBlock, 0, (cudaStream_t)stream>>>(
      n_rows, indptr, indices, data, x, y);
  PHIFEM_CHECK_LAUNCH();
  return PHIFEM_OK;
}
