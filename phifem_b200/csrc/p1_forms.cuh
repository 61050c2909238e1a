// Closed-form building blocks of the P1 forms on simplices (shared by the row-gather and the cell-once kernels):
// cofactors / gradients of the barycentric coordinates, CellDiameter (reference
// demo/strong-dirichlet/flower/main.py:100), SURVEY.md Appendix B.
#pragma once
#include "common.cuh"

namespace phifem {

template <int D>
__device__ __forceinline__ double dot(const double (&a)[D], const double (&b)[D]) {
  double s = a[0] * b[0];
#pragma unroll
  for (int d = 1; d < D; ++d) s += a[d] * b[d];
  return s;
}

// gradients of the barycentric coordinates of the simplex X[0..D], det of the edge matrix
template <int D>
__device__ __forceinline__ void simplex_gradients(const double (&X)[D + 1][D], double (&G)[D + 1][D],
                                                  double& det) {
  double e[D][D];
#pragma unroll
  for (int k = 0; k < D; ++k)
#pragma unroll
    for (int d = 0; d < D; ++d) e[k][d] = X[k + 1][d] - X[0][d];
  if constexpr (D == 2) {
    det = e[0][0] * e[1][1] - e[1][0] * e[0][1];
    const double inv = 1.0 / det;
    G[1][0] = e[1][1] * inv;  G[1][1] = -e[1][0] * inv;
    G[2][0] = -e[0][1] * inv; G[2][1] = e[0][0] * inv;
  } else {
    const double r1[3] = {e[1][1] * e[2][2] - e[1][2] * e[2][1], e[1][2] * e[2][0] - e[1][0] * e[2][2],
                          e[1][0] * e[2][1] - e[1][1] * e[2][0]};
    const double r2[3] = {e[2][1] * e[0][2] - e[2][2] * e[0][1], e[2][2] * e[0][0] - e[2][0] * e[0][2],
                          e[2][0] * e[0][1] - e[2][1] * e[0][0]};
    const double r3[3] = {e[0][1] * e[1][2] - e[0][2] * e[1][1], e[0][2] * e[1][0] - e[0][0] * e[1][2],
                          e[0][0] * e[1][1] - e[0][1] * e[1][0]};
    det = e[0][0] * r1[0] + e[0][1] * r1[1] + e[0][2] * r1[2];
    const double inv = 1.0 / det;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      G[1][d] = r1[d] * inv;
      G[2][d] = r2[d] * inv;
      G[3][d] = r3[d] * inv;
    }
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
    double s = G[1][d];
#pragma unroll
    for (int k = 2; k <= D; ++k) s += G[k][d];
    G[0][d] = -s;
  }
}

template <int D>
__device__ __forceinline__ double diameter2(const double (&X)[D + 1][D]) {
  double h2 = 0.0;  // CellDiameter^2 = max squared vertex distance (main.py:100)
#pragma unroll
  for (int a = 0; a <= D; ++a)
#pragma unroll
    for (int b = a + 1; b <= D; ++b) {
      double s = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double t = X[a][d] - X[b][d];
        s += t * t;
      }
      h2 = fmax(h2, s);
    }
  return h2;
}

template <int D> constexpr double volume_factor() { return D == 2 ? 0.5 : 1.0 / 6.0; }

// Unnormalised gradients: R[k] = det * grad(lambda_k) (cofactors of the edge matrix), edges from X[0].
template <int D>
__device__ __forceinline__ void simplex_cofactors(const double (&X)[D + 1][D], double (&R)[D + 1][D],
                                                  double& det) {
  double e[D][D];
#pragma unroll
  for (int k = 0; k < D; ++k)
#pragma unroll
    for (int d = 0; d < D; ++d) e[k][d] = X[k + 1][d] - X[0][d];
  if constexpr (D == 2) {
    R[1][0] = e[1][1];  R[1][1] = -e[1][0];
    R[2][0] = -e[0][1]; R[2][1] = e[0][0];
    det = e[0][0] * e[1][1] - e[1][0] * e[0][1];
  } else {
    R[1][0] = e[1][1] * e[2][2] - e[1][2] * e[2][1];
    R[1][1] = e[1][2] * e[2][0] - e[1][0] * e[2][2];
    R[1][2] = e[1][0] * e[2][1] - e[1][1] * e[2][0];
    R[2][0] = e[2][1] * e[0][2] - e[2][2] * e[0][1];
    R[2][1] = e[2][2] * e[0][0] - e[2][0] * e[0][2];
    R[2][2] = e[2][0] * e[0][1] - e[2][1] * e[0][0];
    R[3][0] = e[0][1] * e[1][2] - e[0][2] * e[1][1];
    R[3][1] = e[0][2] * e[1][0] - e[0][0] * e[1][2];
    R[3][2] = e[0][0] * e[1][1] - e[0][1] * e[1][0];
    det = e[0][0] * R[1][0] + e[0][1] * R[1][1] + e[0][2] * R[1][2];
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
    double s = R[1][d];
#pragma unroll
    for (int k = 2; k <= D; ++k) s += R[k][d];
    R[0][d] = -s;
  }
}

}  // namespace phifem
