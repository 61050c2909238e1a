// Device helpers shared by the quadrature kernels (assemble_pk.cu, assemble_elasticity.cu): P1 / P2 Lagrange
// tabulation in barycentric form, affine simplex geometry, facet normals.
#pragma once
#include "common.cuh"

namespace phifem {
namespace pk {

constexpr int kBlockPk = 128;
constexpr int kMaxQuadPoints = 128;

// dolfinx / basix local edge order [dep-knowledge, SURVEY.md C.7]: edge e joins vertices ev(e, 0) < ev(e, 1)
template <int D>
__host__ __device__ constexpr int ev(int e, int s) {
  if (D == 2) {
    constexpr int t[3][2] = {{1, 2}, {0, 2}, {0, 1}};
    return t[e][s];
  } else {
    constexpr int t[6][2] = {{2, 3}, {1, 3}, {1, 2}, {0, 3}, {0, 2}, {0, 1}};
    return t[e][s];
  }
}

template <int D, int K>
struct Space {
  static constexpr int NV = D + 1;
  static constexpr int NE = K == 2 ? D * (D + 1) / 2 : 0;
  static constexpr int ND = NV + NE;
};

template <int D>
__device__ __forceinline__ double dotd(const double (&a)[D], const double (&b)[D]) {
  double s = a[0] * b[0];
#pragma unroll
  for (int d = 1; d < D; ++d) s += a[d] * b[d];
  return s;
}

// values and gradients of the P_K Lagrange basis at barycentric point lam; G = grad(lambda)
template <int D, int K>
__device__ __forceinline__ void tabulate(const double (&lam)[D + 1], const double (&G)[D + 1][D],
                                         double (&val)[Space<D, K>::ND], double (&grad)[Space<D, K>::ND][D]) {
  constexpr int NV = D + 1;
  if constexpr (K == 1) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      val[i] = lam[i];
#pragma unroll
      for (int d = 0; d < D; ++d) grad[i][d] = G[i][d];
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      val[i] = lam[i] * (2.0 * lam[i] - 1.0);
      const double c = 4.0 * lam[i] - 1.0;
#pragma unroll
      for (int d = 0; d < D; ++d) grad[i][d] = c * G[i][d];
    }
#pragma unroll
    for (int e = 0; e < Space<D, K>::NE; ++e) {
      const int a = ev<D>(e, 0), b = ev<D>(e, 1);
      val[NV + e] = 4.0 * lam[a] * lam[b];
#pragma unroll
      for (int d = 0; d < D; ++d) grad[NV + e][d] = 4.0 * (lam[a] * G[b][d] + lam[b] * G[a][d]);
    }
  }
}

// Laplacians of the basis (constant on the cell for K <= 2)
template <int D, int K>
__device__ __forceinline__ void laplacians(const double (&G)[D + 1][D], double (&lap)[Space<D, K>::ND]) {
  constexpr int NV = D + 1;
  if constexpr (K == 1) {
#pragma unroll
    for (int i = 0; i < NV; ++i) lap[i] = 0.0;
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) lap[i] = 4.0 * dotd<D>(G[i], G[i]);
#pragma unroll
    for (int e = 0; e < Space<D, K>::NE; ++e) lap[NV + e] = 8.0 * dotd<D>(G[ev<D>(e, 0)], G[ev<D>(e, 1)]);
  }
}

// affine geometry of one simplex: vertex coordinates, grad(lambda), |K|, h_T^2
template <int D>
struct Geometry {
  double X[D + 1][D], G[D + 1][D], vol, h2;
};

template <int D>
__device__ __forceinline__ void geometry_from_vertices(Geometry<D>& g);

template <int D>
__device__ __forceinline__ void load_geometry(const phifem_mesh& m, int64_t c, Geometry<D>& g) {
  constexpr int NV = D + 1;
  int v[NV];
  if (D == 3) {
    const int4 q = __ldg(reinterpret_cast<const int4*>(m.cells) + c);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[D] = q.w;
  } else {
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = __ldg(m.cells + c * NV + k);
  }
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int d = 0; d < D; ++d) g.X[k][d] = __ldg(m.x + (int64_t)v[k] * D + d);
  geometry_from_vertices<D>(g);
}

// grad(lambda), |K|, h_T^2 from the vertex coordinates g.X
template <int D>
__device__ __forceinline__ void geometry_from_vertices(Geometry<D>& g) {
  double e[D][D];
#pragma unroll
  for (int k = 0; k < D; ++k)
#pragma unroll
    for (int d = 0; d < D; ++d) e[k][d] = g.X[k + 1][d] - g.X[0][d];
  double det;
  if constexpr (D == 2) {
    det = e[0][0] * e[1][1] - e[1][0] * e[0][1];
    const double inv = 1.0 / det;
    g.G[1][0] = e[1][1] * inv;  g.G[1][1] = -e[1][0] * inv;
    g.G[2][0] = -e[0][1] * inv; g.G[2][1] = e[0][0] * inv;
    g.vol = 0.5 * fabs(det);
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
      g.G[1][d] = r1[d] * inv;
      g.G[2][d] = r2[d] * inv;
      g.G[3][d] = r3[d] * inv;
    }
    g.vol = fabs(det) * (1.0 / 6.0);
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
    double s = g.G[1][d];
#pragma unroll
    for (int k = 2; k <= D; ++k) s += g.G[k][d];
    g.G[0][d] = -s;
  }
  double h2 = 0.0;  // CellDiameter^2 = max squared vertex distance (main.py:100)
#pragma unroll
  for (int a = 0; a <= D; ++a)
#pragma unroll
    for (int b = a + 1; b <= D; ++b) {
      double s = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double t = g.X[a][d] - g.X[b][d];
        s += t * t;
      }
      h2 = fmax(h2, s);
    }
  g.h2 = h2;
}

// cell-local coefficients of a P_K function (dofmap NULL => vertex dofs = mesh.cells, K == 1 only)
template <int D, int K>
__device__ __forceinline__ void load_dofs(const phifem_mesh& m, const phifem_pk_space& sp,
                                          const double* __restrict__ coef, int64_t c,
                                          double (&out)[Space<D, K>::ND]) {
  constexpr int ND = Space<D, K>::ND;
  const int32_t* dm = sp.dofmap ? sp.dofmap + c * ND : m.cells + c * ND;
#pragma unroll
  for (int k = 0; k < ND; ++k) out[k] = __ldg(coef + __ldg(dm + k));
}

// phi_h, grad(phi_h) at a point from the tabulated basis (KP == KW shares the tables)
template <int D, int KW, int KP>
__device__ __forceinline__ void eval_phi(const double (&lam)[D + 1], const double (&G)[D + 1][D],
                                         const double (&wv)[Space<D, KW>::ND],
                                         const double (&wg)[Space<D, KW>::ND][D],
                                         const double (&pc)[Space<D, KP>::ND], double& ph, double (&gph)[D]) {
  constexpr int NDP = Space<D, KP>::ND;
  ph = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) gph[d] = 0.0;
  if constexpr (KP == KW) {
#pragma unroll
    for (int k = 0; k < NDP; ++k) {
      ph += pc[k] * wv[k];
#pragma unroll
      for (int d = 0; d < D; ++d) gph[d] += pc[k] * wg[k][d];
    }
  } else {
    double pv[NDP], pg[NDP][D];
    tabulate<D, KP>(lam, G, pv, pg);
#pragma unroll
    for (int k = 0; k < NDP; ++k) {
      ph += pc[k] * pv[k];
#pragma unroll
      for (int d = 0; d < D; ++d) gph[d] += pc[k] * pg[k][d];
    }
  }
}

// barycentric point of the cell from a point of local facet o (the facet's vertices in ascending local order)
template <int D>
__device__ __forceinline__ void facet_to_cell(const double* __restrict__ fl, int o, double (&lam)[D + 1]) {
#pragma unroll
  for (int k = 0; k <= D; ++k) {
    double v = 0.0;
    if (k < o) v = fl[k];
    if (k > o) v = fl[k - 1];
    lam[k] = v;
  }
}

template <int D>
__device__ __forceinline__ void facet_normal(const Geometry<D>& g, int o, double (&n)[D], double& area) {
  double Go[D];
#pragma unroll
  for (int d = 0; d < D; ++d) Go[d] = 0.0;
#pragma unroll
  for (int k = 0; k <= D; ++k)
    if (k == o)
#pragma unroll
      for (int d = 0; d < D; ++d) Go[d] = g.G[k][d];
  const double gnorm = sqrt(dotd<D>(Go, Go));
#pragma unroll
  for (int d = 0; d < D; ++d) n[d] = -Go[d] / gnorm;  // outward normal of THIS cell
  area = D * g.vol * gnorm;
}

template <int N>
__device__ __forceinline__ double pick(const double (&a)[N], int i) {
  double v = a[0];
#pragma unroll
  for (int k = 1; k < N; ++k)
    if (k == i) v = a[k];
  return v;
}

// phi_h and grad(phi_h) at a point, from the level set's own tabulation
template <int D, int KP>
__device__ __forceinline__ void eval_phi_only(const double (&lam)[D + 1], const double (&G)[D + 1][D],
                                              const double (&pc)[Space<D, KP>::ND], double& ph, double (&gph)[D]) {
  constexpr int NDP = Space<D, KP>::ND;
  double pv[NDP], pg[NDP][D];
  tabulate<D, KP>(lam, G, pv, pg);
  ph = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) gph[d] = 0.0;
#pragma unroll
  for (int k = 0; k < NDP; ++k) {
    ph += pc[k] * pv[k];
#pragma unroll
    for (int d = 0; d < D; ++d) gph[d] += pc[k] * pg[k][d];
  }
}

}  // namespace pk
}  // namespace phifem
