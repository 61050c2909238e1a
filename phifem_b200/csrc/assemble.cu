// K2/K3/K4: strong-Dirichlet phi-FEM operator, P1 on triangles / tetrahedra, assembled into CSR.
//
// Forms of reference demo/strong-dirichlet/flower/main.py:104-128; closed-form element tensors of
// SURVEY.md Appendix B (exact for P1 phi, P1 w/v, P1 f), so the kernels stay near the fp64/HBM
// ridge instead of looping over quadrature points.  What dolfinx does per entity with one FFCx
// `tabulate_tensor` call plus a `MatSetValuesLocal(ADD_VALUES)` binary search is one thread here:
// gather geometry + phi + f, evaluate the tensor in registers, scatter through a precomputed
// cell -> CSR-slot map with fp64 reductions (REDG.E.ADD.F64) that resolve in L2.
#include "common.cuh"

namespace phifem {
namespace {

constexpr int kBlock = 128;

template <int D>
struct Simplex {
  double G[D + 1][D];  // grad(lambda_i)
  double vol;          // |K|
  double p[D + 1];     // phi at the vertices
  int v[D + 1];
};

template <int D>
__device__ __forceinline__ void load_vertices(const phifem_mesh& m, int64_t c, int (&v)[D + 1],
                                              double (&xc)[D + 1][D]) {
  if (D == 3) {
    const int4 q = __ldg(reinterpret_cast<const int4*>(m.cells) + c);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[D] = q.w;
  } else {
#pragma unroll
    for (int k = 0; k <= D; ++k) v[k] = __ldg(m.cells + c * (D + 1) + k);
  }
#pragma unroll
  for (int k = 0; k <= D; ++k)
#pragma unroll
    for (int d = 0; d < D; ++d) xc[k][d] = __ldg(m.x + (int64_t)v[k] * D + d);
}

template <int D>
__device__ __forceinline__ void gradients(const double (&xc)[D + 1][D], double (&G)[D + 1][D],
                                          double& vol) {
  if (D == 2) {
    const double a0 = xc[1][0] - xc[0][0], a1 = xc[1][1] - xc[0][1];
    const double b0 = xc[2][0] - xc[0][0], b1 = xc[2][1] - xc[0][1];
    const double det = a0 * b1 - b0 * a1;
    const double inv = 1.0 / det;
    G[1][0] = b1 * inv;  G[1][1] = -b0 * inv;
    G[2][0] = -a1 * inv; G[2][1] = a0 * inv;
    vol = 0.5 * fabs(det);
  } else {
    double a[3], b[3], c[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      a[d] = xc[1][d] - xc[0][d];
      b[d] = xc[2][d] - xc[0][d];
      c[d] = xc[D][d] - xc[0][d];
    }
    double r1[3] = {b[1] * c[2] - b[2] * c[1], b[2] * c[0] - b[0] * c[2], b[0] * c[1] - b[1] * c[0]};
    double r2[3] = {c[1] * a[2] - c[2] * a[1], c[2] * a[0] - c[0] * a[2], c[0] * a[1] - c[1] * a[0]};
    double r3[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    const double det = a[0] * r1[0] + a[1] * r1[1] + a[2] * r1[2];
    const double inv = 1.0 / det;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      G[1][d] = r1[d] * inv;
      G[2][d] = r2[d] * inv;
      G[D][d] = r3[d] * inv;
    }
    vol = fabs(det) * (1.0 / 6.0);
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
    double s = 0.0;
#pragma unroll
    for (int k = 1; k <= D; ++k) s += G[k][d];
    G[0][d] = -s;
  }
}

template <int D>
__device__ __forceinline__ double diameter2(const double (&xc)[D + 1][D]) {
  double h2 = 0.0;  // CellDiameter^2 = max squared vertex distance (main.py:100)
#pragma unroll
  for (int a = 0; a <= D; ++a)
#pragma unroll
    for (int b = a + 1; b <= D; ++b) {
      double s = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double e = xc[a][d] - xc[b][d];
        s += e * e;
      }
      h2 = fmax(h2, s);
    }
  return h2;
}

template <int D>
__device__ __forceinline__ double dot(const double (&a)[D], const double (&b)[D]) {
  double s = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) s += a[d] * b[d];
  return s;
}

// ---- K2 + K4: cells of dx((1,2)) and dx(2) -----------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kBlock) k_assemble_cells_p1(
    phifem_mesh m, const double* __restrict__ phi, const double* __restrict__ f,
    const int8_t* __restrict__ ctags, const int32_t* __restrict__ active, int64_t n_active,
    const int32_t* __restrict__ slots, double sigma, double* __restrict__ data, double* __restrict__ b) {
  constexpr int NV = D + 1;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_active) return;
  const int64_t c = __ldg(active + e);
  int v[NV];
  double xc[NV][D];
  load_vertices<D>(m, c, v, xc);
  double p[NV], fv[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    p[k] = __ldg(phi + v[k]);
    fv[k] = __ldg(f + v[k]);
  }
  // slot loads issued early: they do not depend on the arithmetic below
  int sl[NV * NV];
  if (NV * NV % 4 == 0) {
#pragma unroll
    for (int q = 0; q < NV * NV / 4; ++q) {
      const int4 s4 = __ldg(reinterpret_cast<const int4*>(slots + e * NV * NV) + q);
      sl[4 * q] = s4.x; sl[4 * q + 1] = s4.y; sl[4 * q + 2] = s4.z; sl[4 * q + 3] = s4.w;
    }
  } else {
#pragma unroll
    for (int q = 0; q < NV * NV; ++q) sl[q] = __ldg(slots + e * NV * NV + q);
  }
  double G[NV][D], vol;
  gradients<D>(xc, G, vol);
  double g[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NV; ++k) s += p[k] * G[k][d];
    g[d] = s;
  }
  const double gg = dot<D>(g, g);
  double a[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) a[i] = dot<D>(g, G[i]);
  const double cM = vol * (1.0 / ((D + 1) * (D + 2)));
  double P = 0.0, F = 0.0, FP = 0.0;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    P += p[k];
    F += fv[k];
    FP += fv[k] * p[k];
  }
  double mm[NV], mu = 0.0;  // m_i = int lambda_i phi, mu = int phi^2
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    mm[i] = cM * (P + p[i]);
    mu += p[i] * mm[i];
  }
  double stab = 0.0;
  if (ctags[c] == 2) stab = sigma * diameter2<D>(xc) * vol;  // sigma h_T^2 |K| on cut cells
  const double ggM = gg * cM, stab4 = 4.0 * stab;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int j = i; j < NV; ++j) {
      const double val = ggM * (i == j ? 2.0 : 1.0) + a[i] * mm[j] + mm[i] * a[j] +
                         dot<D>(G[i], G[j]) * mu + stab4 * a[i] * a[j];
      atomicAdd(data + sl[i * NV + j], val);
      if (j != i) atomicAdd(data + sl[j * NV + i], val);
    }
  }
  constexpr double fact_d = D == 2 ? 2.0 : 6.0, fact_d3 = D == 2 ? 120.0 : 720.0;
  const double c3 = vol * (fact_d / fact_d3);
  const double fmean2 = 2.0 * stab * F * (1.0 / NV);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const double bi = c3 * ((F * P + FP) + fv[i] * P + F * p[i] + 2.0 * fv[i] * p[i]) - fmean2 * a[i];
    atomicAdd(b + v[i], bi);
  }
}

__device__ __forceinline__ double alpha3(int a, int b, int c) {
  return (double)((1 + (a == b)) * (1 + (a == c) + (b == c)));
}

// ---- one-sided boundary term over ds(100) entities ----------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kBlock) k_assemble_boundary_p1(
    phifem_mesh m, const double* __restrict__ phi, const int32_t* __restrict__ entities,
    int64_t n_entities, const int32_t* __restrict__ slots, double* __restrict__ data) {
  constexpr int NV = D + 1;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_entities) return;
  const int64_t c = __ldg(entities + 2 * e);
  const int o = __ldg(entities + 2 * e + 1);
  int v[NV];
  double xc[NV][D];
  load_vertices<D>(m, c, v, xc);
  double p[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) p[k] = __ldg(phi + v[k]);
  double G[NV][D], vol;
  gradients<D>(xc, G, vol);
  double Go[D];
#pragma unroll
  for (int k = 0; k < NV; ++k)
    if (k == o)
#pragma unroll
      for (int d = 0; d < D; ++d) Go[d] = G[k][d];
  const double gnorm = sqrt(dot<D>(Go, Go));
  double n[D], g[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    n[d] = -Go[d] / gnorm;  // outward normal of THIS cell (tests/test_one_sided_integral.py pins it)
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NV; ++k) s += p[k] * G[k][d];
    g[d] = s;
  }
  const double area = D * vol * gnorm;
  const double gn = dot<D>(g, n);
  constexpr double cfac = D == 2 ? 1.0 / 24.0 : 2.0 / 120.0;  // (d-1)!/(d+2)!
  const double cF = area * cfac;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      double acc = 0.0;
      const double Gnj = dot<D>(G[j], n);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        if (k == o) continue;
        double s = 0.0;
#pragma unroll
        for (int l = 0; l < NV; ++l)
          if (l != o) s += p[l] * alpha3(l, k, i);
        const double tjki = (j != o) ? alpha3(j, k, i) : 0.0;
        acc += p[k] * (gn * tjki + Gnj * s);
      }
      const double val = (i == o) ? 0.0 : -cF * acc;
      atomicAdd(data + __ldg(slots + e * NV * NV + i * NV + j), val);
    }
  }
}

// ---- K3: ghost penalty over interior facets tagged 2 / 3 ------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kBlock) k_assemble_ghost_p1(
    phifem_mesh m, const double* __restrict__ phi, const int32_t* __restrict__ facets, int64_t n_facets,
    const int32_t* __restrict__ slots, double sigma, double* __restrict__ data) {
  constexpr int NV = D + 1, NM = 2 * NV;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_facets) return;
  const int32_t fct = __ldg(facets + e);
  const int2 cc = __ldg(reinterpret_cast<const int2*>(m.f2c) + fct);
  int fvert[D];            // global vertices of the facet, ordered as in cell +
  double Jv[NM][D];        // jump integrand of macro dof a at facet vertex k (affine on the facet)
  double hsum = 0.0, area = 0.0;
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    const int64_t c = side == 0 ? cc.x : cc.y;
    int v[NV];
    double xc[NV][D];
    load_vertices<D>(m, c, v, xc);
    int o = 0;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (__ldg(m.c2f + c * NV + i) == fct) o = i;
    double p[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) p[k] = __ldg(phi + v[k]);
    double G[NV][D], vol;
    gradients<D>(xc, G, vol);
    double Go[D];
#pragma unroll
    for (int k = 0; k < NV; ++k)
      if (k == o)
#pragma unroll
        for (int d = 0; d < D; ++d) Go[d] = G[k][d];
    const double gnorm = sqrt(dot<D>(Go, Go));
    double n[D], g[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      n[d] = -Go[d] / gnorm;
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < NV; ++k) s += p[k] * G[k][d];
      g[d] = s;
    }
    const double gn = dot<D>(g, n);
    hsum += sqrt(diameter2<D>(xc));
    if (side == 0) {
      area = D * vol * gnorm;
      int q = 0;
#pragma unroll
      for (int k = 0; k < NV; ++k)
        if (k != o) {
#pragma unroll
          for (int t = 0; t < D; ++t)
            if (t == q) fvert[t] = v[k];
          ++q;
        }
    }
#pragma unroll
    for (int a = 0; a < NV; ++a) {
      const double Gna = dot<D>(G[a], n);
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const double pk = __ldg(phi + fvert[k]);
        Jv[side * NV + a][k] = (v[a] == fvert[k] ? gn : 0.0) + Gna * pk;
      }
    }
  }
  const double coef = sigma * 0.5 * hsum * area * (1.0 / (D * (D + 1)));
  double Js[NM];
#pragma unroll
  for (int a = 0; a < NM; ++a) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) s += Jv[a][k];
    Js[a] = s;
  }
#pragma unroll
  for (int a = 0; a < NM; ++a) {
#pragma unroll
    for (int bb = a; bb < NM; ++bb) {
      double s = Js[a] * Js[bb];
#pragma unroll
      for (int k = 0; k < D; ++k) s += Jv[a][k] * Jv[bb][k];
      const double val = coef * s;
      atomicAdd(data + __ldg(slots + e * NM * NM + a * NM + bb), val);
      if (bb != a) atomicAdd(data + __ldg(slots + e * NM * NM + bb * NM + a), val);
    }
  }
}

int check_simplex_mesh(const phifem_mesh* m) {
  PHIFEM_CHECK_ARG(m != nullptr && m->x && m->cells, "mesh is null");
  if (m->cell_type != PHIFEM_TRIANGLE && m->cell_type != PHIFEM_TETRAHEDRON) {
    set_error("P1 assembly supports triangles and tetrahedra, got cell type %d", m->cell_type);
    return PHIFEM_ERR_UNSUPPORTED;
  }
  PHIFEM_CHECK_ARG(m->gdim == (m->cell_type == PHIFEM_TRIANGLE ? 2 : 3), "gdim mismatch");
  return PHIFEM_OK;
}

}  // namespace
}  // namespace phifem

using namespace phifem;

extern "C" int phifem_assemble_cells_p1(const phifem_mesh* mesh, const double* phi, const double* f,
                                        const int8_t* cell_tags8, const int32_t* active,
                                        int64_t n_active, const int32_t* slots, double sigma,
                                        double* data, double* b, void* stream) {
  if (int rc = check_simplex_mesh(mesh)) return rc;
  PHIFEM_CHECK_ARG(phi && f && cell_tags8 && data && b, "null pointer");
  PHIFEM_CHECK_ARG(n_active == 0 || (active && slots), "null active / slots");
  if (n_active == 0) return PHIFEM_OK;
  const int grid = (int)((n_active + kBlock - 1) / kBlock);
  cudaStream_t st = (cudaStream_t)stream;
  if (mesh->cell_type == PHIFEM_TRIANGLE)
    k_assemble_cells_p1<2><<<grid, kBlock, 0, st>>>(*mesh, phi, f, cell_tags8, active, n_active, slots,
                                                    sigma, data, b);
  else
    k_assemble_cells_p1<3><<<grid, kBlock, 0, st>>>(*mesh, phi, f, cell_tags8, active, n_active, slots,
                                                    sigma, data, b);
  PHIFEM_CHECK_LAUNCH();
  return PHIFEM_OK;
}

extern "C" int phifem_assemble_boundary_p1(const phifem_mesh* mesh, const double* phi,
                                           const int32_t* entities, int64_t n_entities,
                                           const int32_t* slots, double* data, void* stream) {
  if (int rc = check_simplex_mesh(mesh)) return rc;
  PHIFEM_CHECK_ARG(phi && data, "null pointer");
  PHIFEM_CHECK_ARG(n_entities == 0 || (entities && slots), "null entities / slots");
  if (n_entities == 0) return PHIFEM_OK;
  const int grid = (int)((n_entities + kBlock - 1) / kBlock);
  cudaStream_t st = (cudaStream_t)stream;
  if (mesh->cell_type == PHIFEM_TRIANGLE)
    k_assemble_boundary_p1<2><<<grid, kBlock, 0, st>>>(*mesh, phi, entities, n_entities, slots, data);
  else
    k_assemble_boundary_p1<3><<<grid, kBlock, 0, st>>>(*mesh, phi, entities, n_entities, slots, data);
  PHIFEM_CHECK_LAUNCH();
  return PHIFEM_OK;
}

extern "C" int phifem_assemble_ghost_p1(const phifem_mesh* mesh, const double* phi,
                                        const int32_t* facets, int64_t n_facets, const int32_t* slots,
                                        double sigma, double* data, void* stream) {
  if (int rc = check_simplex_mesh(mesh)) return rc;
  PHIFEM_CHECK_ARG(phi && data && mesh->c2f && mesh->f2c, "null pointer");
  PHIFEM_CHECK_ARG(n_facets == 0 || (facets && slots), "null facets / slots");
  if (n_facets == 0) return PHIFEM_OK;
  const int grid = (int)((n_facets + kBlock - 1) / kBlock);
  cudaStream_t st = (cudaStream_t)stream;
  if (mesh->cell_type == PHIFEM_TRIANGLE)
    k_assemble_ghost_p1<2><<<grid, kBlock, 0, st>>>(*mesh, phi, facets, n_facets, slots, sigma, data);
  else
    k_assemble_ghost_p1<3><<<grid, kBlock, 0, st>>>(*mesh, phi, facets, n_facets, slots, sigma, data);
  PHIFEM_CHECK_LAUNCH();
  return PHIFEM_OK;
}
