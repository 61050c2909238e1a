// A host without Python: cut-cell tags, CSR pattern and assembly of the strong-Dirichlet phi-FEM operator through the C
// ABI of libphifem_b200.so alone (include/phifem_b200.h) -- twice: with the per-entity kernels over slot maps
// (phifem_pattern_create_p1 + phifem_assemble_{cells,boundary,ghost}_p1) and with the benchmarked row-gather path
// (phifem_rows_plan_create + phifem_assemble_rows_p1).  Stand-in for what a C++ / PETSc code base would do with the
// arrays of its own mesh; here the mesh is a unit square of 2 n^2 triangles built on the host.
//
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a examples/capi_host.cu -Iinclude \
//        -Lphifem_b200 -lphifem_b200 -Xlinker -rpath=$PWD/phifem_b200 -o capi_host && ./capi_host 48
//
// Reference calls replaced: compute_tags_measures (src/phifem/mesh_scripts.py:571-653) and assemble_matrix /
// assemble_vector of demo/strong-dirichlet/flower/main.py:121-131, the one-sided ds(100) term included.
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>

#include "phifem_b200.h"

#define CHECK(call)                                                              \
  do {                                                                           \
    int rc__ = (call);                                                           \
    if (rc__ != 0) {                                                             \
      std::fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, phifem_last_error()); \
      return 1;                                                                  \
    }                                                                            \
  } while (0)

template <typename T>
T* to_device(const std::vector<T>& v) {
  T* d = nullptr;
  cudaMalloc(&d, std::max<size_t>(1, v.size()) * sizeof(T));
  cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  return d;
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? std::atoi(argv[1]) : 32;
  // mesh: vertex (i, j) -> i (n + 1) + j at (i / n, j / n); every square split along its "right" diagonal
  const int nvx = (n + 1) * (n + 1), nc = 2 * n * n;
  std::vector<double> x(2 * (size_t)nvx), phi(nvx), f(nvx);
  for (int i = 0; i <= n; ++i)
    for (int j = 0; j <= n; ++j) {
      const int v = i * (n + 1) + j;
      x[2 * v] = (double)i / n;
      x[2 * v + 1] = (double)j / n;
    }
  for (int v = 0; v < nvx; ++v) {
    const double dx = x[2 * v] - 0.503, dy = x[2 * v + 1] - 0.497;
    phi[v] = dx * dx + dy * dy - 0.3 * 0.3;
    f[v] = 1.0 + x[2 * v] + 2.0 * x[2 * v + 1];
  }
  std::vector<int32_t> cells(3 * (size_t)nc);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      const int v00 = i * (n + 1) + j, v10 = v00 + n + 1, v01 = v00 + 1, v11 = v10 + 1, c = 2 * (i * n + j);
      const int t[6] = {v00, v10, v11, v00, v01, v11};
      std::copy(t, t + 6, cells.begin() + 3 * (size_t)c);
    }
  // facets = lexicographic rank of the sorted vertex pair (dolfinx serial numbering); local facet i opposite vertex i
  std::map<std::array<int, 2>, int> facet_of;
  for (int c = 0; c < nc; ++c)
    for (int i = 0; i < 3; ++i) {
      std::array<int, 2> e = {cells[3 * c + (i + 1) % 3], cells[3 * c + (i + 2) % 3]};
      if (e[0] > e[1]) std::swap(e[0], e[1]);
      facet_of.emplace(e, 0);
    }
  int nf = 0;
  for (auto& kv : facet_of) kv.second = nf++;
  std::vector<int32_t> c2f(3 * (size_t)nc), f2c(2 * (size_t)nf, -1), bfacets;
  for (int c = 0; c < nc; ++c)
    for (int i = 0; i < 3; ++i) {
      std::array<int, 2> e = {cells[3 * c + (i + 1) % 3], cells[3 * c + (i + 2) % 3]};
      if (e[0] > e[1]) std::swap(e[0], e[1]);
      const int fc = facet_of[e];
      c2f[3 * c + i] = fc;
      f2c[2 * fc + (f2c[2 * fc] < 0 ? 0 : 1)] = c;  // cells ascending: c grows
    }
  for (int fc = 0; fc < nf; ++fc)
    if (f2c[2 * fc + 1] < 0) bfacets.push_back(fc);

  phifem_mesh mesh = {};
  mesh.cell_type = PHIFEM_TRIANGLE;
  mesh.gdim = 2;
  mesh.n_vertices = nvx;
  mesh.n_cells = nc;
  mesh.n_facets = nf;
  mesh.x = to_device(x);
  mesh.cells = to_device(cells);
  mesh.c2f = to_device(c2f);
  mesh.f2c = to_device(f2c);
  mesh.boundary_facets = to_device(bfacets);
  mesh.n_boundary_facets = (int64_t)bfacets.size();
  {  // once per mesh: the static part of the ds detection on the mesh-boundary facets (mesh_scripts.py:434-452)
    uint32_t* owner = nullptr;
    double* scale = nullptr;
    cudaMalloc(&owner, std::max<size_t>(1, bfacets.size()) * 2 * sizeof(uint32_t));
    cudaMalloc(&scale, std::max<size_t>(1, bfacets.size()) * 4 * sizeof(double));
    CHECK(phifem_boundary_records(&mesh, owner, scale, nullptr));
    mesh.boundary_owner = owner;
    mesh.boundary_scale = scale;
  }
  double* d_phi = to_device(phi);
  double* d_f = to_device(f);

  // P1 level set, detection degree 1: the detection points are the vertices (identity tables)
  const double ftab_h[3][2][3] = {{{0, 1, 0}, {0, 0, 1}}, {{1, 0, 0}, {0, 0, 1}}, {{1, 0, 0}, {0, 1, 0}}};
  std::vector<double> ftab(&ftab_h[0][0][0], &ftab_h[0][0][0] + 18);
  phifem_levelset ls = {};
  ls.mode = 0;
  ls.n_dofs_per_cell = 3;
  ls.n_cell_points = 3;
  ls.n_facet_points = 2;
  ls.coeffs = d_phi;
  ls.facet_table = to_device(ftab);

  int32_t *cell_tags, *facet_tags;
  int8_t *cell_tags8, *facet_tags8;
  uint8_t* scratch;
  int64_t* counters;
  cudaMalloc(&cell_tags, nc * sizeof(int32_t));
  cudaMalloc(&cell_tags8, nc);
  cudaMalloc(&facet_tags, nf * sizeof(int32_t));
  cudaMalloc(&facet_tags8, nf);
  cudaMalloc(&scratch, nvx + 3);
  cudaMalloc(&counters, PHIFEM_N_COUNTERS * sizeof(int64_t));
  cudaMemset(counters, 0, PHIFEM_N_COUNTERS * sizeof(int64_t));
  CHECK(phifem_tag_cells(&mesh, &ls, 0, cell_tags, cell_tags8, scratch, counters, nullptr));
  CHECK(phifem_tag_facets(&mesh, &ls, cell_tags8, facet_tags, facet_tags8, counters, nullptr));
  int64_t cnt[PHIFEM_N_COUNTERS];
  cudaMemcpy(cnt, counters, sizeof(cnt), cudaMemcpyDeviceToHost);

  // ds(100): Gamma_h facets (tag 4) seen from the cells tagged 1 / 2 (src/phifem/mesh_scripts.py:619-622)
  int64_t n_ent = 0;
  int32_t* ents = nullptr;
  CHECK(phifem_integration_entities(&mesh, cell_tags8, facet_tags8, 4, (1u << 1) | (1u << 2), nullptr, 0, &n_ent, nullptr));
  cudaMalloc(&ents, std::max<int64_t>(1, n_ent) * 2 * sizeof(int32_t));
  CHECK(phifem_integration_entities(&mesh, cell_tags8, facet_tags8, 4, (1u << 1) | (1u << 2), ents, n_ent, &n_ent, nullptr));

  phifem_pattern* pat = nullptr;
  CHECK(phifem_pattern_create_p1(&mesh, cell_tags8, facet_tags8, ents, n_ent, &pat, nullptr));
  phifem_pattern_view v;
  CHECK(phifem_pattern_view_of(pat, &v));
  double *data, *b;
  cudaMalloc(&data, std::max<int64_t>(1, v.nnz) * sizeof(double));
  cudaMalloc(&b, v.n_rows * sizeof(double));
  cudaMemset(data, 0, v.nnz * sizeof(double));
  cudaMemset(b, 0, v.n_rows * sizeof(double));
  if (v.n_active)
    CHECK(phifem_assemble_cells_p1(&mesh, d_phi, d_f, cell_tags8, v.active, v.n_active, v.slots_cells, 1.0, data, b,
                                   nullptr));
  if (v.n_entities) CHECK(phifem_assemble_boundary_p1(&mesh, d_phi, ents, v.n_entities, v.slots_boundary, data, nullptr));
  if (v.n_ghost) CHECK(phifem_assemble_ghost_p1(&mesh, d_phi, v.ghost, v.n_ghost, v.slots_ghost, 1.0, data, nullptr));
  if (cudaDeviceSynchronize() != cudaSuccess) {
    std::fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 1;
  }
  std::vector<double> h_data(v.nnz), h_b(v.n_rows);
  std::vector<int32_t> h_idx(v.nnz);
  cudaMemcpy(h_data.data(), data, v.nnz * sizeof(double), cudaMemcpyDeviceToHost);
  cudaMemcpy(h_b.data(), b, v.n_rows * sizeof(double), cudaMemcpyDeviceToHost);
  cudaMemcpy(h_idx.data(), v.indices, v.nnz * sizeof(int32_t), cudaMemcpyDeviceToHost);
  double abs_sum = 0.0, b_sum = 0.0;
  long long idx_sum = 0;
  for (int64_t i = 0; i < v.nnz; ++i) {
    abs_sum += std::fabs(h_data[i]);
    idx_sum += (long long)h_idx[i] * (i % 7 + 1);
  }
  for (double t : h_b) b_sum += t;
  // The benchmarked path: the row-gather plan built on the device (phifem_rows_plan_create) and the row-gather kernels
  // (phifem_assemble_rows_p1: no atomics, no zero-fill).  Same pattern, same operator as the per-entity kernels above.
  phifem_rows_plan_handle* rh = nullptr;
  CHECK(phifem_rows_plan_create(&mesh, cell_tags8, facet_tags8, ents, n_ent, nullptr, 0, &rh, nullptr));
  phifem_rows_plan rplan;
  phifem_rows_plan_info rinfo;
  CHECK(phifem_rows_plan_view(rh, &rplan, &rinfo));
  double *data2, *b2;
  cudaMalloc(&data2, std::max<int64_t>(1, rinfo.nnz) * sizeof(double));
  cudaMalloc(&b2, rinfo.n_rows * sizeof(double));
  cudaMemset(data2, 0xff, rinfo.nnz * sizeof(double));  // NaNs: every entry must be written by the kernels
  cudaMemset(b2, 0, rinfo.n_rows * sizeof(double));
  CHECK(phifem_assemble_rows_p1(&mesh, d_phi, d_f, 1.0, &rplan, data2, b2, nullptr));
  if (cudaDeviceSynchronize() != cudaSuccess) {
    std::fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 1;
  }
  std::vector<double> r_data(rinfo.nnz), r_b(rinfo.n_rows);
  std::vector<int32_t> r_idx(rinfo.nnz);
  cudaMemcpy(r_data.data(), data2, rinfo.nnz * sizeof(double), cudaMemcpyDeviceToHost);
  cudaMemcpy(r_b.data(), b2, rinfo.n_rows * sizeof(double), cudaMemcpyDeviceToHost);
  cudaMemcpy(r_idx.data(), rplan.indices, rinfo.nnz * sizeof(int32_t), cudaMemcpyDeviceToHost);
  double rows_abs_sum = 0.0, rows_b_sum = 0.0, max_diff = 0.0, max_abs = 0.0;
  int same_pattern = rinfo.nnz == v.nnz;
  for (int64_t i = 0; same_pattern && i < v.nnz; ++i) {
    same_pattern = r_idx[i] == h_idx[i];
    rows_abs_sum += std::fabs(r_data[i]);
    max_diff = std::max(max_diff, std::fabs(r_data[i] - h_data[i]));
    max_abs = std::max(max_abs, std::fabs(h_data[i]));
  }
  for (double t : r_b) rows_b_sum += t;
  std::printf("interior=%lld cut=%lld exterior=%lld nnz=%lld n_active=%lld n_ghost=%lld n_entities=%lld "
              "indices_checksum=%lld data_abs_sum=%.17g b_sum=%.17g rows_same_pattern=%d rows_data_abs_sum=%.17g "
              "rows_b_sum=%.17g rows_max_rel_diff=%.3e\n",
              (long long)cnt[PHIFEM_CNT_INTERIOR], (long long)cnt[PHIFEM_CNT_CUT], (long long)cnt[PHIFEM_CNT_EXTERIOR],
              (long long)v.nnz, (long long)v.n_active, (long long)v.n_ghost, (long long)v.n_entities, idx_sum, abs_sum,
              b_sum, same_pattern, rows_abs_sum, rows_b_sum, max_diff / (max_abs > 0 ? max_abs : 1.0));
  phifem_rows_plan_destroy(rh);
  phifem_pattern_destroy(pat);
  return 0;
}
