// Symbolic phase of the P1 strong-Dirichlet operator behind the C ABI: what dolfinx does in
// `create_sparsity_pattern` + `create_matrix` when `assemble_matrix(form(a))` is first called
// (reference demo/strong-dirichlet/flower/main.py:121-123) -- the CSR pattern of the dofs coupled by the cells of
// dx((1,2)) and by the macro elements of the interior facets of dS((2,3)) [dep-knowledge, SURVEY.md C.3] -- plus the
// entity -> CSR-slot maps the per-entity kernels (phifem_assemble_{cells,boundary,ghost}_p1) scatter through.
// The Python package builds the same arrays with torch sort / unique (phifem_b200/assemble.py); this file lets a host
// without Python (a C++ / PETSc code base) run tags -> pattern -> assembly with the shared library alone.
//
// Device algorithm: stream compaction of the active cells and ghost facets (cub::DeviceSelect), one 64-bit key
// row * n_vertices + col per coupled vertex pair, radix sort + unique (cub) = the pattern, binary searches for the row
// pointers and the slot maps.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "common.cuh"

struct phifem_pattern {
  phifem_pattern_view v;
  void* owned[8];
  int n_owned;
};

namespace phifem {
// A private stream-ordered pool per device whose memory stays cached between calls (release threshold = max): the
// default pool hands its memory back to the driver at every synchronisation, and re-allocating the ~10 GB of sort
// scratch of config E costs ten times the sort.  phifem_pattern_release_scratch() trims it.
cudaMemPool_t g_pools[64] = {};
cudaMemPool_t scratch_pool() {  // also used by csrc/rows_plan.cu
  int dev = 0;
  cudaGetDevice(&dev);
  cudaMemPool_t& pool = g_pools[dev & 63];
  if (!pool) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) {
      pool = nullptr;
      return nullptr;
    }
    uint64_t keep = 6ull << 30;  // cached between calls up to 6 GiB (the row-gather plan of 50 M tetrahedra needs ~3);
                                 // anything above goes back to the driver at the next synchronisation
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  return pool;
}

namespace {

constexpr int kBlockSym = 256;

struct IsActiveCell {
  const int8_t* tags;
  __device__ bool operator()(int c) const { return tags[c] == 1 || tags[c] == 2; }
};
struct IsGhostFacet {
  const int8_t* tags;
  const int32_t* f2c;
  __device__ bool operator()(int f) const { return (tags[f] == 2 || tags[f] == 3) && f2c[2 * (int64_t)f + 1] >= 0; }
};

// macro vertex list of an interior facet: [facet vertices in the local order of cell + (= f2c[f][0]), opposite vertex of
// cell +, opposite vertex of cell -] (include/phifem_b200.h, phifem_assemble_ghost_p1)
__global__ void k_macro_vertices(phifem_mesh m, int nv, const int32_t* __restrict__ facets, int64_t n,
                                 int32_t* __restrict__ macro) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int32_t f = facets[e];
  int32_t* out = macro + e * (nv + 1);
  int k = 0;
  for (int side = 0; side < 2; ++side) {
    const int64_t c = m.f2c[2 * (int64_t)f + side];
    int opposite = -1;
    for (int i = 0; i < nv; ++i) {
      const int32_t v = m.cells[c * nv + i];
      if (m.c2f[c * nv + i] == f) opposite = v;  // local facet i is opposite local vertex i
      else if (side == 0) out[k++] = v;
    }
    out[nv - 1 + side] = opposite;
  }
}

// keys row * n_vertices + col of every ordered vertex pair of every listed entity; verts[e * w + i]
__global__ void k_pair_keys(const int32_t* __restrict__ verts, int64_t n, int w, int64_t n_vertices,
                            int64_t* __restrict__ keys) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * w * w) return;
  const int64_t e = t / (w * w);
  const int r = (int)(t - e * w * w);
  keys[t] = (int64_t)verts[e * w + r / w] * n_vertices + verts[e * w + r % w];
}

// the same with an indirection: entity e = cells row list[e * stride] (active cells, one-sided entities)
__global__ void k_cell_pair_keys(const int32_t* __restrict__ cells, int nv, const int32_t* __restrict__ list,
                                 int stride, int64_t n, int64_t n_vertices, int64_t* __restrict__ keys) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * nv * nv) return;
  const int64_t e = t / (nv * nv);
  const int r = (int)(t - e * nv * nv);
  const int64_t c = list[e * stride];
  keys[t] = (int64_t)cells[c * nv + r / nv] * n_vertices + cells[c * nv + r % nv];
}

__device__ __forceinline__ int64_t lower_bound(const int64_t* __restrict__ a, int64_t n, int64_t key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__global__ void k_slots(const int64_t* __restrict__ uniq, int64_t nnz, const int64_t* __restrict__ keys, int64_t n,
                        int32_t* __restrict__ slots) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) slots[t] = (int32_t)lower_bound(uniq, nnz, keys[t]);
}

__global__ void k_row_pointers(const int64_t* __restrict__ uniq, int64_t nnz, int64_t n_rows,
                               int32_t* __restrict__ indptr) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r <= n_rows) indptr[r] = (int32_t)lower_bound(uniq, nnz, r * n_rows);
}

__global__ void k_columns(const int64_t* __restrict__ uniq, int64_t nnz, int64_t n_rows,
                          int32_t* __restrict__ indices) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nnz) indices[t] = (int32_t)(uniq[t] % n_rows);
}

inline unsigned blocks_for(int64_t n) { return (unsigned)((n + kBlockSym - 1) / kBlockSym); }

// ordering of `_compute_integration_entities` (mesh_scripts.py:137-192): cells in first-appearance order of the
// candidate records (key = 2 facet + column), the local facets of a cell ascending
__global__ void k_entity_first_key(const int64_t* __restrict__ rec, int64_t n, unsigned long long* __restrict__ first) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) atomicMin(first + rec[3 * t + 1], (unsigned long long)rec[3 * t]);
}
__global__ void k_entity_sort_keys(const int64_t* __restrict__ rec, int64_t n,
                                   const unsigned long long* __restrict__ first, int64_t* __restrict__ skey,
                                   int64_t* __restrict__ sval) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int64_t cell = rec[3 * t + 1], lf = rec[3 * t + 2];
  skey[t] = (int64_t)first[cell] * 8 + lf;
  sval[t] = cell * 8 + lf;
}
__global__ void k_entity_unpack(const int64_t* __restrict__ sval, int64_t n, int32_t* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  out[2 * t] = (int32_t)(sval[t] >> 3);
  out[2 * t + 1] = (int32_t)(sval[t] & 7);
}

struct Scratch {  // device allocations freed on scope exit (stream-ordered)
  cudaStream_t st;
  cudaMemPool_t pool;
  void* p[16];
  int n = 0;
  Scratch(cudaStream_t s, cudaMemPool_t pl) : st(s), pool(pl) {}
  void* get(size_t bytes) {
    void* q = nullptr;
    if (cudaMallocFromPoolAsync(&q, bytes ? bytes : 1, pool, st) != cudaSuccess) return nullptr;
    p[n++] = q;
    return q;
  }
  ~Scratch() {
    for (int i = 0; i < n; ++i) cudaFreeAsync(p[i], st);
  }
};

}  // namespace
}  // namespace phifem

using namespace phifem;

namespace phifem {
void plan_block_cache_release();  // csrc/rows_plan.cu: parked arrays of destroyed row-gather plans
}
extern "C" void phifem_pattern_release_scratch(void) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (g_pools[dev & 63]) cudaMemPoolTrimTo(g_pools[dev & 63], 0);
  plan_block_cache_release();
}

extern "C" void phifem_pattern_destroy(phifem_pattern* p) {
  if (!p) return;
  for (int i = 0; i < p->n_owned; ++i) cudaFree(p->owned[i]);
  delete p;
}

extern "C" int phifem_pattern_view_of(const phifem_pattern* p, phifem_pattern_view* out) {
  PHIFEM_CHECK_ARG(p && out, "null pointer");
  *out = p->v;
  return PHIFEM_OK;
}

extern "C" int phifem_pattern_create_p1(const phifem_mesh* mesh, const int8_t* cell_tags8, const int8_t* facet_tags8,
                                        const int32_t* entities, int64_t n_entities, phifem_pattern** out,
                                        void* stream) {
  PHIFEM_CHECK_ARG(out != nullptr, "out is null");
  *out = nullptr;
  PHIFEM_CHECK_ARG(mesh != nullptr && mesh->cells && mesh->c2f && mesh->f2c, "mesh / facet connectivity is null");
  if (mesh->cell_type != PHIFEM_TRIANGLE && mesh->cell_type != PHIFEM_TETRAHEDRON) {
    set_error("P1 assembly supports triangles and tetrahedra, got cell type %d", mesh->cell_type);
    return PHIFEM_ERR_UNSUPPORTED;
  }
  PHIFEM_CHECK_ARG(cell_tags8 && facet_tags8, "tag arrays are null");
  PHIFEM_CHECK_ARG(n_entities >= 0 && (n_entities == 0 || entities), "entity list");
  PHIFEM_CHECK_ARG(mesh->n_cells < (1ll << 31) && mesh->n_facets < (1ll << 31) && mesh->n_vertices < (1ll << 31),
                   "int32 index width");
  const int nv = mesh->cell_type == PHIFEM_TRIANGLE ? 3 : 4, nm = nv + 1;
  const int64_t n_rows = mesh->n_vertices;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemPool_t pool = scratch_pool();
  if (!pool) {
    set_error("phifem_pattern_create_p1: cannot create the scratch memory pool");
    return PHIFEM_ERR_CUDA;
  }
  Scratch tmp(st, pool);
  phifem_pattern* pat = new phifem_pattern();
  pat->n_owned = 0;
  auto own = [&](size_t bytes) -> void* {
    void* q = nullptr;
    if (cudaMalloc(&q, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    pat->owned[pat->n_owned++] = q;
    return q;
  };
  auto fail = [&](const char* what) {
    set_error("phifem_pattern_create_p1: %s: %s", what, cudaGetErrorString(cudaGetLastError()));
    phifem_pattern_destroy(pat);
    return PHIFEM_ERR_CUDA;
  };

  // 1. active cells (tags 1 / 2) and ghost-penalty facets (interior facets tagged 2 / 3), ascending
  int32_t* active_full = (int32_t*)tmp.get(sizeof(int32_t) * mesh->n_cells);
  int32_t* ghost_full = (int32_t*)tmp.get(sizeof(int32_t) * mesh->n_facets);
  int64_t* counts = (int64_t*)tmp.get(2 * sizeof(int64_t));
  if (!active_full || !ghost_full || !counts) return fail("scratch allocation");
  {
    thrust::counting_iterator<int> it(0);
    size_t b1 = 0, b2 = 0;
    cub::DeviceSelect::If(nullptr, b1, it, active_full, counts, (int)mesh->n_cells, IsActiveCell{cell_tags8}, st);
    cub::DeviceSelect::If(nullptr, b2, it, ghost_full, counts + 1, (int)mesh->n_facets,
                          IsGhostFacet{facet_tags8, mesh->f2c}, st);
    void* ws = tmp.get(b1 > b2 ? b1 : b2);
    if (!ws) return fail("scratch allocation");
    cub::DeviceSelect::If(ws, b1, it, active_full, counts, (int)mesh->n_cells, IsActiveCell{cell_tags8}, st);
    cub::DeviceSelect::If(ws, b2, it, ghost_full, counts + 1, (int)mesh->n_facets,
                          IsGhostFacet{facet_tags8, mesh->f2c}, st);
  }
  int64_t h_counts[2] = {0, 0};
  if (cudaMemcpyAsync(h_counts, counts, sizeof(h_counts), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess)
    return fail("compaction");
  const int64_t na = h_counts[0], ng = h_counts[1];
  int32_t* active = (int32_t*)own(sizeof(int32_t) * na);
  int32_t* ghost = (int32_t*)own(sizeof(int32_t) * ng);
  if (!active || !ghost) return fail("output allocation");
  cudaMemcpyAsync(active, active_full, sizeof(int32_t) * na, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(ghost, ghost_full, sizeof(int32_t) * ng, cudaMemcpyDeviceToDevice, st);

  // 2. keys of every coupled pair: cells, ghost macro elements, one-sided entities (the latter lie inside active cells:
  //    they add no pattern entry but need slots)
  const int64_t kc = na * nv * nv, kg = ng * nm * nm, kb = n_entities * nv * nv;
  int64_t* keys = (int64_t*)tmp.get(sizeof(int64_t) * (kc + kg + kb));
  int32_t* macro = (int32_t*)tmp.get(sizeof(int32_t) * ng * nm);
  if (!keys || !macro) return fail("scratch allocation");
  if (kc) k_cell_pair_keys<<<blocks_for(kc), kBlockSym, 0, st>>>(mesh->cells, nv, active, 1, na, n_rows, keys);
  if (ng) {
    k_macro_vertices<<<blocks_for(ng), kBlockSym, 0, st>>>(*mesh, nv, ghost, ng, macro);
    k_pair_keys<<<blocks_for(kg), kBlockSym, 0, st>>>(macro, ng, nm, n_rows, keys + kc);
  }
  if (kb)
    k_cell_pair_keys<<<blocks_for(kb), kBlockSym, 0, st>>>(mesh->cells, nv, entities, 2, n_entities, n_rows,
                                                          keys + kc + kg);

  // 3. pattern = sorted unique keys of the cells and macro elements
  const int64_t nk = kc + kg;
  if (nk >= (1ll << 31)) {
    set_error("phifem_pattern_create_p1: %lld coupled pairs (limit 2^31): shard the mesh", (long long)nk);
    phifem_pattern_destroy(pat);
    return PHIFEM_ERR_ARGUMENT;
  }
  int64_t* sorted = (int64_t*)tmp.get(sizeof(int64_t) * nk);
  int64_t* uniq = (int64_t*)tmp.get(sizeof(int64_t) * nk);
  int64_t* d_nnz = (int64_t*)tmp.get(sizeof(int64_t));
  if (!sorted || !uniq || !d_nnz) return fail("scratch allocation");
  int64_t nnz = 0;
  if (nk) {
    int end_bit = 1;
    while (end_bit < 63 && (1ll << end_bit) < n_rows * n_rows) ++end_bit;
    size_t b1 = 0, b2 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, b1, keys, sorted, (int)nk, 0, end_bit, st);
    cub::DeviceSelect::Unique(nullptr, b2, sorted, uniq, d_nnz, (int)nk, st);
    void* ws = tmp.get(b1 > b2 ? b1 : b2);
    if (!ws) return fail("scratch allocation");
    cub::DeviceRadixSort::SortKeys(ws, b1, keys, sorted, (int)nk, 0, end_bit, st);
    cub::DeviceSelect::Unique(ws, b2, sorted, uniq, d_nnz, (int)nk, st);
    if (cudaMemcpyAsync(&nnz, d_nnz, sizeof(nnz), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
      return fail("sort / unique");
  }

  // 4. CSR arrays and slot maps
  int32_t* indptr = (int32_t*)own(sizeof(int32_t) * (n_rows + 1));
  int32_t* indices = (int32_t*)own(sizeof(int32_t) * nnz);
  // three allocations: the kernels read their slot rows with vector loads (cudaMalloc alignment)
  int32_t* slots_c = (int32_t*)own(sizeof(int32_t) * kc);
  int32_t* slots_g = (int32_t*)own(sizeof(int32_t) * kg);
  int32_t* slots_b = (int32_t*)own(sizeof(int32_t) * kb);
  if (!indptr || !indices || !slots_c || !slots_g || !slots_b) return fail("output allocation");
  k_row_pointers<<<blocks_for(n_rows + 1), kBlockSym, 0, st>>>(uniq, nnz, n_rows, indptr);
  if (nnz) k_columns<<<blocks_for(nnz), kBlockSym, 0, st>>>(uniq, nnz, n_rows, indices);
  if (kc) k_slots<<<blocks_for(kc), kBlockSym, 0, st>>>(uniq, nnz, keys, kc, slots_c);
  if (kg) k_slots<<<blocks_for(kg), kBlockSym, 0, st>>>(uniq, nnz, keys + kc, kg, slots_g);
  if (kb) k_slots<<<blocks_for(kb), kBlockSym, 0, st>>>(uniq, nnz, keys + kc + kg, kb, slots_b);
  if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) return fail("pattern kernels");

  pat->v.n_rows = n_rows;
  pat->v.nnz = nnz;
  pat->v.n_active = na;
  pat->v.n_ghost = ng;
  pat->v.n_entities = n_entities;
  pat->v.indptr = indptr;
  pat->v.indices = indices;
  pat->v.active = active;
  pat->v.ghost = ghost;
  pat->v.slots_cells = slots_c;
  pat->v.slots_ghost = slots_g;
  pat->v.slots_boundary = slots_b;
  *out = pat;
  return PHIFEM_OK;
}

extern "C" int phifem_integration_entities_count(const phifem_mesh* mesh, const int8_t* cell_tags8,
                                                 const int8_t* facet_tags8, int32_t facet_tag, uint32_t cell_mask,
                                                 int64_t* n_entities_dev, void* stream) {
  PHIFEM_CHECK_ARG(mesh != nullptr && n_entities_dev != nullptr, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(n_entities_dev, 0, sizeof(int64_t), st);
  if (mesh->n_facets == 0) return PHIFEM_OK;
  return phifem_entity_records(mesh, cell_tags8, facet_tags8, facet_tag, cell_mask, nullptr, 0, n_entities_dev, stream);
}

extern "C" int phifem_integration_entities_fill(const phifem_mesh* mesh, const int8_t* cell_tags8,
                                                const int8_t* facet_tags8, int32_t facet_tag, uint32_t cell_mask,
                                                int32_t* entities, int64_t n, void* stream) {
  PHIFEM_CHECK_ARG(mesh != nullptr, "null pointer");
  PHIFEM_CHECK_ARG(n >= 0 && (n == 0 || entities), "entity buffer");
  if (n == 0 || mesh->n_facets == 0) return PHIFEM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemPool_t pool = scratch_pool();
  if (!pool) {
    set_error("phifem_integration_entities_fill: cannot create the scratch memory pool");
    return PHIFEM_ERR_CUDA;
  }
  Scratch tmp(st, pool);
  auto fail = [&](const char* what) {
    set_error("phifem_integration_entities_fill: %s: %s", what, cudaGetErrorString(cudaGetLastError()));
    return PHIFEM_ERR_CUDA;
  };
  int64_t* d_n = (int64_t*)tmp.get(sizeof(int64_t));
  int64_t* rec = (int64_t*)tmp.get(sizeof(int64_t) * 3 * n);
  unsigned long long* first = (unsigned long long*)tmp.get(sizeof(unsigned long long) * mesh->n_cells);
  int64_t* skey = (int64_t*)tmp.get(sizeof(int64_t) * n);
  int64_t* sval = (int64_t*)tmp.get(sizeof(int64_t) * n);
  int64_t* skey2 = (int64_t*)tmp.get(sizeof(int64_t) * n);
  int64_t* sval2 = (int64_t*)tmp.get(sizeof(int64_t) * n);
  if (!d_n || !rec || !first || !skey || !sval || !skey2 || !sval2) return fail("scratch allocation");
  cudaMemsetAsync(d_n, 0, sizeof(int64_t), st);
  if (int rc = phifem_entity_records(mesh, cell_tags8, facet_tags8, facet_tag, cell_mask, rec, n, d_n, stream))
    return rc;
  cudaMemsetAsync(first, 0xff, sizeof(unsigned long long) * mesh->n_cells, st);
  k_entity_first_key<<<blocks_for(n), kBlockSym, 0, st>>>(rec, n, first);
  k_entity_sort_keys<<<blocks_for(n), kBlockSym, 0, st>>>(rec, n, first, skey, sval);
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, skey, skey2, sval, sval2, (int)n, 0, 40, st);
  void* ws = tmp.get(bytes);
  if (!ws) return fail("scratch allocation");
  cub::DeviceRadixSort::SortPairs(ws, bytes, skey, skey2, sval, sval2, (int)n, 0, 40, st);  // keys < 2^(32 + 1 + 3)
  k_entity_unpack<<<blocks_for(n), kBlockSym, 0, st>>>(sval2, n, entities);
  if (cudaGetLastError() != cudaSuccess) return fail("ordering kernels");
  return PHIFEM_OK;
}

extern "C" int phifem_integration_entities(const phifem_mesh* mesh, const int8_t* cell_tags8, const int8_t* facet_tags8,
                                           int32_t facet_tag, uint32_t cell_mask, int32_t* entities, int64_t capacity,
                                           int64_t* n_entities, void* stream) {
  PHIFEM_CHECK_ARG(mesh != nullptr && n_entities != nullptr, "null pointer");
  PHIFEM_CHECK_ARG(capacity >= 0 && (capacity == 0 || entities), "entity buffer");
  *n_entities = 0;
  if (mesh->n_facets == 0) return PHIFEM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemPool_t pool = scratch_pool();
  if (!pool) {
    set_error("phifem_integration_entities: cannot create the scratch memory pool");
    return PHIFEM_ERR_CUDA;
  }
  int64_t n = 0;
  {
    Scratch tmp(st, pool);
    int64_t* d_n = (int64_t*)tmp.get(sizeof(int64_t));
    if (!d_n) {
      set_error("phifem_integration_entities: scratch allocation: %s", cudaGetErrorString(cudaGetLastError()));
      return PHIFEM_ERR_CUDA;
    }
    if (int rc = phifem_integration_entities_count(mesh, cell_tags8, facet_tags8, facet_tag, cell_mask, d_n, stream))
      return rc;
    if (cudaMemcpyAsync(&n, d_n, sizeof(n), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) {
      set_error("phifem_integration_entities: counting: %s", cudaGetErrorString(cudaGetLastError()));
      return PHIFEM_ERR_CUDA;
    }
  }
  *n_entities = n;
  if (n == 0 || n > capacity) return PHIFEM_OK;  // the caller retries with a buffer of *n_entities pairs
  if (int rc = phifem_integration_entities_fill(mesh, cell_tags8, facet_tags8, facet_tag, cell_mask, entities, n, stream))
    return rc;
  if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
    set_error("phifem_integration_entities: ordering kernels: %s", cudaGetErrorString(cudaGetLastError()));
    return PHIFEM_ERR_CUDA;
  }
  return PHIFEM_OK;
}
