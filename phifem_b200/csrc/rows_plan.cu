// Symbolic phase of the row-gather assembly behind the C ABI: CSR pattern + the row lists of phifem_rows_plan
// (include/phifem_b200.h) built on the device, without Python.
//
// What dolfinx does in `create_sparsity_pattern` + `create_matrix` when `assemble_matrix(form(a))` is first called
// (reference demo/strong-dirichlet/flower/main.py:121-123) [dep-knowledge, SURVEY.md C.3], plus the per-row entity
// lists the kernels of csrc/assemble_rows.cu walk.  The Python package builds the same arrays with torch sort / unique
// passes (phifem_b200/assemble.py + rows.py: every ordered vertex pair of every active cell becomes a 64-bit key:
// 317 M keys at config E); this builder works from the vertex -> cell adjacency instead:
//   1. stream compaction of the active cells / ghost-penalty facets (cub::DeviceSelect);
//   2. one radix sort of (vertex, active cell, local index) keys = the vertex -> cell lists, cells ascending;
//   3. a thread per row merges the vertices of its cells (and the opposite vertex across each of its ghost facets) into
//      a sorted unique list held in local memory: pass 1 counts (-> indptr by a scan), pass 2 writes the column indices,
//      the position of the diagonal and the row's cell records -- the positions of the cell's other vertices inside the
//      column list it has just built -- straight into the sliced-ELLPACK layout;
//   4. surface records (ghost-penalty macro elements, one-sided entities): one sort by (row, sequence), rows ordered
//      along the Morton curve and balanced by record count inside chunks of 4096 rows, as phifem_b200/rows.py does.
// Every array equals the torch-built plan bit for bit (tests/test_gpu_rows_plan_capi.py).
#include <mutex>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "common.cuh"

struct phifem_rows_plan_handle {
  phifem_rows_plan plan;
  phifem_rows_plan_info info;
  void* owned[40];
  int n_owned;
  size_t owned_bytes[40];
  int device;
};

namespace phifem {
cudaMemPool_t scratch_pool();  // csrc/symbolic.cu: private stream-ordered pool, cached between calls

// The arrays a plan OWNS (pattern, row lists, records: 0.55 GB at config E) are cudaMalloc blocks; a destroyed plan parks
// them here and the next plan takes the blocks that fit, so that re-planning -- what a moving interface pays whenever the
// cut pattern changes -- makes no driver allocation at all (73 -> 14 ms at config E; taking them from the stream-ordered
// pool instead made the FIRST plan of a process 0.5-0.9 s: the pool grows by one mapping per array).
// phifem_pattern_release_scratch() empties the cache.
namespace {
struct BlockCache {
  static constexpr int kMax = 96;
  void* ptr[kMax];
  size_t bytes[kMax];
  int n = 0;
  size_t total = 0;
};
BlockCache g_block_cache[64];
std::mutex g_block_cache_mutex;  // plans may be created / destroyed from several host threads
constexpr size_t kBlockCacheLimit = 4ull << 30;
}  // namespace

void* plan_block_take(int dev, size_t bytes) {
  std::lock_guard<std::mutex> lock(g_block_cache_mutex);
  BlockCache& c = g_block_cache[dev & 63];
  int best = -1;
  for (int i = 0; i < c.n; ++i)
    if (c.bytes[i] >= bytes && c.bytes[i] <= bytes + bytes / 4 + 4096 && (best < 0 || c.bytes[i] < c.bytes[best])) best = i;
  if (best < 0) return nullptr;
  void* q = c.ptr[best];
  c.total -= c.bytes[best];
  c.ptr[best] = c.ptr[c.n - 1];
  c.bytes[best] = c.bytes[c.n - 1];
  --c.n;
  return q;
}
void plan_block_give(int dev, void* q, size_t bytes) {
  std::lock_guard<std::mutex> lock(g_block_cache_mutex);
  BlockCache& c = g_block_cache[dev & 63];
  if (c.n >= BlockCache::kMax || c.total + bytes > kBlockCacheLimit) {
    cudaFree(q);
    return;
  }
  c.ptr[c.n] = q;
  c.bytes[c.n] = bytes;
  ++c.n;
  c.total += bytes;
}
void plan_block_cache_release() {  // called by phifem_pattern_release_scratch (csrc/symbolic.cu)
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_block_cache_mutex);
  BlockCache& c = g_block_cache[dev & 63];
  for (int i = 0; i < c.n; ++i) cudaFree(c.ptr[i]);
  c.n = 0;
  c.total = 0;
}
namespace {

constexpr int kB = 256;
constexpr int kMaxNnz = 255;          // positions are uint8
constexpr int kBalanceChunk = 4096;   // phifem_b200/rows.py BALANCE_CHUNK
inline unsigned nblk(int64_t n) { return (unsigned)((n + kB - 1) / kB); }

struct IsActive {
  const int8_t* tags;
  __device__ bool operator()(int c) const { return tags[c] == 1 || tags[c] == 2; }
};
struct IsGhost {
  const int8_t* tags;
  const int32_t* f2c;
  __device__ bool operator()(int f) const { return (tags[f] == 2 || tags[f] == 3) && f2c[2 * (int64_t)f + 1] >= 0; }
};
struct HighWord {
  __device__ int32_t operator()(uint64_t k) const { return (int32_t)(k >> 32); }
};

struct Tmp {  // stream-ordered scratch, freed on scope exit
  cudaStream_t st;
  cudaMemPool_t pool;
  void* p[96];
  int n = 0;
  bool ok = true;
  Tmp(cudaStream_t s, cudaMemPool_t pl) : st(s), pool(pl) {}
  template <typename T> T* get(int64_t count) {
    void* q = nullptr;
    if (n >= 96 || cudaMallocFromPoolAsync(&q, (size_t)(count > 0 ? count : 1) * sizeof(T), pool, st) != cudaSuccess) {
      ok = false;
      return nullptr;
    }
    p[n++] = q;
    return (T*)q;
  }
  ~Tmp() {
    for (int i = 0; i < n; ++i) cudaFreeAsync(p[i], st);
  }
};

__device__ __forceinline__ int64_t lower_bound_u64(const uint64_t* __restrict__ a, int64_t n, uint64_t key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// macro vertex list of an interior facet: [facet vertices in the local order of cell + (= f2c[f][0]), opposite vertex of
// cell +, opposite vertex of cell -]
__global__ void k_macro(phifem_mesh m, int nv, const int32_t* __restrict__ facets, int64_t n, int32_t* __restrict__ macro) {
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
      if (m.c2f[c * nv + i] == f) opposite = v;
      else if (side == 0) out[k++] = v;
    }
    out[nv - 1 + side] = opposite;
  }
}

// one-sided entity (cell, local facet o): [facet vertices in ascending local order, opposite vertex]
__global__ void k_entity_macro(const int32_t* __restrict__ cells, int nv, const int32_t* __restrict__ ents, int64_t n,
                               int32_t* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int64_t c = ents[2 * e];
  const int o = ents[2 * e + 1];
  int k = 0;
  for (int i = 0; i < nv; ++i)
    if (i != o) out[e * nv + k++] = cells[c * nv + i];
  out[e * nv + nv - 1] = cells[c * nv + o];
}

__global__ void k_v2c_keys(const int32_t* __restrict__ cells, int nv, const int32_t* __restrict__ active, int64_t na,
                           uint64_t* __restrict__ keys) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= na * nv) return;
  const int64_t a = t / nv;
  const int i = (int)(t - a * nv);
  const uint32_t v = (uint32_t)cells[(int64_t)active[a] * nv + i];
  keys[t] = ((uint64_t)v << 32) | (uint64_t)(a * 4 + i);
}

__global__ void k_extra_keys(const int32_t* __restrict__ macro, int nv, int64_t ng, uint64_t* __restrict__ keys) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ng) return;
  const uint32_t a = (uint32_t)macro[g * (nv + 1) + nv - 1], b = (uint32_t)macro[g * (nv + 1) + nv];
  keys[2 * g] = ((uint64_t)a << 32) | b;
  keys[2 * g + 1] = ((uint64_t)b << 32) | a;
}

__global__ void k_starts(const uint64_t* __restrict__ sorted, int64_t n, int64_t n_rows, int32_t* __restrict__ ptr) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r <= n_rows) ptr[r] = (int32_t)lower_bound_u64(sorted, n, (uint64_t)r << 32);
}

// sorted unique insertion into nb[0..n)
__device__ __forceinline__ bool insert_sorted(int32_t* nb, int& n, int32_t v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (nb[mid] < v) lo = mid + 1;
    else hi = mid;
  }
  if (lo < n && nb[lo] == v) return true;
  if (n >= kMaxNnz + 1) return false;
  for (int k = n; k > lo; --k) nb[k] = nb[k - 1];
  nb[lo] = v;
  ++n;
  return true;
}
__device__ __forceinline__ int find_sorted(const int32_t* nb, int n, int32_t v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (nb[mid] < v) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// PASS 0: row_nnz[r]; PASS 1: column indices, and for listed rows the diagonal position and the cell records
template <int PASS>
__global__ void __launch_bounds__(128) k_row_pattern(
    const int32_t* __restrict__ cells, int nv, const int8_t* __restrict__ ctags, const int32_t* __restrict__ active,
    const uint64_t* __restrict__ v2c, const int32_t* __restrict__ vptr, const uint64_t* __restrict__ extra,
    const int32_t* __restrict__ xptr, int64_t n_rows, int32_t* __restrict__ row_nnz, int* __restrict__ overflow,
    const int32_t* __restrict__ indptr, int32_t* __restrict__ indices, const int32_t* __restrict__ li_of_row,
    uint8_t* __restrict__ diag_pos, const int32_t* __restrict__ rec_ptr, uint32_t* __restrict__ rec) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int b0 = vptr[r], b1 = vptr[r + 1];
  if (b0 == b1) {
    if (PASS == 0) row_nnz[r] = 0;
    return;
  }
  int32_t nb[kMaxNnz + 1];
  int n = 0;
  bool ok = insert_sorted(nb, n, (int32_t)r);
  for (int k = b0; k < b1; ++k) {
    const uint32_t ai = (uint32_t)v2c[k];
    const int64_t c = active[ai >> 2];
    const int i = ai & 3;
    for (int j = 0; j < nv; ++j)
      if (j != i) ok = insert_sorted(nb, n, cells[c * nv + j]) && ok;
  }
  for (int k = xptr[r]; k < xptr[r + 1]; ++k) ok = insert_sorted(nb, n, (int32_t)(uint32_t)extra[k]) && ok;
  if (PASS == 0) {
    if (!ok || n > kMaxNnz) {
      *overflow = 1;
      n = 0;
    }
    row_nnz[r] = n;
    return;
  }
  const int start = indptr[r];
  for (int k = 0; k < n; ++k) indices[start + k] = nb[k];
  const int li = li_of_row[r];
  if (li < 0) return;
  diag_pos[li] = (uint8_t)find_sorted(nb, n, (int32_t)r);
  const int64_t base = (int64_t)rec_ptr[li >> 5] * 32 + (li & 31);
  for (int k = b0; k < b1; ++k) {
    const uint32_t ai = (uint32_t)v2c[k];
    const int64_t c = active[ai >> 2];
    const int i = ai & 3;
    uint32_t word = ctags[c] == 2 ? (1u << 24) : 0u;
    int byte = 0;
    for (int j = 0; j < nv; ++j)
      if (j != i) word |= (uint32_t)find_sorted(nb, n, cells[c * nv + j]) << (8 * byte++);
    rec[base + (int64_t)(k - b0) * 32] = word;
  }
}

__global__ void k_listed_flags(const int32_t* __restrict__ row_nnz, const uint8_t* __restrict__ mask, int64_t n,
                               uint8_t* __restrict__ flag) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) flag[r] = row_nnz[r] > 0 && (!mask || mask[r]);
}
struct FlagSet {
  const uint8_t* f;
  __device__ bool operator()(int r) const { return f[r] != 0; }
};

__global__ void k_inverse(const int32_t* __restrict__ rows, int64_t n, int32_t* __restrict__ li_of_row) {
  const int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (li < n) li_of_row[rows[li]] = (int32_t)li;
}

// width of slice s = largest record count among its (up to 32) rows
__global__ void k_slice_width(const int32_t* __restrict__ rows, const int32_t* __restrict__ count_of_row, int64_t n_listed,
                              int64_t n_slices, int32_t* __restrict__ width) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slices) return;
  int w = 0;
  for (int l = 0; l < 32 && s * 32 + l < n_listed; ++l) {
    const int c = count_of_row[rows[s * 32 + l]];
    w = c > w ? c : w;
  }
  width[s] = w;
}
__global__ void k_degree(const int32_t* __restrict__ vptr, int64_t n, int32_t* __restrict__ deg) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) deg[r] = vptr[r + 1] - vptr[r];
}

// Morton key of phifem_b200/mesh.py morton_keys: coordinates quantised to 21 bits per axis over [lo, hi], bits
// interleaved (axis 0 least significant); the same IEEE operations in the same order
__global__ void k_morton(const double* __restrict__ x, int d, const int32_t* __restrict__ rows, int64_t n,
                         const double* __restrict__ lohi, uint64_t* __restrict__ key, int32_t* __restrict__ idx) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int bits = 21;
  uint64_t out = 0;
  for (int k = 0; k < d; ++k) {
    const double lo = lohi[k], hi = lohi[3 + k];
    const double span = fmax(hi - lo, 1e-300);
    const double scaled = __dmul_rn(__ddiv_rn(x[(int64_t)rows[t] * d + k] - lo, span), 2097151.0);
    long long q = (long long)scaled;
    q = q < 0 ? 0 : (q > 2097151 ? 2097151 : q);
    for (int bit = 0; bit < bits; ++bit) out |= (uint64_t)((q >> bit) & 1) << (bit * d + k);
  }
  key[t] = out;
  idx[t] = (int32_t)t;
}
__global__ void k_minmax_init(double* lohi) {
  if (threadIdx.x < 3) {
    lohi[threadIdx.x] = 1e308;
    lohi[3 + threadIdx.x] = -1e308;
  }
}
__device__ __forceinline__ void atomic_min_double(double* a, double v) {
  unsigned long long* p = reinterpret_cast<unsigned long long*>(a);
  unsigned long long old = *p;
  while (__longlong_as_double((long long)old) > v) {
    const unsigned long long seen = atomicCAS(p, old, (unsigned long long)__double_as_longlong(v));
    if (seen == old) break;
    old = seen;
  }
}
__device__ __forceinline__ void atomic_max_double(double* a, double v) {
  unsigned long long* p = reinterpret_cast<unsigned long long*>(a);
  unsigned long long old = *p;
  while (__longlong_as_double((long long)old) < v) {
    const unsigned long long seen = atomicCAS(p, old, (unsigned long long)__double_as_longlong(v));
    if (seen == old) break;
    old = seen;
  }
}
__global__ void k_minmax(const double* __restrict__ x, int d, const int32_t* __restrict__ rows, int64_t n,
                         double* __restrict__ lohi) {
  __shared__ double slo[3], shi[3];
  if (threadIdx.x < 3) {
    slo[threadIdx.x] = 1e308;
    shi[threadIdx.x] = -1e308;
  }
  __syncthreads();
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
    for (int k = 0; k < d; ++k) {
      const double v = x[(int64_t)rows[t] * d + k];
      atomic_min_double(slo + k, v);
      atomic_max_double(shi + k, v);
    }
  __syncthreads();
  if (threadIdx.x < d) {
    atomic_min_double(lohi + threadIdx.x, slo[threadIdx.x]);
    atomic_max_double(lohi + 3 + threadIdx.x, shi[threadIdx.x]);
  }
}
__global__ void k_gather32(const int32_t* __restrict__ src, const int32_t* __restrict__ idx, int64_t n,
                           int32_t* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = src[idx[t]];
}
// balance key of phifem_b200/rows.py: chunk * (max + 1) + (max - count), position p in the Morton order
__global__ void k_balance_keys(const int32_t* __restrict__ count, const int32_t* __restrict__ order, int64_t n,
                               const int32_t* __restrict__ maxc, uint64_t* __restrict__ key, int32_t* __restrict__ idx) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int mx = *maxc;
  key[p] = (uint64_t)(p / kBalanceChunk) * (uint64_t)(mx + 1) + (uint64_t)(mx - count[order[p]]);
  idx[p] = order[p];
}

// surface records: sequence t < ng (nv+1): ghost facet g = t / (nv+1), macro index a; then entity e, facet vertex tt
__global__ void k_surface_keys(const int32_t* __restrict__ macro, const int32_t* __restrict__ emacro, int nv, int64_t ng,
                               int64_t ne, const uint8_t* __restrict__ listed, uint64_t* __restrict__ keys) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int d = nv - 1;
  const int64_t n1 = ng * (nv + 1), n = n1 + ne * d;
  if (t >= n) return;
  int32_t row;
  if (t < n1) row = macro[t];
  else {
    const int64_t u = t - n1;
    row = emacro[(u / d) * nv + (u % d)];
  }
  keys[t] = listed[row] ? (((uint64_t)(uint32_t)row << 32) | (uint64_t)t) : ~0ull;  // dropped records sort last
}
__global__ void k_scatter_count(const int32_t* __restrict__ urows, const int32_t* __restrict__ ucount, int64_t nu,
                                int32_t* __restrict__ count_of_row) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u < nu) count_of_row[urows[u]] = ucount[u];
}
__device__ __forceinline__ uint32_t pos_in_row(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                               int32_t row, int32_t col) {
  int lo = indptr[row], hi = indptr[row + 1];
  const int start = lo;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (indices[mid] < col) lo = mid + 1;
    else hi = mid;
  }
  return (uint32_t)(lo - start);
}
__global__ void k_surface_records(const uint64_t* __restrict__ sorted, int64_t n_valid, const int32_t* __restrict__ ustart,
                                  const int32_t* __restrict__ uidx_of_row, const int32_t* __restrict__ li_of_row,
                                  const int32_t* __restrict__ macro, const int32_t* __restrict__ emacro, int nv, int64_t ng,
                                  const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                  const int32_t* __restrict__ rec_ptr, uint32_t* __restrict__ rec) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_valid) return;
  const int d = nv - 1;
  const int32_t row = (int32_t)(sorted[p] >> 32);
  const int64_t t = (int64_t)(sorted[p] & 0xffffffffull);
  const int k = (int)(p - ustart[uidx_of_row[row]]);
  const int li = li_of_row[row];
  uint32_t w0 = 0, w1;
  const int64_t n1 = ng * (nv + 1);
  if (t < n1) {
    const int64_t g = t / (nv + 1);
    const int a = (int)(t - g * (nv + 1));
    int byte = 0;
    for (int j = 0; j <= nv; ++j)
      if (j != a) w0 |= pos_in_row(indptr, indices, row, macro[g * (nv + 1) + j]) << (8 * byte++);
    w1 = (uint32_t)g | ((uint32_t)a << 28);
  } else {
    const int64_t u = t - n1, e = u / d;
    const int tt = (int)(u - e * d);
    w0 = pos_in_row(indptr, indices, row, emacro[e * nv + nv - 1]);  // opposite vertex first
    for (int q = 1; q < d; ++q) w0 |= pos_in_row(indptr, indices, row, emacro[e * nv + (tt + q) % d]) << (8 * q);
    w1 = (uint32_t)(ng + e) | ((uint32_t)tt << 28) | (1u << 31);
  }
  const int64_t at = ((int64_t)rec_ptr[li >> 5] + k) * 32 + (li & 31);
  rec[2 * at] = w0;
  rec[2 * at + 1] = w1;
}
__global__ void k_surface_diag(const int32_t* __restrict__ rows, int64_t n, const int32_t* __restrict__ indptr,
                               const int32_t* __restrict__ indices, uint8_t* __restrict__ diag_pos) {
  const int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (li < n) diag_pos[li] = (uint8_t)pos_in_row(indptr, indices, rows[li], rows[li]);
}
__global__ void k_set_last(int32_t* ptr, int64_t n, const int32_t* last_width) {  // exclusive scan -> ptr[n] = total
  if (threadIdx.x == 0 && blockIdx.x == 0) ptr[n] = (n ? ptr[n - 1] + last_width[n - 1] : 0);
}

}  // namespace
}  // namespace phifem

using namespace phifem;

extern "C" void phifem_rows_plan_destroy(phifem_rows_plan_handle* h) {
  if (!h) return;
  // kernels reading the arrays may still be in flight on any stream; the blocks are parked for the next plan
  if (h->n_owned) cudaDeviceSynchronize();
  for (int i = 0; i < h->n_owned; ++i) plan_block_give(h->device, h->owned[i], h->owned_bytes[i]);
  delete h;
}

extern "C" int phifem_rows_plan_view(const phifem_rows_plan_handle* h, phifem_rows_plan* plan,
                                     phifem_rows_plan_info* info) {
  PHIFEM_CHECK_ARG(h != nullptr, "handle is null");
  if (plan) *plan = h->plan;
  if (info) *info = h->info;
  return PHIFEM_OK;
}

extern "C" int phifem_rows_plan_create(const phifem_mesh* mesh, const int8_t* cell_tags8, const int8_t* facet_tags8,
                                       const int32_t* entities, int64_t n_entities, const uint8_t* row_mask,
                                       int32_t morton_cells, phifem_rows_plan_handle** out, void* stream) {
  PHIFEM_CHECK_ARG(out != nullptr, "out is null");
  *out = nullptr;
  PHIFEM_CHECK_ARG(mesh != nullptr && mesh->x && mesh->cells && mesh->c2f && mesh->f2c, "mesh / facet connectivity is null");
  if (mesh->cell_type != PHIFEM_TRIANGLE && mesh->cell_type != PHIFEM_TETRAHEDRON) {
    set_error("P1 assembly supports triangles and tetrahedra, got cell type %d", mesh->cell_type);
    return PHIFEM_ERR_UNSUPPORTED;
  }
  PHIFEM_CHECK_ARG(cell_tags8 && facet_tags8, "tag arrays are null");
  PHIFEM_CHECK_ARG(n_entities >= 0 && (n_entities == 0 || entities), "entity list");
  PHIFEM_CHECK_ARG(mesh->n_cells < (1ll << 31) && mesh->n_facets < (1ll << 31) && mesh->n_vertices < (1ll << 31),
                   "int32 index width");
  const int nv = mesh->cell_type == PHIFEM_TRIANGLE ? 3 : 4, nm = nv + 1, d = nv - 1;
  const int64_t n_rows = mesh->n_vertices, ne = n_entities;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemPool_t pool = scratch_pool();
  if (!pool) {
    set_error("phifem_rows_plan_create: cannot create the scratch memory pool");
    return PHIFEM_ERR_CUDA;
  }
  Tmp tmp(st, pool);
  phifem_rows_plan_handle* h = new phifem_rows_plan_handle();
  h->n_owned = 0;
  h->device = 0;
  cudaGetDevice(&h->device);
  auto own = [&](size_t bytes) -> void* {
    if (bytes == 0) bytes = 1;
    if (h->n_owned >= 40) return nullptr;
    void* q = plan_block_take(h->device, bytes);  // (a parked block may be up to 25 % larger; it is given back under
                                                  // the requested size, never a larger one)
    if (!q && cudaMalloc(&q, bytes) != cudaSuccess) return nullptr;
    h->owned_bytes[h->n_owned] = bytes;
    h->owned[h->n_owned++] = q;
    return q;
  };
  auto fail = [&](const char* what, int rc = PHIFEM_ERR_CUDA) {
    set_error("phifem_rows_plan_create: %s: %s", what, cudaGetErrorString(cudaGetLastError()));
    phifem_rows_plan_destroy(h);
    return rc;
  };
  auto sync_ok = [&]() { return cudaStreamSynchronize(st) == cudaSuccess && cudaGetLastError() == cudaSuccess; };
  thrust::counting_iterator<int> count_it(0);

  // ---- 1. active cells, ghost-penalty facets ------------------------------------------------------------------
  int32_t* active_full = tmp.get<int32_t>(mesh->n_cells);
  int32_t* ghost_full = tmp.get<int32_t>(mesh->n_facets);
  int64_t* d_counts = tmp.get<int64_t>(4);
  if (!tmp.ok) return fail("scratch allocation");
  {
    size_t b1 = 0, b2 = 0;
    cub::DeviceSelect::If(nullptr, b1, count_it, active_full, d_counts, (int)mesh->n_cells, IsActive{cell_tags8}, st);
    cub::DeviceSelect::If(nullptr, b2, count_it, ghost_full, d_counts + 1, (int)mesh->n_facets,
                          IsGhost{facet_tags8, mesh->f2c}, st);
    void* ws = tmp.get<char>((int64_t)(b1 > b2 ? b1 : b2));
    if (!tmp.ok) return fail("scratch allocation");
    cub::DeviceSelect::If(ws, b1, count_it, active_full, d_counts, (int)mesh->n_cells, IsActive{cell_tags8}, st);
    cub::DeviceSelect::If(ws, b2, count_it, ghost_full, d_counts + 1, (int)mesh->n_facets,
                          IsGhost{facet_tags8, mesh->f2c}, st);
  }
  int64_t h_counts[2] = {0, 0};
  cudaMemcpyAsync(h_counts, d_counts, sizeof(h_counts), cudaMemcpyDeviceToHost, st);
  if (!sync_ok()) return fail("compaction");
  const int64_t na = h_counts[0], ng = h_counts[1];
  if (na >= (1ll << 29) || ng + ne >= (1ll << 28)) {
    set_error("phifem_rows_plan_create: %lld active cells / %lld surface entities exceed the record format",
              (long long)na, (long long)(ng + ne));
    phifem_rows_plan_destroy(h);
    return PHIFEM_ERR_UNSUPPORTED;
  }
  int32_t* active = (int32_t*)own(sizeof(int32_t) * na);
  int32_t* ghost = (int32_t*)own(sizeof(int32_t) * ng);
  int32_t* macro = (int32_t*)own(sizeof(int32_t) * ng * nm);
  int32_t* emacro = (int32_t*)own(sizeof(int32_t) * ne * nv);
  double* work = (double*)own(sizeof(double) * 8 * (ng + ne > 0 ? ng + ne : 1));
  double* work_static = (double*)own(sizeof(double) * 8 * (ng + ne > 0 ? ng + ne : 1));
  if (!active || !ghost || !macro || !emacro || !work || !work_static) return fail("output allocation");
  cudaMemcpyAsync(active, active_full, sizeof(int32_t) * na, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(ghost, ghost_full, sizeof(int32_t) * ng, cudaMemcpyDeviceToDevice, st);
  if (ng) k_macro<<<nblk(ng), kB, 0, st>>>(*mesh, nv, ghost, ng, macro);
  if (ne) k_entity_macro<<<nblk(ne), kB, 0, st>>>(mesh->cells, nv, entities, ne, emacro);

  // ---- 2. vertex -> (active cell, local index) lists, cells ascending; opposite-vertex couplings of the ghost facets
  const int64_t nk = na * nv, nx = 2 * ng;
  int vbits = 1;
  while (vbits < 32 && (1ll << vbits) < n_rows) ++vbits;
  uint64_t* keys = tmp.get<uint64_t>(nk);
  uint64_t* v2c = tmp.get<uint64_t>(nk);
  uint64_t* xkeys = tmp.get<uint64_t>(nx);
  uint64_t* extra = tmp.get<uint64_t>(nx);
  int32_t* vptr = tmp.get<int32_t>(n_rows + 1);
  int32_t* xptr = tmp.get<int32_t>(n_rows + 1);
  if (!tmp.ok) return fail("scratch allocation");
  {
    size_t b1 = 0, b2 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, b1, keys, v2c, (int)nk, 0, 32 + vbits, st);
    cub::DeviceRadixSort::SortKeys(nullptr, b2, xkeys, extra, (int)nx, 0, 32 + vbits, st);
    void* ws = tmp.get<char>((int64_t)(b1 > b2 ? b1 : b2));
    if (!tmp.ok) return fail("scratch allocation");
    if (nk) {
      k_v2c_keys<<<nblk(nk), kB, 0, st>>>(mesh->cells, nv, active, na, keys);
      cub::DeviceRadixSort::SortKeys(ws, b1, keys, v2c, (int)nk, 0, 32 + vbits, st);
    }
    if (nx) {
      k_extra_keys<<<nblk(ng), kB, 0, st>>>(macro, nv, ng, xkeys);
      cub::DeviceRadixSort::SortKeys(ws, b2, xkeys, extra, (int)nx, 0, 32 + vbits, st);
    }
  }
  k_starts<<<nblk(n_rows + 1), kB, 0, st>>>(v2c, nk, n_rows, vptr);
  k_starts<<<nblk(n_rows + 1), kB, 0, st>>>(extra, nx, n_rows, xptr);

  // ---- 3. pattern: pass 1 counts, scan, pass 2 writes columns + cell records --------------------------------------
  int32_t* row_nnz = tmp.get<int32_t>(n_rows + 1);
  int32_t* indptr = (int32_t*)own(sizeof(int32_t) * (n_rows + 1));
  int* d_overflow = tmp.get<int>(1);
  int32_t* d_max = tmp.get<int32_t>(2);
  if (!tmp.ok || !indptr) return fail("allocation");
  cudaMemsetAsync(d_overflow, 0, sizeof(int), st);
  cudaMemsetAsync(row_nnz + n_rows, 0, sizeof(int32_t), st);
  k_row_pattern<0><<<(unsigned)((n_rows + 127) / 128), 128, 0, st>>>(
      mesh->cells, nv, cell_tags8, active, v2c, vptr, extra, xptr, n_rows, row_nnz, d_overflow, nullptr, nullptr,
      nullptr, nullptr, nullptr, nullptr);
  {
    size_t b1 = 0, b2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b1, row_nnz, indptr, (int)(n_rows + 1), st);
    cub::DeviceReduce::Max(nullptr, b2, row_nnz, d_max, (int)n_rows, st);
    void* ws = tmp.get<char>((int64_t)(b1 > b2 ? b1 : b2));
    if (!tmp.ok) return fail("scratch allocation");
    cub::DeviceScan::ExclusiveSum(ws, b1, row_nnz, indptr, (int)(n_rows + 1), st);
    cub::DeviceReduce::Max(ws, b2, row_nnz, d_max, (int)n_rows, st);
  }
  // listed rows of the cell pass: rows with pattern entries (inside the mask), ascending
  uint8_t* listed = tmp.get<uint8_t>(n_rows);
  int32_t* rows_full = tmp.get<int32_t>(n_rows);
  if (!tmp.ok) return fail("scratch allocation");
  k_listed_flags<<<nblk(n_rows), kB, 0, st>>>(row_nnz, row_mask, n_rows, listed);
  {
    size_t b1 = 0;
    cub::DeviceSelect::If(nullptr, b1, count_it, rows_full, d_counts + 2, (int)n_rows, FlagSet{listed}, st);
    void* ws = tmp.get<char>((int64_t)b1);
    if (!tmp.ok) return fail("scratch allocation");
    cub::DeviceSelect::If(ws, b1, count_it, rows_full, d_counts + 2, (int)n_rows, FlagSet{listed}, st);
  }
  int32_t h_nnz = 0, h_max = 0;
  int h_overflow = 0;
  int64_t n_listed = 0;
  cudaMemcpyAsync(&h_nnz, indptr + n_rows, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(&h_max, d_max, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(&h_overflow, d_overflow, sizeof(int), cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(&n_listed, d_counts + 2, sizeof(int64_t), cudaMemcpyDeviceToHost, st);
  if (!sync_ok()) return fail("pattern pass 1");
  if (h_overflow) {
    set_error("phifem_rows_plan_create: a row holds more than %d entries; use the atomic scatter kernels", kMaxNnz);
    phifem_rows_plan_destroy(h);
    return PHIFEM_ERR_UNSUPPORTED;
  }
  const int64_t nnz = h_nnz;
  double* lohi = tmp.get<double>(6);
  uint64_t* mkey = tmp.get<uint64_t>(n_listed);
  uint64_t* mkey2 = tmp.get<uint64_t>(n_listed);
  int32_t* midx = tmp.get<int32_t>(n_listed);
  int32_t* midx2 = tmp.get<int32_t>(n_listed);
  int32_t* cell_rows = (int32_t*)own(sizeof(int32_t) * n_listed);
  if (!tmp.ok || !cell_rows) return fail("allocation");
  auto morton_sort = [&](const int32_t* rows_in, int64_t n, int32_t* order_out) -> bool {
    // order_out[p] = index into rows_in of the p-th row along the Morton curve (stable)
    if (n == 0) return true;
    k_minmax_init<<<1, 32, 0, st>>>(lohi);
    k_minmax<<<(unsigned)(n < 148 * 8 * kB ? nblk(n) : 148 * 8), kB, 0, st>>>(mesh->x, d, rows_in, n, lohi);
    k_morton<<<nblk(n), kB, 0, st>>>(mesh->x, d, rows_in, n, lohi, mkey, midx);
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b, mkey, mkey2, midx, order_out, (int)n, 0, 21 * d, st);
    void* ws = tmp.get<char>((int64_t)b);
    if (!tmp.ok) return false;
    cub::DeviceRadixSort::SortPairs(ws, b, mkey, mkey2, midx, order_out, (int)n, 0, 21 * d, st);
    return true;
  };
  if (morton_cells && n_listed) {
    if (!morton_sort(rows_full, n_listed, midx2)) return fail("scratch allocation");
    k_gather32<<<nblk(n_listed), kB, 0, st>>>(rows_full, midx2, n_listed, cell_rows);
  } else {
    cudaMemcpyAsync(cell_rows, rows_full, sizeof(int32_t) * n_listed, cudaMemcpyDeviceToDevice, st);
  }
  int32_t* li_of_row = tmp.get<int32_t>(n_rows);
  int32_t* deg = tmp.get<int32_t>(n_rows);
  const int64_t ns_c = (n_listed + 31) / 32;
  int32_t* width_c = tmp.get<int32_t>(ns_c + 1);
  int32_t* ptr_c = (int32_t*)own(sizeof(int32_t) * (ns_c + 1));
  uint8_t* diag_c = (uint8_t*)own(n_listed);
  int32_t* indices = (int32_t*)own(sizeof(int32_t) * nnz);
  if (!tmp.ok || !ptr_c || !diag_c || !indices) return fail("allocation");
  cudaMemsetAsync(li_of_row, 0xff, sizeof(int32_t) * n_rows, st);
  if (n_listed) k_inverse<<<nblk(n_listed), kB, 0, st>>>(cell_rows, n_listed, li_of_row);
  k_degree<<<nblk(n_rows), kB, 0, st>>>(vptr, n_rows, deg);
  if (ns_c) k_slice_width<<<nblk(ns_c), kB, 0, st>>>(cell_rows, deg, n_listed, ns_c, width_c);
  {
    size_t b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b, width_c, ptr_c, (int)ns_c, st);
    void* ws = tmp.get<char>((int64_t)b);
    if (!tmp.ok) return fail("scratch allocation");
    if (ns_c) cub::DeviceScan::ExclusiveSum(ws, b, width_c, ptr_c, (int)ns_c, st);
    k_set_last<<<1, 1, 0, st>>>(ptr_c, ns_c, width_c);
  }
  int32_t h_total_c = 0;
  cudaMemcpyAsync(&h_total_c, ptr_c + ns_c, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  if (!sync_ok()) return fail("cell list layout");
  const int64_t total_c = h_total_c;
  if (total_c * 32 >= (1ll << 31)) return fail("record array exceeds 2^31 words", PHIFEM_ERR_UNSUPPORTED);
  uint32_t* rec_c = (uint32_t*)own(sizeof(uint32_t) * 32 * (total_c > 0 ? total_c : 1));
  if (!rec_c) return fail("output allocation");
  cudaMemsetAsync(rec_c, 0xff, sizeof(uint32_t) * 32 * (total_c > 0 ? total_c : 1), st);
  k_row_pattern<1><<<(unsigned)((n_rows + 127) / 128), 128, 0, st>>>(
      mesh->cells, nv, cell_tags8, active, v2c, vptr, extra, xptr, n_rows, row_nnz, d_overflow, indptr, indices,
      li_of_row, diag_c, ptr_c, rec_c);

  // ---- 4. surface list --------------------------------------------------------------------------------------------
  const int64_t n_srec_all = ng * nm + ne * d;
  uint64_t* skeys = tmp.get<uint64_t>(n_srec_all);
  uint64_t* ssorted = tmp.get<uint64_t>(n_srec_all);
  int32_t* urows = tmp.get<int32_t>(n_srec_all);
  int32_t* ucount = tmp.get<int32_t>(n_srec_all + 1);
  int32_t* ustart = tmp.get<int32_t>(n_srec_all + 1);
  if (!tmp.ok) return fail("scratch allocation");
  int64_t n_valid = 0, nu = 0;
  if (n_srec_all) {
    k_surface_keys<<<nblk(n_srec_all), kB, 0, st>>>(macro, emacro, nv, ng, ne, listed, skeys);
    size_t b1 = 0, b2 = 0, b3 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, b1, skeys, ssorted, (int)n_srec_all, 0, 64, st);
    thrust::transform_iterator<HighWord, const uint64_t*> hi_it(ssorted, HighWord());
    cub::DeviceRunLengthEncode::Encode(nullptr, b2, hi_it, urows, ucount, d_counts + 3, (int)n_srec_all, st);
    cub::DeviceScan::ExclusiveSum(nullptr, b3, ucount, ustart, (int)n_srec_all, st);
    size_t b = b1 > b2 ? b1 : b2;
    b = b > b3 ? b : b3;
    void* ws = tmp.get<char>((int64_t)b);
    if (!tmp.ok) return fail("scratch allocation");
    cub::DeviceRadixSort::SortKeys(ws, b1, skeys, ssorted, (int)n_srec_all, 0, 64, st);
    cub::DeviceRunLengthEncode::Encode(ws, b2, hi_it, urows, ucount, d_counts + 3, (int)n_srec_all, st);
    cudaMemcpyAsync(&nu, d_counts + 3, sizeof(int64_t), cudaMemcpyDeviceToHost, st);
    if (!sync_ok()) return fail("surface record sort");
    // the last run holds the dropped records (row word 0xffffffff) when there are any
    int32_t last_row = 0, last_count = 0;
    if (nu) {
      cudaMemcpyAsync(&last_row, urows + nu - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
      cudaMemcpyAsync(&last_count, ucount + nu - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
      if (!sync_ok()) return fail("surface record sort");
    }
    n_valid = n_srec_all;
    if (nu && last_row == -1) {
      --nu;
      n_valid -= last_count;
    }
    if (nu) cub::DeviceScan::ExclusiveSum(ws, b3, ucount, ustart, (int)nu, st);
  }
  const int64_t ns_s = (nu + 31) / 32;
  int32_t* surf_rows = (int32_t*)own(sizeof(int32_t) * nu);
  uint8_t* diag_s = (uint8_t*)own(nu);
  int32_t* ptr_s = (int32_t*)own(sizeof(int32_t) * (ns_s + 1));
  int32_t* order1 = tmp.get<int32_t>(nu);
  int32_t* order2 = tmp.get<int32_t>(nu);
  int32_t* count_of_row = tmp.get<int32_t>(n_rows);
  int32_t* uidx_of_row = tmp.get<int32_t>(n_rows);
  int32_t* sli_of_row = tmp.get<int32_t>(n_rows);
  int32_t* width_s = tmp.get<int32_t>(ns_s + 1);
  if (!tmp.ok || !surf_rows || !diag_s || !ptr_s) return fail("allocation");
  int64_t total_s = 0;
  uint32_t* rec_s = nullptr;
  if (nu) {
    // Morton order of the rows, then stable re-sort by record count (descending) inside chunks of 4096 rows
    uint64_t* bkey = tmp.get<uint64_t>(nu);
    uint64_t* bkey2 = tmp.get<uint64_t>(nu);
    int32_t* bidx = tmp.get<int32_t>(nu);
    if (!tmp.ok) return fail("scratch allocation");
    if (nu > n_listed) return fail("internal: more surface rows than listed rows", PHIFEM_ERR_ARGUMENT);
    if (!morton_sort(urows, nu, order1)) return fail("scratch allocation");
    {
      size_t b1 = 0, b2 = 0;
      cub::DeviceReduce::Max(nullptr, b1, ucount, d_max + 1, (int)nu, st);
      cub::DeviceRadixSort::SortPairs(nullptr, b2, bkey, bkey2, bidx, order2, (int)nu, 0, 64, st);
      void* ws = tmp.get<char>((int64_t)(b1 > b2 ? b1 : b2));
      if (!tmp.ok) return fail("scratch allocation");
      cub::DeviceReduce::Max(ws, b1, ucount, d_max + 1, (int)nu, st);
      k_balance_keys<<<nblk(nu), kB, 0, st>>>(ucount, order1, nu, d_max + 1, bkey, bidx);
      cub::DeviceRadixSort::SortPairs(ws, b2, bkey, bkey2, bidx, order2, (int)nu, 0, 64, st);
    }
    k_gather32<<<nblk(nu), kB, 0, st>>>(urows, order2, nu, surf_rows);
    cudaMemsetAsync(sli_of_row, 0xff, sizeof(int32_t) * n_rows, st);
    k_inverse<<<nblk(nu), kB, 0, st>>>(surf_rows, nu, sli_of_row);
    k_inverse<<<nblk(nu), kB, 0, st>>>(urows, nu, uidx_of_row);
    k_scatter_count<<<nblk(nu), kB, 0, st>>>(urows, ucount, nu, count_of_row);
    k_slice_width<<<nblk(ns_s), kB, 0, st>>>(surf_rows, count_of_row, nu, ns_s, width_s);
    {
      size_t b = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, b, width_s, ptr_s, (int)ns_s, st);
      void* ws = tmp.get<char>((int64_t)b);
      if (!tmp.ok) return fail("scratch allocation");
      cub::DeviceScan::ExclusiveSum(ws, b, width_s, ptr_s, (int)ns_s, st);
      k_set_last<<<1, 1, 0, st>>>(ptr_s, ns_s, width_s);
    }
    int32_t h_total_s = 0;
    cudaMemcpyAsync(&h_total_s, ptr_s + ns_s, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    if (!sync_ok()) return fail("surface list layout");
    total_s = h_total_s;
  } else {
    cudaMemsetAsync(ptr_s, 0, sizeof(int32_t), st);
  }
  if (total_s * 64 >= (1ll << 31)) return fail("record array exceeds 2^31 words", PHIFEM_ERR_UNSUPPORTED);
  rec_s = (uint32_t*)own(sizeof(uint32_t) * 64 * (total_s > 0 ? total_s : 1));
  if (!rec_s) return fail("output allocation");
  cudaMemsetAsync(rec_s, 0xff, sizeof(uint32_t) * 64 * (total_s > 0 ? total_s : 1), st);
  if (nu) {
    k_surface_records<<<nblk(n_valid), kB, 0, st>>>(ssorted, n_valid, ustart, uidx_of_row, sli_of_row, macro, emacro, nv, ng,
                                                   indptr, indices, ptr_s, rec_s);
    k_surface_diag<<<nblk(nu), kB, 0, st>>>(surf_rows, nu, indptr, indices, diag_s);
  }
  if (!sync_ok()) return fail("plan kernels");

  phifem_rows_plan& p = h->plan;
  p.indptr = indptr;
  p.indices = indices;
  p.max_row_nnz = h_max;
  p.reserved = 0;
  p.cells = phifem_row_list{n_listed, cell_rows, diag_c, ptr_c, rec_c};
  p.surface = phifem_row_list{nu, surf_rows, diag_s, ptr_s, rec_s};
  p.n_ghost_facets = ng;
  p.ghost_macro = macro;
  p.n_entities = ne;
  p.entity_macro = emacro;
  p.surface_work = work;
  p.cell_geom = nullptr;
  p.tiles = nullptr;
  p.surface_static = work_static;
  if (int rc = phifem_surface_static_p1(mesh, &p, stream)) {  // mesh-only part of the facet-once records, once per plan
    phifem_rows_plan_destroy(h);
    return rc;
  }
  phifem_rows_plan_info& info = h->info;
  info.n_rows = n_rows;
  info.nnz = nnz;
  info.n_active = na;
  info.n_ghost = ng;
  info.n_entities = ne;
  info.active = active;
  info.ghost = ghost;
  info.n_cell_records = nk;        // before the row mask; the listed rows carry their own
  info.n_surface_records = n_valid;
  info.cells_record_slots = total_c * 32;
  info.surface_record_slots = total_s * 32;
  *out = h;
  return PHIFEM_OK;
}
