// K2/K3/K4: strong-Dirichlet phi-FEM operator, P1 on triangles / tetrahedra, assembled into CSR.
//
// Forms of reference demo/strong-dirichlet/flower/main.py:104-128; closed-form element tensors of
// SURVEY.md Appendix B (exact for P1 phi, P1 w/v, P1 f), so the kernels stay near the fp64/HBM
// ridge instead of looping over quadrature points.  What dolfinx does per entity with one FFCx
// `tabulate_tensor` call plus a `MatSetValuesLocal(ADD_VALUES)` binary search is one thread here.
//
// Two scatter strategies share the same element math (the `emit` functors below):
//   * atomic   -- one thread per entity, fp64 reductions (REDG.E.ADD.F64) through a precomputed
//                 entity -> CSR-slot map.  Simple, but ~8 reductions land on every CSR entry and the
//                 SM issues < 1 reduction lane per clock: 410 M reductions = 2.5 ms at 50.9 M tets.
//   * blocked  -- owner-computes.  Rows are grouped into spatially compact blocks; one CTA per block
//                 evaluates every entity touching its rows (halo entities are recomputed by each
//                 block they touch), drops each contribution into a shared-memory buffer at a position
//                 precomputed so that contributions to the same CSR entry are adjacent, then one
//                 thread per CSR entry sums its segment and stores it.  No atomics, no zero-fill, plain
//                 coalesced row stores, bitwise reproducible.
#include <cuda_pipeline.h>

#include "common.cuh"

namespace phifem {
namespace {

constexpr int kBlock = 128;

template <int D>
__device__ __forceinline__ void load_xc(const phifem_mesh& m, const int (&v)[D + 1], double (&xc)[D + 1][D]) {
#pragma unroll
  for (int k = 0; k <= D; ++k)
#pragma unroll
    for (int d = 0; d < D; ++d) xc[k][d] = __ldg(m.x + (int64_t)v[k] * D + d);
}

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
  load_xc<D>(m, v, xc);
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

// ---- element tensors; `emit.mat(i, j, value)` / `emit.vec(i, value)` receive the entries -------------

// dx((1,2)) stiffness + dx(2) stabilisation (main.py:105,107-112), load vector (:126-128)
template <int D, typename Emit>
__device__ __forceinline__ void cell_tensor(const double (&xc)[D + 1][D], const double (&p)[D + 1],
                                            const double (&fv)[D + 1], bool is_cut, double sigma,
                                            Emit& emit) {
  constexpr int NV = D + 1;
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
  if (is_cut) stab = sigma * diameter2<D>(xc) * vol;  // sigma h_T^2 |K| on cut cells
  const double ggM = gg * cM, stab4 = 4.0 * stab;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int j = i; j < NV; ++j) {
      const double val = ggM * (i == j ? 2.0 : 1.0) + a[i] * mm[j] + mm[i] * a[j] +
                         dot<D>(G[i], G[j]) * mu + stab4 * a[i] * a[j];
      emit.mat(i, j, val);
      if (j != i) emit.mat(j, i, val);
    }
  }
  constexpr double fact_d = D == 2 ? 2.0 : 6.0, fact_d3 = D == 2 ? 120.0 : 720.0;
  const double c3 = vol * (fact_d / fact_d3);
  const double fmean2 = 2.0 * stab * F * (1.0 / NV);
#pragma unroll
  for (int i = 0; i < NV; ++i)
    emit.vec(i, c3 * ((F * P + FP) + fv[i] * P + F * p[i] + 2.0 * fv[i] * p[i]) - fmean2 * a[i]);
}

__device__ __forceinline__ double alpha3(int a, int b, int c) {
  return (double)((1 + (a == b)) * (1 + (a == c) + (b == c)));
}

// -int_F (grad(phi w).n) phi v on local facet o of a cell (main.py:106), row = test
template <int D, typename Emit>
__device__ __forceinline__ void boundary_tensor(const double (&xc)[D + 1][D], const double (&p)[D + 1],
                                                int o, Emit& emit) {
  constexpr int NV = D + 1;
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
      emit.mat(i, j, (i == o) ? 0.0 : -cF * acc);
    }
  }
}

// sigma avg(h_T) int_F jump.jump (main.py:113-118).  The macro element [cell+ vertices, cell- vertices]
// has 2(D+1) dofs but only D+2 distinct vertices (the D facet vertices appear in both halves); the jump
// integrand is linear in the basis function, so the two halves of a shared vertex are summed first and
// (D+2)^2 entries are emitted instead of (2D+2)^2 -- 25 instead of 64 scatter operations per tetrahedron
// facet.  Vertex order of the emitted tensor: facet vertices as ordered in cell +, opposite vertex of
// cell +, opposite vertex of cell -.
template <int D, typename Emit>
__device__ __forceinline__ void ghost_tensor(const phifem_mesh& m, const double* __restrict__ phi,
                                             int32_t fct, double sigma, Emit& emit) {
  constexpr int NV = D + 1, NG = D + 2;
  const int2 cc = __ldg(reinterpret_cast<const int2*>(m.f2c) + fct);
  int fvert[D];            // global vertices of the facet, ordered as in cell +
  double pf[D];            // phi at the facet vertices
  double Jg[NG][D];        // jump integrand of each distinct vertex at facet vertex k (affine on the facet)
#pragma unroll
  for (int a = 0; a < NG; ++a)
#pragma unroll
    for (int k = 0; k < D; ++k) Jg[a][k] = 0.0;
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
            if (t == q) {
              fvert[t] = v[k];
              pf[t] = p[k];
            }
          ++q;
        }
    }
#pragma unroll
    for (int a = 0; a < NV; ++a) {
      const double Gna = dot<D>(G[a], n);
      // destination: the facet vertex this dof sits on, or this side's opposite vertex
      int t = D + side;
#pragma unroll
      for (int k = 0; k < D; ++k)
        if (a != o && v[a] == fvert[k]) t = k;
#pragma unroll
      for (int tt = 0; tt < NG; ++tt)
        if (tt == t)
#pragma unroll
          for (int k = 0; k < D; ++k) Jg[tt][k] += (tt == k && t < D ? gn : 0.0) + Gna * pf[k];
    }
  }
  const double coef = sigma * 0.5 * hsum * area * (1.0 / (D * (D + 1)));
  double Js[NG];
#pragma unroll
  for (int a = 0; a < NG; ++a) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) s += Jg[a][k];
    Js[a] = s;
  }
#pragma unroll
  for (int a = 0; a < NG; ++a) {
#pragma unroll
    for (int bb = a; bb < NG; ++bb) {
      double s = Js[a] * Js[bb];
#pragma unroll
      for (int k = 0; k < D; ++k) s += Jg[a][k] * Jg[bb][k];
      const double val = coef * s;
      emit.mat(a, bb, val);
      if (bb != a) emit.mat(bb, a, val);
    }
  }
}

// ---- scatter strategy 1: fp64 reductions through an entity -> CSR-slot map -------------------------------
template <int N>
struct AtomicEmit {
  double* data;
  double* b;
  const int* verts;
  int sl[N * N];  // CSR slots of this entity, fetched before the arithmetic starts (independent loads)
  __device__ __forceinline__ void load_slots(const int32_t* __restrict__ slots) {
    if (N * N % 4 == 0) {
#pragma unroll
      for (int q = 0; q < N * N / 4; ++q) {
        const int4 s4 = __ldg(reinterpret_cast<const int4*>(slots) + q);
        sl[4 * q] = s4.x; sl[4 * q + 1] = s4.y; sl[4 * q + 2] = s4.z; sl[4 * q + 3] = s4.w;
      }
    } else {
#pragma unroll
      for (int q = 0; q < N * N; ++q) sl[q] = __ldg(slots + q);
    }
  }
  __device__ __forceinline__ void mat(int i, int j, double v) const { atomicAdd(data + sl[i * N + j], v); }
  __device__ __forceinline__ void vec(int i, double v) const { atomicAdd(b + verts[i], v); }
};

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
  AtomicEmit<NV> emit{data, b, v};
  emit.load_slots(slots + e * NV * NV);
  cell_tensor<D>(xc, p, fv, ctags[c] == 2, sigma, emit);
}

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
  AtomicEmit<NV> emit{data, nullptr, v};
  emit.load_slots(slots + e * NV * NV);
  boundary_tensor<D>(xc, p, o, emit);
}

template <int D>
__global__ void __launch_bounds__(kBlock) k_assemble_ghost_p1(
    phifem_mesh m, const double* __restrict__ phi, const int32_t* __restrict__ facets, int64_t n_facets,
    const int32_t* __restrict__ slots, double sigma, double* __restrict__ data) {
  constexpr int NM = D + 2;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_facets) return;
  AtomicEmit<NM> emit{data, nullptr, nullptr};
  emit.load_slots(slots + e * NM * NM);
  ghost_tensor<D>(m, phi, __ldg(facets + e), sigma, emit);
}

// ---- scatter strategy 2: owner-computes row blocks, shared-memory segmented reduction -----------------------
constexpr int kBlockedThreads = 512;   // one CTA per SM with the full shared-memory buffer
constexpr int kBlockedThreadsHalf = 256;  // two co-resident CTAs per SM with half-size buffers

// positions are int16 pairs packed in 32-bit words, column-major over the instances of a kind:
// word q of instance i at pos[q * n_instances + i]; entry e sits in word e/2, half e%2; < 0 = not ours
template <int NWORDS>
struct BufferEmit {
  double* buf;
  int nmat;      // entries per row of the local tensor (mat(i,j) -> entry i*nmat+j)
  int vec_base;  // first entry of the vector part
  unsigned int w[NWORDS];
  __device__ __forceinline__ void put(int e, double v) const {
    const int pos = (int)(short)(w[e >> 1] >> (16 * (e & 1)));
    if (pos >= 0) buf[pos] = v;
  }
  __device__ __forceinline__ void mat(int i, int j, double v) const { put(i * nmat + j, v); }
  __device__ __forceinline__ void vec(int i, double v) const { put(vec_base + i, v); }
};

struct BlockDesc {
  int n_contrib, s_begin, n_seg_pad, c_begin, c_end, g_begin, g_end, b_begin, b_end;
};
__device__ __forceinline__ BlockDesc load_desc(const phifem_blocked_plan& pl, int blk) {
  const int4* p = reinterpret_cast<const int4*>(pl.block_desc + (int64_t)blk * PHIFEM_BLOCK_DESC_INTS);
  const int4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
  return BlockDesc{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x};
}

template <int CW>
struct CellRecord {
  unsigned int w[CW];
  int4 v;
};
template <int CW>
__device__ __forceinline__ void load_record(const phifem_blocked_plan& pl, int i, CellRecord<CW>& r) {
#pragma unroll
  for (int q = 0; q < CW; ++q)
    r.w[q] = __ldg(reinterpret_cast<const unsigned int*>(pl.cell_pos) + (int64_t)q * pl.n_cell_inst + i);
  r.v = __ldg(reinterpret_cast<const int4*>(pl.cell_verts) + i);
}

// Latency plan (16 warps per SM is all the register file allows for the fp64 element math):
//  * the instance record a thread needs next -- also across the block boundary -- is loaded into
//    registers before the current instance is evaluated, so only the L2-resident vertex gathers sit
//    on the critical path of phase 1;
//  * the segment tables of the block are copied global -> shared with cp.async (LDGSTS) while
//    phase 1 runs, so phase 2 touches shared memory only, with 4 loads in flight per thread.
template <int D, int THREADS>
__global__ void __launch_bounds__(THREADS, kBlockedThreads / THREADS) k_assemble_blocked_p1(
    phifem_mesh m, const double* __restrict__ phi, const double* __restrict__ f, double sigma,
    phifem_blocked_plan pl, double* __restrict__ data, double* __restrict__ b) {
  constexpr int NV = D + 1, NM = D + 2;  // NM: distinct vertices of a facet macro element
  constexpr int CW = (NV * NV + NV + 1) / 2, GW = (NM * NM + 1) / 2, BW = (NV * NV + 1) / 2;
  extern __shared__ __align__(16) double buf[];
  int32_t* s_dest = reinterpret_cast<int32_t*>(buf + ((pl.capacity + 1) & ~1));
  int16_t* s_start = reinterpret_cast<int16_t*>(s_dest + pl.max_segments);
  const int tid = threadIdx.x;
  int blk = blockIdx.x;
  if (blk >= pl.n_blocks) return;
  BlockDesc d = load_desc(pl, blk);
  CellRecord<CW> cur;
  int have = -1;  // instance whose record sits in `cur`
  if (d.c_begin + tid < d.c_end) {
    have = d.c_begin + tid;
    load_record<CW>(pl, have, cur);
  }
  for (; blk < pl.n_blocks; blk += gridDim.x) {
    const bool more = blk + (int)gridDim.x < pl.n_blocks;
    BlockDesc nd = d;
    if (more) nd = load_desc(pl, blk + gridDim.x);
    // stage this block's segment tables (16-byte chunks; ranges are padded to 8 entries by the plan)
    {
      const int4* g_start = reinterpret_cast<const int4*>(pl.seg_start + d.s_begin);
      const int4* g_dest = reinterpret_cast<const int4*>(pl.seg_dest + d.s_begin);
      for (int q = tid; q < d.n_seg_pad / 8; q += THREADS)
        __pipeline_memcpy_async(reinterpret_cast<int4*>(s_start) + q, g_start + q, 16);
      for (int q = tid; q < d.n_seg_pad / 4; q += THREADS)
        __pipeline_memcpy_async(reinterpret_cast<int4*>(s_dest) + q, g_dest + q, 16);
      __pipeline_commit();
    }
    // phase 1: evaluate every entity touching the block's rows, park the entries we own in shared memory
    for (int i = d.c_begin + tid; i < d.c_end;) {
      BufferEmit<CW> emit{buf, NV, NV * NV};
#pragma unroll
      for (int q = 0; q < CW; ++q) emit.w[q] = cur.w[q];
      const int4 vq = cur.v;
      const int ni = i + THREADS;
      if (ni < d.c_end) {
        have = ni;
        load_record<CW>(pl, ni, cur);
      } else if (more && nd.c_begin + tid < nd.c_end) {
        have = nd.c_begin + tid;
        load_record<CW>(pl, have, cur);
      }
      const bool is_cut = vq.x < 0;  // tag-2 flag rides in the sign bit of the first vertex id
      int v[NV];
      v[0] = vq.x & 0x7fffffff; v[1] = vq.y; v[2] = vq.z;
      if (NV == 4) v[NV - 1] = vq.w;
      double xc[NV][D], p[NV], fv[NV];
      load_xc<D>(m, v, xc);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        p[k] = __ldg(phi + v[k]);
        fv[k] = __ldg(f + v[k]);
      }
      cell_tensor<D>(xc, p, fv, is_cut, sigma, emit);
      i = ni;
    }
    if (more && nd.c_begin + tid < nd.c_end && have != nd.c_begin + tid) {
      have = nd.c_begin + tid;  // threads idle in this block still prefetch for the next one
      load_record<CW>(pl, have, cur);
    }
    for (int i = d.g_begin + tid; i < d.g_end; i += THREADS) {
      BufferEmit<GW> emit{buf, NM, 0};
#pragma unroll
      for (int q = 0; q < GW; ++q)
        emit.w[q] = __ldg(reinterpret_cast<const unsigned int*>(pl.ghost_pos) + (int64_t)q * pl.n_ghost_inst + i);
      ghost_tensor<D>(m, phi, __ldg(pl.ghost_facet + i), sigma, emit);
    }
    for (int i = d.b_begin + tid; i < d.b_end; i += THREADS) {
      BufferEmit<BW> emit{buf, NV, 0};
#pragma unroll
      for (int q = 0; q < BW; ++q)
        emit.w[q] = __ldg(reinterpret_cast<const unsigned int*>(pl.bnd_pos) + (int64_t)q * pl.n_bnd_inst + i);
      const int64_t c = __ldg(pl.bnd_entity + 2 * (int64_t)i);
      const int o = __ldg(pl.bnd_entity + 2 * (int64_t)i + 1);
      int v[NV];
      double xc[NV][D], p[NV];
      load_vertices<D>(m, c, v, xc);
#pragma unroll
      for (int k = 0; k < NV; ++k) p[k] = __ldg(phi + v[k]);
      boundary_tensor<D>(xc, p, o, emit);
    }
    __pipeline_wait_prior(0);
    __syncthreads();
    // phase 2: one thread per CSR entry / load-vector row sums its segment in a fixed order and stores it
    // (pad entries carry start == n_contrib, so they are empty and the last real segment ends there)
    for (int s = tid; s < d.n_seg_pad - 1; s += THREADS) {
      const int lo = s_start[s], hi = s_start[s + 1];
      if (hi <= lo) continue;
      double acc = 0.0;
      int k = lo;
      for (; k + 4 <= hi; k += 4) {
        const double v0 = buf[k], v1 = buf[k + 1], v2 = buf[k + 2], v3 = buf[k + 3];
        acc = (((acc + v0) + v1) + v2) + v3;
      }
      for (; k < hi; ++k) acc += buf[k];
      const int32_t dst = s_dest[s];
      if (dst >= 0) data[dst] = acc;
      else b[dst & 0x7fffffff] = acc;
    }
    __syncthreads();
    d = nd;
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
  if (n_active == 0) return PHIFEM_OK;  // nothing to add (an empty pattern has no storage to point at)
  PHIFEM_CHECK_ARG(phi && f && cell_tags8 && data && b, "null pointer");
  PHIFEM_CHECK_ARG(active && slots, "null active / slots");
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
  if (n_entities == 0) return PHIFEM_OK;
  PHIFEM_CHECK_ARG(phi && data, "null pointer");
  PHIFEM_CHECK_ARG(entities && slots, "null entities / slots");
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
  if (n_facets == 0) return PHIFEM_OK;
  PHIFEM_CHECK_ARG(phi && data && mesh->c2f && mesh->f2c, "null pointer");
  PHIFEM_CHECK_ARG(facets && slots, "null facets / slots");
  const int grid = (int)((n_facets + kBlock - 1) / kBlock);
  cudaStream_t st = (cudaStream_t)stream;
  if (mesh->cell_type == PHIFEM_TRIANGLE)
    k_assemble_ghost_p1<2><<<grid, kBlock, 0, st>>>(*mesh, phi, facets, n_facets, slots, sigma, data);
  else
    k_assemble_ghost_p1<3><<<grid, kBlock, 0, st>>>(*mesh, phi, facets, n_facets, slots, sigma, data);
  PHIFEM_CHECK_LAUNCH();
  return PHIFEM_OK;
}

extern "C" int phifem_assemble_blocked_p1(const phifem_mesh* mesh, const double* phi, const double* f,
                                          double sigma, const phifem_blocked_plan* plan, double* data,
                                          double* b, void* stream) {
  if (int rc = check_simplex_mesh(mesh)) return rc;
  PHIFEM_CHECK_ARG(plan != nullptr, "plan is null");
  if (plan->n_blocks == 0) return PHIFEM_OK;
  PHIFEM_CHECK_ARG(phi && f && data && b, "null pointer");
  PHIFEM_CHECK_ARG(plan->n_blocks == 0 || (plan->block_desc && plan->seg_start && plan->seg_dest),
                   "plan arrays are null");
  PHIFEM_CHECK_ARG(plan->n_ghost_inst == 0 || (mesh->c2f && mesh->f2c), "ghost facets need c2f / f2c");
  PHIFEM_CHECK_ARG(plan->capacity > 0 && plan->capacity <= 32767, "plan.capacity out of range");
  if (plan->n_blocks == 0) return PHIFEM_OK;
  PHIFEM_CHECK_ARG(plan->max_segments > 0 && plan->max_segments % 8 == 0, "plan.max_segments");
  const size_t smem = (size_t)((plan->capacity + 1) & ~1) * sizeof(double) +
                      (size_t)plan->max_segments * (sizeof(int32_t) + sizeof(int16_t)) + 16;
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = kNumSMs;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // buffers up to ~110 KB: two CTAs of 256 threads share an SM (one reduces while the other computes)
  const bool half = smem <= 110 * 1024;
  const int per_sm = half ? 2 : 1;
  const int grid = plan->n_blocks < sms * per_sm ? plan->n_blocks : sms * per_sm;
  cudaError_t err = cudaSuccess;
  auto launch = [&](auto kernel, int threads) {
    err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err == cudaSuccess) kernel<<<grid, threads, smem, st>>>(*mesh, phi, f, sigma, *plan, data, b);
  };
  if (mesh->cell_type == PHIFEM_TRIANGLE) {
    if (half) launch(k_assemble_blocked_p1<2, kBlockedThreadsHalf>, kBlockedThreadsHalf);
    else launch(k_assemble_blocked_p1<2, kBlockedThreads>, kBlockedThreads);
  } else {
    if (half) launch(k_assemble_blocked_p1<3, kBlockedThreadsHalf>, kBlockedThreadsHalf);
    else launch(k_assemble_blocked_p1<3, kBlockedThreads>, kBlockedThreads);
  }
  if (err != cudaSuccess) {
    set_error("phifem_assemble_blocked_p1: cannot reserve %zu bytes of shared memory: %s", smem,
              cudaGetErrorString(err));
    return PHIFEM_ERR_CUDA;
  }
  PHIFEM_CHECK_LAUNCH();
  return PHIFEM_OK;
}
