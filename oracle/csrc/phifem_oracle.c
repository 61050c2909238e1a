/*
 * CPU oracle in C (TEST INFRASTRUCTURE + the CPU baseline leg of bench.py -- never the product).
 *
 * Plain-C restatement, for P1 level sets / P1 operators on triangles and tetrahedra, of
 *   - reference src/phifem/mesh_scripts.py `_compute_detection_vector` (:95-134), `_tag_cells`
 *     (:284-390, without single_layer_cut) and `_tag_facets` (:393-558) with detection degree 1;
 *   - the element tensors and ADD-scatter of the strong-Dirichlet forms,
 *     reference demo/strong-dirichlet/flower/main.py:104-128 (closed forms of SURVEY.md Appendix B).
 * It mirrors oracle/tags.py and oracle/assembly.py operation by operation (same sequential sums;
 * compile with -ffp-contract=off) and is validated against them in tests/test_oracle_native.py.
 * OpenMP over cells / facets so that the baseline uses all host cores, like an MPI run of the
 * reference would.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static const int TRI_FV[3][3] = {{1, 2, -1}, {0, 2, -1}, {0, 1, -1}};
static const int TET_FV[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* threads of the following calls (the CPU baseline is reported on 1 core and on all cores) */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

static double detj_abs(const double* x, int gdim, const int32_t* v) {
  if (gdim == 2) {
    const double* p0 = x + 2 * (int64_t)v[0];
    const double* p1 = x + 2 * (int64_t)v[1];
    const double* p2 = x + 2 * (int64_t)v[2];
    double j00 = p1[0] - p0[0], j01 = p2[0] - p0[0], j10 = p1[1] - p0[1], j11 = p2[1] - p0[1];
    return fabs(j00 * j11 - j01 * j10);
  }
  const double* p0 = x + 3 * (int64_t)v[0];
  double a[3], b[3], c[3];
  for (int d = 0; d < 3; ++d) {
    a[d] = x[3 * (int64_t)v[1] + d] - p0[d];
    b[d] = x[3 * (int64_t)v[2] + d] - p0[d];
    c[d] = x[3 * (int64_t)v[3] + d] - p0[d];
  }
  double det = (a[0] * (b[1] * c[2] - c[1] * b[2]) - b[0] * (a[1] * c[2] - c[1] * a[2])) +
               c[0] * (a[1] * b[2] - b[1] * a[2]);
  return fabs(det);
}

static int classify(double num, double den) {
  double d = (den > 0.0) ? num / den : 0.5;
  if (d > -1.0 && d < 1.0) return 2;
  if (d == 1.0) return 3;
  if (d == -1.0) return 1;
  return 0;
}

/* mesh_scripts.py:95-134 + :343-347, detection degree 1 (points = vertices), weights 1 */
void oracle_tag_cells_p1(const double* x, int gdim, const int32_t* cells, int64_t n_cells,
                         const double* phi, int32_t* tags) {
  const int nv = gdim + 1;
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < n_cells; ++c) {
    const int32_t* v = cells + c * nv;
    double s = detj_abs(x, gdim, v);
    double num = 0.0, den = 0.0;
    for (int k = 0; k < nv; ++k) {
      double t = phi[v[k]] * s;
      num = num + t;
      den = den + fabs(t);
    }
    tags[c] = classify(num, den);
  }
}

static double facet_scale(const double* x, int gdim, const int32_t* v, int lf) {
  if (gdim == 2) {
    const double* a = x + 2 * (int64_t)v[TRI_FV[lf][0]];
    const double* b = x + 2 * (int64_t)v[TRI_FV[lf][1]];
    double d0 = b[0] - a[0], d1 = b[1] - a[1];
    return sqrt(d0 * d0 + d1 * d1);
  }
  const double* a = x + 3 * (int64_t)v[TET_FV[lf][0]];
  const double* b = x + 3 * (int64_t)v[TET_FV[lf][1]];
  const double* c = x + 3 * (int64_t)v[TET_FV[lf][2]];
  double e1[3], e2[3];
  for (int d = 0; d < 3; ++d) {
    e1[d] = b[d] - a[d];
    e2[d] = c[d] - a[d];
  }
  double cx = e1[1] * e2[2] - e1[2] * e2[1];
  double cy = e1[2] * e2[0] - e1[0] * e2[2];
  double cz = e1[0] * e2[1] - e1[1] * e2[0];
  return sqrt((cx * cx + cy * cy) + cz * cz);
}

/* ds detection of the owner of a mesh-boundary facet (mesh_scripts.py:434-452) */
static int owner_ds_cut(const double* x, int gdim, const int32_t* cells, const int32_t* c2f,
                        const int32_t* f2c, const double* phi, int64_t c) {
  const int nv = gdim + 1, nvf = gdim;
  const int32_t* v = cells + c * nv;
  double num = 0.0, den = 0.0;
  int32_t last = -1;
  for (int round = 0; round < nv; ++round) {
    int lf = -1;
    int32_t best = INT32_MAX;
    for (int i = 0; i < nv; ++i) {
      int32_t f = c2f[c * nv + i];
      if (f2c[2 * (int64_t)f + 1] < 0 && f > last && f < best) {
        best = f;
        lf = i;
      }
    }
    if (lf < 0) break;
    last = best;
    double sc = facet_scale(x, gdim, v, lf);
    double fn = 0.0, fd = 0.0;
    for (int q = 0; q < nvf; ++q) {
      int loc = gdim == 2 ? TRI_FV[lf][q] : TET_FV[lf][q];
      double t = phi[v[loc]] * sc;
      fn = fn + t;
      fd = fd + fabs(t);
    }
    num = num + fn;
    den = den + fd;
  }
  double d = (den > 0.0) ? num / den : 0.5;
  return d > -1.0 && d < 1.0;
}

/* mesh_scripts.py:393-558 as a per-facet decision over the adjacent cell tags */
void oracle_tag_facets_p1(const double* x, int gdim, const int32_t* cells, const int32_t* c2f,
                          const int32_t* f2c, int64_t n_cells, int64_t n_facets, const double* phi,
                          const int32_t* ctags, int32_t* ftags) {
  int any_ext = 0;
#pragma omp parallel for reduction(| : any_ext) schedule(static)
  for (int64_t c = 0; c < n_cells; ++c) any_ext |= (ctags[c] == 3);
#pragma omp parallel for schedule(static)
  for (int64_t f = 0; f < n_facets; ++f) {
    const int32_t c0 = f2c[2 * f], c1 = f2c[2 * f + 1];
    const int t0 = ctags[c0], bnd = c1 < 0, t1 = bnd ? 0 : ctags[c1];
    const int inI = (t0 == 1) | (t1 == 1), inC = (t0 == 2) | (t1 == 2), inE = (t0 == 3) | (t1 == 3);
    int k = 0;
    if (bnd) k = owner_ds_cut(x, gdim, cells, c2f, f2c, phi, c0);
    const int cut_bnd = bnd && k;
    const int uncut_bnd = bnd && !k && !inE && !inI;
    const int int_bnd = inI && inC;
    int boundary = any_ext ? ((inE && inC) || uncut_bnd) : bnd;
    const int direct = inE && inI;
    const int cutf = (inC && !(boundary || int_bnd || direct || uncut_bnd)) || cut_bnd;
    const int rem = int_bnd || boundary || direct;
    const int interior = inI && !rem, exterior = inE && !rem;
    boundary = boundary && !cutf;
    int tag = 0;
    if (exterior) tag = 5;
    if (interior) tag = 1;
    if (int_bnd) tag = 3;
    if (cutf) tag = 2;
    if (boundary) tag = 4;
    if (direct) tag = 6;
    ftags[f] = tag;
  }
}

/* ---- strong-Dirichlet operator, closed forms -------------------------------- */
typedef struct {
  double G[4][3];
  double vol, h2;
  double p[4];
  int32_t v[4];
} simplex_t;

static void simplex_setup(const double* x, int D, const int32_t* v, const double* phi, simplex_t* s) {
  const int nv = D + 1;
  double xc[4][3];
  for (int k = 0; k < nv; ++k) {
    s->v[k] = v[k];
    s->p[k] = phi[v[k]];
    for (int d = 0; d < D; ++d) xc[k][d] = x[(int64_t)D * v[k] + d];
  }
  if (D == 2) {
    double a0 = xc[1][0] - xc[0][0], a1 = xc[1][1] - xc[0][1];
    double b0 = xc[2][0] - xc[0][0], b1 = xc[2][1] - xc[0][1];
    double det = a0 * b1 - b0 * a1;
    s->G[1][0] = b1 / det;  s->G[1][1] = -b0 / det;
    s->G[2][0] = -a1 / det; s->G[2][1] = a0 / det;
    s->vol = 0.5 * fabs(det);
  } else {
    double a[3], b[3], c[3];
    for (int d = 0; d < 3; ++d) {
      a[d] = xc[1][d] - xc[0][d];
      b[d] = xc[2][d] - xc[0][d];
      c[d] = xc[3][d] - xc[0][d];
    }
    double r1[3] = {b[1] * c[2] - b[2] * c[1], b[2] * c[0] - b[0] * c[2], b[0] * c[1] - b[1] * c[0]};
    double r2[3] = {c[1] * a[2] - c[2] * a[1], c[2] * a[0] - c[0] * a[2], c[0] * a[1] - c[1] * a[0]};
    double r3[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    double det = a[0] * r1[0] + a[1] * r1[1] + a[2] * r1[2];
    for (int d = 0; d < 3; ++d) {
      s->G[1][d] = r1[d] / det;
      s->G[2][d] = r2[d] / det;
      s->G[3][d] = r3[d] / det;
    }
    s->vol = fabs(det) / 6.0;
  }
  for (int d = 0; d < D; ++d) {
    double t = 0.0;
    for (int k = 1; k < nv; ++k) t += s->G[k][d];
    s->G[0][d] = -t;
  }
  s->h2 = 0.0;
  for (int a = 0; a < nv; ++a)
    for (int b = a + 1; b < nv; ++b) {
      double t = 0.0;
      for (int d = 0; d < D; ++d) t += (xc[a][d] - xc[b][d]) * (xc[a][d] - xc[b][d]);
      if (t > s->h2) s->h2 = t;
    }
}

static double dotD(const double* a, const double* b, int D) {
  double s = 0.0;
  for (int d = 0; d < D; ++d) s += a[d] * b[d];
  return s;
}

static void atomic_add(double* p, double v) {
#pragma omp atomic
  *p += v;
}

void oracle_assemble_cells_p1(const double* x, int D, const int32_t* cells, const double* phi,
                              const double* f, const int32_t* ctags, const int32_t* active,
                              int64_t n_active, const int32_t* slots, double sigma, double* data,
                              double* b) {
  const int nv = D + 1;
  const double fact_d = D == 2 ? 2.0 : 6.0, fact_d3 = D == 2 ? 120.0 : 720.0;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < n_active; ++e) {
    const int64_t c = active[e];
    simplex_t s;
    simplex_setup(x, D, cells + c * nv, phi, &s);
    double g[3] = {0, 0, 0}, a[4], mm[4], fv[4];
    for (int d = 0; d < D; ++d)
      for (int k = 0; k < nv; ++k) g[d] += s.p[k] * s.G[k][d];
    const double gg = dotD(g, g, D);
    const double cM = s.vol / ((D + 1) * (D + 2));
    double P = 0, F = 0, FP = 0, mu = 0;
    for (int k = 0; k < nv; ++k) {
      fv[k] = f[s.v[k]];
      a[k] = dotD(g, s.G[k], D);
      P += s.p[k];
      F += fv[k];
      FP += fv[k] * s.p[k];
    }
    for (int k = 0; k < nv; ++k) {
      mm[k] = cM * (P + s.p[k]);
      mu += s.p[k] * mm[k];
    }
    const double stab = ctags[c] == 2 ? sigma * s.h2 * s.vol : 0.0;
    for (int i = 0; i < nv; ++i)
      for (int j = 0; j < nv; ++j) {
        double val = gg * cM * (i == j ? 2.0 : 1.0) + a[i] * mm[j] + mm[i] * a[j] +
                     dotD(s.G[i], s.G[j], D) * mu + 4.0 * stab * a[i] * a[j];
        atomic_add(data + slots[e * nv * nv + i * nv + j], val);
      }
    const double c3 = s.vol * fact_d / fact_d3;
    for (int i = 0; i < nv; ++i) {
      double bi = c3 * ((F * P + FP) + fv[i] * P + F * s.p[i] + 2.0 * fv[i] * s.p[i]) -
                  2.0 * stab * (F / nv) * a[i];
      atomic_add(b + s.v[i], bi);
    }
  }
}

static double alpha3(int a, int b, int c) { return (double)((1 + (a == b)) * (1 + (a == c) + (b == c))); }

void oracle_assemble_boundary_p1(const double* x, int D, const int32_t* cells, const double* phi,
                                 const int32_t* entities, int64_t n_entities, const int32_t* slots,
                                 double* data) {
  const int nv = D + 1;
  const double cfac = D == 2 ? 1.0 / 24.0 : 2.0 / 120.0;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < n_entities; ++e) {
    const int64_t c = entities[2 * e];
    const int o = entities[2 * e + 1];
    simplex_t s;
    simplex_setup(x, D, cells + c * nv, phi, &s);
    const double gnorm = sqrt(dotD(s.G[o], s.G[o], D));
    double n[3], g[3] = {0, 0, 0};
    for (int d = 0; d < D; ++d) {
      n[d] = -s.G[o][d] / gnorm;
      for (int k = 0; k < nv; ++k) g[d] += s.p[k] * s.G[k][d];
    }
    const double cF = D * s.vol * gnorm * cfac, gn = dotD(g, n, D);
    for (int i = 0; i < nv; ++i)
      for (int j = 0; j < nv; ++j) {
        double acc = 0.0, Gnj = dotD(s.G[j], n, D);
        for (int k = 0; k < nv; ++k) {
          if (k == o) continue;
          double t = 0.0;
          for (int l = 0; l < nv; ++l)
            if (l != o) t += s.p[l] * alpha3(l, k, i);
          acc += s.p[k] * (gn * (j != o ? alpha3(j, k, i) : 0.0) + Gnj * t);
        }
        atomic_add(data + slots[e * nv * nv + i * nv + j], i == o ? 0.0 : -cF * acc);
      }
  }
}

/* slots[n_facets, (nv+1)^2] runs over the DISTINCT vertices of the macro element: facet vertices as
 * ordered in cell +, opposite vertex of cell +, opposite vertex of cell - (include/phifem_b200.h).  The
 * (2nv)^2 macro entries of the reference are still computed and added one by one. */
void oracle_assemble_ghost_p1(const double* x, int D, const int32_t* cells, const int32_t* c2f,
                              const int32_t* f2c, const double* phi, const int32_t* facets,
                              int64_t n_facets, const int32_t* slots, double sigma, double* data) {
  const int nv = D + 1, nm = 2 * nv, ng = D + 2;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < n_facets; ++e) {
    const int32_t fct = facets[e];
    double Jv[8][3], Js[8], hsum = 0.0, area = 0.0;
    int32_t fvert[3] = {0, 0, 0};
    int target[8];
    for (int side = 0; side < 2; ++side) {
      const int64_t c = f2c[2 * (int64_t)fct + side];
      simplex_t s;
      simplex_setup(x, D, cells + c * nv, phi, &s);
      int o = 0;
      for (int i = 0; i < nv; ++i)
        if (c2f[c * nv + i] == fct) o = i;
      const double gnorm = sqrt(dotD(s.G[o], s.G[o], D));
      double n[3], g[3] = {0, 0, 0};
      for (int d = 0; d < D; ++d) {
        n[d] = -s.G[o][d] / gnorm;
        for (int k = 0; k < nv; ++k) g[d] += s.p[k] * s.G[k][d];
      }
      const double gn = dotD(g, n, D);
      hsum += sqrt(s.h2);
      if (side == 0) {
        area = D * s.vol * gnorm;
        int q = 0;
        for (int k = 0; k < nv; ++k)
          if (k != o) fvert[q++] = s.v[k];
      }
      for (int a = 0; a < nv; ++a) {
        double Gna = dotD(s.G[a], n, D);
        target[side * nv + a] = D + side;
        for (int k = 0; k < D; ++k) {
          Jv[side * nv + a][k] = (s.v[a] == fvert[k] ? gn : 0.0) + Gna * phi[fvert[k]];
          if (s.v[a] == fvert[k]) target[side * nv + a] = k;
        }
      }
    }
    const double coef = sigma * 0.5 * hsum * area / (D * (D + 1));
    for (int a = 0; a < nm; ++a) {
      Js[a] = 0.0;
      for (int k = 0; k < D; ++k) Js[a] += Jv[a][k];
    }
    for (int a = 0; a < nm; ++a)
      for (int bb = 0; bb < nm; ++bb) {
        double t = Js[a] * Js[bb];
        for (int k = 0; k < D; ++k) t += Jv[a][k] * Jv[bb][k];
        atomic_add(data + slots[e * ng * ng + target[a] * ng + target[bb]], coef * t);
      }
  }
}
