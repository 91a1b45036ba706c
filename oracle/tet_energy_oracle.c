/* CPU oracle #2 for the tssplat geometry-energy hot path -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain C, fp64, matrix-free, OpenMP.  It is the checker and the timed CPU baseline
 * (bench.py cpu_baseline / --impl reference), never the product: nothing under tssplat_b200/ or
 * tet_spheres/ links, loads or calls it.
 *
 * PARITY UNPINNED: the reference has no golden vectors for this path and the Laplacian weights
 * live in un-vendored libpgo (see oracle/tet_energy_oracle.py header).  Assumption restated here:
 * L = face-adjacency graph Laplacian over tets (L_tt = #face-neighbours, L_ts = -1), optionally
 * row-scaled by 1/#neighbours, applied to each of the 9 entries of F.
 *
 * What it restates (paths relative to /root/reference), written independently of the sparse
 * operator form in tet_energy_oracle.py so the two can cross-check each other:
 *   F_t = Ds_t Dm_t^-1                         geometry/mesh_utils.py:38-69
 *   E   = c1 * 1/2 * ||L F||^2  +  c2 * sum_t max(-det F_t, 0)^order
 *                                               tssplat_ext/tet_spheres/tet_spheres_cuda.cu:118-195
 *         (1/2 x^T G^T L^T L G x  ==  1/2 ||L G x||^2; the 0.5 is cu:157, c1/c2 cu:191)
 *   dE/dx = gradH * G^T ( c1 * L^T L F + c2 * D ),  D_t = -p(-J)^(p-1) cof(F_t) if J<0 else 0
 *                                               tet_spheres_cuda.cu:197-263, :68-102, :32-46
 *   order not in {2,4}: zero energy, zero gradient  (cu:57-63, 83-89)
 *
 * Build: see oracle/Makefile.  Three builds of this one file: the fp64 checker (portable -O2), and two
 * TIMING builds for bench.py's CPU arms (-O3 -march=x86-64-v3, used only when the host has AVX2+FMA):
 * fp64, and fp32 arithmetic (-DTSO_REAL=float: the reference's own precision, tet_spheres.cpp:43-45).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef TSO_REAL
#define TSO_REAL double
#endif
typedef TSO_REAL real;   /* arithmetic type of the per-tet passes (the checker build uses double) */

typedef struct {
  int n, nele, scale;
  int *tets;      /* nele*4 */
  int *nbr;       /* nele*4, -1 = boundary; face k is opposite local vertex k */
  real *B;        /* nele*9, row-major Dm^-1 */
  real *w;        /* nele, Laplacian row scale (1 or 1/deg) */
  int *deg;       /* nele */
  int *inc_ptr;   /* n+1  : vertex -> incident (tet*4+slot) */
  int *inc;       /* nele*4 */
  real *F, *H, *P;    /* nele*9 scratch */
  real *contrib;      /* nele*12 */
} TsoOracle;

typedef struct { int64_t a, b, c; int owner; } FaceKey;

static int face_cmp(const void *p, const void *q) {
  const FaceKey *x = (const FaceKey *)p, *y = (const FaceKey *)q;
  if (x->a != y->a) return x->a < y->a ? -1 : 1;
  if (x->b != y->b) return x->b < y->b ? -1 : 1;
  if (x->c != y->c) return x->c < y->c ? -1 : 1;
  return 0;
}

static void sort3(int64_t *v) {
  int64_t t;
  if (v[0] > v[1]) { t = v[0]; v[0] = v[1]; v[1] = t; }
  if (v[1] > v[2]) { t = v[1]; v[1] = v[2]; v[2] = t; }
  if (v[0] > v[1]) { t = v[0]; v[0] = v[1]; v[1] = t; }
}

static int inv3(const double *m, double *o) {
  double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  double d = m[0] * c00 + m[1] * c01 + m[2] * c02;
  if (d == 0.0) return 1;
  double id = 1.0 / d;
  o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  return 0;
}

void tso_destroy(TsoOracle *o) {
  if (!o) return;
  free(o->tets); free(o->nbr); free(o->B); free(o->w); free(o->deg); free(o->inc_ptr); free(o->inc);
  free(o->F); free(o->H); free(o->P); free(o->contrib); free(o);
}

/* rest: float32 [n*3] (the reference receives float32 and widens: tet_spheres.cpp:251-254) */
TsoOracle *tso_create(const float *rest, const int *tets, int n, int nele, int laplacian_scale) {
  static const int FACE[4][3] = {{1, 2, 3}, {0, 3, 2}, {0, 1, 3}, {0, 2, 1}};
  TsoOracle *o = (TsoOracle *)calloc(1, sizeof(TsoOracle));
  o->n = n; o->nele = nele; o->scale = laplacian_scale;
  o->tets = (int *)malloc(sizeof(int) * 4 * (size_t)nele);
  memcpy(o->tets, tets, sizeof(int) * 4 * (size_t)nele);
  o->nbr = (int *)malloc(sizeof(int) * 4 * (size_t)nele);
  o->B = (real *)malloc(sizeof(real) * 9 * (size_t)nele);
  o->w = (real *)malloc(sizeof(real) * (size_t)nele);
  o->deg = (int *)malloc(sizeof(int) * (size_t)nele);
  o->F = (real *)malloc(sizeof(real) * 9 * (size_t)nele);
  o->H = (real *)malloc(sizeof(real) * 9 * (size_t)nele);
  o->P = (real *)malloc(sizeof(real) * 9 * (size_t)nele);
  o->contrib = (real *)malloc(sizeof(real) * 12 * (size_t)nele);
  for (int t = 0; t < nele; t++) {
    const int *v = tets + 4 * t;
    for (int k = 0; k < 4; k++) if (v[k] < 0 || v[k] >= n) { tso_destroy(o); return NULL; }
    double Dm[9];
    for (int r = 0; r < 3; r++)
      for (int k = 0; k < 3; k++)
        Dm[3 * r + k] = (double)rest[3 * v[k + 1] + r] - (double)rest[3 * v[0] + r];
    double Bd[9];
    if (inv3(Dm, Bd)) { tso_destroy(o); return NULL; }
    for (int i = 0; i < 9; i++) o->B[9 * t + i] = (real)Bd[i];     /* fp64 -> working precision (tet_spheres.cpp:43-45) */
  }
  /* face adjacency by sorting face keys */
  FaceKey *fk = (FaceKey *)malloc(sizeof(FaceKey) * 4 * (size_t)nele);
  for (int t = 0; t < nele; t++)
    for (int k = 0; k < 4; k++) {
      int64_t f[3] = {tets[4 * t + FACE[k][0]], tets[4 * t + FACE[k][1]], tets[4 * t + FACE[k][2]]};
      sort3(f);
      FaceKey *e = fk + 4 * (size_t)t + k;
      e->a = f[0]; e->b = f[1]; e->c = f[2]; e->owner = 4 * t + k;
    }
  qsort(fk, 4 * (size_t)nele, sizeof(FaceKey), face_cmp);
  for (size_t i = 0; i < 4 * (size_t)nele; i++) o->nbr[i] = -1;
  for (size_t i = 0; i + 1 < 4 * (size_t)nele; i++) {
    if (face_cmp(fk + i, fk + i + 1) == 0) {
      if (i + 2 < 4 * (size_t)nele && face_cmp(fk + i, fk + i + 2) == 0) { free(fk); tso_destroy(o); return NULL; }
      o->nbr[fk[i].owner] = fk[i + 1].owner / 4;
      o->nbr[fk[i + 1].owner] = fk[i].owner / 4;
      i++;
    }
  }
  free(fk);
  for (int t = 0; t < nele; t++) {
    int d = 0;
    for (int k = 0; k < 4; k++) d += o->nbr[4 * t + k] >= 0;
    o->deg[t] = d;
    o->w[t] = (real)(laplacian_scale ? (d > 0 ? 1.0 / d : 0.0) : 1.0);
  }
  /* vertex incidence CSR */
  o->inc_ptr = (int *)calloc((size_t)n + 1, sizeof(int));
  o->inc = (int *)malloc(sizeof(int) * 4 * (size_t)nele);
  for (int i = 0; i < 4 * nele; i++) o->inc_ptr[tets[i] + 1]++;
  for (int i = 0; i < n; i++) o->inc_ptr[i + 1] += o->inc_ptr[i];
  int *cur = (int *)malloc(sizeof(int) * (size_t)n);
  memcpy(cur, o->inc_ptr, sizeof(int) * (size_t)n);
  for (int i = 0; i < 4 * nele; i++) o->inc[cur[tets[i]]++] = i;
  free(cur);
  return o;
}

int tso_real_bytes(void) { return (int)sizeof(real); }

int tso_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static inline real det3(const real *F) {
  return F[0] * (F[4] * F[8] - F[5] * F[7]) - F[1] * (F[3] * F[8] - F[5] * F[6]) + F[2] * (F[3] * F[7] - F[4] * F[6]);
}

/* d det / dF, row-major (tet_spheres_cuda.cu:32-46) */
static inline void cof3(const real *F, real *C) {
  C[0] = F[4] * F[8] - F[5] * F[7]; C[1] = F[5] * F[6] - F[3] * F[8]; C[2] = F[3] * F[7] - F[4] * F[6];
  C[3] = F[2] * F[7] - F[1] * F[8]; C[4] = F[0] * F[8] - F[2] * F[6]; C[5] = F[1] * F[6] - F[0] * F[7];
  C[6] = F[1] * F[5] - F[2] * F[4]; C[7] = F[2] * F[3] - F[0] * F[5]; C[8] = F[0] * F[4] - F[1] * F[3];
}

/* x: float32 [n*3].  terms[0] = 1/2||LF||^2 (unweighted by c1), terms[1] = sum barrier (unweighted),
 * terms[2] = sum AMIPS (only written by tso_energy_grad_ex; unweighted).  grad: double [n*3] or NULL.
 * AMIPS (c3; no counterpart in the reference): psi = tr(F^T F) / (3 det(F)^(2/3)) - 1 for
 * det F > 0, else 0;  d psi / dF = 2 / (3 J^(2/3)) * (F - tr / (3 J) * cof F). */
static int tso_energy_grad_impl(TsoOracle *o, const float *x, double c1, double c2, double c3, int order, double gradH,
                                double *terms, double *grad, int nthreads);

int tso_energy_grad(TsoOracle *o, const float *x, double c1, double c2, int order, double gradH,
                    double *terms, double *grad, int nthreads) {
  double t3[3];
  int rc = tso_energy_grad_impl(o, x, c1, c2, 0.0, order, gradH, t3, grad, nthreads);
  if (terms) { terms[0] = t3[0]; terms[1] = t3[1]; }
  return rc;
}

int tso_energy_grad_ex(TsoOracle *o, const float *x, double c1, double c2, double c3, int order, double gradH,
                       double *terms3, double *grad, int nthreads) {
  return tso_energy_grad_impl(o, x, c1, c2, c3, order, gradH, terms3, grad, nthreads);
}

static int tso_energy_grad_impl(TsoOracle *o, const float *x, double c1, double c2, double c3, int order, double gradH,
                                double *terms, double *grad, int nthreads) {
  const int nele = o->nele, n = o->n;
  double sm = 0.0, bar = 0.0, ami = 0.0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
#pragma omp for schedule(static)
    for (int t = 0; t < nele; t++) {
      const int *v = o->tets + 4 * t;
      const real *B = o->B + 9 * t;
      real Ds[9];
      for (int r = 0; r < 3; r++)
        for (int k = 0; k < 3; k++) Ds[3 * r + k] = (real)x[3 * v[k + 1] + r] - (real)x[3 * v[0] + r];
      real *F = o->F + 9 * t;
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) F[3 * r + c] = Ds[3 * r] * B[c] + Ds[3 * r + 1] * B[3 + c] + Ds[3 * r + 2] * B[6 + c];
    }
#pragma omp for schedule(static) reduction(+ : sm)
    for (int t = 0; t < nele; t++) {
      real *H = o->H + 9 * t;
      const real *F = o->F + 9 * t;
      for (int i = 0; i < 9; i++) H[i] = o->deg[t] * F[i];
      for (int k = 0; k < 4; k++) {
        int s = o->nbr[4 * t + k];
        if (s >= 0) for (int i = 0; i < 9; i++) H[i] -= o->F[9 * s + i];
      }
      real e = 0;
      for (int i = 0; i < 9; i++) { H[i] *= o->w[t]; e += H[i] * H[i]; }
      sm += 0.5 * (double)e;
    }
#pragma omp for schedule(static) reduction(+ : bar, ami)
    for (int t = 0; t < nele; t++) {
      /* P = c1 * (L^T H)_t + c2 * D_t ;  (L^T H)_t = deg_t w_t H_t - sum_s w_s H_s */
      real P[9];
      const real *H = o->H + 9 * t;         /* H_t = w_t (deg_t F_t - sum F_s), stored scaled */
      for (int i = 0; i < 9; i++) P[i] = o->deg[t] * o->w[t] * H[i];
      for (int k = 0; k < 4; k++) {
        int s = o->nbr[4 * t + k];
        if (s >= 0) for (int i = 0; i < 9; i++) P[i] -= o->w[s] * o->H[9 * s + i];
      }
      for (int i = 0; i < 9; i++) P[i] *= (real)c1;
      const real *F = o->F + 9 * t;
      real J = det3(F);
      if (J < 0) {
        real m = -J, e = 0, coef = 0, C[9];
        if (order == 2) { e = m * m; coef = 2 * m; }
        else if (order == 4) { e = m * m * m * m; coef = 4 * m * m * m; }
        bar += (double)e;
        cof3(F, C);
        for (int i = 0; i < 9; i++) P[i] += (real)c2 * (-coef) * C[i];
      } else if (c3 != 0.0 && J > 0) {
        real tr = 0, C[9];
        for (int i = 0; i < 9; i++) tr += F[i] * F[i];
        const double j23 = pow((double)J, 2.0 / 3.0);
        ami += (double)tr / (3.0 * j23) - 1.0;
        cof3(F, C);
        const real a = (real)(2.0 / (3.0 * j23)), b = (real)((double)tr / (3.0 * (double)J));
        for (int i = 0; i < 9; i++) P[i] += (real)c3 * a * (F[i] - b * C[i]);
      }
      /* dE/dx_k = P a_k,  a_k = row k-1 of B (k=1..3), a_0 = -(a_1+a_2+a_3) */
      const real *B = o->B + 9 * t;
      real *cb = o->contrib + 12 * t;
      for (int r = 0; r < 3; r++) {
        real s0 = 0;
        for (int k = 0; k < 3; k++) {
          real g = P[3 * r] * B[3 * k] + P[3 * r + 1] * B[3 * k + 1] + P[3 * r + 2] * B[3 * k + 2];
          cb[3 * (k + 1) + r] = g;
          s0 -= g;
        }
        cb[r] = s0;
      }
    }
    if (grad) {
#pragma omp for schedule(static)
      for (int v = 0; v < n; v++) {
        real g0 = 0, g1 = 0, g2 = 0;
        for (int e = o->inc_ptr[v]; e < o->inc_ptr[v + 1]; e++) {
          const real *cb = o->contrib + 3 * (size_t)o->inc[e];
          g0 += cb[0]; g1 += cb[1]; g2 += cb[2];
        }
        grad[3 * v] = gradH * (double)g0; grad[3 * v + 1] = gradH * (double)g1; grad[3 * v + 2] = gradH * (double)g2;
      }
    }
  }
  if (terms) { terms[0] = sm; terms[1] = bar; terms[2] = ami; }
  (void)c2;
  return 0;
}
