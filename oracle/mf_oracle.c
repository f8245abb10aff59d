/* TEST INFRASTRUCTURE — CPU restatement (plain C, sum-factorised, OpenMP over
 * cells) of the reference's matrix-free operator path in 3-D.  It is the
 * checker for the CUDA library and the "port" CPU baseline of bench.py; the
 * product never links or calls it.
 *
 * PARITY PIN STATUS: the reference has no operator-level golden vectors
 * (SURVEY.md §8c).  This file is validated against oracle/np_oracle.py (dense,
 * no sum factorisation) to <= 1e-13 and through it against the 2-D Q1
 * Turek-Hron end-to-end goldens of /root/reference/tests/solid (tests/test_golden_turek.py).
 *
 * Reference lines restated (relative to /root/reference):
 *   include/solid_mechanics/mf_nh_operator.h:392-523   cache()
 *   include/solid_mechanics/mf_nh_operator.h:575-647   vmult_add / cell loop
 *   include/solid_mechanics/mf_nh_operator.h:651-1019  do_operation_on_cell
 *   include/solid_mechanics/mf_nh_operator.h:1023-1060 compute_diagonal
 *   include/solid_mechanics/material.h:78-94           material constants
 *   include/heat_transfer/heat_transfer.h:205-250      Laplace+mass operator
 * deal.II semantics (FEEvaluation evaluate/integrate, Jacobian conventions,
 * constraints) are restated from SURVEY.md Appendix A.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXN 8

/* ------------------------------------------------------------------ */
/* 1-D basis: FE_Q(p) on Gauss-Lobatto points, Gauss quadrature on [0,1] */

static void legendre(int n, double x, double *P, double *dP, double *ddP)
{
  /* P_n, P_n', P_n'' at x in (-1,1) */
  double p0 = 1.0, p1 = x;
  if (n == 0) { *P = 1.0; *dP = 0.0; *ddP = 0.0; return; }
  for (int k = 2; k <= n; ++k) {
    double pk = ((2.0 * k - 1.0) * x * p1 - (k - 1.0) * p0) / k;
    p0 = p1; p1 = pk;
  }
  *P = p1;
  *dP = n * (x * p1 - p0) / (x * x - 1.0);
  *ddP = (2.0 * x * (*dP) - n * (n + 1.0) * p1) / (1.0 - x * x);
}

static void gauss_points(int nq, double *x, double *w)
{
  const double pi = 3.14159265358979323846;
  for (int i = 0; i < nq; ++i) {
    double z = -cos(pi * (i + 0.75) / (nq + 0.5));
    for (int it = 0; it < 100; ++it) {
      double P, dP, ddP;
      legendre(nq, z, &P, &dP, &ddP);
      double dz = P / dP;
      z -= dz;
      if (fabs(dz) < 1e-16) break;
    }
    double P, dP, ddP;
    legendre(nq, z, &P, &dP, &ddP);
    x[i] = z;
    w[i] = 2.0 / ((1.0 - z * z) * dP * dP);
  }
  for (int i = 0; i < nq / 2; ++i) { /* symmetrise */
    double a = 0.5 * (x[nq - 1 - i] - x[i]);
    x[i] = -a; x[nq - 1 - i] = a;
    double ww = 0.5 * (w[i] + w[nq - 1 - i]);
    w[i] = w[nq - 1 - i] = ww;
  }
  if (nq % 2) x[nq / 2] = 0.0;
  for (int i = 0; i < nq; ++i) { x[i] = 0.5 * (x[i] + 1.0); w[i] *= 0.5; }
}

static void gll_points(int p, double *x)
{
  const double pi = 3.14159265358979323846;
  x[0] = -1.0; x[p] = 1.0;
  for (int i = 1; i < p; ++i) {
    double z = -cos(pi * i / p);
    for (int it = 0; it < 100; ++it) {
      double P, dP, ddP;
      legendre(p, z, &P, &dP, &ddP);
      double dz = dP / ddP;
      z -= dz;
      if (fabs(dz) < 1e-16) break;
    }
    x[i] = z;
  }
  for (int i = 0; i <= p / 2; ++i) {
    double a = 0.5 * (x[p - i] - x[i]);
    x[i] = -a; x[p - i] = a;
  }
  for (int i = 0; i <= p; ++i) x[i] = 0.5 * (x[i] + 1.0);
}

static void lagrange(int n, const double *nodes, double x, double *v, double *d)
{
  for (int i = 0; i < n; ++i) {
    double den = 1.0, num = 1.0, s = 0.0;
    for (int j = 0; j < n; ++j) if (j != i) { den *= nodes[i] - nodes[j]; num *= x - nodes[j]; }
    for (int m = 0; m < n; ++m) if (m != i) {
      double t = 1.0;
      for (int j = 0; j < n; ++j) if (j != i && j != m) t *= x - nodes[j];
      s += t;
    }
    v[i] = num / den; d[i] = s / den;
  }
}

/* S[q*n+i] = l_i(xi_q); D[q*n+i] = l_i'(xi_q); C[q*nq+r] = collocation
 * derivative on the Gauss points (= D S^-1 when nq == n) */
void orc_shape_1d(int p, int nq, double *S, double *D, double *C, double *w, double *nodes, double *qpts)
{
  int n = p + 1;
  double nd[MAXN], xq[MAXN], ww[MAXN], v[MAXN], d[MAXN];
  gll_points(p, nd);
  gauss_points(nq, xq, ww);
  for (int q = 0; q < nq; ++q) {
    lagrange(n, nd, xq[q], v, d);
    for (int i = 0; i < n; ++i) { S[q * n + i] = v[i]; D[q * n + i] = d[i]; }
    lagrange(nq, xq, xq[q], v, d);
    for (int r = 0; r < nq; ++r) C[q * nq + r] = d[r];
    if (w) w[q] = ww[q];
    if (qpts) qpts[q] = xq[q];
  }
  if (nodes) for (int i = 0; i < n; ++i) nodes[i] = nd[i];
}

/* ------------------------------------------------------------------ */
/* Q1 mapping metric: invjac[e][f] = d xi_f / d x_e (deal.II inverse_jacobian,
 * mf_nh_operator.h:872-887), JxW = det(J) w_q */

static inline double det3(const double *A)
{
  return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) +
         A[2] * (A[3] * A[7] - A[4] * A[6]);
}
static inline void inv3(const double *A, double *B, double *det)
{
  double d = det3(A), id = 1.0 / d;
  B[0] = (A[4] * A[8] - A[5] * A[7]) * id; B[1] = (A[2] * A[7] - A[1] * A[8]) * id; B[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  B[3] = (A[5] * A[6] - A[3] * A[8]) * id; B[4] = (A[0] * A[8] - A[2] * A[6]) * id; B[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  B[6] = (A[3] * A[7] - A[4] * A[6]) * id; B[7] = (A[1] * A[6] - A[0] * A[7]) * id; B[8] = (A[0] * A[4] - A[1] * A[3]) * id;
  *det = d;
}

void orc_q1_metric(int64_t ncell, const double *vtx /*[cell][8][3]*/, int nq, double *invjac /*[cell][nq3][9]*/,
                   double *jxw /*[cell][nq3]*/)
{
  double xq[MAXN], w[MAXN];
  gauss_points(nq, xq, w);
  int nq3 = nq * nq * nq;
#pragma omp parallel for
  for (int64_t c = 0; c < ncell; ++c) {
    const double *v = vtx + c * 24;
    for (int q = 0; q < nq3; ++q) {
      int qi = q % nq, qj = (q / nq) % nq, qk = q / (nq * nq);
      double xi[3] = {xq[qi], xq[qj], xq[qk]};
      double J[9] = {0};
      for (int m = 0; m < 8; ++m) {
        int b[3] = {m & 1, (m >> 1) & 1, (m >> 2) & 1};
        double l[3], dl[3];
        for (int d = 0; d < 3; ++d) { l[d] = b[d] ? xi[d] : 1.0 - xi[d]; dl[d] = b[d] ? 1.0 : -1.0; }
        double g[3] = {dl[0] * l[1] * l[2], l[0] * dl[1] * l[2], l[0] * l[1] * dl[2]};
        for (int e = 0; e < 3; ++e) for (int f = 0; f < 3; ++f) J[e * 3 + f] += v[m * 3 + e] * g[f];
      }
      double Ji[9], det;
      inv3(J, Ji, &det); /* Ji[f][e] = d xi_f / d x_e */
      double *o = invjac + (c * nq3 + q) * 9;
      for (int e = 0; e < 3; ++e) for (int f = 0; f < 3; ++f) o[e * 3 + f] = Ji[f * 3 + e];
      jxw[c * nq3 + q] = det * w[qi] * w[qj] * w[qk];
    }
  }
}

/* ------------------------------------------------------------------ */
/* cell data + operators                                                */

typedef struct {
  int p, n, nq, ncomp;
  int64_t ncell, ndofs;
  const uint32_t *dof_idx;   /* [cell][ncomp*n^3], component-major, lexicographic */
  const uint8_t *constrained; /* [ndofs] */
  const double *invjac;      /* [cell][nq3][9] reference mapping */
  const double *jxw;         /* [cell][nq3] */
  const uint8_t *material_id; /* [cell] or NULL */
  double S[MAXN * MAXN], D[MAXN * MAXN], C[MAXN * MAXN], w[MAXN];
} orc_cells;

orc_cells *orc_cells_create(int p, int nq, int ncomp, int64_t ncell, int64_t ndofs, const uint32_t *dof_idx,
                            const uint8_t *constrained, const double *invjac, const double *jxw,
                            const uint8_t *material_id)
{
  orc_cells *c = (orc_cells *)calloc(1, sizeof(orc_cells));
  c->p = p; c->n = p + 1; c->nq = nq; c->ncomp = ncomp; c->ncell = ncell; c->ndofs = ndofs;
  c->dof_idx = dof_idx; c->constrained = constrained; c->invjac = invjac; c->jxw = jxw;
  c->material_id = material_id;
  orc_shape_1d(p, nq, c->S, c->D, c->C, c->w, NULL, NULL);
  return c;
}
void orc_cells_destroy(orc_cells *c) { free(c); }

/* apply M (rows x cols, row-major) along direction dir of a tensor with
 * extents ext[3] (x fastest); output extents replace ext[dir] by rows.
 * transpose: apply M^T (then M is cols x rows as stored, rows_out = cols). */
static inline void sweep(const double *M, int rows, int cols, int transpose, const double *in, double *out,
                         const int *ext, int dir)
{
  int nin = ext[dir];
  int nout = transpose ? cols : rows;
  int stride = 1;
  for (int d = 0; d < dir; ++d) stride *= ext[d];
  int outer = 1;
  for (int d = dir + 1; d < 3; ++d) outer *= ext[d];
  for (int o = 0; o < outer; ++o)
    for (int s = 0; s < stride; ++s) {
      const double *pi = in + (size_t)o * nin * stride + s;
      double *po = out + (size_t)o * nout * stride + s;
      for (int r = 0; r < nout; ++r) {
        double acc = 0.0;
        for (int k = 0; k < nin; ++k) acc += (transpose ? M[k * cols + r] : M[r * cols + k]) * pi[k * stride];
        po[r * stride] = acc;
      }
    }
}

#define MAXP3 (MAXN * MAXN * MAXN)

/* phi.evaluate(values|gradients): u[n^3] -> val[nq3], G[3][nq3] (reference
 * gradients d/dxi_f), collocation-derivative variant */
static void evaluate(const orc_cells *cd, const double *u, double *val, double *G)
{
  int n = cd->n, nq = cd->nq;
  double t1[MAXP3], t2[MAXP3];
  int e0[3] = {n, n, n}, e1[3] = {nq, n, n}, e2[3] = {nq, nq, n}, e3[3] = {nq, nq, nq};
  sweep(cd->S, nq, n, 0, u, t1, e0, 0);
  sweep(cd->S, nq, n, 0, t1, t2, e1, 1);
  sweep(cd->S, nq, n, 0, t2, val, e2, 2);
  if (G) for (int f = 0; f < 3; ++f) sweep(cd->C, nq, nq, 0, val, G + (size_t)f * nq * nq * nq, e3, f);
}

/* phi.integrate(values|gradients): exact transpose */
static void integrate(const orc_cells *cd, const double *val, const double *G, double *r)
{
  int n = cd->n, nq = cd->nq, nq3 = nq * nq * nq;
  double acc[MAXP3], t1[MAXP3], t2[MAXP3];
  int e3[3] = {nq, nq, nq}, e2[3] = {nq, nq, nq}, e1[3] = {nq, nq, n}, e0[3] = {nq, n, n};
  (void)e2;
  if (val) memcpy(acc, val, sizeof(double) * nq3); else memset(acc, 0, sizeof(double) * nq3);
  if (G) for (int f = 0; f < 3; ++f) {
    sweep(cd->C, nq, nq, 1, G + (size_t)f * nq3, t1, e3, f);
    for (int q = 0; q < nq3; ++q) acc[q] += t1[q];
  }
  sweep(cd->S, nq, n, 1, acc, t1, e3, 2);
  sweep(cd->S, nq, n, 1, t1, t2, e1, 1);
  sweep(cd->S, nq, n, 1, t2, r, e0, 0);
}

/* ---------------------------- NeoHook ------------------------------ */

enum { LVL_SCALAR_REFERENTIAL = 0, LVL_TENSOR2 = 1, LVL_TENSOR4 = 2, LVL_TENSOR4_NS = 3 };

typedef struct {
  double kappa, c_1, mu, rho, rho_alpha, lambda; /* material.h:252-258 */
} orc_mat;

typedef struct {
  const orc_cells *cd;
  orc_mat mat[2]; /* [0] material, [1] material_inclusion */
  int level, formulation;
  /* tables, [cell][q] (mf_nh_operator.h:207-211) */
  double *cached_scalar, *cached_second_scalar, *cached_tensor2, *cached_tensor4, *cached_tensor4_ns;
  double *cached_position; /* scalar_referential: [cell][3 n^3] */
  /* MappingQEulerian metric (A.11): invjac_cur = F^-T invjac_ref, JxW_cur = J JxW_ref */
  double *invjac_cur, *jxw_cur;
  double *diagonal, *inverse_diagonal;
} orc_nh;

static void set_mat(orc_mat *m, double mu, double nu, double rho, double alpha_1)
{ /* material.h:78-94 */
  m->kappa = (2.0 * mu * (1.0 + nu)) / (3.0 * (1.0 - 2.0 * nu));
  m->c_1 = mu / 2.0; m->mu = mu; m->rho = rho; m->rho_alpha = rho * alpha_1;
  m->lambda = (2.0 * mu * nu) / (1.0 - 2.0 * nu);
}

orc_nh *orc_nh_create(const orc_cells *cd, double mu, double nu, double rho, double alpha_1, double mu_incl,
                      int level, int formulation)
{
  orc_nh *op = (orc_nh *)calloc(1, sizeof(orc_nh));
  op->cd = cd; op->level = level; op->formulation = formulation;
  set_mat(&op->mat[0], mu, nu, rho, alpha_1);
  set_mat(&op->mat[1], mu_incl, nu, rho, alpha_1);
  size_t nq3 = (size_t)cd->nq * cd->nq * cd->nq, N = (size_t)cd->ncell * nq3;
  /* initialize(): mf_nh_operator.h:346-387 */
  op->cached_scalar = (double *)calloc(N, sizeof(double));
  op->cached_second_scalar = (double *)calloc(N, sizeof(double));
  if (level == LVL_TENSOR2 || level == LVL_TENSOR4) op->cached_tensor2 = (double *)calloc(N * 9, sizeof(double));
  if (level == LVL_TENSOR4) op->cached_tensor4 = (double *)calloc(N * 36, sizeof(double));
  if (level == LVL_TENSOR4_NS) op->cached_tensor4_ns = (double *)calloc(N * 81, sizeof(double));
  if (level == LVL_SCALAR_REFERENTIAL)
    op->cached_position = (double *)calloc((size_t)cd->ncell * 3 * cd->n * cd->n * cd->n, sizeof(double));
  if (level == LVL_TENSOR2 || level == LVL_TENSOR4) {
    op->invjac_cur = (double *)calloc(N * 9, sizeof(double));
    op->jxw_cur = (double *)calloc(N, sizeof(double));
  }
  return op;
}

void orc_nh_destroy(orc_nh *op)
{
  free(op->cached_scalar); free(op->cached_second_scalar); free(op->cached_tensor2); free(op->cached_tensor4);
  free(op->cached_tensor4_ns); free(op->cached_position); free(op->invjac_cur); free(op->jxw_cur);
  free(op->diagonal); free(op->inverse_diagonal); free(op);
}

static const int VI[6] = {0, 1, 2, 0, 0, 1}, VJ[6] = {0, 1, 2, 1, 2, 2}; /* symmetric index pairs */

/* cache(): mf_nh_operator.h:392-523.  Returns number of q-points with det F <= 0 */
int64_t orc_nh_cache(orc_nh *op, const double *displacement)
{
  const orc_cells *cd = op->cd;
  const int n3 = cd->n * cd->n * cd->n, nq3 = cd->nq * cd->nq * cd->nq;
  int64_t bad = 0;
#pragma omp parallel for reduction(+ : bad)
  for (int64_t cell = 0; cell < cd->ncell; ++cell) {
    const orc_mat *m = &op->mat[(cd->material_id && cd->material_id[cell] != 0) ? 1 : 0]; /* :402-404 */
    double u[3][MAXP3], G[3][3 * MAXP3], val[MAXP3];
    for (int c = 0; c < 3; ++c) {
      for (int a = 0; a < n3; ++a) u[c][a] = displacement[cd->dof_idx[(cell * 3 + c) * n3 + a]]; /* read_dof_values_plain :410 */
      evaluate(cd, u[c], val, G[c]);
      if (op->level == LVL_SCALAR_REFERENTIAL)
        memcpy(op->cached_position + (cell * 3 + c) * n3, u[c], sizeof(double) * n3); /* :416-418 */
    }
    for (int q = 0; q < nq3; ++q) {
      const size_t iq = (size_t)cell * nq3 + q;
      const double *ij = cd->invjac + iq * 9;
      double F[9];
      for (int c = 0; c < 3; ++c)
        for (int e = 0; e < 3; ++e) {
          double s = 0.0;
          for (int f = 0; f < 3; ++f) s += ij[e * 3 + f] * G[c][f * nq3 + q]; /* get_gradient */
          F[c * 3 + e] = s + (c == e ? 1.0 : 0.0);                         /* Kinematics::F */
        }
      double Finv[9], detF;
      inv3(F, Finv, &detF);
      if (!(detF > 0)) ++bad; /* :430-435 */
      if (op->invjac_cur) {  /* deformed-frame metric */
        double *o = op->invjac_cur + iq * 9;
        for (int e = 0; e < 3; ++e) for (int f = 0; f < 3; ++f) {
          double s = 0.0;
          for (int mm = 0; mm < 3; ++mm) s += Finv[mm * 3 + e] * ij[mm * 3 + f];
          o[e * 3 + f] = s;
        }
        op->jxw_cur[iq] = detF * cd->jxw[iq];
      }
      if (op->formulation == 0) { /* :419-439 */
        op->cached_scalar[iq] = pow(detF, -1.0 / 3.0);
        memcpy(op->cached_tensor2 + iq * 9, F, sizeof(F));
        continue;
      }
      const double scalar = m->mu - 2.0 * m->lambda * log(detF); /* :455-456 */
      double tau[9];
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double b = 0.0;
        for (int k = 0; k < 3; ++k) b += F[i * 3 + k] * F[j * 3 + k];
        tau[i * 3 + j] = m->mu * b - (i == j ? scalar : 0.0);
      }
      switch (op->level) {
      case LVL_SCALAR_REFERENTIAL: op->cached_scalar[iq] = scalar; break;
      case LVL_TENSOR2:
        op->cached_scalar[iq] = scalar * 2.0 / detF;
        op->cached_second_scalar[iq] = 1.0 / detF;
        for (int i = 0; i < 9; ++i) op->cached_tensor2[iq * 9 + i] = tau[i] / detF;
        break;
      case LVL_TENSOR4: {
        op->cached_second_scalar[iq] = 1.0 / detF;
        for (int i = 0; i < 9; ++i) op->cached_tensor2[iq * 9 + i] = tau[i] / detF;
        double *C4 = op->cached_tensor4 + iq * 36; /* :489-493 */
        const double a = scalar * 2.0 / detF, b = m->lambda * 2.0 / detF;
        for (int I = 0; I < 6; ++I) for (int J = 0; J < 6; ++J) {
          int i = VI[I], j = VJ[I], k = VI[J], l = VJ[J];
          double S4 = 0.5 * ((i == k && j == l) + (i == l && j == k));
          double IxI = (i == j && k == l);
          C4[I * 6 + J] = a * S4 + b * IxI;
        }
        break;
      }
      case LVL_TENSOR4_NS: { /* :497-514 */
        double *A = op->cached_tensor4_ns + iq * 81;
        const double lnJ = log(detF);
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) for (int l = 0; l < 3; ++l) {
          /* F_inv_t[i][j] = Finv[j][i] */
          double t1 = Finv[j * 3 + i] * Finv[l * 3 + k];          /* outer_product(F^-T, F^-T) */
          double t2 = Finv[l * 3 + i] * Finv[j * 3 + k];          /* outer_product_iljk(F^-T, F^-1): A_il B_jk */
          double t3 = (i == k && j == l);                         /* IxI_ikjl */
          A[((i * 3 + j) * 3 + k) * 3 + l] = 2.0 * m->lambda * t1 + (m->mu - 2.0 * m->lambda * lnJ) * t2 + m->mu * t3;
        }
        break;
      }
      }
    }
  }
  return bad;
}

/* do_operation_on_cell (mf_nh_operator.h:651-1019) on local values u[3][n3] -> r[3][n3] */
static void nh_cell_op(const orc_nh *op, int64_t cell, double (*u)[MAXP3], double (*r)[MAXP3])
{
  const orc_cells *cd = op->cd;
  const int nq3 = cd->nq * cd->nq * cd->nq, n3 = cd->n * cd->n * cd->n;
  const orc_mat *m = &op->mat[(cd->material_id && cd->material_id[cell] != 0) ? 1 : 0]; /* :656-658 */
  const int deformed = (op->level == LVL_TENSOR2 || op->level == LVL_TENSOR4);
  const double *invjac = (deformed ? op->invjac_cur : cd->invjac) + (size_t)cell * nq3 * 9;
  const double *jxw = (deformed ? op->jxw_cur : cd->jxw) + (size_t)cell * nq3;
  double val[3][MAXP3], G[3][3 * MAXP3], Gu[3][3 * MAXP3], tmpv[MAXP3];
  for (int c = 0; c < 3; ++c) evaluate(cd, u[c], val[c], G[c]); /* :697 */
  if (op->formulation == 1 && op->level == LVL_SCALAR_REFERENTIAL)
    for (int c = 0; c < 3; ++c) /* second evaluate :861-867 */
      evaluate(cd, op->cached_position + ((size_t)cell * 3 + c) * n3, tmpv, Gu[c]);

  for (int q = 0; q < nq3; ++q) {
    const size_t iq = (size_t)cell * nq3 + q;
    const double *ij = invjac + q * 9;
    double g[9], sub_g[9], sub_v[3];
    for (int c = 0; c < 3; ++c) for (int e = 0; e < 3; ++e) {
      double s = 0.0;
      for (int f = 0; f < 3; ++f) s += ij[e * 3 + f] * G[c][f * nq3 + q];
      g[c * 3 + e] = s;
    }
    if (op->formulation == 0) { /* :701-840 */
      const double *F = op->cached_tensor2 + iq * 9;
      const double detF = det3(F), cs = op->cached_scalar[iq];
      double tau_bar[9], tau[9], s[9], jc[9], tau_iso[9];
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double b = 0.0;
        for (int k = 0; k < 3; ++k) b += (F[i * 3 + k] * cs) * (F[j * 3 + k] * cs);
        tau_bar[i * 3 + j] = b * (2.0 * m->c_1);
        s[i * 3 + j] = 0.5 * (g[i * 3 + j] + g[j * 3 + i]);
      }
      const double tr_tau_bar = tau_bar[0] + tau_bar[4] + tau_bar[8];
      const double tr_tau_bar_dim = tr_tau_bar * (1.0 / 3.0);
      const double dPsi = (m->kappa / 2.0) * (detF - 1.0 / detF);
      const double dPsiJ = dPsi * detF;
      const double d2Psi = (m->kappa / 2.0) * (1.0 + 1.0 / (detF * detF));
      const double tr = s[0] + s[4] + s[8];
      const double tmp = detF * (dPsi + detF * d2Psi) * tr;
      double tis = 0.0;
      for (int i = 0; i < 9; ++i) {
        int dg = (i % 4 == 0);
        tau[i] = tau_bar[i] + (dg ? dPsiJ - tr_tau_bar_dim : 0.0);
        tau_iso[i] = tau_bar[i] - (dg ? tr_tau_bar_dim : 0.0);
      }
      for (int i = 0; i < 9; ++i) tis += tau_iso[i] * s[i];
      for (int i = 0; i < 9; ++i) {
        int dg = (i % 4 == 0);
        double dev = s[i] - (dg ? tr * (1.0 / 3.0) : 0.0);
        jc[i] = s[i] * (-dPsiJ * 2.0) + (dg ? tmp : 0.0) + ((2.0 / 3.0) * tr_tau_bar) * dev -
                ((2.0 / 3.0) * tr) * tau_iso[i] - (dg ? (2.0 / 3.0) * tis : 0.0);
      }
      const double inv = 1.0 / detF;
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double geo = 0.0;
        for (int k = 0; k < 3; ++k) geo += g[i * 3 + k] * tau[k * 3 + j];
        sub_g[i * 3 + j] = (jc[i * 3 + j] + geo) * inv;
      }
      for (int c = 0; c < 3; ++c) sub_v[c] = val[c][q] * m->rho_alpha * inv;
    } else switch (op->level) {
    case LVL_SCALAR_REFERENTIAL: { /* :848-943 */
      double F[9], Finv[9], detF, tau[9], gx[9], jc[9], queued[9];
      for (int c = 0; c < 3; ++c) for (int e = 0; e < 3; ++e) {
        double s = 0.0;
        for (int f = 0; f < 3; ++f) s += ij[e * 3 + f] * Gu[c][f * nq3 + q];
        F[c * 3 + e] = s + (c == e ? 1.0 : 0.0);
      }
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double b = 0.0;
        for (int k = 0; k < 3; ++k) b += F[i * 3 + k] * F[j * 3 + k];
        tau[i * 3 + j] = m->mu * b - (i == j ? op->cached_scalar[iq] : 0.0);
      }
      inv3(F, Finv, &detF);
      for (int d = 0; d < 3; ++d) for (int e = 0; e < 3; ++e) {
        double s = 0.0;
        for (int f = 0; f < 3; ++f) s += Finv[f * 3 + e] * g[d * 3 + f]; /* :910-922 */
        gx[d * 3 + e] = s;
      }
      const double trg = gx[0] + gx[4] + gx[8];
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
        jc[i * 3 + j] = (2.0 * op->cached_scalar[iq]) * 0.5 * (gx[i * 3 + j] + gx[j * 3 + i]) +
                        (i == j ? 2.0 * m->lambda * trg : 0.0);
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double s = 0.0;
        for (int k = 0; k < 3; ++k) s += gx[i * 3 + k] * tau[k * 3 + j];
        queued[i * 3 + j] = jc[i * 3 + j] + s;
      }
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double s = 0.0;
        for (int k = 0; k < 3; ++k) s += queued[i * 3 + k] * Finv[j * 3 + k]; /* * transpose(F_inv) :938 */
        sub_g[i * 3 + j] = s;
      }
      for (int c = 0; c < 3; ++c) sub_v[c] = val[c][q] * m->rho_alpha;
      break;
    }
    case LVL_TENSOR2: { /* :945-974 */
      const double c1 = op->cached_scalar[iq], c2 = op->cached_second_scalar[iq];
      const double *T = op->cached_tensor2 + iq * 9;
      const double tmp = 2.0 * m->lambda * c2 * (g[0] + g[4] + g[8]);
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double s = 0.0;
        for (int k = 0; k < 3; ++k) s += g[i * 3 + k] * T[k * 3 + j];
        sub_g[i * 3 + j] = c1 * 0.5 * (g[i * 3 + j] + g[j * 3 + i]) + (i == j ? tmp : 0.0) + s;
      }
      for (int c = 0; c < 3; ++c) sub_v[c] = val[c][q] * m->rho_alpha * c2;
      break;
    }
    case LVL_TENSOR4: { /* :976-994 */
      const double c2 = op->cached_second_scalar[iq];
      const double *T = op->cached_tensor2 + iq * 9, *C4 = op->cached_tensor4 + iq * 36;
      double sv[6];
      for (int I = 0; I < 6; ++I) sv[I] = 0.5 * (g[VI[I] * 3 + VJ[I]] + g[VJ[I] * 3 + VI[I]]);
      for (int I = 0; I < 6; ++I) {
        double s = 0.0;
        for (int J = 0; J < 6; ++J) s += C4[I * 6 + J] * sv[J] * (J < 3 ? 1.0 : 2.0);
        sub_g[VI[I] * 3 + VJ[I]] = s; sub_g[VJ[I] * 3 + VI[I]] = s;
      }
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double s = 0.0;
        for (int k = 0; k < 3; ++k) s += g[i * 3 + k] * T[k * 3 + j];
        sub_g[i * 3 + j] += s;
      }
      for (int c = 0; c < 3; ++c) sub_v[c] = val[c][q] * m->rho_alpha * c2;
      break;
    }
    default: { /* LVL_TENSOR4_NS :996-1010 */
      const double *A = op->cached_tensor4_ns + iq * 81;
      for (int i = 0; i < 9; ++i) {
        double s = 0.0;
        for (int k = 0; k < 9; ++k) s += A[i * 9 + k] * g[k]; /* double_contract<2,0,3,1> */
        sub_g[i] = s;
      }
      for (int c = 0; c < 3; ++c) sub_v[c] = val[c][q] * m->rho_alpha;
    }
    }
    /* submit_gradient / submit_value */
    for (int c = 0; c < 3; ++c) {
      for (int f = 0; f < 3; ++f) {
        double s = 0.0;
        for (int e = 0; e < 3; ++e) s += ij[e * 3 + f] * sub_g[c * 3 + e];
        G[c][f * nq3 + q] = s * jxw[q];
      }
      val[c][q] = sub_v[c] * jxw[q];
    }
  }
  for (int c = 0; c < 3; ++c) integrate(cd, val[c], G[c], r[c]); /* :1018 */
}

/* vmult: dst = 0; vmult_add (:543-625), single rank */
void orc_nh_vmult(const orc_nh *op, double *dst, const double *src)
{
  const orc_cells *cd = op->cd;
  const int n3 = cd->n * cd->n * cd->n;
  memset(dst, 0, sizeof(double) * cd->ndofs);
#pragma omp parallel for schedule(static)
  for (int64_t cell = 0; cell < cd->ncell; ++cell) {
    double u[3][MAXP3], r[3][MAXP3];
    const uint32_t *idx = cd->dof_idx + cell * 3 * n3;
    for (int c = 0; c < 3; ++c) for (int a = 0; a < n3; ++a) {
      uint32_t d = idx[c * n3 + a];
      u[c][a] = cd->constrained[d] ? 0.0 : src[d]; /* read_dof_values, A.3 */
    }
    nh_cell_op(op, cell, u, r);
    for (int c = 0; c < 3; ++c) for (int a = 0; a < n3; ++a) {
      uint32_t d = idx[c * n3 + a];
      if (!cd->constrained[d]) {
#pragma omp atomic
        dst[d] += r[c][a];
      }
    }
  }
  for (int64_t i = 0; i < cd->ndofs; ++i) if (cd->constrained[i]) dst[i] += src[i]; /* :623-624 */
}


/* ------------------------------------------------------------------------------------------------------------------
 * SIMD-over-cells variant of vmult for the benchmark level (tensor2 caching, formulation 1) - the way deal.II's
 * MatrixFree runs the reference's operator on a CPU: VL cells per batch, one SIMD lane per cell (VectorizedArray<double>),
 * every per-cell scalar of nh_cell_op becomes a lane array, tables stored [batch][q][field][lane].  Same arithmetic as
 * orc_nh_vmult (tests/test_oracle.py compares the two); used by bench.py's CPU arms so that the reported baseline is a
 * vectorised CPU implementation, not a scalar one.  GCC function multi-versioning picks AVX-512 / AVX2 / baseline code
 * at run time (the library is built on a machine that may differ from the GPU box's host).
 * Scatter: entries touched by the cells of one thread only are added plainly, the others atomically. */
#define VL 8
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define ORC_CLONES __attribute__((target_clones("avx512f", "avx2,fma", "default")))
#else
#define ORC_CLONES
#endif

typedef struct {
  int64_t nbatch;
  int nthreads;
  double *tab;        /* [batch][q][NFB][VL]: K_cur (9), JxW_cur, c1, c2, T (9) */
  uint8_t *shared;    /* [ndofs] 1: touched by cells of more than one thread */
} orc_simd;
enum { NFB = 21 };

static inline void sweep_v(const double *M, int rows, int cols, int transpose, const double *in, double *out, const int *ext, int dir)
{
  int nin = ext[dir];
  int nout = transpose ? cols : rows;
  int stride = 1;
  for (int d = 0; d < dir; ++d) stride *= ext[d];
  int outer = 1;
  for (int d = dir + 1; d < 3; ++d) outer *= ext[d];
  for (int o = 0; o < outer; ++o)
    for (int s = 0; s < stride; ++s) {
      const double *pi = in + ((size_t)o * nin * stride + s) * VL;
      double *po = out + ((size_t)o * nout * stride + s) * VL;
      for (int r = 0; r < nout; ++r) {
        double acc[VL];
        for (int l = 0; l < VL; ++l) acc[l] = 0.0;
        for (int k = 0; k < nin; ++k) {
          const double mk = transpose ? M[k * cols + r] : M[r * cols + k];
          const double *x = pi + (size_t)k * stride * VL;
#pragma omp simd
          for (int l = 0; l < VL; ++l) acc[l] += mk * x[l];
        }
        for (int l = 0; l < VL; ++l) po[(size_t)r * stride * VL + l] = acc[l];
      }
    }
}

ORC_CLONES
static void nh_batch_tensor2(const orc_nh *op, const double *tab, const double *lam2 /*[VL] 2 lambda*/, const double *rha /*[VL] rho alpha*/,
                             double (*u)[MAXP3 * VL], double (*r)[MAXP3 * VL], double *work)
{
  const orc_cells *cd = op->cd;
  const int n = cd->n, nq = cd->nq, nq3 = nq * nq * nq;
  double *val = work, *G = work + 3 * (size_t)MAXP3 * VL, *t1 = G + 9 * (size_t)MAXP3 * VL, *t2 = t1 + (size_t)MAXP3 * VL;
  int e0[3] = {n, n, n}, e1[3] = {nq, n, n}, e2[3] = {nq, nq, n}, e3[3] = {nq, nq, nq};
  for (int c = 0; c < 3; ++c) { /* evaluate(values | gradients) :697 */
    double *v = val + (size_t)c * MAXP3 * VL;
    sweep_v(cd->S, nq, n, 0, u[c], t1, e0, 0);
    sweep_v(cd->S, nq, n, 0, t1, t2, e1, 1);
    sweep_v(cd->S, nq, n, 0, t2, v, e2, 2);
    for (int f = 0; f < 3; ++f) sweep_v(cd->C, nq, nq, 0, v, G + ((size_t)c * 3 + f) * MAXP3 * VL, e3, f);
  }
  for (int q = 0; q < nq3; ++q) { /* :945-974 */
    const double *t = tab + (size_t)q * NFB * VL;
    double g[9][VL], sg[9][VL];
    for (int c = 0; c < 3; ++c)
      for (int e = 0; e < 3; ++e) {
#pragma omp simd
        for (int l = 0; l < VL; ++l) {
          double s = 0.0;
          for (int f = 0; f < 3; ++f) s += t[(e * 3 + f) * VL + l] * G[(((size_t)c * 3 + f) * MAXP3 + q) * VL + l];
          g[c * 3 + e][l] = s;
        }
      }
    const double *jxw = t + 9 * VL, *c1 = t + 10 * VL, *c2 = t + 11 * VL, *T = t + 12 * VL;
    double tmp[VL];
#pragma omp simd
    for (int l = 0; l < VL; ++l) tmp[l] = lam2[l] * c2[l] * (g[0][l] + g[4][l] + g[8][l]);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
#pragma omp simd
        for (int l = 0; l < VL; ++l) {
          double s = 0.0;
          for (int k = 0; k < 3; ++k) s += g[i * 3 + k][l] * T[(k * 3 + j) * VL + l];
          sg[i * 3 + j][l] = c1[l] * 0.5 * (g[i * 3 + j][l] + g[j * 3 + i][l]) + (i == j ? tmp[l] : 0.0) + s;
        }
      }
    for (int c = 0; c < 3; ++c) {
      for (int f = 0; f < 3; ++f) {
        double *o = G + (((size_t)c * 3 + f) * MAXP3 + q) * VL;
#pragma omp simd
        for (int l = 0; l < VL; ++l) {
          double s = 0.0;
          for (int e = 0; e < 3; ++e) s += t[(e * 3 + f) * VL + l] * sg[c * 3 + e][l];
          o[l] = s * jxw[l];
        }
      }
      double *v = val + ((size_t)c * MAXP3 + q) * VL;
#pragma omp simd
      for (int l = 0; l < VL; ++l) v[l] = v[l] * rha[l] * c2[l] * jxw[l];
    }
  }
  for (int c = 0; c < 3; ++c) { /* integrate :1018 */
    double *acc = val + (size_t)c * MAXP3 * VL;
    for (int f = 0; f < 3; ++f) {
      sweep_v(cd->C, nq, nq, 1, G + ((size_t)c * 3 + f) * MAXP3 * VL, t1, e3, f);
      for (int i = 0; i < nq3 * VL; ++i) acc[i] += t1[i];
    }
    int x1[3] = {nq, nq, n}, x0[3] = {nq, n, n};
    sweep_v(cd->S, nq, n, 1, acc, t1, e3, 2);
    sweep_v(cd->S, nq, n, 1, t1, t2, x1, 1);
    sweep_v(cd->S, nq, n, 1, t2, r[c], x0, 0);
  }
}

/* batched tables + shared-entry flags; call after orc_nh_cache (tensor2, formulation 1, one material) */
orc_simd *orc_nh_simd_create(const orc_nh *op)
{
  const orc_cells *cd = op->cd;
  if (op->level != LVL_TENSOR2 || op->formulation != 1) return NULL;
  const int nq3 = cd->nq * cd->nq * cd->nq, n3 = cd->n * cd->n * cd->n;
  orc_simd *s = (orc_simd *)calloc(1, sizeof(orc_simd));
  s->nbatch = (cd->ncell + VL - 1) / VL;
#ifdef _OPENMP
  s->nthreads = omp_get_max_threads();
#else
  s->nthreads = 1;
#endif
  s->tab = (double *)calloc((size_t)s->nbatch * nq3 * NFB * VL, sizeof(double));
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < s->nbatch; ++b)
    for (int l = 0; l < VL; ++l) {
      int64_t cell = b * VL + l;
      if (cell >= cd->ncell) cell = cd->ncell - 1; /* tail lanes repeat the last cell; their result is dropped */
      for (int q = 0; q < nq3; ++q) {
        const size_t iq = (size_t)cell * nq3 + q;
        double *t = s->tab + ((size_t)b * nq3 + q) * NFB * VL + l;
        for (int i = 0; i < 9; ++i) t[i * VL] = op->invjac_cur[iq * 9 + i];
        t[9 * VL] = op->jxw_cur[iq];
        t[10 * VL] = op->cached_scalar[iq];
        t[11 * VL] = op->cached_second_scalar[iq];
        for (int i = 0; i < 9; ++i) t[(12 + i) * VL] = op->cached_tensor2[iq * 9 + i];
      }
    }
  /* entries touched by batches of different threads (static schedule: thread = batch / chunk) */
  s->shared = (uint8_t *)calloc(cd->ndofs, 1);
  int32_t *first = (int32_t *)malloc(sizeof(int32_t) * cd->ndofs);
  for (int64_t i = 0; i < cd->ndofs; ++i) first[i] = -1;
  const int64_t chunk = (s->nbatch + s->nthreads - 1) / s->nthreads;
  for (int64_t cell = 0; cell < cd->ncell; ++cell) {
    const int32_t th = (int32_t)((cell / VL) / chunk);
    const uint32_t *idx = cd->dof_idx + cell * 3 * n3;
    for (int a = 0; a < 3 * n3; ++a) {
      if (first[idx[a]] < 0) first[idx[a]] = th;
      else if (first[idx[a]] != th) s->shared[idx[a]] = 1;
    }
  }
  free(first);
  return s;
}
void orc_nh_simd_destroy(orc_simd *s) { if (s) { free(s->tab); free(s->shared); free(s); } }

void orc_nh_vmult_simd(const orc_nh *op, const orc_simd *s, double *dst, const double *src)
{
  const orc_cells *cd = op->cd;
  const int n3 = cd->n * cd->n * cd->n, nq3 = cd->nq * cd->nq * cd->nq;
  memset(dst, 0, sizeof(double) * cd->ndofs);
  const int64_t chunk = (s->nbatch + s->nthreads - 1) / s->nthreads;
#pragma omp parallel num_threads(s->nthreads)
  {
    double (*u)[MAXP3 * VL] = (double (*)[MAXP3 * VL])aligned_alloc(64, sizeof(double) * 3 * MAXP3 * VL);
    double (*r)[MAXP3 * VL] = (double (*)[MAXP3 * VL])aligned_alloc(64, sizeof(double) * 3 * MAXP3 * VL);
    double *work = (double *)aligned_alloc(64, sizeof(double) * 14 * MAXP3 * VL);
#ifdef _OPENMP
    const int th = omp_get_thread_num();
#else
    const int th = 0;
#endif
    const int64_t b0 = th * chunk, b1 = (b0 + chunk < s->nbatch) ? b0 + chunk : s->nbatch;
    for (int64_t b = b0; b < b1; ++b) {
      for (int l = 0; l < VL; ++l) {
        int64_t cell = b * VL + l;
        if (cell >= cd->ncell) cell = cd->ncell - 1;
        const uint32_t *idx = cd->dof_idx + cell * 3 * n3;
        for (int c = 0; c < 3; ++c)
          for (int a = 0; a < n3; ++a) {
            const uint32_t d = idx[c * n3 + a];
            u[c][(size_t)a * VL + l] = cd->constrained[d] ? 0.0 : src[d];
          }
      }
      double lam2[VL], rha[VL];
      for (int l = 0; l < VL; ++l) { /* material per cell (:656-658), one lane each */
        int64_t cell = b * VL + l;
        if (cell >= cd->ncell) cell = cd->ncell - 1;
        const orc_mat *m = &op->mat[(cd->material_id && cd->material_id[cell] != 0) ? 1 : 0];
        lam2[l] = 2.0 * m->lambda;
        rha[l] = m->rho_alpha;
      }
      nh_batch_tensor2(op, s->tab + (size_t)b * nq3 * NFB * VL, lam2, rha, u, r, work);
      for (int l = 0; l < VL; ++l) {
        const int64_t cell = b * VL + l;
        if (cell >= cd->ncell) break;
        const uint32_t *idx = cd->dof_idx + cell * 3 * n3;
        for (int c = 0; c < 3; ++c)
          for (int a = 0; a < n3; ++a) {
            const uint32_t d = idx[c * n3 + a];
            if (cd->constrained[d]) continue;
            const double v = r[c][(size_t)a * VL + l];
            if (s->shared[d]) {
#pragma omp atomic
              dst[d] += v;
            } else {
              dst[d] += v;
            }
          }
      }
    }
    free(u); free(r); free(work);
  }
  for (int64_t i = 0; i < cd->ndofs; ++i) if (cd->constrained[i]) dst[i] += src[i]; /* :623-624 */
}

/* which clone the dispatcher picks on this machine (reported next to the CPU baseline) */
const char *orc_simd_isa(void)
{
#if defined(__x86_64__) && defined(__GNUC__)
  __builtin_cpu_init();
  if (__builtin_cpu_supports("avx512f")) return "avx512f (8 doubles per vector)";
  if (__builtin_cpu_supports("avx2")) return "avx2+fma (4 doubles per vector)";
  return "baseline x86-64 (as compiled)";
#else
  return "scalar";
#endif
}

/* compute_diagonal (:1023-1060) via MatrixFreeTools::compute_diagonal semantics (A.4) */
void orc_nh_compute_diagonal(orc_nh *op, double *diag, double *inv_diag)
{
  const orc_cells *cd = op->cd;
  const int n3 = cd->n * cd->n * cd->n;
  memset(diag, 0, sizeof(double) * cd->ndofs);
#pragma omp parallel for schedule(static)
  for (int64_t cell = 0; cell < cd->ncell; ++cell) {
    double u[3][MAXP3], r[3][MAXP3];
    const uint32_t *idx = cd->dof_idx + cell * 3 * n3;
    for (int i = 0; i < 3 * n3; ++i) {
      memset(u, 0, sizeof(u));
      u[i / n3][i % n3] = 1.0;
      nh_cell_op(op, cell, u, r);
      uint32_t d = idx[i];
      if (!cd->constrained[d]) {
#pragma omp atomic
        diag[d] += r[i / n3][i % n3];
      }
    }
  }
  for (int64_t i = 0; i < cd->ndofs; ++i) {
    if (cd->constrained[i]) diag[i] = 1.0; /* :1043-1044 */
    inv_diag[i] = 1.0 / diag[i];
  }
}

/* ---------------------------- Laplace ------------------------------ */
/* (M + dt*coef*K) u, heat_transfer.h:234-250 */
static void laplace_cell_op(const orc_cells *cd, int64_t cell, double delta_t, const double *coef, const double *u, double *r)
{
  const int nq3 = cd->nq * cd->nq * cd->nq;
  double val[MAXP3], G[3 * MAXP3];
  evaluate(cd, u, val, G);
  for (int q = 0; q < nq3; ++q) {
    const size_t iq = (size_t)cell * nq3 + q;
    const double *ij = cd->invjac + iq * 9;
    const double jxw = cd->jxw[iq], k = (coef ? coef[iq] : 1.0) * delta_t;
    double g[3], o[3];
    for (int e = 0; e < 3; ++e) g[e] = ij[e * 3 + 0] * G[q] + ij[e * 3 + 1] * G[nq3 + q] + ij[e * 3 + 2] * G[2 * nq3 + q];
    for (int f = 0; f < 3; ++f) o[f] = (ij[0 * 3 + f] * g[0] + ij[1 * 3 + f] * g[1] + ij[2 * 3 + f] * g[2]) * k * jxw;
    for (int f = 0; f < 3; ++f) G[f * nq3 + q] = o[f];
    val[q] *= jxw;
  }
  integrate(cd, val, G, r);
}

void orc_laplace_vmult(const orc_cells *cd, double delta_t, const double *coef, double *dst, const double *src)
{
  const int n3 = cd->n * cd->n * cd->n;
  memset(dst, 0, sizeof(double) * cd->ndofs);
#pragma omp parallel for schedule(static)
  for (int64_t cell = 0; cell < cd->ncell; ++cell) {
    double u[MAXP3], r[MAXP3];
    const uint32_t *idx = cd->dof_idx + cell * n3;
    for (int a = 0; a < n3; ++a) u[a] = cd->constrained[idx[a]] ? 0.0 : src[idx[a]];
    laplace_cell_op(cd, cell, delta_t, coef, u, r);
    for (int a = 0; a < n3; ++a) if (!cd->constrained[idx[a]]) {
#pragma omp atomic
      dst[idx[a]] += r[a];
    }
  }
  for (int64_t i = 0; i < cd->ndofs; ++i) if (cd->constrained[i]) dst[i] += src[i];
}

void orc_laplace_compute_diagonal(const orc_cells *cd, double delta_t, const double *coef, double *diag, double *inv_diag)
{
  const int n3 = cd->n * cd->n * cd->n;
  memset(diag, 0, sizeof(double) * cd->ndofs);
#pragma omp parallel for schedule(static)
  for (int64_t cell = 0; cell < cd->ncell; ++cell) {
    double u[MAXP3], r[MAXP3];
    const uint32_t *idx = cd->dof_idx + cell * n3;
    for (int i = 0; i < n3; ++i) {
      memset(u, 0, sizeof(u));
      u[i] = 1.0;
      laplace_cell_op(cd, cell, delta_t, coef, u, r);
      if (!cd->constrained[idx[i]]) {
#pragma omp atomic
        diag[idx[i]] += r[i];
      }
    }
  }
  for (int64_t i = 0; i < cd->ndofs; ++i) {
    if (cd->constrained[i]) diag[i] = 1.0; /* heat_transfer.h:220 */
    inv_diag[i] = 1.0 / diag[i];
  }
}

int orc_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}


/* Host memory bandwidth, STREAM triad convention (a = b + s c, 24 B per element counted), all OpenMP threads, first-touch
 * placement by the same static schedule; best of `reps` passes in GB/s.  Reported next to the CPU arm's algorithmic GB/s so that
 * the arm can be judged against what the host's memory system delivers (bench.py cpu_baseline). */
double orc_stream_triad_gbs(long n, int reps)
{
  double *a = (double *)malloc((size_t)n * sizeof(double)), *b = (double *)malloc((size_t)n * sizeof(double)),
         *c = (double *)malloc((size_t)n * sizeof(double));
  if (!a || !b || !c) {
    free(a);
    free(b);
    free(c);
    return -1.0;
  }
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    a[i] = 0.0;
    b[i] = 1.0 + 1e-9 * (double)i;
    c[i] = 2.0;
  }
  double best = 0.0;
  for (int r = 0; r < reps; ++r) {
    const double t0 = omp_get_wtime();
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) a[i] = b[i] + 3.0 * c[i];
    const double t1 = omp_get_wtime();
    const double gbs = 24.0 * (double)n / (t1 - t0) / 1e9;
    if (gbs > best) best = gbs;
  }
  volatile double sink = a[n / 2];
  (void)sink;
  free(a);
  free(b);
  free(c);
  return best;
}
