/* oracle/dogleg_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * CPU restatement of libdogleg's hot path; see dogleg_oracle.h. Scalar,
 * sequential-order double arithmetic exactly as the reference accumulates it,
 * so that oracle-vs-reference differences stay at round-off of the Cholesky
 * back-end only.
 */
#include "dogleg_oracle.h"
#include "sparse_chol_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdio.h>

/* ================================================================ kernels */

double orc_norm2(const double* v, int n)
{
  double s = 0.0;
  for(int i = 0; i < n; i++) s += v[i] * v[i];
  return s;
}
double orc_inner(const double* a, const double* b, int n)
{
  double s = 0.0;
  for(int i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}

/* a2: Jt_x = Jt * x, measurements visited in order, scatter-add per entry */
void orc_Jt_times_x(double* Jt_x, int Nstate, int Nmeas,
                    const int* Jp, const int* Ji, const double* Jx, const double* x)
{
  for(int k = 0; k < Nstate; k++) Jt_x[k] = 0.0;
  for(int j = 0; j < Nmeas; j++)
    for(int q = Jp[j]; q < Jp[j+1]; q++)
      Jt_x[Ji[q]] += x[j] * Jx[q];
}

/* a5: sum over measurements of (gradient_j . v)^2 */
double orc_norm2_J_times_v(int Nmeas, const int* Jp, const int* Ji,
                           const double* Jx, const double* v)
{
  double total = 0.0;
  for(int j = 0; j < Nmeas; j++)
  {
    double dot = 0.0;
    for(int q = Jp[j]; q < Jp[j+1]; q++) dot += v[Ji[q]] * Jx[q];
    total += dot * dot;
  }
  return total;
}

/* a16 */
void orc_dense_Jt_times_x(double* Jt_x, const double* J, const double* x, int Nmeas, int Nstate)
{
  for(int k = 0; k < Nstate; k++)
  {
    double s = 0.0;
    for(int i = 0; i < Nmeas; i++) s += J[(size_t)i * Nstate + k] * x[i];
    Jt_x[k] = s;
  }
}
double orc_dense_norm2_J_times_v(const double* J, const double* v, int Nmeas, int Nstate)
{
  double total = 0.0;
  for(int i = 0; i < Nmeas; i++)
  {
    double dot = orc_inner(v, J + (size_t)i * Nstate, Nstate);
    total += dot * dot;
  }
  return total;
}

/* a17: rank-1 accumulation into row-first packed upper storage, then +lambda */
void orc_dense_JtJ_packed_upper(double* JtJ, const double* J, int Nmeas, int Nstate, double lambda)
{
  const size_t sz = (size_t)Nstate * (Nstate + 1) / 2;
  memset(JtJ, 0, sz * sizeof(double));
  for(int i = 0; i < Nmeas; i++)
  {
    const double* g = J + (size_t)i * Nstate;
    size_t at = 0;
    for(int r = 0; r < Nstate; r++)
      for(int c = r; c < Nstate; c++, at++)
        JtJ[at] += g[c] * g[r];
  }
  if(lambda > 0.0)
  {
    size_t at = 0;
    for(int r = 0; r < Nstate; r++) { JtJ[at] += lambda; at += Nstate - r; }
  }
}

/* a19 */
double orc_xt_Apacked_upper_x(const double* v, const double* A, int N)
{
  double s = 0.0;
  size_t at = 0;
  for(int r = 0; r < N; r++)
  {
    s += A[at++] * v[r] * v[r];
    for(int c = r + 1; c < N; c++, at++) s += 2. * A[at] * v[c] * v[r];
  }
  return s;
}
double orc_xt_A_x(const double* v, const double* A, int N)
{
  double s = 0.0;
  for(int r = 0; r < N; r++)
    for(int c = 0; c < N; c++)
      s += A[(size_t)r * N + c] * v[r] * v[c];
  return s;
}

void orc_sparse_JtJ_dense(double* JtJ, int Nstate, int Nmeas,
                          const int* Jp, const int* Ji, const double* Jx, double lambda)
{
  memset(JtJ, 0, (size_t)Nstate * Nstate * sizeof(double));
  for(int j = 0; j < Nmeas; j++)
    for(int a = Jp[j]; a < Jp[j+1]; a++)
      for(int b = Jp[j]; b < Jp[j+1]; b++)
        JtJ[(size_t)Ji[a] * Nstate + Ji[b]] += Jx[a] * Jx[b];
  for(int k = 0; k < Nstate; k++) JtJ[(size_t)k * Nstate + k] += lambda;
}

/* ------------------------------------------------------ dense Cholesky */

/* index of (row r, col c>=r) in row-first packed upper == column-major packed
 * lower element (c, r) */
static inline size_t pu(int r, int c, int n) { return (size_t)r * n - (size_t)r * (r - 1) / 2 + (c - r); }

/* LAPACK dpptrf 'L' (unblocked, right-looking): for each column scale by the
 * square root of the pivot then rank-1 update the trailing packed triangle */
int orc_pptrf_lower(double* ap, int n)
{
  for(int j = 0; j < n; j++)
  {
    double ajj = ap[pu(j, j, n)];
    if(!(ajj > 0.0) || !isfinite(ajj)) return j + 1;
    ajj = sqrt(ajj);
    ap[pu(j, j, n)] = ajj;
    for(int i = j + 1; i < n; i++) ap[pu(j, i, n)] /= ajj;
    for(int c = j + 1; c < n; c++)
    {
      const double lcj = ap[pu(j, c, n)];
      for(int i = c; i < n; i++) ap[pu(c, i, n)] -= ap[pu(j, i, n)] * lcj;
    }
  }
  return 0;
}
void orc_pptrs_lower(const double* ap, int n, double* b)
{
  for(int j = 0; j < n; j++)
  {
    b[j] /= ap[pu(j, j, n)];
    for(int i = j + 1; i < n; i++) b[i] -= ap[pu(j, i, n)] * b[j];
  }
  for(int j = n - 1; j >= 0; j--)
  {
    for(int i = j + 1; i < n; i++) b[j] -= ap[pu(j, i, n)] * b[i];
    b[j] /= ap[pu(j, j, n)];
  }
}
/* full storage, row-first; factor kept where fortran 'L' puts it: element
 * (row r, col c) of the row-first array with c >= r */
int orc_potrf_rowfirst(double* a, int n)
{
  for(int j = 0; j < n; j++)
  {
    double ajj = a[(size_t)j * n + j];
    if(!(ajj > 0.0) || !isfinite(ajj)) return j + 1;
    ajj = sqrt(ajj);
    a[(size_t)j * n + j] = ajj;
    for(int i = j + 1; i < n; i++) a[(size_t)j * n + i] /= ajj;
    for(int c = j + 1; c < n; c++)
    {
      const double lcj = a[(size_t)j * n + c];
      for(int i = c; i < n; i++) a[(size_t)c * n + i] -= a[(size_t)j * n + i] * lcj;
    }
  }
  return 0;
}
void orc_potrs_rowfirst(const double* a, int n, double* b)
{
  for(int j = 0; j < n; j++)
  {
    b[j] /= a[(size_t)j * n + j];
    for(int i = j + 1; i < n; i++) b[i] -= a[(size_t)j * n + i] * b[j];
  }
  for(int j = n - 1; j >= 0; j--)
  {
    for(int i = j + 1; i < n; i++) b[j] -= a[(size_t)j * n + i] * b[i];
    b[j] /= a[(size_t)j * n + j];
  }
}

/* ============================================================ step logic */

/* a6: k = -|g|^2 / |J g|^2 ; cauchy = k g ; |cauchy|^2 = k^2 |g|^2 */
double orc_cauchy(double* updateCauchy, double* norm2_updateCauchy,
                  const double* Jt_x, double norm2_J_Jt_x, int Nstate)
{
  const double g2 = orc_norm2(Jt_x, Nstate);
  const double k  = -g2 / norm2_J_Jt_x;
  *norm2_updateCauchy = k * k * g2;
  for(int i = 0; i < Nstate; i++) updateCauchy[i] = k * Jt_x[i];
  return k;
}

/* a10: the point on the segment cauchy -> gn that sits on the trust-region
 * boundary. With d = a-b: l2 = |d|^2, neg_c = d.a,
 * k = (neg_c + sqrt(neg_c^2 - l2 (|a|^2 - delta^2))) / l2 */
double orc_interpolate(double* step, double* norm2_step,
                       const double* a, double norm2a,
                       const double* b, double trustregion, int Nstate)
{
  double l2 = 0.0, neg_c = 0.0;
  for(int i = 0; i < Nstate; i++)
  {
    const double d = a[i] - b[i];
    l2    += d * d;
    neg_c += d * a[i];
  }
  double disc = neg_c * neg_c - l2 * (norm2a - trustregion * trustregion);
  if(disc < 0.0) disc = 0.0;
  const double k = (neg_c + sqrt(disc)) / l2;
  double n2 = 0.0;
  for(int i = 0; i < Nstate; i++)
  {
    step[i] = a[i] + k * (b[i] - a[i]);
    n2 += step[i] * step[i];
  }
  *norm2_step = n2;
  return k;
}

/* =========================================================== whole solve */

enum { ST_DENSE = 0, ST_SPARSE = 1, ST_PRODUCTS = 2 };

typedef struct
{
  double* p; double* x; double norm2_x;
  cholmod_sparse Jt;            /* sparse */
  double* J;                    /* dense */
  double* JtJ;                  /* products */
  double* Jt_x;
  double* cauchy; double* gn; double* step_to_here;
  double  norm2_cauchy, norm2_gn, norm2_step_to_here;
  int have_cauchy, have_gn, have_factorization, edge;
} opoint;

typedef struct
{
  int type, N, M, nnz;
  dogleg_callback_t* f; dogleg_callback_dense_t* fd; dogleg_callback_dense_products_t* fp;
  void* cookie;
  const dogleg_parameters2_t* prm;
  double lambda;
  orc_factor* F;                /* sparse */
  double* fact;                 /* dense / products */
  orc_result_t* res;
} octx;

static opoint* point_new(const octx* c)
{
  opoint* pt = calloc(1, sizeof(*pt));
  const int N = c->N, M = c->M;
  pt->p = calloc(N, sizeof(double));
  pt->Jt_x = calloc(N, sizeof(double));
  pt->cauchy = calloc(N, sizeof(double));
  pt->gn = calloc(N, sizeof(double));
  pt->step_to_here = calloc(N, sizeof(double));
  if(c->type != ST_PRODUCTS) pt->x = calloc(M ? M : 1, sizeof(double));
  if(c->type == ST_SPARSE)
  {
    pt->Jt.nrow = N; pt->Jt.ncol = M; pt->Jt.nzmax = c->nnz;
    pt->Jt.p = calloc(M + 1, sizeof(int));
    pt->Jt.i = calloc(c->nnz, sizeof(int));
    pt->Jt.x = calloc(c->nnz, sizeof(double));
    pt->Jt.stype = 0; pt->Jt.itype = CHOLMOD_INT; pt->Jt.xtype = CHOLMOD_REAL;
    pt->Jt.dtype = CHOLMOD_DOUBLE; pt->Jt.sorted = 1; pt->Jt.packed = 1;
  }
  else if(c->type == ST_DENSE) pt->J = calloc((size_t)M * N, sizeof(double));
  else pt->JtJ = calloc((size_t)N * N, sizeof(double));
  return pt;
}
static void point_free(opoint* pt)
{
  free(pt->p); free(pt->x); free(pt->Jt_x); free(pt->cauchy); free(pt->gn);
  free(pt->step_to_here); free(pt->Jt.p); free(pt->Jt.i); free(pt->Jt.x);
  free(pt->J); free(pt->JtJ); free(pt);
}

/* a15: dogleg.c:1004-1083 */
static int evaluate(opoint* pt, octx* c)
{
  pt->norm2_x = -1.0;
  pt->have_cauchy = pt->have_gn = pt->have_factorization = pt->edge = 0;
  if(c->type == ST_SPARSE)
  {
    c->f(pt->p, pt->x, &pt->Jt, c->cookie);
    orc_Jt_times_x(pt->Jt_x, c->N, c->M, pt->Jt.p, pt->Jt.i, pt->Jt.x, pt->x);
    pt->norm2_x = orc_norm2(pt->x, c->M);
  }
  else if(c->type == ST_DENSE)
  {
    c->fd(pt->p, pt->x, pt->J, c->cookie);
    orc_dense_Jt_times_x(pt->Jt_x, pt->J, pt->x, c->M, c->N);
    pt->norm2_x = orc_norm2(pt->x, c->M);
  }
  else
    c->fp(pt->p, &pt->norm2_x, pt->Jt_x, pt->JtJ, c->cookie);

  for(int i = 0; i < c->N; i++)
    if(fabs(pt->Jt_x[i]) > c->prm->Jt_x_threshold) return 0;
  return 1;   /* gradient below threshold everywhere */
}

/* |J v|^2 for whichever representation is present; <0 on unsupported layout */
static double norm2_Jv(const opoint* pt, const octx* c, const double* v)
{
  if(c->type == ST_SPARSE)
    return orc_norm2_J_times_v(c->M, pt->Jt.p, pt->Jt.i, pt->Jt.x, v);
  if(c->type == ST_DENSE)
    return orc_dense_norm2_J_times_v(pt->J, v, c->M, c->N);
  if(c->prm->JtJ_packed && c->prm->JtJ_upper) return orc_xt_Apacked_upper_x(v, pt->JtJ, c->N);
  if(!c->prm->JtJ_packed)                     return orc_xt_A_x(v, pt->JtJ, c->N);
  return -1.0;
}

/* a7 / a17-a19: dogleg.c:634-820, lambda ladder included */
static int factorize(opoint* pt, octx* c)
{
  if(pt->have_factorization) return 1;
  const int N = c->N;
  if(c->type == ST_SPARSE)
  {
    if(!c->F) c->F = orc_analyze(N, c->M, pt->Jt.p, pt->Jt.i, c->res ? c->res->perm : NULL);
    for(;;)
    {
      orc_factorize(c->F, pt->Jt.p, pt->Jt.i, pt->Jt.x, c->lambda, c->res ? c->res->use_ll : 0);
      if(c->F->minor == N) break;
      c->lambda = c->lambda == 0.0 ? 1e-10 : c->lambda * 10.0;
      if(!isfinite(c->lambda)) return 0;
    }
  }
  else
  {
    const int packed = c->type == ST_DENSE || c->prm->JtJ_packed;
    const size_t sz = packed ? (size_t)N * (N + 1) / 2 : (size_t)N * N;
    if(!c->fact) c->fact = calloc(sz, sizeof(double));
    if(c->type == ST_PRODUCTS && packed && !c->prm->JtJ_upper) return 0; /* oracle: unsupported */
    for(;;)
    {
      if(c->type == ST_DENSE)
        orc_dense_JtJ_packed_upper(c->fact, pt->J, c->M, N, c->lambda);
      else
      {
        memcpy(c->fact, pt->JtJ, sz * sizeof(double));
        if(c->lambda > 0.0)
        {
          if(packed) { size_t at = 0; for(int r = 0; r < N; r++) { c->fact[at] += c->lambda; at += N - r; } }
          else       for(int r = 0; r < N; r++) c->fact[(size_t)r * (N + 1)] += c->lambda;
        }
      }
      const int info = packed ? orc_pptrf_lower(c->fact, N) : orc_potrf_rowfirst(c->fact, N);
      if(info == 0) break;
      c->lambda = c->lambda == 0.0 ? 1e-10 : c->lambda * 10.0;
      if(!isfinite(c->lambda)) return 0;
    }
  }
  pt->have_factorization = 1;
  return 1;
}

/* a8: dogleg.c:822-908: solve (JtJ + lambda I) u = Jt_x, gn = -u */
static int gauss_newton(opoint* pt, octx* c)
{
  if(pt->have_gn) return 1;
  if(!factorize(pt, c)) return 0;
  const int N = c->N;
  if(c->type == ST_SPARSE) orc_solve(c->F, pt->Jt_x, pt->gn, 1);
  else
  {
    memcpy(pt->gn, pt->Jt_x, N * sizeof(double));
    if(c->type == ST_DENSE || c->prm->JtJ_packed) orc_pptrs_lower(c->fact, N, pt->gn);
    else                                          orc_potrs_rowfirst(c->fact, N, pt->gn);
  }
  for(int i = 0; i < N; i++) pt->gn[i] *= -1.0;
  pt->norm2_gn = orc_norm2(pt->gn, N);
  pt->have_gn = 1;
  return 1;
}

/* a12 + a11: dogleg.c:1172-1297 and :1085-1165 */
static int take_step(double* expected, opoint* from, opoint* to, double delta, octx* c, orc_trial_t* rec)
{
  const int N = c->N;
  double* step = to->step_to_here;
  if(!from->have_cauchy)
  {
    const double jg2 = norm2_Jv(from, c, from->Jt_x);
    if(c->type == ST_PRODUCTS && jg2 < 0.0 && c->prm->JtJ_packed && !c->prm->JtJ_upper) return 0;
    orc_cauchy(from->cauchy, &from->norm2_cauchy, from->Jt_x, jg2, N);
    from->have_cauchy = 1;
  }
  rec->step_len_cauchy = sqrt(from->norm2_cauchy);

  if(from->norm2_cauchy >= delta * delta)
  {
    /* clipped steepest descent; note the UNCLIPPED length is what is recorded */
    rec->step_type = 0;
    to->norm2_step_to_here = from->norm2_cauchy;
    const double s = delta / sqrt(from->norm2_cauchy);
    for(int i = 0; i < N; i++) step[i] = s * from->cauchy[i];
    from->edge = 1;
  }
  else
  {
    if(!gauss_newton(from, c)) return 0;
    rec->step_len_gn = sqrt(from->norm2_gn);
    if(from->norm2_gn <= delta * delta)
    {
      rec->step_type = 1;
      to->norm2_step_to_here = from->norm2_gn;
      memcpy(step, from->gn, N * sizeof(double));
      from->edge = 0;
    }
    else
    {
      rec->step_type = 2;
      rec->k_cauchy_to_gn = orc_interpolate(step, &to->norm2_step_to_here,
                                            from->cauchy, from->norm2_cauchy, from->gn, delta, N);
      rec->step_len_interpolated = sqrt(to->norm2_step_to_here);
      from->edge = 1;
    }
  }
  rec->norm2_step = to->norm2_step_to_here;
  for(int i = 0; i < N; i++) to->p[i] = from->p[i] + step[i];

  *expected = -2.0 * orc_inner(from->Jt_x, step, N) - norm2_Jv(from, c, step);
  rec->expected_improvement = *expected;

  for(int i = 0; i < N; i++)
    if(fabs(step[i]) > c->prm->update_threshold) return 1;
  *expected = -1.0;     /* sentinel: step too small, we are done */
  return 1;
}

static orc_trial_t* next_record(octx* c, orc_trial_t* scratch)
{
  orc_trial_t* r = scratch;
  if(c->res && c->res->trials && c->res->Ntrials < c->res->Ntrials_max)
    r = &c->res->trials[c->res->Ntrials];
  memset(r, 0, sizeof(*r));
  r->norm2x_after = r->step_len_gn = r->step_len_interpolated = r->k_cauchy_to_gn = INFINITY;
  r->observed_improvement = r->rho = r->trustregion_after = INFINITY;
  return r;
}
static void commit_record(octx* c) { if(c->res) c->res->Ntrials++; }

/* a14 + a13: dogleg.c:1359-1476 and :1303-1356 */
static double run(double* p, octx* c)
{
  const dogleg_parameters2_t* P = c->prm;
  opoint* before = point_new(c);
  opoint* after  = point_new(c);
  memcpy(before->p, p, c->N * sizeof(double));
  if(c->res) c->res->Ntrials = 0;

  double delta = P->trustregion0;
  int steps = 0, ok = 1;
  int done = evaluate(before, c);

  while(ok && !done && steps < P->max_iterations)
  {
    for(;;)
    {
      orc_trial_t scratch, *rec = next_record(c, &scratch);
      rec->iteration = steps;
      rec->trustregion_before = delta;
      rec->norm2x_before = before->norm2_x;

      double expected;
      if(!take_step(&expected, before, after, delta, c, rec)) { ok = 0; break; }
      if(expected < 0.0) { rec->accepted = 1; commit_record(c); done = 1; break; }

      const int zero_gradient = evaluate(after, c);
      rec->norm2x_after = after->norm2_x;

      const double observed = before->norm2_x - after->norm2_x;
      const double rho = observed / expected;
      rec->observed_improvement = observed; rec->rho = rho;
      if(rho < P->trustregion_decrease_threshold)
      {
        if(!before->edge) delta = sqrt(before->norm2_gn);
        delta *= P->trustregion_decrease_factor;
      }
      else if(rho > P->trustregion_increase_threshold && before->edge)
        delta *= P->trustregion_increase_factor;
      rec->trustregion_after = delta;

      if(rho > 0.0)
      {
        rec->accepted = 1; commit_record(c);
        steps++;
        opoint* t = after; after = before; before = t;
        if(zero_gradient) done = 1;
        break;
      }
      rec->accepted = 0; commit_record(c);
      if(delta < P->trustregion_threshold) { done = 1; break; }
    }
  }

  double ret = -1.0;
  if(ok)
  {
    ret = before->norm2_x;
    memcpy(p, before->p, c->N * sizeof(double));
  }
  if(c->res) { c->res->norm2_x = ret; c->res->accepted_steps = steps; c->res->lambda = c->lambda; }
  point_free(before); point_free(after);
  orc_free(c->F); free(c->fact);
  return ret;
}

static dogleg_parameters2_t defaults(void)
{
  dogleg_parameters2_t d;
  memset(&d, 0, sizeof(d));
  d.max_iterations = 100; d.trustregion0 = 1.0e3;
  d.trustregion_decrease_factor = 0.1;  d.trustregion_decrease_threshold = 0.25;
  d.trustregion_increase_factor = 2;    d.trustregion_increase_threshold = 0.75;
  d.Jt_x_threshold = 1e-8; d.update_threshold = 1e-8; d.trustregion_threshold = 1e-8;
  return d;
}

double orc_optimize_sparse(double* p, unsigned Nstate, unsigned Nmeas, unsigned NJnnz,
                           dogleg_callback_t* f, void* cookie,
                           const dogleg_parameters2_t* parameters, orc_result_t* result)
{
  if(NJnnz == 0 || !f) return -1.0;
  dogleg_parameters2_t d = defaults();
  octx c = { .type = ST_SPARSE, .N = (int)Nstate, .M = (int)Nmeas, .nnz = (int)NJnnz,
             .f = f, .cookie = cookie, .prm = parameters ? parameters : &d, .res = result };
  return run(p, &c);
}
double orc_optimize_dense(double* p, unsigned Nstate, unsigned Nmeas,
                          dogleg_callback_dense_t* f, void* cookie,
                          const dogleg_parameters2_t* parameters, orc_result_t* result)
{
  if(!f) return -1.0;
  dogleg_parameters2_t d = defaults();
  octx c = { .type = ST_DENSE, .N = (int)Nstate, .M = (int)Nmeas,
             .fd = f, .cookie = cookie, .prm = parameters ? parameters : &d, .res = result };
  return run(p, &c);
}
double orc_optimize_dense_products(double* p, unsigned Nstate,
                                   dogleg_callback_dense_products_t* f, void* cookie,
                                   const dogleg_parameters2_t* parameters, orc_result_t* result)
{
  if(!f) return -1.0;
  dogleg_parameters2_t d = defaults();
  octx c = { .type = ST_PRODUCTS, .N = (int)Nstate, .M = 0,
             .fp = f, .cookie = cookie, .prm = parameters ? parameters : &d, .res = result };
  return run(p, &c);
}
