/* tests/support/problems.c -- see problems.h */
#define _GNU_SOURCE
#include "problems.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static dlb_problem* alloc_problem(int N, int M)
{
  dlb_problem* P = calloc(1, sizeof(*P));
  P->N = N; P->M = M;
  P->b = calloc(M ? M : 1, sizeof(double));
  P->p_true = calloc(N, sizeof(double));
  P->p0 = calloc(N, sizeof(double));
  return P;
}

static inline double phi(double p)  { return p + 0.1 * p * p * p; }
static inline double dphi(double p) { return 1.0 + 0.3 * p * p; }

/* fills Ax, p_true, p0, b once the pattern (Ap, Ai) exists */
static void finish_sparse(dlb_problem* P, uint64_t seed)
{
  const int N = P->N, M = P->M;
  P->Ax = malloc(sizeof(double) * (size_t)P->nnz);
  for(int64_t q = 0; q < P->nnz; q++) P->Ax[q] = dlb_uniform(seed, 1, (uint64_t)q);
  for(int k = 0; k < N; k++)
  {
    P->p_true[k] = dlb_uniform(seed, 2, k);
    P->p0[k]     = P->p_true[k] + 0.5 * dlb_uniform(seed, 3, k);
  }
  for(int j = 0; j < M; j++)
  {
    double s = 0.0;
    for(int q = P->Ap[j]; q < P->Ap[j+1]; q++) s += P->Ax[q] * phi(P->p_true[P->Ai[q]]);
    P->b[j] = s + 0.01 * dlb_uniform(seed, 4, j);
  }
}

dlb_problem* dlb_problem_random_sparse(int N, int M, int nnz_per_meas, uint64_t seed)
{
  if(nnz_per_meas > N) nnz_per_meas = N;
  dlb_problem* P = alloc_problem(N, M);
  P->nnz = (int64_t)M * nnz_per_meas;
  P->Ap = malloc(sizeof(int) * (M + 1));
  P->Ai = malloc(sizeof(int) * (size_t)P->nnz);
  char* used = calloc(N, 1);
  for(int j = 0; j < M; j++)
  {
    P->Ap[j] = j * nnz_per_meas;
    int* col = P->Ai + (size_t)j * nnz_per_meas;
    /* a local cluster plus a few far entries, all distinct, then sorted */
    int base = (int)(((uint64_t)j * (uint64_t)N) / (uint64_t)M);
    int cnt = 0; uint64_t t = 0;
    while(cnt < nnz_per_meas)
    {
      uint64_t h = dlb_splitmix64(dlb_splitmix64(seed + 77) ^ ((uint64_t)j << 20) ^ t++);
      int r = (cnt < (2 * nnz_per_meas + 2) / 3) ? (base + (int)(h % (uint64_t)(2 * nnz_per_meas))) % N
                                                 : (int)(h % (uint64_t)N);
      if(used[r]) continue;
      used[r] = 1; col[cnt++] = r;
    }
    for(int a = 1; a < cnt; a++) { int v = col[a], b = a - 1; while(b >= 0 && col[b] > v) { col[b+1] = col[b]; b--; } col[b+1] = v; }
    for(int a = 0; a < cnt; a++) used[col[a]] = 0;
  }
  P->Ap[M] = (int)P->nnz;
  free(used);
  finish_sparse(P, seed);
  return P;
}

/* Ragged columns: lengths 0..kmax. The first N measurements touch one state each (they anchor
 * every state: full rank), then a run of 200 empty measurements, then 200 measurements that
 * alternate between one fixed row list and an empty one (a periodic run with an empty class), then
 * random row lists of random length with every 5th measurement empty. An empty measurement still
 * has a residual (x_j = -b_j): it counts in |x|^2 and nowhere else. */
dlb_problem* dlb_problem_ragged(int N, int M, int kmax, uint64_t seed)
{
  if(kmax > N) kmax = N;
  if(M < N + 400) M = N + 400;
  dlb_problem* P = alloc_problem(N, M);
  P->Ap = malloc(sizeof(int) * (M + 1));
  P->Ai = malloc(sizeof(int) * (size_t)M * (kmax > 0 ? kmax : 1));
  char* used = calloc(N, 1);
  int nnz = 0;
  for(int j = 0; j < M; j++)
  {
    P->Ap[j] = nnz;
    int* col = P->Ai + nnz;
    int len;
    if(j < N) { col[0] = j; nnz++; continue; }
    if(j < N + 200) continue;
    if(j < N + 400) { if((j - N) & 1) continue; len = kmax < 3 ? kmax : 3; for(int a = 0; a < len; a++) col[a] = (a * (N - 1)) / (len > 1 ? len - 1 : 1); nnz += len; continue; }
    const uint64_t h0 = dlb_splitmix64(dlb_splitmix64(seed + 91) ^ (uint64_t)j);
    len = (j % 5 == 0) ? 0 : (int)(h0 % (uint64_t)(kmax + 1));
    int cnt = 0; uint64_t t = 0;
    while(cnt < len)
    {
      const uint64_t h = dlb_splitmix64(h0 ^ (0x51ull << 40) ^ t++);
      const int r = (int)(h % (uint64_t)N);
      if(used[r]) continue;
      used[r] = 1; col[cnt++] = r;
    }
    for(int a = 1; a < cnt; a++) { int v = col[a], b = a - 1; while(b >= 0 && col[b] > v) { col[b+1] = col[b]; b--; } col[b+1] = v; }
    for(int a = 0; a < cnt; a++) used[col[a]] = 0;
    nnz += cnt;
  }
  P->Ap[M] = nnz;
  P->nnz = nnz;
  free(used);
  finish_sparse(P, seed);
  return P;
}

/* measurement (f,c,k,xy) touches intrinsics
 * {12c+xy, 12c+2+xy, 12c+4..12c+11}, extrinsics 12ncam+6(c-1)..+5 for c>0, frame
 * pose 12ncam+6(ncam-1)+6f..+5, and the last two (global) states. f-major order. */
dlb_problem* dlb_problem_mrcal(int ncam, int nframes, int npts, uint64_t seed)
{
  const int N = 12 * ncam + 6 * (ncam - 1) + 6 * nframes + 2;
  const int64_t M64 = (int64_t)ncam * nframes * npts * 2;
  const int M = (int)M64;
  dlb_problem* P = alloc_problem(N, M);
  P->nnz = (int64_t)nframes * npts * 2 * (18 + 24 * (int64_t)(ncam - 1));
  P->Ap = malloc(sizeof(int) * ((size_t)M + 1));
  P->Ai = malloc(sizeof(int) * (size_t)P->nnz);
  const int ext0 = 12 * ncam, frm0 = ext0 + 6 * (ncam - 1), glob0 = frm0 + 6 * nframes;
  int64_t q = 0; int j = 0;
  for(int f = 0; f < nframes; f++)
    for(int c = 0; c < ncam; c++)
      for(int k = 0; k < npts; k++)
        for(int xy = 0; xy < 2; xy++, j++)
        {
          P->Ap[j] = (int)q;
          P->Ai[q++] = 12 * c + xy;
          P->Ai[q++] = 12 * c + 2 + xy;
          for(int t = 4; t < 12; t++) P->Ai[q++] = 12 * c + t;
          if(c > 0) for(int t = 0; t < 6; t++) P->Ai[q++] = ext0 + 6 * (c - 1) + t;
          for(int t = 0; t < 6; t++) P->Ai[q++] = frm0 + 6 * f + t;
          P->Ai[q++] = glob0; P->Ai[q++] = glob0 + 1;
        }
  P->Ap[M] = (int)q;
  finish_sparse(P, seed);
  return P;
}

/* cameras (9 states each) first, then points (3 each).
 * Point m is seen by obs_per_point distinct cameras inside a window of
 * 'window' consecutive cameras (on a ring) centred at floor(m ncams/npoints);
 * longrange_permille/1000 of the observations go to a uniformly random camera
 * instead. Each observation gives two measurements (9 camera + 3 point entries). */
dlb_problem* dlb_problem_ba(int ncams, int npoints, int obs_per_point, int window,
                            int longrange_permille, uint64_t seed)
{
  if(window > ncams) window = ncams;
  if(obs_per_point > window) obs_per_point = window;
  const int N = 9 * ncams + 3 * npoints;
  const int M = 2 * obs_per_point * npoints;
  dlb_problem* P = alloc_problem(N, M);
  P->nnz = (int64_t)M * 12;
  P->Ap = malloc(sizeof(int) * ((size_t)M + 1));
  P->Ai = malloc(sizeof(int) * (size_t)P->nnz);
  int* cams = malloc(sizeof(int) * obs_per_point);
  int64_t q = 0; int j = 0;
  for(int m = 0; m < npoints; m++)
  {
    const int centre = (int)(((int64_t)m * ncams) / npoints);
    int cnt = 0; uint64_t t = 0;
    while(cnt < obs_per_point)
    {
      uint64_t h = dlb_splitmix64(dlb_splitmix64(seed + 99) ^ ((uint64_t)m << 16) ^ t++);
      int c;
      if((int)((h >> 40) % 1000) < longrange_permille) c = (int)((h >> 8) % (uint64_t)ncams);
      else c = ((centre - window / 2 + (int)(h % (uint64_t)window)) % ncams + ncams) % ncams;
      int dup = 0;
      for(int a = 0; a < cnt; a++) if(cams[a] == c) dup = 1;
      if(!dup) cams[cnt++] = c;
    }
    for(int a = 1; a < cnt; a++) { int v = cams[a], b = a - 1; while(b >= 0 && cams[b] > v) { cams[b+1] = cams[b]; b--; } cams[b+1] = v; }
    for(int a = 0; a < cnt; a++)
      for(int xy = 0; xy < 2; xy++, j++)
      {
        P->Ap[j] = (int)q;
        for(int t9 = 0; t9 < 9; t9++) P->Ai[q++] = 9 * cams[a] + t9;
        for(int t3 = 0; t3 < 3; t3++) P->Ai[q++] = 9 * ncams + 3 * m + t3;
      }
  }
  P->Ap[M] = (int)q;
  free(cams);
  finish_sparse(P, seed);
  return P;
}

dlb_problem* dlb_problem_dense(int N, int M, uint64_t seed)
{
  dlb_problem* P = alloc_problem(N, M);
  P->nnz = 0;
  P->Adense = malloc(sizeof(double) * (size_t)M * N);
  for(size_t q = 0; q < (size_t)M * N; q++) P->Adense[q] = dlb_uniform(seed, 1, q);
  for(int k = 0; k < N; k++)
  {
    P->p_true[k] = dlb_uniform(seed, 2, k);
    P->p0[k]     = P->p_true[k] + 0.5 * dlb_uniform(seed, 3, k);
  }
  for(int i = 0; i < M; i++)
  {
    double s = 0.0;
    for(int k = 0; k < N; k++) s += P->Adense[(size_t)i * N + k] * phi(P->p_true[k]);
    P->b[i] = s + 0.01 * dlb_uniform(seed, 4, i);
  }
  return P;
}

/* ---- the reference's sample problem (sample.c:25-80, 351-371): a 6-parameter
 * surface a b x^2 + b c y^2 + c x y + d x + e y + f on a 10x10 grid with
 * +-0.5 uniform noise from glibc random() seeded with 0, true values 1..6,
 * start point random()/RAND_MAX - 0.1. Stored as: Ax[2i]=x_i, Ax[2i+1]=y_i. */
dlb_problem* dlb_problem_sample(void)
{
  const int N = 6, W = 10, M = W * W;
  dlb_problem* P = alloc_problem(N, M);
  P->kind = 1;
  P->nnz = (int64_t)M * N;
  P->Ax = malloc(sizeof(double) * 2 * M);
  srandom(0);
  int i = 0;
  for(int ix = 0; ix < W; ix++)
    for(int iy = 0; iy < W; iy++, i++)
    {
      P->Ax[2*i]   = -10 + ix * 2.0;
      P->Ax[2*i+1] = -10 + iy * 2.0;
    }
  for(i = 0; i < N; i++) P->p_true[i] = 1.0 + i;
  for(i = 0; i < M; i++)
  {
    const double x = P->Ax[2*i], y = P->Ax[2*i+1];
    P->b[i] = 1.0*2.0 * x*x + 2.0*3.0 * y*y + 3.0 * x*y + 4.0 * x + 5.0 * y + 6.0 +
              ((double)random() / (double)RAND_MAX - 0.5) * 1.0;
  }
  for(i = 0; i < N; i++) P->p0[i] = ((double)random() / (double)RAND_MAX - 0.1) * 1.0;
  return P;
}

/* the measurements [col_begin, col_begin+ncols) of a sparse problem as a problem of its own
 * (same states, same p0/p_true): what one rank of a row-sharded solve evaluates */
dlb_problem* dlb_problem_slice(const dlb_problem* P, int col_begin, int ncols)
{
  if(P->kind != 0 || col_begin < 0 || col_begin + ncols > P->M) return NULL;
  if(P->Adense)
  { /* dense: rows [col_begin, col_begin + ncols) of A and b */
    dlb_problem* Q = alloc_problem(P->N, ncols);
    Q->nnz = 0;
    Q->Adense = malloc(sizeof(double) * (size_t)(ncols ? ncols : 1) * P->N);
    memcpy(Q->Adense, P->Adense + (size_t)col_begin * P->N, sizeof(double) * (size_t)ncols * P->N);
    memcpy(Q->b, P->b + col_begin, sizeof(double) * (size_t)ncols);
    memcpy(Q->p_true, P->p_true, sizeof(double) * P->N);
    memcpy(Q->p0, P->p0, sizeof(double) * P->N);
    return Q;
  }
  if(!P->Ap) return NULL;
  dlb_problem* Q = alloc_problem(P->N, ncols);
  const int q0 = P->Ap[col_begin];
  Q->nnz = P->Ap[col_begin + ncols] - q0;
  Q->Ap = malloc(sizeof(int) * ((size_t)ncols + 1));
  Q->Ai = malloc(sizeof(int) * (size_t)(Q->nnz ? Q->nnz : 1));
  Q->Ax = malloc(sizeof(double) * (size_t)(Q->nnz ? Q->nnz : 1));
  for(int j = 0; j <= ncols; j++) Q->Ap[j] = P->Ap[col_begin + j] - q0;
  memcpy(Q->Ai, P->Ai + q0, sizeof(int) * (size_t)Q->nnz);
  memcpy(Q->Ax, P->Ax + q0, sizeof(double) * (size_t)Q->nnz);
  memcpy(Q->b, P->b + col_begin, sizeof(double) * (size_t)ncols);
  memcpy(Q->p_true, P->p_true, sizeof(double) * P->N);
  memcpy(Q->p0, P->p0, sizeof(double) * P->N);
  return Q;
}

void dlb_problem_free(dlb_problem* P)
{
  if(!P) return;
  free(P->Ap); free(P->Ai); free(P->Ax); free(P->Adense); free(P->b); free(P->p_true); free(P->p0);
  free(P->trace_p); free(P->trace_norm2x); free(P);
}
void dlb_problem_trace(dlb_problem* P, int on, int cap)
{
  P->trace_on = on;
  if(on && cap > P->trace_cap)
  {
    P->trace_p = realloc(P->trace_p, sizeof(double) * (size_t)cap * P->N);
    P->trace_norm2x = realloc(P->trace_norm2x, sizeof(double) * cap);
    P->trace_cap = cap;
  }
  P->ncalls = 0;
}
void dlb_problem_reset(dlb_problem* P) { P->ncalls = 0; P->cb_seconds = 0.0; }

static void record(dlb_problem* P, const double* p, double norm2x)
{
  if(P->trace_on && P->ncalls < P->trace_cap)
  {
    memcpy(P->trace_p + (size_t)P->ncalls * P->N, p, sizeof(double) * P->N);
    P->trace_norm2x[P->ncalls] = norm2x;
  }
  P->ncalls++;
}

/* sample.c model: residual and gradient of measurement i */
static inline double sample_eval(const dlb_problem* P, const double* p, int i, double* g)
{
  const double x = P->Ax[2*i], y = P->Ax[2*i+1];
  g[0] = p[1]*x*x;
  g[1] = p[0]*x*x + p[2] * y*y;
  g[2] = p[1] * y*y + x*y;
  g[3] = x; g[4] = y; g[5] = 1.0;
  return p[0] * p[1] * x*x + p[1] * p[2] * y*y + p[2] * x*y + p[3] * x + p[4] * y + p[5] - P->b[i];
}

void dlb_cb_sparse(const double* p, double* x, cholmod_sparse* Jt, void* cookie)
{
  dlb_problem* P = cookie;
  const double t0 = now_s();
  int* Jp = Jt->p; int* Ji = Jt->i; double* Jx = Jt->x;
  double n2 = 0.0;
  if(P->kind == 1)
  {
    int q = 0;
    for(int i = 0; i < P->M; i++)
    {
      double g[6];
      x[i] = sample_eval(P, p, i, g);
      Jp[i] = q;
      for(int k = 0; k < 6; k++, q++) { Ji[q] = k; Jx[q] = g[k]; }
    }
    Jp[P->M] = q;
  }
  else
  {
    const int M = P->M;
    const int waves = (P->progress && M >= 4096) ? 16 : 1;
    for(int wv = 0; wv < waves; wv++)
    {
      const int ja = (int)((long long)M * wv / waves), jb = (int)((long long)M * (wv + 1) / waves);
#ifdef _OPENMP
      #pragma omp parallel for schedule(static) num_threads(P->nthreads > 0 ? P->nthreads : omp_get_max_threads())
#endif
      for(int j = ja; j < jb; j++)
      {
        double s = 0.0;
        Jp[j] = P->Ap[j];
        for(int q = P->Ap[j]; q < P->Ap[j+1]; q++)
        {
          const double pk = p[P->Ai[q]];
          Ji[q] = P->Ai[q];
          Jx[q] = P->Ax[q] * dphi(pk);
          s += P->Ax[q] * phi(pk);
        }
        x[j] = s - P->b[j];
      }
      if(P->progress) P->progress((size_t)P->Ap[jb], (size_t)jb);       /* columns [0, jb) are final */
    }
    Jp[M] = P->Ap[M];
  }
  if(P->trace_on) { for(int i = 0; i < P->M; i++) n2 += x[i] * x[i]; }
  record(P, p, n2);
  P->cb_seconds += now_s() - t0;
}

void dlb_cb_dense(const double* p, double* x, double* J, void* cookie)
{
  dlb_problem* P = cookie;
  const double t0 = now_s();
  const int N = P->N, M = P->M;
  if(P->kind == 1)
    for(int i = 0; i < M; i++) x[i] = sample_eval(P, p, i, J + (size_t)i * N);
  else if(P->Adense)
  {
#ifdef _OPENMP
    #pragma omp parallel for schedule(static) num_threads(P->nthreads > 0 ? P->nthreads : omp_get_max_threads())
#endif
    for(int i = 0; i < M; i++)
    {
      double s = 0.0;
      for(int k = 0; k < N; k++)
      {
        const double a = P->Adense[(size_t)i * N + k];
        J[(size_t)i * N + k] = a * dphi(p[k]);
        s += a * phi(p[k]);
      }
      x[i] = s - P->b[i];
    }
  }
  else
  { /* densified sparse problem: the cross-check route of SURVEY.md 8c */
    memset(J, 0, sizeof(double) * (size_t)M * N);
    for(int j = 0; j < M; j++)
    {
      double s = 0.0;
      for(int q = P->Ap[j]; q < P->Ap[j+1]; q++)
      {
        const double pk = p[P->Ai[q]];
        J[(size_t)j * N + P->Ai[q]] = P->Ax[q] * dphi(pk);
        s += P->Ax[q] * phi(pk);
      }
      x[j] = s - P->b[j];
    }
  }
  double n2 = 0.0;
  if(P->trace_on) for(int i = 0; i < M; i++) n2 += x[i] * x[i];
  record(P, p, n2);
  P->cb_seconds += now_s() - t0;
}

void dlb_cb_products(const double* p, double* norm2x, double* xtJ, double* JtJ, void* cookie)
{
  dlb_problem* P = cookie;
  const double t0 = now_s();
  const int N = P->N, M = P->M;
  const size_t sz = P->packed ? (size_t)N * (N + 1) / 2 : (size_t)N * N;
  double* g = malloc(sizeof(double) * N);
  memset(xtJ, 0, sizeof(double) * N);
  memset(JtJ, 0, sizeof(double) * sz);
  double n2 = 0.0;
  for(int i = 0; i < M; i++)
  {
    double xi;
    if(P->kind == 1) xi = sample_eval(P, p, i, g);
    else
    {
      double s = 0.0;
      memset(g, 0, sizeof(double) * N);
      if(P->Adense)
        for(int k = 0; k < N; k++) { const double a = P->Adense[(size_t)i * N + k]; g[k] = a * dphi(p[k]); s += a * phi(p[k]); }
      else
        for(int q = P->Ap[i]; q < P->Ap[i+1]; q++) { const double pk = p[P->Ai[q]]; g[P->Ai[q]] = P->Ax[q] * dphi(pk); s += P->Ax[q] * phi(pk); }
      xi = s - P->b[i];
    }
    n2 += xi * xi;
    for(int k = 0; k < N; k++) xtJ[k] += xi * g[k];
    if(P->packed && P->upper)
    {
      size_t at = 0;
      for(int k = 0; k < N; k++) for(int l = k; l < N; l++, at++) JtJ[at] += g[k] * g[l];
    }
    else if(P->packed)
    {
      size_t at = 0;
      for(int k = 0; k < N; k++) for(int l = 0; l <= k; l++, at++) JtJ[at] += g[k] * g[l];
    }
    else
      for(int k = 0; k < N; k++) for(int l = 0; l < N; l++) JtJ[(size_t)k * N + l] += g[k] * g[l];
  }
  *norm2x = n2;
  free(g);
  record(P, p, n2);
  P->cb_seconds += now_s() - t0;
}

void* dlb_cb_sparse_ptr(void)   { return (void*)&dlb_cb_sparse; }
void* dlb_cb_dense_ptr(void)    { return (void*)&dlb_cb_dense; }
void* dlb_cb_products_ptr(void) { return (void*)&dlb_cb_products; }
