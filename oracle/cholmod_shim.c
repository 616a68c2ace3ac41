/* oracle/cholmod_shim.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Implements the 13 cholmod_* entry points libdogleg calls on
 * top of the oracle's CPU restatement (sparse_chol_oracle.c), so that the
 * UNMODIFIED reference /root/reference/dogleg.c can be compiled and linked into
 * oracle/_ref/libdogleg_ref.so and run its sparse path in this image, which has
 * no SuiteSparse. The reference's own control flow, SpMV kernels, trust-region
 * logic and lambda ladder are therefore the real thing; only the arithmetic the
 * reference delegates to CHOLMOD is the restatement.
 *
 * Behaviour follows CHOLMOD with the settings the reference uses
 * (dogleg.c:1595-1611): simplicial LDL', int32 indices, factorization of
 * beta*I + A*A' for an unsymmetric A (stype 0).
 *
 * Test hooks (not CHOLMOD API): orc_shim_set_permutation(), orc_shim_set_ll().
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define _POSIX_C_SOURCE 200809L
#include <stdarg.h>
#include <time.h>
#include "cholmod.h"
#include "sparse_chol_oracle.h"

static int (*shim_printf)(const char*, ...) = NULL;
static const int* hook_perm   = NULL;
static int        hook_perm_n = 0;
static int        hook_is_ll  = 0;

/* seconds spent inside cholmod_analyze since the last reset (bench.py reports the reference both with
 * and without its per-solve symbolic analysis) */
static double analyze_seconds = 0.0;
double orc_shim_analyze_seconds(int reset) { const double s = analyze_seconds; if(reset) analyze_seconds = 0.0; return s; }
void orc_shim_set_permutation(const int* perm, int n) { hook_perm = perm; hook_perm_n = n; }
void orc_shim_set_ll(int is_ll)                       { hook_is_ll = is_ll; }
void SuiteSparse_config_printf_func_set(int (*f)(const char*, ...)) { shim_printf = f; }

int cholmod_start(cholmod_common* c)
{
  memset(c, 0, sizeof(*c));
  c->supernodal = CHOLMOD_AUTO;
  c->itype = CHOLMOD_INT;
  c->dtype = CHOLMOD_DOUBLE;
  c->postorder = 1;
  c->final_asis = 1; c->final_pack = 1; c->final_monotonic = 1;
  return 1;
}
int cholmod_finish(cholmod_common* c) { (void)c; return 1; }

cholmod_sparse* cholmod_allocate_sparse(size_t nrow, size_t ncol, size_t nzmax,
                                        int sorted, int packed, int stype, int xtype,
                                        cholmod_common* c)
{
  (void)c;
  cholmod_sparse* A = calloc(1, sizeof(*A));
  if(!A) return NULL;
  if(nzmax == 0) nzmax = 1;
  A->nrow = nrow; A->ncol = ncol; A->nzmax = nzmax;
  A->p = calloc(ncol + 1, sizeof(int));
  A->i = calloc(nzmax, sizeof(int));
  A->x = xtype == CHOLMOD_PATTERN ? NULL : calloc(nzmax, sizeof(double));
  A->nz = NULL; A->z = NULL;
  A->stype = stype; A->itype = CHOLMOD_INT; A->xtype = xtype; A->dtype = CHOLMOD_DOUBLE;
  A->sorted = sorted; A->packed = packed;
  return A;
}
int cholmod_free_sparse(cholmod_sparse** A, cholmod_common* c)
{
  (void)c;
  if(A && *A) { free((*A)->p); free((*A)->i); free((*A)->x); free((*A)->nz); free(*A); *A = NULL; }
  return 1;
}
cholmod_dense* cholmod_allocate_dense(size_t nrow, size_t ncol, size_t d, int xtype,
                                      cholmod_common* c)
{
  (void)c;
  cholmod_dense* X = calloc(1, sizeof(*X));
  if(!X) return NULL;
  X->nrow = nrow; X->ncol = ncol; X->d = d; X->nzmax = d * ncol;
  X->x = calloc(X->nzmax ? X->nzmax : 1, sizeof(double));
  X->xtype = xtype; X->dtype = CHOLMOD_DOUBLE;
  return X;
}
int cholmod_free_dense(cholmod_dense** X, cholmod_common* c)
{
  (void)c;
  if(X && *X) { free((*X)->x); free(*X); *X = NULL; }
  return 1;
}

/* the cholmod_factor we hand out carries the oracle factor in ->z (unused for
 * real matrices) and mirrors n, minor, Perm, p, i, x for callers that look */
static orc_factor* OF(cholmod_factor* L) { return (orc_factor*)L->z; }

cholmod_factor* cholmod_analyze(cholmod_sparse* A, cholmod_common* c)
{
  (void)c;
  if(A->stype != 0 || A->itype != CHOLMOD_INT) return NULL;
  const int n = (int)A->nrow, m = (int)A->ncol;
  const int* perm = (hook_perm && hook_perm_n == n) ? hook_perm : NULL;
  struct timespec ts0, ts1;
  clock_gettime(CLOCK_MONOTONIC, &ts0);
  orc_factor* F = orc_analyze(n, m, (const int*)A->p, (const int*)A->i, perm);
  clock_gettime(CLOCK_MONOTONIC, &ts1);
  analyze_seconds += (double)(ts1.tv_sec - ts0.tv_sec) + 1e-9 * (double)(ts1.tv_nsec - ts0.tv_nsec);
  cholmod_factor* L = calloc(1, sizeof(*L));
  L->n = n; L->minor = n;
  L->z = F;
  L->Perm = F->perm; L->IPerm = F->iperm; L->ColCount = F->colcount;
  L->p = F->Lp; L->i = F->Li; L->x = F->Lx; L->nz = F->Lnz; L->nzmax = F->Lp[n];
  L->ordering = perm ? CHOLMOD_GIVEN : CHOLMOD_AMD;
  L->is_ll = 0; L->is_super = 0; L->is_monotonic = 1;
  L->itype = CHOLMOD_INT; L->xtype = CHOLMOD_REAL; L->dtype = CHOLMOD_DOUBLE;
  return L;
}

int cholmod_factorize_p(cholmod_sparse* A, double beta[2], int* fset, size_t fsize,
                        cholmod_factor* L, cholmod_common* c)
{
  (void)fset; (void)fsize; (void)c;
  orc_factor* F = OF(L);
  orc_factorize(F, (const int*)A->p, (const int*)A->i, (const double*)A->x,
                beta ? beta[0] : 0.0, hook_is_ll);
  L->minor = F->minor;
  L->is_ll = F->is_ll;
  if(F->minor < F->n && shim_printf)
    shim_printf("CHOLMOD warning: not positive definite (oracle shim), minor %d", F->minor);
  return 1;
}
int cholmod_factorize(cholmod_sparse* A, cholmod_factor* L, cholmod_common* c)
{
  double beta[2] = {0.0, 0.0};
  return cholmod_factorize_p(A, beta, NULL, 0, L, c);
}
int cholmod_free_factor(cholmod_factor** L, cholmod_common* c)
{
  (void)c;
  if(L && *L) { orc_free(OF(*L)); free(*L); *L = NULL; }
  return 1;
}

cholmod_dense* cholmod_solve(int sys, cholmod_factor* L, cholmod_dense* B, cholmod_common* c)
{
  if(sys != CHOLMOD_A) return NULL;
  orc_factor* F = OF(L);
  const int n = F->n;
  if((int)B->nrow != n) return NULL;
  cholmod_dense* X = cholmod_allocate_dense(n, B->ncol, n, CHOLMOD_REAL, c);
  if((int)B->d == n)
    orc_solve(F, (const double*)B->x, (double*)X->x, (int)B->ncol);
  else
    for(size_t k = 0; k < B->ncol; k++)
      orc_solve(F, (const double*)B->x + k * B->d, (double*)X->x + k * n, 1);
  return X;
}

cholmod_sparse* cholmod_spsolve(int sys, cholmod_factor* L, cholmod_sparse* B, cholmod_common* c)
{
  if(sys != CHOLMOD_A) return NULL;
  orc_factor* F = OF(L);
  const int n = F->n, ncol = (int)B->ncol;
  const int* Bp = B->p; const int* Bi = B->i; const double* Bx = B->x;
  double* b = calloc((size_t)n * ncol, sizeof(double));
  double* x = calloc((size_t)n * ncol, sizeof(double));
  for(int k = 0; k < ncol; k++)
    for(int q = Bp[k]; q < Bp[k+1]; q++) b[(size_t)k * n + Bi[q]] = Bx[q];
  orc_solve(F, b, x, ncol);
  size_t nz = 0;
  for(size_t q = 0; q < (size_t)n * ncol; q++) if(x[q] != 0.0) nz++;
  cholmod_sparse* X = cholmod_allocate_sparse(n, ncol, nz, 1, 1, 0, CHOLMOD_REAL, c);
  int* Xp = X->p; int* Xi = X->i; double* Xx = X->x;
  nz = 0;
  for(int k = 0; k < ncol; k++)
  {
    Xp[k] = (int)nz;
    for(int i = 0; i < n; i++)
      if(x[(size_t)k * n + i] != 0.0) { Xi[nz] = i; Xx[nz] = x[(size_t)k * n + i]; nz++; }
  }
  Xp[ncol] = (int)nz;
  free(b); free(x);
  return X;
}
