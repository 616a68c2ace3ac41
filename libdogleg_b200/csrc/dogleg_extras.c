/* dogleg_extras.c -- API surface around the hot path: the gradient tester
 * (reference dogleg.c:373-522). A developer aid that evaluates the user's
 * callback twice and prints a vnlog table; it does no linear algebra, so it is
 * plain host C here as well (out of the per-iteration path).
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "dogleg.h"

#define SAY(fmt, ...) fprintf(stderr, "libdogleg at %s:%d: " fmt "\n", __FILE__, __LINE__, ## __VA_ARGS__)
#define GRADTEST_DELTA 1e-6

static cholmod_sparse* scratch_Jt(unsigned int Nstate, unsigned int Nmeas, unsigned int NJnnz)
{
  cholmod_sparse* Jt = calloc(1, sizeof(*Jt));
  if(!Jt) return NULL;
  Jt->nrow = Nstate; Jt->ncol = Nmeas; Jt->nzmax = NJnnz;
  Jt->p = calloc((size_t)Nmeas + 1, sizeof(int));
  Jt->i = calloc(NJnnz, sizeof(int));
  Jt->x = calloc(NJnnz, sizeof(double));
  Jt->stype = 0; Jt->itype = CHOLMOD_INT; Jt->xtype = CHOLMOD_REAL; Jt->dtype = CHOLMOD_DOUBLE;
  Jt->sorted = 1; Jt->packed = 1;
  return Jt;
}
static void scratch_free(cholmod_sparse* Jt) { if(Jt) { free(Jt->p); free(Jt->i); free(Jt->x); free(Jt); } }

static double sparse_entry(const cholmod_sparse* Jt, unsigned int var, unsigned int meas)
{
  const int* p = Jt->p; const int* i = Jt->i; const double* x = Jt->x;
  for(int q = p[meas]; q < p[meas + 1]; q++) if((unsigned int)i[q] == var) return x[q];
  return 0.0;
}

static void gradient_table(unsigned int var, const double* p0,
                           unsigned int Nstate, unsigned int Nmeas, unsigned int NJnnz,
                           dogleg_callback_t* f, dogleg_callback_dense_t* f_dense, void* cookie)
{
  double* xm = malloc(sizeof(double) * (Nmeas ? Nmeas : 1));
  double* xp = malloc(sizeof(double) * (Nmeas ? Nmeas : 1));
  double* p  = malloc(sizeof(double) * Nstate);
  cholmod_sparse *Jm = NULL, *Jp = NULL;
  double *Dm = NULL, *Dp = NULL;
  if(!xm || !xp || !p) { SAY("out of memory"); goto done; }
  memcpy(p, p0, sizeof(double) * Nstate);

  printf("# ivar imeasurement gradient_reported gradient_observed error error_relative\n");
  /* central difference: evaluate at p - delta/2 and p + delta/2, stepping the variable exactly as the
   * reference does (-= delta/2, then += delta: dogleg.c:424-428), so that the printed error columns --
   * pure cancellation noise around 1e-8 -- come out the same */
  if(f)
  {
    Jm = scratch_Jt(Nstate, Nmeas, NJnnz); Jp = scratch_Jt(Nstate, Nmeas, NJnnz);
    if(!Jm || !Jp) { SAY("out of memory"); goto done; }
    p[var] -= GRADTEST_DELTA / 2.0; f(p, xm, Jm, cookie);
    p[var] += GRADTEST_DELTA;       f(p, xp, Jp, cookie);
  }
  else
  {
    Dm = malloc(sizeof(double) * (size_t)Nmeas * Nstate); Dp = malloc(sizeof(double) * (size_t)Nmeas * Nstate);
    if(!Dm || !Dp) { SAY("out of memory"); goto done; }
    p[var] -= GRADTEST_DELTA / 2.0; f_dense(p, xm, Dm, cookie);
    p[var] += GRADTEST_DELTA;       f_dense(p, xp, Dp, cookie);
  }
  for(unsigned int i = 0; i < Nmeas; i++)
  {
    const double observed = (xp[i] - xm[i]) / GRADTEST_DELTA;
    const double reported = f ? (sparse_entry(Jm, var, i) + sparse_entry(Jp, var, i)) / 2.0
                              : (Dm[var + (size_t)i * Nstate] + Dp[var + (size_t)i * Nstate]) / 2.0;
    const double mag = fabs(reported) + fabs(observed);
    const double err = fabs(reported - observed);
    printf("%d %d %.6g %.6g %.6g %.6g\n", var, i, reported, observed, err, mag == 0.0 ? 0.0 : err / (mag / 2.0));
  }
done:
  scratch_free(Jm); scratch_free(Jp); free(Dm); free(Dp); free(xm); free(xp); free(p);
}

void dogleg_testGradient(unsigned int var, const double* p0,
                         unsigned int Nstate, unsigned int Nmeas, unsigned int NJnnz,
                         dogleg_callback_t* f, void* cookie)
{
  if(NJnnz == 0) { SAY("I must have NJnnz > 0, instead I have %d", NJnnz); return; }
  gradient_table(var, p0, Nstate, Nmeas, NJnnz, f, NULL, cookie);
}
void dogleg_testGradient_dense(unsigned int var, const double* p0,
                               unsigned int Nstate, unsigned int Nmeas,
                               dogleg_callback_dense_t* f, void* cookie)
{
  gradient_table(var, p0, Nstate, Nmeas, 0, NULL, f, cookie);
}
void dogleg_testGradient_dense_products(unsigned int var, const double* p0,
                                        unsigned int Nstate, unsigned int Nmeas,
                                        dogleg_callback_dense_products_t* f, void* cookie)
{
  /* unimplemented in the reference too (it exits, dogleg.c:440-446); we report and return */
  (void)var; (void)p0; (void)Nstate; (void)Nmeas; (void)f; (void)cookie;
  SAY("dogleg_testGradient_dense_products() is not implemented (nor is it in the reference)");
}
