/* oracle/dogleg_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of libdogleg's per-iteration hot path (SURVEY.md
 * section 8a rows a2-a19), each function citing the reference lines it follows.
 * Pinned by tests/ against (1) the golden convergence log and trace of the
 * reference's sample problem (tests/golden/) and (2) the unmodified reference
 * compiled into oracle/_ref/ (dense and dense-products paths run on LAPACK; the
 * sparse path runs on oracle/cholmod_shim.c because CHOLMOD is not available:
 * for that part parity with real CHOLMOD is UNPINNED, see
 * sparse_chol_oracle.h).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may link or call this.
 */
#ifndef ORC_DOGLEG_H
#define ORC_DOGLEG_H
#include "dogleg.h"            /* types only: callbacks, dogleg_parameters2_t */
#ifdef __cplusplus
extern "C" {
#endif

/* ---- kernels (reference dogleg.c:186-347) ---- */
double orc_norm2(const double* v, int n);                                     /* :190-196 */
double orc_inner(const double* a, const double* b, int n);                    /* :197-203 */
void   orc_Jt_times_x(double* Jt_x, int Nstate, int Nmeas,                    /* :249-261 */
                      const int* Jp, const int* Ji, const double* Jx, const double* x);
double orc_norm2_J_times_v(int Nmeas, const int* Jp, const int* Ji,           /* :262-281 */
                           const double* Jx, const double* v);
void   orc_dense_Jt_times_x(double* Jt_x, const double* J, const double* x,   /* :284-292 */
                            int Nmeas, int Nstate);
double orc_dense_norm2_J_times_v(const double* J, const double* v,            /* :293-306 */
                                 int Nmeas, int Nstate);
void   orc_dense_JtJ_packed_upper(double* JtJ, const double* J,               /* :214-220,709-723 */
                                  int Nmeas, int Nstate, double lambda);
double orc_xt_Apacked_upper_x(const double* v, const double* A, int N);       /* :309-332 */
double orc_xt_A_x(const double* v, const double* A, int N);                   /* :335-347 */

/* sparse JtJ (dense N x N row-first, both triangles) -- what CHOLMOD forms
 * implicitly inside cholmod_factorize (dogleg.c:656-665) */
void   orc_sparse_JtJ_dense(double* JtJ, int Nstate, int Nmeas,
                            const int* Jp, const int* Ji, const double* Jx, double lambda);

/* packed Cholesky in LAPACK's dpptrf/dpptrs sense for the row-first-upper ==
 * column-major-lower triangle libdogleg uses (dogleg.c:779-784, 872-877).
 * Returns 0, or k+1 if the leading minor of order k+1 is not positive definite */
int    orc_pptrf_lower(double* ap, int n);
void   orc_pptrs_lower(const double* ap, int n, double* b);
/* full-storage variants (dpotrf/dpotrs 'L' on column-major == row-first upper) */
int    orc_potrf_rowfirst(double* a, int n);
void   orc_potrs_rowfirst(const double* a, int n, double* b);

/* ---- step logic ---- */
/* Cauchy step, dogleg.c:529-617. Returns k; writes update and its norm2 */
double orc_cauchy(double* updateCauchy, double* norm2_updateCauchy,
                  const double* Jt_x, double norm2_J_Jt_x, int Nstate);
/* dog-leg interpolation, dogleg.c:927-998. Returns k in [0,1] */
double orc_interpolate(double* step, double* norm2_step,
                       const double* cauchy, double norm2_cauchy,
                       const double* gn, double trustregion, int Nstate);

/* ---- whole solves with a trial-by-trial record ---- */
typedef struct
{
  int    iteration, accepted;
  int    step_type;                 /* 0 cauchy, 1 gaussnewton, 2 interpolated */
  double norm2x_before, norm2x_after;
  double step_len_cauchy, step_len_gn, step_len_interpolated, k_cauchy_to_gn;
  double norm2_step;                /* as stored in norm2_step_to_here */
  double expected_improvement, observed_improvement, rho;
  double trustregion_before, trustregion_after;
} orc_trial_t;

typedef struct
{
  double  norm2_x;                  /* return value of the solve */
  int     accepted_steps;
  int     Ntrials, Ntrials_max;
  orc_trial_t* trials;              /* caller-provided storage, may be NULL */
  double  lambda;
  int     use_ll;                   /* sparse: 1 = LL' rule (pivot<=0 fails), 0 = CHOLMOD LDL' rule */
  const int* perm;                  /* sparse: injected ordering or NULL */
} orc_result_t;

double orc_optimize_sparse(double* p, unsigned Nstate, unsigned Nmeas, unsigned NJnnz,
                           dogleg_callback_t* f, void* cookie,
                           const dogleg_parameters2_t* parameters, orc_result_t* result);
double orc_optimize_dense(double* p, unsigned Nstate, unsigned Nmeas,
                          dogleg_callback_dense_t* f, void* cookie,
                          const dogleg_parameters2_t* parameters, orc_result_t* result);
double orc_optimize_dense_products(double* p, unsigned Nstate,
                                   dogleg_callback_dense_products_t* f, void* cookie,
                                   const dogleg_parameters2_t* parameters, orc_result_t* result);

#ifdef __cplusplus
}
#endif
#endif
