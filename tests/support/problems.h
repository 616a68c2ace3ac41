/* tests/support/problems.h -- synthetic least-squares problems shared by the
 * parity tests and bench.py. One set of C callbacks with the libdogleg
 * signatures, so the very same function pointers can be handed to the product
 * (libdogleg.so), to the unmodified reference (oracle/_ref/libdogleg_ref.so)
 * and to the oracle restatement, and every evaluation is recorded in a trace.
 *
 * Model:  x(p) = A phi(p) - b,  phi_k(p) = p_k + 0.1 p_k^3,
 *                        J_ik = A_ik (1 + 0.3 p_k^2),
 *                        b = A phi(p_true) + 0.01 u,  p0 = p_true + 0.5 u
 * with A of a fixed sparsity pattern; all random numbers are a splitmix64
 * counter hash of (seed, index) so C, CUDA and numpy agree bit for bit.
 */
#ifndef DLB_PROBLEMS_H
#define DLB_PROBLEMS_H
#include <stdint.h>
#include <stddef.h>
#include "dogleg.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct dlb_problem
{
  int      kind;                /* 0 synthetic sparse/dense model, 1 reference sample.c problem */
  int      N, M;
  int64_t  nnz;
  int*     Ap;  int* Ai; double* Ax;   /* A' in CCS (N rows, M columns), sparse kinds */
  double*  Adense;                     /* M x N row-first, dense kinds (may be NULL)   */
  double*  b;  double* p_true; double* p0;
  /* trace of every callback evaluation */
  int      trace_on, ncalls, trace_cap;
  double*  trace_p;             /* trace_cap x N */
  double*  trace_norm2x;        /* trace_cap */
  double   cb_seconds;          /* wall time spent inside callbacks */
  int      packed, upper;       /* layout for the dense-products callback */
  int      nthreads;            /* OpenMP threads inside the callbacks (0 = default) */
  /* dogleg_gpu_host_progress of the library under test, or NULL: the sparse callback then fills its outputs in
   * 16 waves, front to back, and announces every finished wave (include/dogleg_gpu.h) */
  void   (*progress)(size_t n_values_final, size_t n_x_final);
} dlb_problem;

static inline uint64_t dlb_splitmix64(uint64_t z)
{
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
/* uniform in (-1,1) from (seed, stream, index) */
static inline double dlb_uniform(uint64_t seed, uint64_t stream, uint64_t idx)
{
  uint64_t h = dlb_splitmix64(dlb_splitmix64(seed * 0x100000001B3ull + stream) ^ idx);
  return ((double)(h >> 11) + 0.5) * (2.0 / 9007199254740992.0) - 1.0;
}

dlb_problem* dlb_problem_sample(void);   /* the reference's sample.c fit (srandom(0)) */
dlb_problem* dlb_problem_random_sparse(int N, int M, int nnz_per_meas, uint64_t seed);
dlb_problem* dlb_problem_ragged(int N, int M, int kmax, uint64_t seed);   /* column lengths 0..kmax */
dlb_problem* dlb_problem_mrcal(int ncam, int nframes, int npts, uint64_t seed);
dlb_problem* dlb_problem_ba(int ncams, int npoints, int obs_per_point, int window,
                            int longrange_permille, uint64_t seed);
dlb_problem* dlb_problem_dense(int N, int M, uint64_t seed);
dlb_problem* dlb_problem_slice(const dlb_problem* P, int col_begin, int ncols);
void         dlb_problem_free(dlb_problem* P);
void         dlb_problem_trace(dlb_problem* P, int on, int cap);
void         dlb_problem_reset(dlb_problem* P);   /* clears trace and timer */

/* cookie = dlb_problem* for all three */
void dlb_cb_sparse  (const double* p, double* x, cholmod_sparse* Jt, void* cookie);
void dlb_cb_dense   (const double* p, double* x, double* J, void* cookie);
void dlb_cb_products(const double* p, double* norm2x, double* xtJ, double* JtJ, void* cookie);

/* addresses of the callbacks, for ctypes */
void* dlb_cb_sparse_ptr(void);
void* dlb_cb_dense_ptr(void);
void* dlb_cb_products_ptr(void);

#ifdef __cplusplus
}
#endif
#endif
