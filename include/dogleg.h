/* dogleg.h -- public C API of the B200-native dog-leg solver.
 *
 * This header is the drop-in boundary: every type below has the size, member
 * order and bit positions of libdogleg ABI 2 (reference dogleg.h:11-210), and
 * every prototype has the reference's name and signature (dogleg.h:214-392), so
 * a program written against libdogleg (its sample.c and test-misc.c included)
 * recompiles and relinks against this library unchanged. What differs is below
 * the API: the per-iteration linear algebra runs in sm_100a CUDA kernels
 * (see dogleg_gpu.h), CHOLMOD/LAPACK are not used.
 *
 * Conventions (same as the reference):
 *   - the cost being minimised is norm2(x(p)), x has Nmeas entries, p has Nstate
 *   - sparse callbacks fill the TRANSPOSED Jacobian Jt as a CCS matrix:
 *     Nstate rows, Nmeas columns, column j = gradient of measurement j,
 *     32-bit indices, p[Nmeas] must be written too
 *   - dense callbacks fill J row-first: J[i*Nstate + k] = d x_i / d p_k
 *   - "packed upper" symmetric storage is row-first: [A B C / B D E / C E F]
 *     is stored as A B C D E F
 */
#pragma once

#include <stddef.h>
#include <stdbool.h>
#include <cholmod.h>   /* SuiteSparse's if installed, else compat/cholmod.h */

#ifdef __cplusplus
#define DOGLEG_STATIC_ASSERT(c,msg) static_assert(c,msg)
extern "C" {
#else
#define DOGLEG_STATIC_ASSERT(c,msg) _Static_assert(c,msg)
#endif

/* ------------------------------------------------------------------ callbacks */

/* reference dogleg.h:11-20 */
typedef void (dogleg_callback_t)(const double*   p,      /* in  (Nstate,)             */
                                 double*         x,      /* out (Nmeas,)              */
                                 cholmod_sparse* Jt,     /* out (Nstate,Nmeas) CCS    */
                                 void*           cookie);

/* reference dogleg.h:21-30 */
typedef void (dogleg_callback_dense_t)(const double* p,  /* in  (Nstate,)             */
                                       double*       x,  /* out (Nmeas,)              */
                                       double*       J,  /* out (Nmeas,Nstate) rows   */
                                       void*         cookie);

/* reference dogleg.h:34-45. The user hands back the products instead of x, J */
typedef void (dogleg_callback_dense_products_t)(const double* p,      /* in  (Nstate,) */
                                                double*       norm2x, /* out scalar    */
                                                double*       xtJ,    /* out (Nstate,) */
                                                double*       JtJ,    /* out symmetric, maybe packed */
                                                void*         cookie);

/* ------------------------------------------------------------ operating point */

/* reference dogleg.h:48-105; 104 bytes. During a solve the vectors live in HBM;
 * the host pointers below are pinned mirrors which the library brings up to
 * date before control returns to the caller (and p before every callback). */
typedef struct
{
  double* p;

  double* x;
  double  norm2_x;
  union
  {
    cholmod_sparse* Jt;       /* DOGLEG_SPARSE                      */
    double*         J_dense;  /* DOGLEG_DENSE, one gradient per row */
    double*         JtJ;      /* DOGLEG_DENSE_PRODUCTS              */
  };
  double* Jt_x;

  /* cached per point so that a rejected step can be retried cheaply */
  double* updateCauchy;
  union
  {
    cholmod_dense* updateGN_cholmoddense;
    double*        updateGN_dense;
  };
  double norm2_updateCauchy, norm2_updateGN;

  /* validity bits, in declaration order from bit 0 */
  union
  {
    int dummy_bits[3];
    struct
    {
      bool have_updateCauchy          : 1;
      bool have_updateGN              : 1;
      bool have_factorization         : 1;
      bool have_x                     : 1;
      bool have_J                     : 1;
      bool have_Jtx                   : 1;
      bool have_JtJ                   : 1;
      bool have_step_to_here          : 1;
      bool didStepToEdgeOfTrustRegion : 1;
    };
  };

  double* step_to_here;
  double  norm2_step_to_here;
} dogleg_operatingPoint_t;

/* ----------------------------------------------------------------- parameters */

#define DOGLEG_DEBUG_VNLOG_BIT 30
#define DOGLEG_DEBUG_VNLOG     (1 << DOGLEG_DEBUG_VNLOG_BIT)

/* reference dogleg.h:112-152; 72 bytes */
typedef struct
{
  int max_iterations;
  union
  {
    int dogleg_debug;          /* legacy view of the same word */
    struct
    {
      bool debug       : 1;    /* human-readable log on stderr               */
      bool JtJ_packed  : 1;    /* dense-products: JtJ is a packed triangle   */
      bool JtJ_upper   : 1;    /* ... the (row-first) upper one              */
      int  dummy       : DOGLEG_DEBUG_VNLOG_BIT-3;
      bool debug_vnlog : 1;    /* vnlog table on stdout; bit 30              */
    };
  };

  double trustregion0;

  double trustregion_decrease_factor;
  double trustregion_decrease_threshold;
  double trustregion_increase_factor;
  double trustregion_increase_threshold;

  double Jt_x_threshold;
  double update_threshold;
  double trustregion_threshold;
} dogleg_parameters2_t;
DOGLEG_STATIC_ASSERT(offsetof(dogleg_parameters2_t,trustregion0) == 2*sizeof(int),
                     "dogleg_parameters2_t must keep the libdogleg 0.16 layout");

typedef enum {
  DOGLEG_DENSE          = 0,
  DOGLEG_SPARSE         = 1,
  DOGLEG_DENSE_PRODUCTS = 2 } dogleg_solve_type_t;
DOGLEG_STATIC_ASSERT(sizeof(dogleg_solve_type_t) == sizeof(int),
                     "dogleg_solve_type_t must be int-sized");

/* -------------------------------------------------------------------- context */

/* reference dogleg.h:166-210. 'common' is not used for arithmetic here: its
 * dlb_private member (shim builds) or the side table keyed on the context
 * address (real-SuiteSparse builds) leads to the device-side state. */
typedef struct
{
  cholmod_common common;

  union
  {
    dogleg_callback_t*                f;
    dogleg_callback_dense_t*          f_dense;
    dogleg_callback_dense_products_t* f_dense_products;
  };
  void* cookie;

  dogleg_operatingPoint_t* beforeStep;   /* the current point between steps   */
  dogleg_operatingPoint_t* afterStep;    /* scratch while a step is tried     */

  union
  {
    cholmod_factor* factorization;        /* sparse: descriptor of the device factor */
    double*         factorization_dense;  /* dense: host mirror of the Cholesky factor */
  };

  double lambda;                          /* sticky diagonal loading, starts at 0 */

  dogleg_solve_type_t solve_type;
  int Nstate, Nmeasurements;

  const dogleg_parameters2_t* parameters;
} dogleg_solverContext_t;

/* ------------------------------------------------------------- parameter API */

void dogleg_getDefaultParameters(dogleg_parameters2_t* parameters);

/* legacy process-global settings used when 'parameters' is NULL */
void dogleg_setMaxIterations(int n);
void dogleg_setTrustregionUpdateParameters(double downFactor, double downThreshold,
                                           double upFactor,   double upThreshold);
void dogleg_setDebug(int debug);                 /* 0, DOGLEG_DEBUG_VNLOG, or any other bit */
void dogleg_setInitialTrustregion(double t);
void dogleg_setThresholds(double Jt_x, double update, double trustregion); /* <=0: leave alone */

/* -------------------------------------------------------------------- solvers */

/* All return norm2(x) at the optimum, or a negative number on error. p is
 * in/out. If returnContext is non-NULL the caller owns the context afterwards
 * and releases it with dogleg_freeContext(). */
double dogleg_optimize(double* p, unsigned int Nstate,
                       unsigned int Nmeas, unsigned int NJnnz,
                       dogleg_callback_t* f, void* cookie,
                       dogleg_solverContext_t** returnContext);
double dogleg_optimize2(double* p, unsigned int Nstate,
                        unsigned int Nmeas, unsigned int NJnnz,
                        dogleg_callback_t* f, void* cookie,
                        const dogleg_parameters2_t* parameters,
                        dogleg_solverContext_t** returnContext);

double dogleg_optimize_dense(double* p, unsigned int Nstate,
                             unsigned int Nmeas,
                             dogleg_callback_dense_t* f, void* cookie,
                             dogleg_solverContext_t** returnContext);
double dogleg_optimize_dense2(double* p, unsigned int Nstate,
                              unsigned int Nmeas,
                              dogleg_callback_dense_t* f, void* cookie,
                              const dogleg_parameters2_t* parameters,
                              dogleg_solverContext_t** returnContext);
double dogleg_optimize_dense_products(double* p, unsigned int Nstate,
                                      dogleg_callback_dense_products_t* f, void* cookie,
                                      const dogleg_parameters2_t* parameters,
                                      dogleg_solverContext_t** returnContext);

/* Make sure the Cholesky factor of JtJ (+lambda I) at 'point' exists. */
bool dogleg_computeJtJfactorization(dogleg_operatingPoint_t* point, dogleg_solverContext_t* ctx);

/* vnlog table of reported vs central-difference gradients for one variable */
void dogleg_testGradient(unsigned int var, const double* p0,
                         unsigned int Nstate, unsigned int Nmeas, unsigned int NJnnz,
                         dogleg_callback_t* f, void* cookie);
void dogleg_testGradient_dense(unsigned int var, const double* p0,
                               unsigned int Nstate, unsigned int Nmeas,
                               dogleg_callback_dense_t* f, void* cookie);
void dogleg_testGradient_dense_products(unsigned int var, const double* p0,
                                        unsigned int Nstate, unsigned int Nmeas,
                                        dogleg_callback_dense_products_t* f, void* cookie);

void dogleg_freeContext(dogleg_solverContext_t** ctx);

/* ---------------------------------------------- outlier helpers (experimental) */

bool dogleg_getOutliernessFactors(double* factors,      /* out: Nfeatures          */
                                  double* scale,        /* in/out: <0 => recompute */
                                  int featureSize,
                                  int Nfeatures,
                                  int NoutlierFeatures,
                                  dogleg_operatingPoint_t* point,
                                  dogleg_solverContext_t* ctx);

struct dogleg_outliers_t
{
  unsigned char marked : 1;
};

bool dogleg_markOutliers(struct dogleg_outliers_t* markedOutliers,   /* in/out */
                         double* scale,                              /* in/out */
                         int*    Noutliers,                          /* in/out */
                         double (getConfidence)(int i_feature_exclude),
                         int featureSize,
                         int Nfeatures,
                         dogleg_operatingPoint_t* point,
                         dogleg_solverContext_t* ctx);

void dogleg_reportOutliers(double (getConfidence)(int i_feature_exclude),
                           double* scale,
                           int featureSize,
                           int Nfeatures,
                           int Noutliers,
                           dogleg_operatingPoint_t* point,
                           dogleg_solverContext_t* ctx);

double dogleg_getOutliernessTrace_newFeature_sparse(const double*            JqueryFeature,
                                                    int                      istateActive,
                                                    int                      NstateActive,
                                                    int                      featureSize,
                                                    int                      NoutlierFeatures,
                                                    dogleg_operatingPoint_t* point,
                                                    dogleg_solverContext_t*  ctx);

#ifdef __cplusplus
}
#endif
