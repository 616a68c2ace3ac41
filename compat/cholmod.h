/* compat/cholmod.h -- layout shim for builds on machines without SuiteSparse.
 *
 * libdogleg's public header embeds CHOLMOD types (a cholmod_common by value at
 * the head of dogleg_solverContext_t, cholmod_sparse* for Jt, cholmod_dense*
 * for the Gauss-Newton update, cholmod_factor* for the factorization; see
 * /root/reference/dogleg.h:8,18,66,77,168,190). This image has no SuiteSparse,
 * so this file declares the subset of the CHOLMOD surface libdogleg touches
 *  with struct layouts that follow the published
 * SuiteSparse 4.x/5.x headers as far as they are known here [ext: unverified,
 * no cholmod.h is available offline to diff against].
 *
 * When a real <cholmod.h> exists, put its directory ahead of compat/ on the
 * include path and this file is never seen.
 *
 * Two consumers:
 *   - the product (libdogleg_b200/csrc): only the TYPES and constants; the
 *     arithmetic CHOLMOD used to do is done by our CUDA kernels.
 *   - the oracle (oracle/cholmod_shim.c): implements the cholmod_* FUNCTIONS
 *     declared at the bottom on the CPU so that the unmodified reference
 *     dogleg.c links and runs its sparse path (test infrastructure only).
 */
#ifndef DLB_COMPAT_CHOLMOD_H
#define DLB_COMPAT_CHOLMOD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHOLMOD_VER_CODE(main,sub) ((main) * 1000 + (sub))
#define CHOLMOD_MAIN_VERSION 5
#define CHOLMOD_SUB_VERSION  2
#define CHOLMOD_VERSION CHOLMOD_VER_CODE(CHOLMOD_MAIN_VERSION,CHOLMOD_SUB_VERSION)
#define DLB_CHOLMOD_IS_SHIM 1

/* xtype / dtype / itype */
#define CHOLMOD_PATTERN 0
#define CHOLMOD_REAL    1
#define CHOLMOD_COMPLEX 2
#define CHOLMOD_ZOMPLEX 3
#define CHOLMOD_DOUBLE  0
#define CHOLMOD_SINGLE  4
#define CHOLMOD_INT     0
#define CHOLMOD_LONG    2

/* systems cholmod_solve() understands */
#define CHOLMOD_A    0
#define CHOLMOD_LDLt 1
#define CHOLMOD_LD   2
#define CHOLMOD_DLt  3
#define CHOLMOD_L    4
#define CHOLMOD_Lt   5
#define CHOLMOD_D    6
#define CHOLMOD_P    7
#define CHOLMOD_Pt   8

/* supernodal strategy */
#define CHOLMOD_SIMPLICIAL 0
#define CHOLMOD_AUTO       1
#define CHOLMOD_SUPERNODAL 2

/* ordering tags stored in cholmod_factor.ordering */
#define CHOLMOD_NATURAL 0
#define CHOLMOD_GIVEN   1
#define CHOLMOD_AMD     2

typedef struct cholmod_sparse_struct
{
  size_t nrow, ncol, nzmax;
  void  *p, *i, *nz, *x, *z;
  int    stype, itype, xtype, dtype, sorted, packed;
} cholmod_sparse;

typedef struct cholmod_dense_struct
{
  size_t nrow, ncol, nzmax, d;
  void  *x, *z;
  int    xtype, dtype;
} cholmod_dense;

typedef struct cholmod_factor_struct
{
  size_t n, minor;
  void  *Perm, *ColCount, *IPerm;
  /* simplicial part */
  size_t nzmax;
  void  *p, *i, *x, *z, *nz, *next, *prev;
  /* supernodal part */
  size_t nsuper, ssize, xsize, maxcsize, maxesize;
  void  *super, *pi, *px, *s;
  int    ordering, is_ll, is_super, is_monotonic;
  int    itype, xtype, dtype;
  int    useGPU;
} cholmod_factor;

/* Only .supernodal is ever written by libdogleg (reference dogleg.c:1599).
 * The leading members follow the order of the published struct; the tail is an
 * opaque reserve so that code which memsets/copies a cholmod_common is safe.
 * .dlb_private is where each implementation hangs its own state. */
typedef struct cholmod_common_struct
{
  double dbound;
  double grow0, grow1;
  size_t grow2;
  size_t maxrank;
  double supernodal_switch;
  int    supernodal;
  int    final_asis, final_super, final_ll, final_pack, final_monotonic, final_resymbol;
  double zrelax[3];
  size_t nrelax[3];
  int    prefer_zomplex, prefer_upper, quick_return_if_not_posdef, prefer_binary;
  int    print, precise;
  int    try_catch;
  void (*error_handler)(int status, const char* file, int line, const char* message);
  int    nmethods, current, selected;
  int    postorder, default_nesdis;
  int    itype, dtype;
  int    no_workspace_reallocate;
  int    status;
  void*  dlb_private;
  char   dlb_reserved[3072];
} cholmod_common;

/* ---- the functions libdogleg calls ---- */
int             cholmod_start (cholmod_common* c);
int             cholmod_finish(cholmod_common* c);
cholmod_sparse* cholmod_allocate_sparse(size_t nrow, size_t ncol, size_t nzmax,
                                        int sorted, int packed, int stype, int xtype,
                                        cholmod_common* c);
int             cholmod_free_sparse(cholmod_sparse** A, cholmod_common* c);
cholmod_dense*  cholmod_allocate_dense(size_t nrow, size_t ncol, size_t d, int xtype,
                                       cholmod_common* c);
int             cholmod_free_dense(cholmod_dense** X, cholmod_common* c);
cholmod_factor* cholmod_analyze(cholmod_sparse* A, cholmod_common* c);
int             cholmod_factorize(cholmod_sparse* A, cholmod_factor* L, cholmod_common* c);
int             cholmod_factorize_p(cholmod_sparse* A, double beta[2], int* fset, size_t fsize,
                                    cholmod_factor* L, cholmod_common* c);
int             cholmod_free_factor(cholmod_factor** L, cholmod_common* c);
cholmod_dense*  cholmod_solve(int sys, cholmod_factor* L, cholmod_dense* B, cholmod_common* c);
cholmod_sparse* cholmod_spsolve(int sys, cholmod_factor* L, cholmod_sparse* B, cholmod_common* c);
void            SuiteSparse_config_printf_func_set(int (*f)(const char*, ...));

#ifdef __cplusplus
}
#endif
#endif
