/* oracle/sparse_chol_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of what SuiteSparse CHOLMOD does for libdogleg
 * (reference call sites: dogleg.c:650-665 analyze/factorize, :853-856 solve).
 * CHOLMOD itself is a third-party dependency that is NOT vendored in
 * /root/reference and NOT installed in this image (no version is pinned by the
 * reference: Makefile:23 links -lcholmod, dogleg.c:17-19,1603-1610 support
 * <=2.2, 2.3-3.x and >=4.0). What is restated here is its published simplicial
 * algorithm, which is the one libdogleg selects (supernodal=0, dogleg.c:1599):
 *   analyze   : fill-reducing ordering of A*A', elimination tree, column counts
 *   factorize : up-looking row-by-row LDL' (or LL') of beta*I + A*A', the
 *               product A*A' formed on the fly one column at a time
 *   solve     : x = P' (L D L')^-1 P b
 * PARITY UNPINNED against real CHOLMOD (none available); pinned instead against
 * the reference's dense LAPACK path on densified problems (tests/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * legs may link or call this.
 */
#ifndef ORC_SPARSE_CHOL_H
#define ORC_SPARSE_CHOL_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct
{
  int     n;
  int*    perm;      /* perm[k]  = original index eliminated k-th               */
  int*    iperm;     /* iperm[i] = position of original index i                 */
  int*    parent;    /* elimination tree of the permuted matrix, -1 at roots    */
  int*    colcount;  /* nnz in each column of L, diagonal included              */
  int*    Lp;        /* n+1 column pointers                                     */
  int*    Li;        /* row indices; the first entry of a column is its diagonal*/
  double* Lx;        /* LDL': Lx[Lp[k]] = D(k,k), unit diagonal implied         */
                     /* LL' : Lx[Lp[k]] = L(k,k)                                */
  int*    Lnz;       /* entries filled so far per column (numeric phase)        */
  int     is_ll;
  int     minor;     /* == n on success, else the first failed pivot            */
  /* row-major copy of the permuted A used to form A*A' on the fly */
  int*    Rp; int* Rj; int* Rsrc;
  int     m;
} orc_factor;

/* A is n x m in CCS (Ap has m+1 entries, Ai row indices < n).
 * user_perm == NULL selects the oracle's own exact minimum-degree ordering. */
orc_factor* orc_analyze(int n, int m, const int* Ap, const int* Ai, const int* user_perm);

/* numeric factorization of beta*I + A*A'. is_ll=0: LDL' where only an exactly
 * zero pivot is a failure (CHOLMOD simplicial rule); is_ll=1: LL' where a
 * pivot <= 0 or non-finite is a failure (LAPACK dpptrf rule). Returns 1 if the
 * call itself succeeded (look at F->minor for definiteness). */
int orc_factorize(orc_factor* F, const int* Ap, const int* Ai, const double* Ax,
                  double beta, int is_ll);

/* B, X are n x nrhs column-major with leading dimension n */
void orc_solve(const orc_factor* F, const double* B, double* X, int nrhs);

void orc_free(orc_factor* F);

/* lower-triangular pattern of (A*A')(perm,perm) as CSC; caller frees *Cp,*Ci */
void orc_aat_lower_pattern(int n, int m, const int* Ap, const int* Ai, const int* iperm,
                           int** Cp, int** Ci);

/* exact minimum degree on the graph of A*A' (ties -> lowest index) */
void orc_min_degree(int n, int m, const int* Ap, const int* Ai, int* perm);

#ifdef __cplusplus
}
#endif
#endif
