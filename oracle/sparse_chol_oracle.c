/* oracle/sparse_chol_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * See sparse_chol_oracle.h for what this restates and why parity with real
 * CHOLMOD is unpinned. Written for clarity, single-threaded, O(sum nnz_col^2)
 * like the simplicial row factorization it follows.
 */
#include "sparse_chol_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdint.h>

static void* xcalloc(size_t n, size_t sz) { void* p = calloc(n ? n : 1, sz); if(!p) abort(); return p; }

/* ------------------------------------------------------------- ordering */

/* Exact minimum degree on an explicit bitset elimination graph. Fine for the
 * oracle's sizes (n up to a few thousand). */
void orc_min_degree(int n, int m, const int* Ap, const int* Ai, int* perm)
{
  const int W = (n + 63) / 64;
  uint64_t* adj = xcalloc((size_t)n * W, sizeof(uint64_t));
#define ADJ(i) (adj + (size_t)(i) * W)
  for(int j = 0; j < m; j++)
    for(int a = Ap[j]; a < Ap[j+1]; a++)
      for(int b = Ap[j]; b < Ap[j+1]; b++)
        if(Ai[a] != Ai[b]) ADJ(Ai[a])[Ai[b] >> 6] |= 1ull << (Ai[b] & 63);

  char* gone = xcalloc(n, 1);
  int*  deg  = xcalloc(n, sizeof(int));
  for(int i = 0; i < n; i++)
  {
    int d = 0;
    for(int w = 0; w < W; w++) d += __builtin_popcountll(ADJ(i)[w]);
    deg[i] = d;
  }
  for(int k = 0; k < n; k++)
  {
    int best = -1;
    for(int i = 0; i < n; i++)
      if(!gone[i] && (best < 0 || deg[i] < deg[best])) best = i;
    perm[k] = best;
    gone[best] = 1;
    const uint64_t* nb = ADJ(best);
    for(int w = 0; w < W; w++)
    {
      uint64_t bits = nb[w];
      while(bits)
      {
        int u = (w << 6) + __builtin_ctzll(bits);
        bits &= bits - 1;
        uint64_t* au = ADJ(u);
        int d = 0;
        for(int v = 0; v < W; v++) { au[v] |= nb[v]; }
        au[u    >> 6] &= ~(1ull << (u    & 63));
        au[best >> 6] &= ~(1ull << (best & 63));
        for(int v = 0; v < W; v++) d += __builtin_popcountll(au[v]);
        deg[u] = d;
      }
    }
  }
#undef ADJ
  free(adj); free(gone); free(deg);
}

/* ------------------------------------------------------------- analyze */

orc_factor* orc_analyze(int n, int m, const int* Ap, const int* Ai, const int* user_perm)
{
  orc_factor* F = xcalloc(1, sizeof(*F));
  F->n = n; F->m = m; F->minor = n;
  F->perm   = xcalloc(n, sizeof(int));
  F->iperm  = xcalloc(n, sizeof(int));
  F->parent = xcalloc(n, sizeof(int));
  F->colcount = xcalloc(n, sizeof(int));
  F->Lp     = xcalloc(n + 1, sizeof(int));
  F->Lnz    = xcalloc(n, sizeof(int));

  if(user_perm) memcpy(F->perm, user_perm, n * sizeof(int));
  else          orc_min_degree(n, m, Ap, Ai, F->perm);
  for(int k = 0; k < n; k++) F->iperm[F->perm[k]] = k;

  /* R = A(perm,:) stored by rows: for permuted row k the columns j it touches
   * and where the value sits in Ax (CHOLMOD builds F = A(p,:)' the same way) */
  const int nnz = Ap[m];
  F->Rp = xcalloc(n + 1, sizeof(int));
  F->Rj = xcalloc(nnz, sizeof(int));
  F->Rsrc = xcalloc(nnz, sizeof(int));
  for(int q = 0; q < nnz; q++) F->Rp[F->iperm[Ai[q]] + 1]++;
  for(int k = 0; k < n; k++) F->Rp[k+1] += F->Rp[k];
  int* fill = xcalloc(n, sizeof(int));
  for(int j = 0; j < m; j++)
    for(int q = Ap[j]; q < Ap[j+1]; q++)
    {
      int k = F->iperm[Ai[q]];
      int dst = F->Rp[k] + fill[k]++;
      F->Rj[dst] = j; F->Rsrc[dst] = q;
    }
  free(fill);

  /* elimination tree (Liu, with path compression) on the pattern of C = A A' */
  int* ancestor = xcalloc(n, sizeof(int));
  for(int k = 0; k < n; k++)
  {
    F->parent[k] = -1; ancestor[k] = -1;
    for(int r = F->Rp[k]; r < F->Rp[k+1]; r++)
    {
      int j = F->Rj[r];
      for(int q = Ap[j]; q < Ap[j+1]; q++)
      {
        int i = F->iperm[Ai[q]];
        while(i != -1 && i < k)
        {
          int inext = ancestor[i];
          ancestor[i] = k;
          if(inext == -1) F->parent[i] = k;
          i = inext;
        }
      }
    }
  }
  free(ancestor);

  /* column counts: one symbolic up-looking sweep (row k of L = etree reach) */
  int* mark = xcalloc(n, sizeof(int));
  for(int k = 0; k < n; k++) mark[k] = -1;
  for(int k = 0; k < n; k++)
  {
    mark[k] = k;
    F->colcount[k]++;                       /* diagonal */
    for(int r = F->Rp[k]; r < F->Rp[k+1]; r++)
    {
      int j = F->Rj[r];
      for(int q = Ap[j]; q < Ap[j+1]; q++)
      {
        int i = F->iperm[Ai[q]];
        for(; i < k && mark[i] != k; i = F->parent[i]) { mark[i] = k; F->colcount[i]++; }
      }
    }
  }
  free(mark);
  for(int k = 0; k < n; k++) F->Lp[k+1] = F->Lp[k] + F->colcount[k];
  F->Li = xcalloc(F->Lp[n], sizeof(int));
  F->Lx = xcalloc(F->Lp[n], sizeof(double));
  return F;
}

/* ----------------------------------------------------------- factorize */

int orc_factorize(orc_factor* F, const int* Ap, const int* Ai, const double* Ax,
                  double beta, int is_ll)
{
  const int n = F->n;
  double* x    = xcalloc(n, sizeof(double));
  int*    mark = xcalloc(n, sizeof(int));
  int*    stack = xcalloc(n, sizeof(int));
  int*    path  = xcalloc(n, sizeof(int));
  F->is_ll = is_ll; F->minor = n;
  for(int k = 0; k < n; k++) { mark[k] = -1; F->Lnz[k] = 0; }

  for(int k = 0; k < n; k++)
  {
    /* column k of the upper triangle of beta*I + A A', formed on the fly, and
     * its etree reach (= pattern of row k of L) in topological order */
    int top = n;
    mark[k] = k;
    for(int r = F->Rp[k]; r < F->Rp[k+1]; r++)
    {
      const int    j   = F->Rj[r];
      const double akj = Ax[F->Rsrc[r]];
      for(int q = Ap[j]; q < Ap[j+1]; q++)
      {
        int i = F->iperm[Ai[q]];
        if(i > k) continue;
        x[i] += Ax[q] * akj;
        int len = 0;
        for(; mark[i] != k; i = F->parent[i]) { path[len++] = i; mark[i] = k; }
        while(len > 0) stack[--top] = path[--len];
      }
    }
    double d = x[k] + beta;
    x[k] = 0.0;
    F->Li[F->Lp[k]] = k;
    F->Lnz[k] = 1;

    for(; top < n; top++)
    {
      const int i  = stack[top];
      const int p0 = F->Lp[i], p1 = p0 + F->Lnz[i];
      double lki;
      if(is_ll)
      {
        lki = x[i] / F->Lx[p0];
        x[i] = 0.0;
        for(int p = p0 + 1; p < p1; p++) x[F->Li[p]] -= F->Lx[p] * lki;
        d -= lki * lki;
      }
      else
      {
        const double yi = x[i];
        x[i] = 0.0;
        for(int p = p0 + 1; p < p1; p++) x[F->Li[p]] -= F->Lx[p] * yi;
        lki = yi / F->Lx[p0];
        d -= lki * yi;
      }
      F->Li[p1] = k; F->Lx[p1] = lki; F->Lnz[i]++;
    }

    if(is_ll)
    {
      if(!(d > 0.0) || !isfinite(d)) { F->minor = k; break; }
      F->Lx[F->Lp[k]] = sqrt(d);
    }
    else
    {
      if(d == 0.0 || !isfinite(d)) { F->minor = k; break; }
      F->Lx[F->Lp[k]] = d;
    }
  }
  free(x); free(mark); free(stack); free(path);
  return 1;
}

/* --------------------------------------------------------------- solve */

void orc_solve(const orc_factor* F, const double* B, double* X, int nrhs)
{
  const int n = F->n;
  double* y = xcalloc(n, sizeof(double));
  for(int c = 0; c < nrhs; c++)
  {
    const double* b = B + (size_t)c * n;
    for(int k = 0; k < n; k++) y[k] = b[F->perm[k]];
    for(int k = 0; k < n; k++)
    {
      const int p0 = F->Lp[k], p1 = p0 + F->Lnz[k];
      if(F->is_ll) y[k] /= F->Lx[p0];
      for(int p = p0 + 1; p < p1; p++) y[F->Li[p]] -= F->Lx[p] * y[k];
    }
    if(!F->is_ll) for(int k = 0; k < n; k++) y[k] /= F->Lx[F->Lp[k]];
    for(int k = n - 1; k >= 0; k--)
    {
      const int p0 = F->Lp[k], p1 = p0 + F->Lnz[k];
      for(int p = p0 + 1; p < p1; p++) y[k] -= F->Lx[p] * y[F->Li[p]];
      if(F->is_ll) y[k] /= F->Lx[p0];
    }
    double* xo = X + (size_t)c * n;
    for(int k = 0; k < n; k++) xo[F->perm[k]] = y[k];
  }
  free(y);
}

void orc_free(orc_factor* F)
{
  if(!F) return;
  free(F->perm); free(F->iperm); free(F->parent); free(F->colcount);
  free(F->Lp); free(F->Li); free(F->Lx); free(F->Lnz);
  free(F->Rp); free(F->Rj); free(F->Rsrc);
  free(F);
}

/* -------------------------------------------------- pattern of A A' (tests) */

void orc_aat_lower_pattern(int n, int m, const int* Ap, const int* Ai, const int* iperm,
                           int** Cp_out, int** Ci_out)
{
  /* rows of permuted A */
  const int nnz = Ap[m];
  int* Rp = xcalloc(n + 1, sizeof(int));
  int* Rj = xcalloc(nnz, sizeof(int));
  for(int q = 0; q < nnz; q++) Rp[(iperm ? iperm[Ai[q]] : Ai[q]) + 1]++;
  for(int k = 0; k < n; k++) Rp[k+1] += Rp[k];
  int* fill = xcalloc(n, sizeof(int));
  for(int j = 0; j < m; j++)
    for(int q = Ap[j]; q < Ap[j+1]; q++)
    {
      int k = iperm ? iperm[Ai[q]] : Ai[q];
      Rj[Rp[k] + fill[k]++] = j;
    }
  free(fill);

  int* mark = xcalloc(n, sizeof(int));
  int* Cp = xcalloc(n + 1, sizeof(int));
  for(int pass = 0; pass < 2; pass++)
  {
    int* Ci = pass ? *Ci_out : NULL;
    for(int k = 0; k < n; k++) mark[k] = -1;
    int cnt = 0;
    for(int k = 0; k < n; k++)
    {
      int start = cnt;
      for(int r = Rp[k]; r < Rp[k+1]; r++)
      {
        int j = Rj[r];
        for(int q = Ap[j]; q < Ap[j+1]; q++)
        {
          int i = iperm ? iperm[Ai[q]] : Ai[q];
          if(i >= k && mark[i] != k) { mark[i] = k; if(Ci) Ci[cnt] = i; cnt++; }
        }
      }
      if(Ci)
      { /* insertion sort of the column */
        for(int a = start + 1; a < cnt; a++)
        {
          int v = Ci[a], b = a - 1;
          while(b >= start && Ci[b] > v) { Ci[b+1] = Ci[b]; b--; }
          Ci[b+1] = v;
        }
      }
      Cp[k+1] = cnt;
    }
    if(!pass) *Ci_out = xcalloc(cnt, sizeof(int));
  }
  free(mark); free(Rp); free(Rj);
  *Cp_out = Cp;
}
