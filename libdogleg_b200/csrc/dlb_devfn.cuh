// dlb_devfn.cuh -- device functions shared by the per-level kernels (dlb_front.cu, dlb_sparse.cu) and
// the persistent trial kernel (dlb_trial.cu): one implementation of the pivot elimination, of the
// block gather of the extend-add and of the quadratic form over the class blocks, so that both
// schedules perform the same operations in the same order.
#pragma once
#include "dlb_common.cuh"
#include "dlb_device.h"

// -DDLB_TRIAL_DEBUG (dlb_trial.cu): SM-cycle stamps of CTA 0 / thread 0 at DBG_MARK() points
#ifdef DLB_ELIM_DBG
static __device__ unsigned long long g_elim_dbg[256];
static __device__ int g_elim_n;
#define DBG_MARK() do { if(blockIdx.x == 0 && threadIdx.x == 0 && g_elim_n < 255) { g_elim_dbg[g_elim_n++] = (unsigned long long)clock64(); } } while(0)
#else
#define DBG_MARK() do { } while(0)
#endif

// Inactive lanes / list slots load from here instead of being predicated off: `cond ? *p : 0.0` made ptxas
// funnel the predicated loads through ONE temporary register (load, select, reuse), which serialises
// loads that are independent -- measured 340 cycles per load in a batch of 32. Unconditional loads from a
// clamped address keep their own destination registers and pipeline.
static __device__ double g_dlb_zero[8];

__device__ __forceinline__ void pair_from_index(int q, int& a, int& b)
{
  a = (int)((sqrt(8.0 * (double)q + 1.0) - 1.0) * 0.5);
  while((a + 1) * (a + 2) / 2 <= q) a++;
  while(a * (a + 1) / 2 > q) a--;
  b = q - a * (a + 1) / 2;
}

// FP64 tensor-core tile update D = A*B + C (mma.sync m8n8k4, SASS DMMA)
__device__ __forceinline__ void devfn_dmma(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Eliminate the first nc pivot columns of the column-major lower-triangular front A (r columns, leading
// dimension ld >= nrows; shared or global memory), blocked by 8 columns, LEFT-looking inside the front. At
// the 2-8 warps per scheduler these fronts run with, everything is a latency chain (measured on B200,
// profiles/micro/latency.cu: dependent DFMA 9, DMMA 26, LDS 29, 64-bit shuffle 30, rsqrt 53 cycles), so per
// block of 8 pivots:
//   1. the block's 8 columns (rows b0..nrows) receive the updates of ALL earlier pivots at once, as 8x8
//      tiles on the FP64 tensor cores with the K loop running over the earlier columns and the tile in
//      registers: C is read and written once per block (the right-looking form of round 2's first half
//      re-read and re-wrote the whole trailing front after every block: 6 shared-memory accesses per two
//      DMMA); a warp owns up to four row tiles at a time and shares their B fragment;
//   2. after a barrier EVERY warp factorizes the 8x8 diagonal block redundantly in registers (36 values,
//      fully unrolled, no shuffles, no shared-memory round trip): the chain from one pivot to the next is
//      rsqrt -> multiply -> multiply-add;
//   3. each thread substitutes the 8 columns of its rows below the block with that register copy; barrier.
// The columns behind the pivots (the update matrix, r > nc) get all nc pivots in one pass at the end, K = nc.
// Two block barriers per 8 pivots. Pivots are formed as d*rsqrt(d) (divergence 7).
// nrows = r: the plain front. nrows = r + 1: the front carries a right-hand side as an extra ROW, which
// the elimination turns into the forward substitution y = L^-1 b (its first nc entries) and the
// updated right-hand side of the parent (the rest). dinv (nc entries, may be NULL) receives 1/L_jj.
// All threads of the CTA call it; returns the first failing column (pivot <= 0 or not finite) or -1.
template<int NT>
__device__ __forceinline__ int front_eliminate(double* A, int r, int nc, int ld, int nrows, int tid, double* dinv)
{
  constexpr int NW = NT / 32;
  const int lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, tt = lane & 3;
  int fail = -1;
  __syncthreads();
  DBG_MARK();
  for(int b0 = 0; b0 < nc && fail < 0; b0 += 8)
  {
    const int bw = nc - b0 < 8 ? nc - b0 : 8;
    // ---- 1. rows [b0, nrows) of the block's columns -= A[rows, 0:b0] A[b0:b0+8, 0:b0]' ----
    if(b0 > 0)
    {
      const int ntr = (nrows - b0 + 7) >> 3;
      const int rb = b0 + g < r ? b0 + g : r - 1;                    // B fragment: the block's own rows
      const int cj = b0 + 2 * tt;
      // tiles dealt round-robin over the warps (a front of <= 8 row tiles: one per warp); two accumulator sets
      // over the even / odd K steps break the DMMA dependency chain of a warp with a single tile
      for(int ti0 = w; ti0 < ntr; ti0 += 4 * NW)
      {
        int ri[4], ric[4]; double c0[4], c1[4], d0[4], d1[4];
#pragma unroll
        for(int u = 0; u < 4; u++)
        {
          ri[u] = b0 + 8 * (ti0 + NW * u) + g;
          ric[u] = ri[u] < nrows ? ri[u] : nrows - 1;
          c0[u] = 0.0; c1[u] = 0.0; d0[u] = 0.0; d1[u] = 0.0;
        }
        for(int k = 0; k < b0; k += 8)
        {                                                             // b0 is a multiple of 8: no K remainder
          const size_t ko = (size_t)(k + tt) * ld, kp = ko + 4 * (size_t)ld;
          const double vb = A[rb + ko], wb = A[rb + kp];
          double va[4], wa[4];
#pragma unroll
          for(int u = 0; u < 4; u++) { va[u] = A[ric[u] + ko]; wa[u] = A[ric[u] + kp]; }
#pragma unroll
          for(int u = 0; u < 4; u++) { devfn_dmma(c0[u], c1[u], va[u], vb); devfn_dmma(d0[u], d1[u], wa[u], wb); }
        }
#pragma unroll
        for(int u = 0; u < 4; u++)
          if(ti0 + NW * u < ntr && ri[u] < nrows)
          {
            if(cj < b0 + bw && (ri[u] >= cj || ri[u] >= r))         A[ri[u] + (size_t)cj * ld]       -= c0[u] + d0[u];
            if(cj + 1 < b0 + bw && (ri[u] >= cj + 1 || ri[u] >= r)) A[ri[u] + (size_t)(cj + 1) * ld] -= c1[u] + d1[u];
          }
      }
      __syncthreads();
    }
    DBG_MARK();
    // ---- 2. the diagonal block, in every lane's registers (identity beyond a short last block) ----
    double L[36], rs[8];
#pragma unroll
    for(int i = 0; i < 8; i++)
#pragma unroll
      for(int j = 0; j <= i; j++)
      { // clamped, unconditional loads (see g_dlb_zero); rows beyond the block become identity rows
        const double val = A[(b0 + (i < bw ? i : 0)) + (size_t)(b0 + (i < bw ? j : 0)) * ld];
        L[i * (i + 1) / 2 + j] = i < bw ? val : (i == j ? 1.0 : 0.0);
      }
#pragma unroll
    for(int j = 0; j < 8; j++)
    {
      const double d = L[j * (j + 1) / 2 + j];
      if((!(d > 0.0) || isinf(d)) && fail < 0) fail = b0 + j;
      rs[j] = rsqrt(d);                              // no sqrt -> division chain per pivot
      L[j * (j + 1) / 2 + j] = d * rs[j];
#pragma unroll
      for(int i = j + 1; i < 8; i++) L[i * (i + 1) / 2 + j] *= rs[j];
#pragma unroll
      for(int i = j + 1; i < 8; i++)
#pragma unroll
        for(int c = j + 1; c <= i; c++)
          L[i * (i + 1) / 2 + c] = fma(-L[i * (i + 1) / 2 + j], L[c * (c + 1) / 2 + j], L[i * (i + 1) / 2 + c]);
    }
    if(fail >= 0) break;                             // the same decision in every thread
    DBG_MARK();
    // ---- 3. rows below the block ----
    for(int i = b0 + bw + tid; i < nrows; i += NT)
    {
      double x[8], v[8];
#pragma unroll
      for(int c = 0; c < 8; c++) v[c] = A[i + (size_t)(b0 + (c < bw ? c : 0)) * ld];
#pragma unroll
      for(int c = 0; c < 8; c++)
      {
        double t = c < bw ? v[c] : 0.0;
#pragma unroll
        for(int cp = 0; cp < c; cp++) t = fma(-x[cp], L[c * (c + 1) / 2 + cp], t);
        x[c] = t * rs[c];
        if(c < bw) A[i + (size_t)(b0 + c) * ld] = x[c];
      }
    }
    if(w == NW - 1 && lane < bw)
    { // the factorized block (row = lane) and the reciprocal pivots: nothing reads them before the final barrier
#pragma unroll
      for(int i = 0; i < 8; i++)
        if(i == lane)
        {
#pragma unroll
          for(int j = 0; j <= i; j++) A[(b0 + i) + (size_t)(b0 + j) * ld] = L[i * (i + 1) / 2 + j];
          if(dinv) dinv[b0 + i] = rs[i];
        }
    }
    __syncthreads();
    DBG_MARK();
  }
  // ---- the update matrix: rows [nc, nrows) x columns [nc, r), row >= column, -= A[rows, 0:nc] A[cols, 0:nc]' ----
  // A warp owns a ROW of tiles (its A fragments are loaded once per K step), four tiles of the row at a time.
  if(fail < 0 && r > nc)
  {
    const int t0 = nc;
    const int ntr = (nrows - t0 + 7) >> 3, ntc = (r - t0 + 7) >> 3;
    for(int ti = ntr - 1 - w; ti >= 0; ti -= NW)
    {
      const int ri = t0 + 8 * ti + g;
      const int ric = ri < nrows ? ri : nrows - 1;
      const int ntj = ti < ntc ? ti + 1 : ntc;                       // tiles of this row on or below the diagonal
      for(int tj0 = 0; tj0 < ntj; tj0 += 4)
      {
        int rjc[4]; double c0[4], c1[4];
#pragma unroll
        for(int u = 0; u < 4; u++)
        {
          const int tj = tj0 + u < ntj ? tj0 + u : ntj - 1;
          const int rj = t0 + 8 * tj + g;
          rjc[u] = rj < r ? rj : r - 1;
          c0[u] = 0.0; c1[u] = 0.0;
        }
        for(int k = 0; k < nc; k += 4)
        {
          const bool kon = k + tt < nc;
          const size_t ko = (size_t)(kon ? k + tt : 0) * ld;
          const double xa = A[ric + ko];
          const double va = kon ? xa : 0.0;                           // zero A fragment beyond the last pivot
          double vb[4];
#pragma unroll
          for(int u = 0; u < 4; u++) vb[u] = A[rjc[u] + ko];
#pragma unroll
          for(int u = 0; u < 4; u++) devfn_dmma(c0[u], c1[u], va, vb[u]);
        }
#pragma unroll
        for(int u = 0; u < 4; u++)
          if(tj0 + u < ntj && ri < nrows)
          {
            const int cj = t0 + 8 * (tj0 + u) + 2 * tt;
            if(cj < r && (ri >= cj || ri >= r))         A[ri + (size_t)cj * ld]       -= c0[u];
            if(cj + 1 < r && (ri >= cj + 1 || ri >= r)) A[ri + (size_t)(cj + 1) * ld] -= c1[u];
          }
      }
    }
    __syncthreads();
    DBG_MARK();
  }
  return fail;
}

// Block gather (extend-add of update matrices / of forward-solve vectors): warp `wid` of `nw` takes the
// targets t0+wid, t0+wid+nw, ... Blocks of >= 16 entries: every lane owns up to GE entries of the
// block at a time and walks the source list with all of them in flight (children ascending, so every
// entry is summed in list order). Smaller blocks (the 9 x 1 pieces of a forward-solve vector): the lanes are
// split into 32 / entries slots, slot s takes the sources s, s + slots, ..., eight rounds of loads in flight,
// and the slots are folded in ascending order. Deterministic, no atomics.
#define DLB_GE 4
__device__ __forceinline__ void gather_targets(const DlbGather& G, long long t0, long long t1, double* pool,
                                               int accumulate, long long wid, long long nw, int lane)
{
  for(long long t = t0 + wid; t < t1; t += nw)
  {
    const long long q0 = G.src_ptr[t], q1 = G.src_ptr[t+1];
    const int h = G.h[t];
    const bool tri = G.w[t] < 0;
    const int w = tri ? -G.w[t] : G.w[t];
    const int ldd = G.ld[t];
    double* dst = pool + G.dst[t];
    const int ne = h * w;
    if(ne >= 16)
      for(int e0 = 0; e0 < ne; e0 += 32 * DLB_GE)
      {
        int so[DLB_GE], sj[DLB_GE]; bool on[DLB_GE]; double acc[DLB_GE];
#pragma unroll
        for(int u = 0; u < DLB_GE; u++)
        {
          const int e = e0 + 32 * u + lane;
          const int j = e / h, i = e - j * h;
          on[u] = e < ne && !(tri && i < j);
          so[u] = i; sj[u] = j; acc[u] = 0.0;
        }
        // the descriptors of up to 32 sources are fetched by the lanes at once and handed round by
        // shuffles: the value loads of a batch are independent of any index load (one round of
        // memory latency per batch instead of two per source)
        for(long long qb = q0; qb < q1; qb += 32)
        {
          const int nq = (int)(q1 - qb < 32 ? q1 - qb : 32);
          const long long mybase = lane < nq ? G.gs_base[qb + lane] : 0;
          const int myld = lane < nq ? G.gs_ld[qb + lane] : 0;
          for(int q8 = 0; q8 < nq; q8 += 8)
          { // 8 sources x DLB_GE entries: all loads first (unconditional: slots beyond the list and
            // entries outside the block read a zero), then the additions in list order
            double vals[8][DLB_GE];
#pragma unroll
            for(int qq = 0; qq < 8; qq++)
            {
              const bool qon = q8 + qq < nq;
              const long long base = __shfl_sync(0xffffffffu, mybase, (q8 + qq) & 31);
              const int ld = __shfl_sync(0xffffffffu, myld, (q8 + qq) & 31);      // fronts have < 46341 rows: 32-bit offsets
              const double* src = pool + base;
#pragma unroll
              for(int u = 0; u < DLB_GE; u++)
              {
                const double* p = (qon && on[u]) ? src + (so[u] + sj[u] * ld) : g_dlb_zero;
                vals[qq][u] = *p;
              }
            }
#pragma unroll
            for(int qq = 0; qq < 8; qq++)
#pragma unroll
              for(int u = 0; u < DLB_GE; u++) acc[u] += vals[qq][u];
          }
        }
#pragma unroll
        for(int u = 0; u < DLB_GE; u++)
          if(on[u])
          {
            double* d = dst + (so[u] + sj[u] * ldd);
            *d = accumulate ? *d + acc[u] : acc[u];
          }
      }
    else
    {
      const int nslots = 32 / ne, slot = lane / ne, e = lane - slot * ne;
      const int j = e / h, i = e - j * h;
      const bool live = slot < nslots && !(tri && i < j);
      double acc = 0.0;
      for(long long qb = q0; qb < q1; qb += 8 * nslots)
      {
        long long base[8]; int ld[8]; bool ok[8];
#pragma unroll
        for(int rr = 0; rr < 8; rr++)
        {
          const long long q = qb + (long long)rr * nslots + slot;
          ok[rr] = live && q < q1;
          base[rr] = G.gs_base[ok[rr] ? q : q0]; ld[rr] = G.gs_ld[ok[rr] ? q : q0];
        }
        double v[8];
#pragma unroll
        for(int rr = 0; rr < 8; rr++)
        {
          const double* p = ok[rr] ? pool + base[rr] + i + (long long)j * ld[rr] : g_dlb_zero;
          v[rr] = *p;
        }
#pragma unroll
        for(int rr = 0; rr < 8; rr++) acc += v[rr];
      }
      for(int sidx = 1; sidx < nslots; sidx++)
      {
        const double other = __shfl_sync(0xffffffffu, acc, (e + sidx * ne) & 31);
        if(slot == 0) acc += other;
      }
      if(slot == 0 && live)
      {
        double* d = dst + i + (long long)j * ldd;
        *d = accumulate ? *d + acc : acc;
      }
    }
  }
}

// v'(Jt Jt')v over the class blocks: warp `wid` of `nw` takes the tasks listA[0..nA) ++ listB[0..nB);
// per task sum_{a>=b} (2 - [a==b]) v[row_a] G_ab v[row_b], the lanes over the packed pairs.
// Returns this lane's partial (to be summed over all lanes of all warps in a fixed order). v and the gather
// pool are deliberately not __restrict__: the persistent kernel reads what other CTAs wrote before a grid
// barrier, which must never go through the non-coherent (ld.global.nc) path.
__device__ __forceinline__ double quadform_partial(const DlbSparseDev& S, const int* __restrict__ listA, int nA,
                                                   const int* __restrict__ listB, int nB, const double* __restrict__ Gpart,
                                                   const double* v, int wid, int nw, int lane)
{
  double total = 0.0;
  for(int i = wid; i < nA + nB; i += nw)
  {
    const int t = i < nA ? listA[i] : listB[i - nA];
    const int c = S.task_cls[t];
    const int r0 = S.cls_ptr[c], k = S.cls_ptr[c+1] - r0;
    const double* G = Gpart + S.task_Goff[t];
    const int npairs = k * (k + 1) / 2;
    int a = 0, b = lane;
    while(b > a) { b -= a + 1; a++; }
    double s = 0.0;
    if(k <= 32)
    { // the class's entries of v are fetched once (lane = slot) and handed round by shuffles: the loop
      // below only has the independent loads of the block
      const double vl = lane < k ? v[S.cls_rows[r0 + lane]] : 0.0;
      for(int q0 = 0; q0 < npairs; q0 += 32 * 6)
      { // 6 x 32 pairs per batch (a 24-row class is 300 pairs): the block's loads are issued together
        double gq[6];
#pragma unroll
        for(int u = 0; u < 6; u++) { const double* p = q0 + 32 * u + lane < npairs ? G + (q0 + 32 * u + lane) : g_dlb_zero; gq[u] = *p; }
#pragma unroll
        for(int u = 0; u < 6; u++)
        {
          const double va = __shfl_sync(0xffffffffu, vl, a & 31), vb = __shfl_sync(0xffffffffu, vl, b & 31);
          if(q0 + 32 * u + lane < npairs)
          {
            const double term = va * gq[u] * vb;
            s += a == b ? term : 2.0 * term;
          }
          b += 32;
          while(b > a) { b -= a + 1; a++; }
        }
      }
    }
    else
      for(int q = lane; q < npairs; q += 32)
      {
        const double va = v[S.cls_rows[r0 + a]], vb = v[S.cls_rows[r0 + b]];
        const double term = va * G[q] * vb;
        s += a == b ? term : 2.0 * term;
        b += 32;
        while(b > a) { b -= a + 1; a++; }
      }
    total += s;
  }
  return total;
}
