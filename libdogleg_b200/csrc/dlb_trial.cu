// dlb_trial.cu -- one whole trial step of the dog-leg loop in ONE persistent kernel.
//
// For problems whose elimination tree is small (every front fits in shared memory: calibration
// problems such as the mrcal-shaped config, 200 frame fronts + one 68-wide root) an iteration is
// a few MFLOP; launched as ~30 kernels with 5 host round trips it was pure latency (round 1:
// factorization 150 us, solves 74 us, |Jv|^2 2 x 21 us, 100 us of host syncs for 6 MFLOP). Here the
// whole sequence between two evaluations of the user callback runs as one cooperative launch with
// grid-wide barriers between the phases:
//
//   Cauchy step            |J g|^2 = g'(JtJ)g on the class blocks, k, k*g     dogleg.c:529-617
//   need Gauss-Newton?     |cauchy|^2 < Delta^2                               dogleg.c:1192-1219
//   factorization + solve  level by level: block gather of the children's update matrices and
//                          forward-solve vectors, then one CTA per front: element assembly, extend-add,
//                          lambda, pivot elimination, forward substitution (roots: backward at once);
//                          backward sweep; gn = -P'z, |gn|^2                  dogleg.c:640-678, 839-866
//   step selection         Cauchy-clipped / GN / interpolated, p + step       dogleg.c:927-998, 1192-1259
//   expected improvement   Jt_x . step, |J step|^2, max|step|                  dogleg.c:1085-1165, 1289-1296
//
// The scalars the host automaton branches on (and, for host callbacks, the new p) are written
// straight into mapped pinned memory followed by a sequence number the host spins on: one host
// round trip per trial step. Every sum has a fixed order (per-CTA partials folded in CTA order by
// every CTA): results are bit-reproducible and independent of the grid size only through the
// partial sums, which is why the grid size is a pure function of the problem (never of the device load).
#ifdef DLB_TRIAL_DEBUG
#define DLB_ELIM_DBG 1
#endif
#include "dlb_common.cuh"
#include "dlb_device.h"
#include "dlb_devfn.cuh"
#include <climits>

// grid-wide barrier over a self-restoring counter (the scheme cooperative groups uses): CTA 0 adds
// 0x80000000 - (n-1), everybody else 1, so one full round flips the top bit and leaves the low bits
// as they were; no reset between barriers or launches. Needs all CTAs co-resident (cooperative launch).
__device__ __forceinline__ void grid_barrier(unsigned int* bar)
{
  __syncthreads();
  if(threadIdx.x == 0)
  {
    const unsigned int nb = blockIdx.x == 0 ? 0x80000000u - (gridDim.x - 1) : 1u;
    __threadfence();
    const unsigned int old = atomicAdd(bar, nb);
    while((((old ^ *(volatile unsigned int*)bar) & 0x80000000u) == 0)) { }
    __threadfence();
  }
  __syncthreads();
}

// sum (or max) over the CTAs of part[cta * DLB_TRIAL_PART + k], folded in CTA order by the calling CTA;
// the result is returned to every thread. sh: 33 doubles.
template<int NT>
__device__ __forceinline__ double fold_sum(const double* part, int k, double* sh)
{
  double a = 0.0;
  for(unsigned int b = threadIdx.x; b < gridDim.x; b += NT) a += ((const volatile double*)part)[(size_t)b * DLB_TRIAL_PART + k];
  a = block_sum(a, sh);
  if(threadIdx.x == 0) sh[32] = a;
  __syncthreads();
  a = sh[32];
  __syncthreads();
  return a;
}
template<int NT>
__device__ __forceinline__ double fold_max(const double* part, int k, double* sh)
{
  double a = 0.0;
  for(unsigned int b = threadIdx.x; b < gridDim.x; b += NT) a = fmax(a, ((const volatile double*)part)[(size_t)b * DLB_TRIAL_PART + k]);
  a = block_max(a, sh);
  if(threadIdx.x == 0) sh[32] = a;
  __syncthreads();
  a = sh[32];
  __syncthreads();
  return a;
}
template<int NT>
__device__ __forceinline__ void put_partial(double* part, int k, double v, double* sh, bool is_max = false)
{
  v = is_max ? block_max(v, sh) : block_sum(v, sh);
  if(threadIdx.x == 0) part[(size_t)blockIdx.x * DLB_TRIAL_PART + k] = v;
}

// DOGLEG_GPU_TRIAL_PROF=1: CTA 0 leaves a globaltimer stamp at every phase boundary (T.prof)
__device__ __forceinline__ void prof_mark(const DlbTrial& T, int& np)
{
  if(T.prof && blockIdx.x == 0 && threadIdx.x == 0 && np < DLB_TRIAL_PROF_MAX)
  {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    T.prof[np++] = t;
  }
}

// ---- one front: assembly, elimination with the right-hand side carried as an extra row (= forward
// substitution), and for a root the backward substitution at once ----
// A: (r+1) x r in shared memory, leading dimension ld = (r+1)|1 (odd: rows and columns conflict-free);
// row r is the right-hand side. dinv: nc reciprocal pivots behind it.
template<int NT>
__device__ __noinline__ void trial_front(const DlbFrontDev* Fp, const DlbTrial* Tp, int s,
                                         double* A, double* dinv, double* sh_red, int* npp)
{
  const DlbFrontDev& F = *Fp; const DlbTrial& T = *Tp; int& np = *npp;
  extern __shared__ double sm[];
  A = sm + (A - sm); dinv = sm + (dinv - sm);      // tell the compiler these are shared-memory addresses (LDS, not generic loads)
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int c0 = F.sn_first[s], nc = F.sn_first[s+1] - c0;
  const int rp = F.rows_ptr[s], r = F.rows_ptr[s+1] - rp;
  const int ld = (r + 1) | 1;
  double* Ag = T.fronts + F.front_off[s];
  for(int idx = tid; idx < ld * r; idx += NT) A[idx] = 0.0;
  __syncthreads();
  // elements: every entry of the front that receives class blocks sums its sources (classes ascending,
  // tasks ascending: the order of the per-level kernels) from the precomputed element lists; two entries
  // and eight sources per thread in flight (unconditional loads from clamped addresses, see g_dlb_zero)
  for(int d0 = T.eg_ptr[s] + tid; d0 < T.eg_ptr[s+1]; d0 += 2 * NT)
  {
    int q0[2], q1[2], idx[2][8]; double g[2][8], acc[2]; unsigned int dd[2];
#pragma unroll
    for(int u = 0; u < 2; u++)
    {
      const bool on = d0 + u * NT < T.eg_ptr[s+1];
      const int d = on ? d0 + u * NT : d0;
      const int a = T.eg_sptr[d], b = T.eg_sptr[d + 1];
      dd[u] = T.eg_dst[d];
      q0[u] = a; q1[u] = on ? b : a;
      acc[u] = 0.0;
    }
    const int nmax = max(q1[0] - q0[0], q1[1] - q0[1]);
    for(int qb = 0; qb < nmax; qb += 8)
    {
#pragma unroll
      for(int u = 0; u < 2; u++)
#pragma unroll
        for(int e = 0; e < 8; e++)
        {
          const int v = T.eg_src[q0[u] + qb + e < q1[u] ? q0[u] + qb + e : q0[u]];
          idx[u][e] = q0[u] + qb + e < q1[u] ? v : -1;
        }
#pragma unroll
      for(int u = 0; u < 2; u++)
#pragma unroll
        for(int e = 0; e < 8; e++) { const double* p = idx[u][e] >= 0 ? T.Gpart + idx[u][e] : g_dlb_zero; g[u][e] = *p; }
#pragma unroll
      for(int u = 0; u < 2; u++)
#pragma unroll
        for(int e = 0; e < 8; e++) acc[u] += g[u][e];
    }
#pragma unroll
    for(int u = 0; u < 2; u++) if(q1[u] > q0[u]) A[(dd[u] & 0xffffu) + (dd[u] >> 16) * ld] = acc[u];
  }
  __syncthreads();
  prof_mark(T, np);     // elements
  // children: gathered into the front's temporary, or pulled one after the other
  const long long toff = F.heavy_tmp_off[s];
  const bool gathered = F.sg_flag[s] != 0;
  double* yg = T.ywork + rp;
  if(toff >= 0)
  {
    const double* T0 = F.heavy_tmp + toff;
    for(int idx0 = tid; idx0 < r * r; idx0 += 8 * NT)
    {
      double t[8];
#pragma unroll
      for(int u = 0; u < 8; u++) t[u] = T0[idx0 + u * NT < r * r ? idx0 + u * NT : idx0];
#pragma unroll
      for(int u = 0; u < 8; u++)
      {
        const int idx = idx0 + u * NT;
        const int j = idx / r, i = idx - j * r;
        if(idx < r * r && i >= j) A[i + j * ld] += t[u];
      }
    }
  }
  else
    for(int ch = F.child_ptr[s]; ch < F.child_ptr[s+1]; ch++)
    {
      const int c   = F.child_list[ch];
      const int ncc = F.sn_first[c+1] - F.sn_first[c];
      const int rc  = F.rows_ptr[c+1] - F.rows_ptr[c];
      const int nb  = rc - ncc;
      const double* U = T.fronts + F.front_off[c];
      const int* rel = F.rel + F.rows_ptr[c] + ncc;
      for(int idx = tid; idx < nb * nb; idx += NT)
      {
        const int j = idx / nb, i = idx - j * nb;
        if(i >= j) A[rel[i] + rel[j] * ld] += U[(ncc + i) + (size_t)(ncc + j) * rc];
      }
      __syncthreads();
    }
  // the right-hand side row: P b on the pivot columns, plus the children's contributions
  for(int i = tid; i < r; i += NT) A[r + i * ld] = (i < nc ? T.Jtx[F.perm[c0 + i]] : 0.0) + (gathered ? yg[i] : 0.0);
  for(int j = tid; j < nc; j += NT) A[j + j * ld] += T.lambda;
  __syncthreads();
  if(!gathered)
    for(int ch = F.child_ptr[s]; ch < F.child_ptr[s+1]; ch++)
    {
      const int c   = F.child_list[ch];
      const int ncc = F.sn_first[c+1] - F.sn_first[c];
      const int rc  = F.rows_ptr[c+1] - F.rows_ptr[c];
      const double* yc = T.ywork + F.rows_ptr[c];
      const int* rel = F.rel + F.rows_ptr[c];
      for(int i = ncc + tid; i < rc; i += NT) A[r + rel[i] * ld] += yc[i];
      __syncthreads();
    }
  prof_mark(T, np);     // children + right-hand side + lambda
  const int fail = front_eliminate<NT>(A, r, nc, ld, r + 1, tid, dinv);
  if(fail >= 0)
  {
    if(tid == 0) atomicMin(T.minor, (long long)(c0 + fail));
    return;
  }
  prof_mark(T, np);     // elimination + forward substitution
  // the factor panel and the update matrix go to HBM: the parent's extend-add, the backward sweep and
  // later multi-RHS solves (outlier helpers) read them there; so does the solved / updated right-hand side
  for(int j = tid >> 5; j < r; j += NT / 32)
    for(int i = lane; i < r; i += 32) Ag[i + (size_t)j * r] = A[i + j * ld];
  for(int i = tid; i < r; i += NT) yg[i] = A[r + i * ld];
  if(F.sn_parent[s] >= 0)
  {
    for(int i = tid; i < nc; i += NT) T.zperm[c0 + i] = A[r + i * ld];
    prof_mark(T, np);   // store
    return;
  }
  prof_mark(T, np);     // store
  // ---- a root (r == nc): backward substitution right away, x = L^-T y, in place in row r ----
  for(int b0 = ((nc - 1) / 32) * 32; b0 >= 0; b0 -= 32)
  {
    const int bw = nc - b0 < 32 ? nc - b0 : 32;
    for(int cc = w; cc < bw; cc += NT / 32)
    {
      double acc = 0.0;
      for(int i = b0 + bw + lane; i < r; i += 32) acc = fma(A[i + (b0 + cc) * ld], A[r + i * ld], acc);
      acc = warp_sum_all(acc);
      if(lane == 0) sh_red[cc] = acc;
    }
    __syncthreads();
    if(w == 0)
    {
      double v = lane < bw ? A[r + (b0 + lane) * ld] - sh_red[lane] : 0.0;
      const double di = lane < bw ? dinv[b0 + lane] : 0.0;
      for(int j = bw - 1; j >= 0; j--)
      {
        const double xj = __shfl_sync(0xffffffffu, v * di, j);
        if(lane == j) v = xj;
        else if(lane < j) v = fma(-A[(b0 + j) + (b0 + lane) * ld], xj, v);
      }
      if(lane < bw) A[r + (b0 + lane) * ld] = v;
    }
    __syncthreads();
  }
  for(int i = tid; i < nc; i += NT) T.zperm[c0 + i] = A[r + i * ld];
}

// backward substitution of a non-root front: x(own columns) = L11^-T (y - L21' x(rows below)).
// P: the r x nc panel staged in shared memory (leading dimension r|1), xs: r entries, dinv: nc reciprocals.
template<int NT>
__device__ __noinline__ void trial_front_bwd(const DlbFrontDev* Fp, const DlbTrial* Tp, int s, double* P, double* sh_red)
{
  const DlbFrontDev& F = *Fp; const DlbTrial& T = *Tp;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int c0 = F.sn_first[s], nc = F.sn_first[s+1] - c0;
  const int rp = F.rows_ptr[s], r = F.rows_ptr[s+1] - rp;
  const int ld = r | 1;
  double* xs = P + (size_t)ld * nc;
  double* dinv = xs + r;
  const double* Ag = T.fronts + F.front_off[s];
  const int* rows = F.rows + rp;
  for(int idx = tid; idx < r * nc; idx += NT) { const int j = idx / r, i = idx - j * r; P[i + j * ld] = Ag[idx]; }
  for(int i = tid; i < r; i += NT) xs[i] = T.zperm[rows[i]];
  for(int j = tid; j < nc; j += NT) dinv[j] = 1.0 / Ag[j + (size_t)j * r];
  __syncthreads();
  for(int b0 = ((nc - 1) / 32) * 32; b0 >= 0; b0 -= 32)
  {
    const int bw = nc - b0 < 32 ? nc - b0 : 32;
    for(int cc = w; cc < bw; cc += NT / 32)
    {
      double acc = 0.0;
      for(int i = b0 + bw + lane; i < r; i += 32) acc = fma(P[i + (b0 + cc) * ld], xs[i], acc);
      acc = warp_sum_all(acc);
      if(lane == 0) sh_red[cc] = acc;
    }
    __syncthreads();
    if(w == 0)
    {
      double v = lane < bw ? xs[b0 + lane] - sh_red[lane] : 0.0;
      const double di = lane < bw ? dinv[b0 + lane] : 0.0;
      for(int j = bw - 1; j >= 0; j--)
      {
        const double xj = __shfl_sync(0xffffffffu, v * di, j);
        if(lane == j) v = xj;
        else if(lane < j) v = fma(-P[(b0 + j) + (b0 + lane) * ld], xj, v);
      }
      if(lane < bw) xs[b0 + lane] = v;
    }
    __syncthreads();
  }
  for(int i = tid; i < nc; i += NT) T.zperm[c0 + i] = xs[i];
  __syncthreads();
}

// The phases are separate (non-inlined) functions working on shared-memory copies of the kernel's
// parameter structs: inlined into one body, the register allocator (128 registers: two CTAs per SM)
// reused address registers as load destinations, which serialised loads that are independent --
// measured 11 000 cycles for a batch of 32 independent L2 loads in the block gather.
// consecutive targets go to different CTAs (warp w of CTA b takes target w * gridDim + b): a level with a
// few hundred targets then loads all SMs instead of filling the first CTAs warp by warp -- the gather
// is bound by the L2 bandwidth a single SM can pull (~27 B/clock measured, profiles/micro/gather_probe.cu)
// `skip`: targets another list has already dealt to the warps 0 .. skip-1 in this phase: this list starts behind them,
// so that the (few) forward-solve targets run beside the front targets instead of after them
__device__ __noinline__ void ph_gather(const DlbGather* G, long long t0, long long t1, double* pool, int accumulate, int lane,
                                       long long skip)
{
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  const long long wid = (long long)(threadIdx.x >> 5) * gridDim.x + blockIdx.x;
  gather_targets(*G, t0, t1, pool, accumulate, (wid + nw - skip % nw) % nw, nw, lane);
}
__device__ __noinline__ double ph_quadform(const DlbSparseDev* S, const double* Gpart, const double* v, int wid, int nw, int lane)
{
  return quadform_partial(*S, S->big_tasks, S->nbig, S->asm_small_tasks, S->nasm_small, Gpart, v, wid, nw, lane);
}
struct TrialShared { DlbSparseDev S; DlbFrontDev F; DlbTrial T; };

template<int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? 2 : 1)
k_trial(DlbSparseDev S, DlbFrontDev F, DlbTrial T)
{
  __shared__ TrialShared shp;
  if(threadIdx.x == 0) shp.S = S;
  if(threadIdx.x == 32) shp.F = F;
  if(threadIdx.x == 64 % NT) shp.T = T;
  __syncthreads();
  int np = 1;
  if(T.prof && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); T.prof[0] = t; }
#define PROF_MARK() prof_mark(T, np)
  extern __shared__ double sm[];
  __shared__ double sh_red[33];
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = blockIdx.x * (NT / 32) + (tid >> 5), nw = gridDim.x * (NT / 32);
  const int gtid = blockIdx.x * NT + tid, gnt = gridDim.x * NT;
  const int N = S.n;
  double* part = T.part;
  const double d2 = T.delta * T.delta;

  // the failure flag alternates between two slots from launch to launch: this launch uses T.minor and
  // resets the other one for the next (nobody touches it now)
  if(blockIdx.x == 0 && tid == 0) *T.minor_next = LLONG_MAX;

  // ---- phase 0: the Cauchy quadratic form and -- speculatively, before it is known whether the Cauchy
  // step stays inside the trust region (it almost always does) -- the leaf level of the factorization;
  // leaves have no children, so neither needs anything another CTA produces ----
  const bool factorize = !T.have_gn;
  if(!T.have_cauchy) put_partial<NT>(part, 0, ph_quadform(&shp.S, T.Gpart, T.Jtx, wid, nw, lane), sh_red);
  if(factorize)
    for(int q = T.level_ptr[0] + blockIdx.x; q < T.level_ptr[1]; q += gridDim.x)
    {
      trial_front<NT>(&shp.F, &shp.T, F.level_sn[q], sm, sm + (size_t)(T.max_rows + 2) * T.max_rows, sh_red, &np);
      __syncthreads();
    }
  if(!T.have_cauchy || factorize) grid_barrier(T.bar);
  PROF_MARK();                                    // phase 0

  // ---- Cauchy step ----
  double n2c, kc = 0.0;
  if(!T.have_cauchy)
  {
    const double jg2 = fold_sum<NT>(part, 0, sh_red);
    const double g2 = T.norm2_Jtx;
    kc = -g2 / jg2;
    n2c = kc * kc * g2;
    // written for later launches (a retry from the same point); this launch recomputes kc * Jt_x[i] where
    // it needs it, so no barrier separates this loop from its readers
    for(int i = gtid; i < N; i += gnt) T.cauchy[i] = kc * T.Jtx[i];
    if(blockIdx.x == 0 && tid == 0) { T.sc->norm2_JJtx = jg2; T.sc->k_cauchy = kc; T.sc->norm2_cauchy = n2c; }
  }
  else n2c = T.norm2_cauchy;

  // ---- Gauss-Newton step: the rest of the factorization + solves, if the Cauchy step stays inside the trust region ----
  const bool want_gn = !(n2c >= d2);
  double n2gn = T.have_gn ? T.norm2_gn : 0.0;
  long long minor_seen = LLONG_MAX;
  if(want_gn && factorize)
  {
    minor_seen = *(volatile long long*)T.minor;
    for(int l = 1; l < T.nlev && minor_seen == LLONG_MAX; l++)
    {
      const long long g0 = T.level_gt[2*l], g1 = T.level_gt[2*l+1], g2 = T.level_gt[2*l+2];
      const long long s0 = T.level_sg[2*l], s1 = T.level_sg[2*l+1], s2 = T.level_sg[2*l+2];
      if(g2 > g0 || s2 > s0)
      { // pass 1: chunks of the long source lists into scratch; the level's temporaries start from zero
        for(long long i = gtid; i < T.level_tmp[l]; i += gnt) F.heavy_tmp[i] = 0.0;
        for(int q = T.level_ptr[l] + blockIdx.x; q < T.level_ptr[l+1]; q += gridDim.x)
        { // ... and so do the right-hand-side rows of the fronts whose children's vectors are gathered
          const int sq = F.level_sn[q];
          if(F.sg_flag[sq]) for(int i = F.rows_ptr[sq] + tid; i < F.rows_ptr[sq+1]; i += NT) T.ywork[i] = 0.0;
        }
        ph_gather(&shp.F.fg, g0, g1, T.fronts, 0, lane, 0);
        ph_gather(&shp.F.sg, s0, s1, T.ywork, 0, lane, g1 - g0);
        grid_barrier(T.bar);
        PROF_MARK();                              // gather pass 1
        ph_gather(&shp.F.fg, g1, g2, T.fronts, 1, lane, 0);
        ph_gather(&shp.F.sg, s1, s2, T.ywork, 1, lane, g2 - g1);
        grid_barrier(T.bar);
        PROF_MARK();                              // gather pass 2
      }
      for(int q = T.level_ptr[l] + blockIdx.x; q < T.level_ptr[l+1]; q += gridDim.x)
      {
        trial_front<NT>(&shp.F, &shp.T, F.level_sn[q], sm, sm + (size_t)(T.max_rows + 2) * T.max_rows, sh_red, &np);
        __syncthreads();
      }
      grid_barrier(T.bar);
      PROF_MARK();                                // fronts of the level
      minor_seen = *(volatile long long*)T.minor;
    }
    if(minor_seen == LLONG_MAX)
      for(int l = T.nlev - 2; l >= 0; l--)
      {
        for(int q = T.level_ptr[l] + blockIdx.x; q < T.level_ptr[l+1]; q += gridDim.x)
        {
          const int s = F.level_sn[q];
          if(F.sn_parent[s] < 0) continue;           // a root below the top level: already solved
          trial_front_bwd<NT>(&shp.F, &shp.T, s, sm, sh_red);
        }
        grid_barrier(T.bar);
        PROF_MARK();                              // backward sweep of the level
      }
  }
  if(minor_seen != LLONG_MAX)
  { // not positive definite: the host loads the diagonal and launches again (dogleg.c:668-677)
    if(blockIdx.x == 0 && tid == 0)
    {
      T.sc->minor = minor_seen; T.sc->trial_flags = 0.0;
      DlbPublished* pub = T.pub;
      pub->sc = *T.sc;
      __threadfence_system();
      *(volatile unsigned long long*)&pub->seq = T.seq;
    }
    return;
  }
  const bool fresh_gn = want_gn && factorize;

  if(T.small_tail)
  {
    // ---- small problems: every CTA forms gn, the step and their dot products by itself (N doubles each in
    // shared memory, fixed-order block reductions: identical in all CTAs), so the only grid-wide step left
    // is the quadratic form of the expected improvement; CTA 0 writes the vectors ----
    double* sgn = sm; double* sstep = sm + N;
    if(want_gn)
    {
      if(fresh_gn)
      {
        double n2 = 0.0;
        for(int k = tid; k < N; k += NT) { const double v = -T.zperm[k]; sgn[F.perm[k]] = v; n2 = fma(v, v, n2); }
        n2 = block_sum(n2, sh_red);
        if(tid == 0) sh_red[32] = n2;
        __syncthreads();
        n2gn = sh_red[32];
        __syncthreads();
        if(blockIdx.x == 0) { for(int i = tid; i < N; i += NT) T.gn[i] = sgn[i]; if(tid == 0) T.sc->norm2_gn = n2gn; }
      }
      else { for(int i = tid; i < N; i += NT) sgn[i] = T.gn[i]; __syncthreads(); }
    }
    const int type = !want_gn ? DLB_TRIAL_CAUCHY : (n2gn <= d2 ? DLB_TRIAL_GN : DLB_TRIAL_INTERP);
    double kI = 0.0, disc_raw = 0.0;
    if(type == DLB_TRIAL_INTERP)
    {
      double l2 = 0.0, negc = 0.0;
      for(int i = tid; i < N; i += NT)
      {
        const double a = T.have_cauchy ? T.cauchy[i] : kc * T.Jtx[i], d = a - sgn[i];
        l2 = fma(d, d, l2);
        negc = fma(d, a, negc);
      }
      l2 = block_sum(l2, sh_red);
      if(tid == 0) sh_red[32] = l2;
      __syncthreads();
      l2 = sh_red[32];
      __syncthreads();
      negc = block_sum(negc, sh_red);
      if(tid == 0) sh_red[32] = negc;
      __syncthreads();
      negc = sh_red[32];
      __syncthreads();
      disc_raw = negc * negc - l2 * (n2c - d2);
      kI = (negc + sqrt(disc_raw < 0.0 ? 0.0 : disc_raw)) / l2;
    }
    const double scale = type == DLB_TRIAL_CAUCHY ? T.delta / sqrt(n2c) : 1.0;
    double n2 = 0.0, gd = 0.0, mx = 0.0;
    for(int i = tid; i < N; i += NT)
    {
      const double a = type == DLB_TRIAL_GN ? 0.0 : (T.have_cauchy ? T.cauchy[i] : kc * T.Jtx[i]);
      double sv;
      if(type == DLB_TRIAL_CAUCHY)  sv = scale * a;
      else if(type == DLB_TRIAL_GN) sv = sgn[i];
      else sv = a + kI * (sgn[i] - a);
      sstep[i] = sv;
      n2 = fma(sv, sv, n2);
      gd = fma(T.Jtx[i], sv, gd);
      mx = fmax(mx, fabs(sv));
      if(blockIdx.x == 0)
      {
        const double pn = T.p_from[i] + sv;
        T.step[i] = sv; T.p_to[i] = pn;
        if(T.h_p_to) T.h_p_to[i] = pn;
      }
    }
    __syncthreads();
    put_partial<NT>(part, 7, ph_quadform(&shp.S, T.Gpart, sstep, wid, nw, lane), sh_red);
    grid_barrier(T.bar);
    PROF_MARK();                                  // step + quadratic form
    if(blockIdx.x == 0)
    {
      n2 = block_sum(n2, sh_red); __syncthreads();
      gd = block_sum(gd, sh_red); __syncthreads();
      mx = block_max(mx, sh_red); __syncthreads();
      const double js2 = fold_sum<NT>(part, 7, sh_red);
      if(tid == 0)
      {
        DlbScalars* sc = T.sc;
        sc->norm2_step   = type == DLB_TRIAL_CAUCHY ? n2c : (type == DLB_TRIAL_GN ? n2gn : n2);
        sc->Jtx_dot_step = gd; sc->maxabs_step = mx; sc->norm2_Jstep = js2;
        if(type == DLB_TRIAL_INTERP) { sc->k_interp = kI; sc->discriminant = disc_raw; }
        sc->step_type = (double)type;
        sc->trial_flags = fresh_gn ? 1.0 : 0.0;
        sc->minor = -1;
        DlbPublished* pub = T.pub;
        pub->sc = *sc;
        __threadfence_system();
        *(volatile unsigned long long*)&pub->seq = T.seq;
        PROF_MARK();                              // publish
        if(T.prof) T.prof[DLB_TRIAL_PROF_MAX] = (unsigned long long)np;
      }
    }
    return;
  }

  // ---- large state vectors: the same steps grid-wide ----
  if(!T.have_cauchy) grid_barrier(T.bar);         // the Cauchy vector written above is read by other threads below
  if(fresh_gn)
  {
    double n2 = 0.0;
    for(int k = gtid; k < N; k += gnt)
    {
      const double v = -T.zperm[k];
      T.gn[F.perm[k]] = v;
      n2 = fma(v, v, n2);
    }
    put_partial<NT>(part, 1, n2, sh_red);
    grid_barrier(T.bar);
    n2gn = fold_sum<NT>(part, 1, sh_red);
    if(blockIdx.x == 0 && tid == 0) T.sc->norm2_gn = n2gn;
  }
  // step selection (dogleg.c:1192-1255)
  const int type = !want_gn ? DLB_TRIAL_CAUCHY : (n2gn <= d2 ? DLB_TRIAL_GN : DLB_TRIAL_INTERP);
  double kI = 0.0, disc_raw = 0.0;
  if(type == DLB_TRIAL_INTERP)
  {
    double l2 = 0.0, negc = 0.0;
    for(int i = gtid; i < N; i += gnt)
    {
      const double a = T.cauchy[i], d = a - T.gn[i];
      l2 = fma(d, d, l2);
      negc = fma(d, a, negc);
    }
    put_partial<NT>(part, 2, l2, sh_red);
    put_partial<NT>(part, 3, negc, sh_red);
    grid_barrier(T.bar);
    l2 = fold_sum<NT>(part, 2, sh_red);
    negc = fold_sum<NT>(part, 3, sh_red);
    disc_raw = negc * negc - l2 * (n2c - d2);
    kI = (negc + sqrt(disc_raw < 0.0 ? 0.0 : disc_raw)) / l2;
  }
  {
    const double scale = type == DLB_TRIAL_CAUCHY ? T.delta / sqrt(n2c) : 1.0;
    double n2 = 0.0, gd = 0.0, mx = 0.0;
    for(int i = gtid; i < N; i += gnt)
    {
      double sv;
      if(type == DLB_TRIAL_CAUCHY)  sv = scale * T.cauchy[i];
      else if(type == DLB_TRIAL_GN) sv = T.gn[i];
      else { const double a = T.cauchy[i]; sv = a + kI * (T.gn[i] - a); }
      const double pn = T.p_from[i] + sv;
      T.step[i] = sv;
      T.p_to[i] = pn;
      if(T.h_p_to) T.h_p_to[i] = pn;
      n2 = fma(sv, sv, n2);
      gd = fma(T.Jtx[i], sv, gd);
      mx = fmax(mx, fabs(sv));
    }
    put_partial<NT>(part, 4, n2, sh_red);
    put_partial<NT>(part, 5, gd, sh_red);
    put_partial<NT>(part, 6, mx, sh_red, true);
  }
  grid_barrier(T.bar);
  // expected improvement: |J step|^2 = step'(JtJ)step on the class blocks
  put_partial<NT>(part, 7, ph_quadform(&shp.S, T.Gpart, T.step, wid, nw, lane), sh_red);
  grid_barrier(T.bar);
  if(blockIdx.x == 0)
  {
    const double n2s = fold_sum<NT>(part, 4, sh_red), gd = fold_sum<NT>(part, 5, sh_red);
    const double mx = fold_max<NT>(part, 6, sh_red), js2 = fold_sum<NT>(part, 7, sh_red);
    if(tid == 0)
    {
      DlbScalars* sc = T.sc;
      // the reference stores the UNCLIPPED cauchy length and the cached GN length (dogleg.c:1200, 1228)
      sc->norm2_step   = type == DLB_TRIAL_CAUCHY ? n2c : (type == DLB_TRIAL_GN ? n2gn : n2s);
      sc->Jtx_dot_step = gd; sc->maxabs_step = mx; sc->norm2_Jstep = js2;
      if(type == DLB_TRIAL_INTERP) { sc->k_interp = kI; sc->discriminant = disc_raw; }
      sc->step_type = (double)type;
      sc->trial_flags = fresh_gn ? 1.0 : 0.0;
      sc->minor = -1;
      DlbPublished* pub = T.pub;
      pub->sc = *sc;
      __threadfence_system();
      *(volatile unsigned long long*)&pub->seq = T.seq;
      PROF_MARK();
      if(T.prof) T.prof[DLB_TRIAL_PROF_MAX] = (unsigned long long)np;
    }
  }
#undef PROF_MARK
}

// ------------------------------------------------------------------ launcher
template<int NT>
static int trial_grid_limit(size_t smem, int sm_count)
{
  static DlbPerDeviceOnce attr_once;
  if(attr_once.first()) cudaFuncSetAttribute(k_trial<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int per_sm = 0;
  if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trial<NT>, NT, smem) != cudaSuccess) return 0;
  return per_sm * sm_count;
}
size_t dlb_trial_smem_bytes(int max_rows) { return sizeof(double) * ((size_t)(max_rows + 2) * max_rows + 2 * (size_t)max_rows + 8); }
int dlb_trial_threads(int max_rows) { return max_rows > 112 ? 512 : 256; }
int dlb_trial_max_grid(int max_rows, int sm_count)
{
  const size_t smem = dlb_trial_smem_bytes(max_rows);
  return max_rows > 112 ? trial_grid_limit<512>(smem, sm_count) : trial_grid_limit<256>(smem, sm_count);
}
int dlb_launch_trial(const DlbSparseDev& S, const DlbFrontDev& F, const DlbTrial& T, int grid, cudaStream_t st)
{
  const size_t smem = dlb_trial_smem_bytes(T.max_rows);
  void* args[3] = { (void*)&S, (void*)&F, (void*)&T };
  cudaError_t rc;
  if(T.max_rows > 112) rc = cudaLaunchCooperativeKernel((const void*)k_trial<512>, dim3(grid), dim3(512), args, smem, st);
  else                rc = cudaLaunchCooperativeKernel((const void*)k_trial<256>, dim3(grid), dim3(256), args, smem, st);
  return rc == cudaSuccess ? 0 : -1;
}

// -DDLB_TRIAL_DEBUG: SM-cycle stamps inside front_eliminate of CTA 0 (diagonal block / substitution /
// trailing update per 8-column block), printed after every trial when DOGLEG_GPU_TRIAL_PROF=1
extern "C" void dlb_trial_dbg_dump()
{
#ifdef DLB_TRIAL_DEBUG
  unsigned long long h[256]; int n = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, g_elim_n, sizeof(int));
  cudaMemcpyFromSymbol(h, g_elim_dbg, sizeof(h));
  fprintf(stderr, "elim marks, SM cycles (%d):", n);
  for(int i = 1; i < n && i < 255; i++) fprintf(stderr, " %lld", (long long)(h[i] - h[i-1]));
  fprintf(stderr, "\n");
  n = 0; cudaMemcpyToSymbol(g_elim_n, &n, sizeof(int));
#endif
}
