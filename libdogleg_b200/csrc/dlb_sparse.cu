// dlb_sparse.cu -- streaming kernels over the CCS Jacobian Jt (values only).
//
// Layout in HBM: the Jacobian values stay exactly as the callback wrote them
// (CCS order, column j = measurement j). The fixed sparsity pattern is NOT read
// per nonzero: measurement columns with identical row lists form a "pattern
// class" (dlb_symbolic.h); a task is (class, contiguous range of its member
// columns) and is processed by one CTA: each warp takes a sub-range, with
// lane == slot inside the column, and the warps' partials are combined in shared
// memory in warp order, so
// - value loads are coalesced (a column is contiguous),
// - the per-nonzero index traffic of CCS (4 B/nnz) is replaced by 8 B/column,
// - all sums run in a fixed order: no atomics, bit-reproducible results.
//
// Replaces: mul_spmatrix_densevector + norm2 (reference dogleg.c:249-261,
// 190-196, called at :1025-1027), norm2_mul_spmatrix_t_densevector (:262-281,
// called at :566, :1109) and the on-the-fly Jt*Jt' formation inside
// cholmod_factorize (:656-665).
#include "dlb_common.cuh"
#include "dlb_device.h"
#include "dlb_devfn.cuh"

// ---------------------------------------------------------------- gradient
#define TASK_WARPS (DLB_NT / 32)

// sub-range of a task's member columns handled by warp w (multiple of 4 columns per warp)
__device__ __forceinline__ void warp_range(int m0, int m1, int w, int& a, int& b)
{
  int per = (m1 - m0 + TASK_WARPS - 1) / TASK_WARPS;
  per = (per + 3) & ~3;
  a = min(m1, m0 + w * per);
  b = min(m1, a + per);
}

// ------------------------------------------------------------- warp-level cp.async pipeline
// The gradient and |Jv|^2 kernels give every big task to ONE warp, which streams the task's
// member columns through shared memory in batches of PIPE_BC columns, PIPE_NST batches deep
// (LDGSTS, 8 bytes per lane and column): 32 columns (~5 KB) stay in flight per warp without
// holding registers, ~2400 warps are resident on the GPU -- enough outstanding bytes for the
// HBM latency-bandwidth product. (A register loop with 4 loads in flight and the columns of a
// task split over the 8 warps of a CTA topped out near 3 TB/s: every warp then saw only ~50
// columns per task, too few to amortise the dependent index -> value load chain.)
#define PIPE_BC 16
#define PIPE_NST 3
#define PIPE_LD 33                      // odd stride: conflict-free for lane = slot and lane = column
#define PIPE_WARP_DOUBLES (PIPE_NST * PIPE_BC * PIPE_LD + PIPE_NST * PIPE_BC + 32)
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template<int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// The column positions (and x indices) of a batch are fetched one batch ahead (pipe_fetch), so
// that issuing a batch never waits for an index load.
struct PipeIdx { unsigned int pos; int col; };
__device__ __forceinline__ PipeIdx pipe_fetch(const DlbSparseDev& S, bool want_col, int m, int cnt, int lane)
{
  PipeIdx ix = {0u, 0};
  if(lane < cnt) { ix.pos = S.mem_pos[m + lane]; if(want_col) ix.col = S.mem_col[m + lane]; }
  return ix;
}
// columns [m, m+cnt) of the task into stage 'st' of the warp's tile (+ their x into xs if x != NULL)
__device__ __forceinline__ void pipe_issue(const PipeIdx ix, const double* __restrict__ Jx, const double* __restrict__ x,
                                           double* tile, double* xs, int st, int cnt, int k, int lane)
{
  if(cnt > 0)
  {
    const unsigned int pos = ix.pos;
    if(x && lane < cnt) cp_async8(xs + st * PIPE_BC + lane, x + ix.col);
    double* dst = tile + st * PIPE_BC * PIPE_LD + lane;
#pragma unroll 4
    for(int c = 0; c < cnt; c++)
    {
      const unsigned int p = __shfl_sync(0xffffffffu, pos, c);
      if(lane < k) cp_async8(dst + c * PIPE_LD, Jx + p + lane);
    }
  }
  cp_async_commit();
}

// gpart[task_goff[t] + a] = sum over the task's member columns of J(a,col)*x[col]
// n2part[cta]             = sum over the CTA's tasks of x[col]^2
// Big task bt is handled by warp (bt / gridDim.x) of CTA (bt % gridDim.x): the tasks spread
// evenly over the CTAs whatever their number.
__global__ void __launch_bounds__(DLB_NT)
k_sparse_grad(DlbSparseDev S, const double* __restrict__ Jx, const double* __restrict__ x,
              double* __restrict__ gpart, double* __restrict__ n2part)
{
  __shared__ double sh[32];
  extern __shared__ double sh_pipe[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double* tile = sh_pipe + (size_t)w * PIPE_WARP_DOUBLES;
  double* xs = tile + PIPE_NST * PIPE_BC * PIPE_LD;
  double n2 = 0.0;
  for(int bt = w * gridDim.x + blockIdx.x; bt < S.ngj_big; bt += TASK_WARPS * gridDim.x)
  {
    const int t = S.gj_big_tasks[bt];
    const int c = S.task_cls[t];
    const int k = S.cls_ptr[c+1] - S.cls_ptr[c];
    const long long goff = S.task_goff[t];
    const int m0 = S.task_m0[t], m1 = S.task_m1[t];
    if(k <= 32)
    {
      const int nb = (m1 - m0 + PIPE_BC - 1) / PIPE_BC;
      auto count = [&](int b) { return b < nb ? min(PIPE_BC, m1 - m0 - b * PIPE_BC) : 0; };
      PipeIdx ix = pipe_fetch(S, true, m0, count(0), lane);
      for(int b = 0; b < PIPE_NST - 1; b++)
      {
        const PipeIdx nx = pipe_fetch(S, true, m0 + (b + 1) * PIPE_BC, count(b + 1), lane);
        pipe_issue(ix, Jx, x, tile, xs, b, count(b), k, lane);
        ix = nx;
      }
      double acc = 0.0;
      for(int b = 0; b < nb; b++)
      {
        const int bn = b + PIPE_NST - 1;
        const PipeIdx nx = pipe_fetch(S, true, m0 + (bn + 1) * PIPE_BC, count(bn + 1), lane);
        pipe_issue(ix, Jx, x, tile, xs, bn % PIPE_NST, count(bn), k, lane);
        ix = nx;
        cp_async_wait<PIPE_NST - 1>();
        __syncwarp();
        const int st = b % PIPE_NST, cnt = min(PIPE_BC, m1 - m0 - b * PIPE_BC);
        const double* tl = tile + st * PIPE_BC * PIPE_LD + lane;
        const double* xb = xs + st * PIPE_BC;
        if(lane < k) for(int cc = 0; cc < cnt; cc++) acc = fma(tl[cc * PIPE_LD], xb[cc], acc);
        if(lane < cnt) n2 = fma(xb[lane], xb[lane], n2);
        __syncwarp();
      }
      cp_async_wait<0>();
      if(lane < k) gpart[goff + lane] = acc;
    }
    else
    { // long columns: the lanes stride over the slots, one column at a time
      for(int a0 = 0; a0 < k; a0 += 32)
      {
        const int a = a0 + lane;
        double acc = 0.0;
        for(int m = m0; m < m1; m++)
        {
          const double xv = x[S.mem_col[m]];
          if(a < k) acc = fma(ldg_stream(Jx + S.mem_pos[m] + a), xv, acc);
        }
        if(a < k) gpart[goff + a] = acc;
      }
      for(int m = m0 + lane; m < m1; m += 32) { const double xv = x[S.mem_col[m]]; n2 = fma(xv, xv, n2); }
    }
  }
  n2 = block_sum(n2, sh);
  if(threadIdx.x == 0) n2part[blockIdx.x] = n2;
}

// Jt_x[i] = sum of its partial entries in a fixed order. States that occur in few (class, slot)
// pairs are summed by one warp each; states that occur in many (the "border" states every
// measurement touches) get a whole CTA (heavy list). Then |x|^2, |Jt x|^2, max|Jt x|.
__device__ __forceinline__ double grad_entry_sum(const DlbSparseDev& S, const double* __restrict__ gpart, int q)
{
  const double* src = gpart + S.ginv_off[q];       // the entry's slot in the class's first task
  const int c = S.ginv_cls[q];                     // -1: the class has a single task (nothing else to look up)
  if(c < 0) return *src;
  const int k = S.cls_ptr[c+1] - S.cls_ptr[c];
  const int nt = S.gp_count[c];
  double s0 = 0.0;
  for(int t = 0; t < nt; t++) s0 += src[(size_t)t * k];
  return s0;
}
__global__ void __launch_bounds__(DLB_NT)
k_sparse_grad_reduce(DlbSparseDev S, const double* __restrict__ gpart, const double* __restrict__ n2part,
                     int n2count, double* __restrict__ Jtx, double* part, unsigned int* counter, DlbScalars* sc,
                     DlbPublished* pub, unsigned long long seq, int n2_behind_Jtx)
{
  __shared__ double shb[32];
  const int lane = threadIdx.x & 31;
  double g2 = 0.0, gmax = 0.0, n2 = 0.0;
  // heavy states: one CTA each
  for(int h = blockIdx.x; h < S.nheavy; h += gridDim.x)
  {
    const int i = S.heavy_state[h];
    double s = 0.0;
    for(int q = S.ginv_ptr[i] + threadIdx.x; q < S.ginv_ptr[i+1]; q += DLB_NT) s += grad_entry_sum(S, gpart, q);
    s = block_sum(s, shb);
    if(threadIdx.x == 0) { Jtx[i] = s; g2 = fma(s, s, g2); gmax = fmax(gmax, fabs(s)); }
    __syncthreads();
  }
  // medium states (DLB_LIGHT_MAX .. heavy_threshold-1 entries): one warp each
  for(int m = blockIdx.x * TASK_WARPS + (threadIdx.x >> 5); m < S.nmedium; m += gridDim.x * TASK_WARPS)
  {
    const int i = S.medium_state[m];
    double s = 0.0;
    for(int q = S.ginv_ptr[i] + lane; q < S.ginv_ptr[i+1]; q += 32) s += grad_entry_sum(S, gpart, q);
    s = warp_sum(s);
    if(lane == 0) { Jtx[i] = s; g2 = fma(s, s, g2); gmax = fmax(gmax, fabs(s)); }
  }
  // light states (the coordinates of a bundle-adjustment point occur in 4 classes): one thread each
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < S.n; i += gridDim.x * blockDim.x)
  {
    const int q0 = S.ginv_ptr[i], q1 = S.ginv_ptr[i+1];
    if(q1 - q0 >= DLB_LIGHT_MAX) continue;
    double s = 0.0;
    for(int q = q0; q < q1; q++) s += grad_entry_sum(S, gpart, q);
    Jtx[i] = s; g2 = fma(s, s, g2); gmax = fmax(gmax, fabs(s));
  }
  for(int t = blockIdx.x * blockDim.x + threadIdx.x; t < n2count; t += gridDim.x * blockDim.x) n2 += n2part[t];
  double out[5];
  if(grid_reduce5(n2, g2, 0.0, 0.0, gmax, part, counter, out))
  {
    sc->norm2_x = out[0]; sc->norm2_Jtx = out[1]; sc->maxabs_Jtx = out[4];
    if(n2_behind_Jtx) Jtx[S.n] = out[0];            // row-sharded: |x|^2 travels with the gradient in one all-reduce
    if(pub)
    { // the host spins on the sequence number in mapped pinned memory: no D2H copy, no stream sync
      pub->sc = *sc;
      __threadfence_system();
      *(volatile unsigned long long*)&pub->seq = seq;
    }
  }
}

// ------------------------------------------------------------------ |J v|^2
// sum over the member columns of (sum_a J(a,col) v[row_a])^2, one warp per big task through the
// cp.async pipeline; lane = column sums its k products from the tile (no shuffle tree per column)
__global__ void __launch_bounds__(DLB_NT)
k_sparse_jv(DlbSparseDev S, const double* __restrict__ Jx, const double* __restrict__ v,
            double* part, unsigned int* counter, const double* add0, const double* add1, double* dst)
{
  extern __shared__ double sh_pipe[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double* tile = sh_pipe + (size_t)w * PIPE_WARP_DOUBLES;
  double* vs = tile + PIPE_NST * PIPE_BC * PIPE_LD + PIPE_NST * PIPE_BC;
  double total = 0.0;           // per lane, over all the warp's tasks
  for(int bt = w * gridDim.x + blockIdx.x; bt < S.ngj_big; bt += TASK_WARPS * gridDim.x)
  {
    const int t = S.gj_big_tasks[bt];
    const int c  = S.task_cls[t];
    const int m0 = S.task_m0[t], m1 = S.task_m1[t];
    const int r0 = S.cls_ptr[c], k = S.cls_ptr[c+1] - r0;
    if(k <= 32)
    {
      __syncwarp();
      vs[lane] = lane < k ? v[S.cls_rows[r0 + lane]] : 0.0;
      const int nb = (m1 - m0 + PIPE_BC - 1) / PIPE_BC;
      auto count = [&](int b) { return b < nb ? min(PIPE_BC, m1 - m0 - b * PIPE_BC) : 0; };
      PipeIdx ix = pipe_fetch(S, false, m0, count(0), lane);
      for(int b = 0; b < PIPE_NST - 1; b++)
      {
        const PipeIdx nx = pipe_fetch(S, false, m0 + (b + 1) * PIPE_BC, count(b + 1), lane);
        pipe_issue(ix, Jx, NULL, tile, NULL, b, count(b), k, lane);
        ix = nx;
      }
      for(int b = 0; b < nb; b++)
      {
        const int bn = b + PIPE_NST - 1;
        const PipeIdx nx = pipe_fetch(S, false, m0 + (bn + 1) * PIPE_BC, count(bn + 1), lane);
        pipe_issue(ix, Jx, NULL, tile, NULL, bn % PIPE_NST, count(bn), k, lane);
        ix = nx;
        cp_async_wait<PIPE_NST - 1>();
        __syncwarp();
        const int st = b % PIPE_NST, cnt = min(PIPE_BC, m1 - m0 - b * PIPE_BC);
        if(lane < cnt)
        {
          const double* tl = tile + st * PIPE_BC * PIPE_LD + lane * PIPE_LD;
          double d = 0.0;
          for(int a = 0; a < k; a++) d = fma(tl[a], vs[a], d);
          total = fma(d, d, total);
        }
        __syncwarp();
      }
      cp_async_wait<0>();
    }
    else
      for(int m = m0; m < m1; m++)
      {
        const unsigned int p = S.mem_pos[m];
        double d = 0.0;
        for(int a = lane; a < k; a += 32) d = fma(ldg_stream(Jx + p + a), v[S.cls_rows[r0 + a]], d);
        d = warp_sum_all(d);
        if(lane == 0) total = fma(d, d, total);
      }
  }
  double out[5];
  if(grid_reduce5(total, 0.0, 0.0, 0.0, 0.0, part, counter, out))
    *dst = out[0] + (add0 ? *add0 : 0.0) + (add1 ? *add1 : 0.0);
}

// ------------------------------------------------- range tasks: contiguous column ranges
// In a calibration problem the measurement columns come in long runs whose pattern classes
// repeat with a short period (x-row, y-row, x-row, ...). Reading one class at a time touches
// every other 144..192-byte column: measured on this GPU that costs ~30% of the HBM bandwidth
// (profiles/micro/stride_read.cu). A range task is a contiguous piece of such a run, handled by
// one warp that reads it as ONE contiguous stream: no per-column index (position and x index
// follow from the period), every warp load is a full coalesced line, and the lanes own fixed
// entries of the period (class slot pairs), so the gradient needs no cross-lane reduction.
// The stream is moved by the TMA unit: one lane issues a 1-D bulk copy (cp.async.bulk, SASS
// UBLKCP) of up to RANGE_CHUNK bytes per stage into the warp's shared-memory ring and arms an
// mbarrier with the byte count; RANGE_NST-1 chunks (~8 KB) per warp are in flight with no
// registers held and one instruction per chunk. Bulk copies need 16-byte aligned addresses and
// sizes: a chunk is the aligned superset of its periods (the Jacobian buffers are padded).
#define RANGE_CHUNK 2048
#define RANGE_NST 4
#define RANGE_STAGE_DOUBLES (RANGE_CHUNK / 8 + 4)
#define RANGE_WARP_DOUBLES (RANGE_NST * RANGE_STAGE_DOUBLES + 128 + 4)           // ring + v entries + mbarriers; even: 16-byte aligned rings
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, int count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity)
{
  asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
               :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// class index inside the period of entry e (koff ascending, P <= 4), without indexing a local array
__device__ __forceinline__ int range_ci(int e, int P, int k1, int k2, int k3)
{
  int c = 0;
  if(P > 1 && e >= k1) c = 1;
  if(P > 2 && e >= k2) c = 2;
  if(P > 3 && e >= k3) c = 3;
  return c;
}
__device__ __forceinline__ int pick4(int c, int v0, int v1, int v2, int v3) { return c == 0 ? v0 : (c == 1 ? v1 : (c == 2 ? v2 : v3)); }
__device__ __forceinline__ long long pick4(int c, long long v0, long long v1, long long v2, long long v3) { return c == 0 ? v0 : (c == 1 ? v1 : (c == 2 ? v2 : v3)); }

// One warp's ring: chunk c of a task = periods [c*cp, min(nper,(c+1)*cp)); returns in 'head' the
// offset (doubles) of the first period inside the stage buffer
template<int STAGE>
struct RangeRing
{
  double* ring; unsigned long long* bars; unsigned phase_bits; int lane;
  __device__ __forceinline__ void issue(const double* Jx, unsigned long long pos, int nperiods, int K, int st)
  {
    if(lane == 0 && nperiods > 0)
    {
      const unsigned long long b0 = pos * 8ull, b1 = b0 + (unsigned long long)nperiods * K * 8ull;
      const unsigned long long a0 = b0 & ~15ull, a1 = (b1 + 15ull) & ~15ull;
      mbar_expect_tx(bars + st, (unsigned)(a1 - a0));
      bulk_g2s(ring + st * STAGE, (const char*)Jx + a0, (unsigned)(a1 - a0), bars + st);
    }
  }
  __device__ __forceinline__ const double* wait(unsigned long long pos, int st)
  {
    mbar_wait(bars + st, (phase_bits >> st) & 1u);
    phase_bits ^= 1u << st;
    return ring + st * STAGE + (pos & 1ull);
  }
};

// NU = ceil(longest period / 32): entries per lane. CHUNK / NST: bytes per stage and stages of the
// warp's ring (a <4096, 3> geometry was measured in round 2, profiles/r02_variants.txt: no gain).
template<int NU, int CHUNK, int NST>
__global__ void __launch_bounds__(DLB_NT)
k_range_grad(DlbSparseDev S, const double* __restrict__ Jx, const double* __restrict__ x,
             double* __restrict__ gpart, double* __restrict__ n2part)
{
  constexpr int STAGE = CHUNK / 8 + 4, WARP_DOUBLES = NST * STAGE + 128 + 4;
  __shared__ double sh[32];
  extern __shared__ __align__(16) double sh_range[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  RangeRing<STAGE> R;
  R.ring = sh_range + (size_t)w * WARP_DOUBLES;
  R.bars = (unsigned long long*)(R.ring + NST * STAGE + 128);
  R.phase_bits = 0; R.lane = lane;
  if(lane == 0) for(int st = 0; st < NST; st++) mbar_init(R.bars + st, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int wg = blockIdx.x * TASK_WARPS + w, nw = gridDim.x * TASK_WARPS;
  double n2 = 0.0;
  for(int i = wg; i < S.nrange; i += nw)
  {
    const DlbRangeTask* rp = S.rtasks + i;
    const int K = rp->Ktot, P = rp->P, nper = rp->ncols / P, j0 = rp->j0;
    const int k1 = rp->koff[1], k2 = rp->koff[2], k3 = rp->koff[3];
    const unsigned long long pos0 = rp->pos0;
    int ci[NU]; bool on[NU]; double acc[NU];
#pragma unroll
    for(int u = 0; u < NU; u++)
    {
      const int e = lane + 32 * u;
      on[u] = e < K; ci[u] = range_ci(e, P, k1, k2, k3); acc[u] = 0.0;
    }
    const double* xb = x + j0;
    // a chunk holds at most 32 measurement columns: their x fit one register per lane
    const int cp = max(1, min((CHUNK - 16) / (8 * K), 32 / P)), nch = (nper + cp - 1) / cp;
    auto cnt = [&](int c) { return c < nch ? min(cp, nper - c * cp) : 0; };
    for(int c = 0; c < NST - 1; c++) R.issue(Jx, pos0 + (unsigned long long)c * cp * K, cnt(c), K, c);
    for(int c = 0; c < nch; c++)
    {
      const int cn = c + NST - 1;
      R.issue(Jx, pos0 + (unsigned long long)cn * cp * K, cnt(cn), K, cn % NST);
      const int nq = cnt(c), qb = c * cp;
      // the x of the chunk's columns: one coalesced load (in flight while the chunk is waited
      // for), handed to the consuming lanes by shuffles instead of one load per FMA
      const double xr = lane < nq * P ? xb[qb * P + lane] : 0.0;
      if(lane < nq * P) n2 = fma(xr, xr, n2);
      const double* tl = R.wait(pos0 + (unsigned long long)c * cp * K, c % NST) + lane;
      int q = 0;
      for(; q + 4 <= nq; q += 4)
      { // 4 periods at a time: all shared-memory loads are issued before the FMAs
        double tv[4][NU], xv[4][NU];
#pragma unroll
        for(int qq = 0; qq < 4; qq++)
#pragma unroll
          for(int u = 0; u < NU; u++)
          {
            tv[qq][u] = on[u] ? tl[(q + qq) * K + 32 * u] : 0.0;
            xv[qq][u] = __shfl_sync(0xffffffffu, xr, (q + qq) * P + ci[u]);
          }
#pragma unroll
        for(int qq = 0; qq < 4; qq++)
#pragma unroll
          for(int u = 0; u < NU; u++) acc[u] = fma(tv[qq][u], xv[qq][u], acc[u]);
      }
      for(; q < nq; q++)
#pragma unroll
        for(int u = 0; u < NU; u++)
        {
          const double xq = __shfl_sync(0xffffffffu, xr, q * P + ci[u]);
          if(on[u]) acc[u] = fma(tl[q * K + 32 * u], xq, acc[u]);
        }
      __syncwarp();
    }
#pragma unroll
    for(int u = 0; u < NU; u++)
      if(on[u])
      {
        const int e = lane + 32 * u;
        const long long go = pick4(ci[u], rp->goff[0], rp->goff[1], rp->goff[2], rp->goff[3]);
        gpart[go + e - pick4(ci[u], 0, k1, k2, k3)] = acc[u];
      }
  }
  n2 = block_sum(n2, sh);
  if(threadIdx.x == 0) n2part[blockIdx.x] = n2;
}

template<int NU>
__global__ void __launch_bounds__(DLB_NT)
k_range_jv(DlbSparseDev S, const double* __restrict__ Jx, const double* __restrict__ v,
           double* part, unsigned int* counter, const double* add0, const double* add1, double* dst)
{
  extern __shared__ __align__(16) double sh_range[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  RangeRing<RANGE_STAGE_DOUBLES> R;
  R.ring = sh_range + (size_t)w * RANGE_WARP_DOUBLES;
  double* vs = R.ring + RANGE_NST * RANGE_STAGE_DOUBLES;
  R.bars = (unsigned long long*)(vs + 128);
  R.phase_bits = 0; R.lane = lane;
  if(lane == 0) for(int st = 0; st < RANGE_NST; st++) mbar_init(R.bars + st, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int wg = blockIdx.x * TASK_WARPS + w, nw = gridDim.x * TASK_WARPS;
  double total = 0.0;
  for(int i = wg; i < S.nrange; i += nw)
  {
    const DlbRangeTask* rp = S.rtasks + i;
    const int K = rp->Ktot, P = rp->P, nper = rp->ncols / P;
    const int k1 = rp->koff[1], k2 = rp->koff[2], k3 = rp->koff[3];
    const unsigned long long pos0 = rp->pos0;
    const int cp = max(1, (RANGE_CHUNK - 16) / (8 * K)), nch = (nper + cp - 1) / cp;
    auto cnt = [&](int c) { return c < nch ? min(cp, nper - c * cp) : 0; };
    for(int c = 0; c < RANGE_NST - 1; c++) R.issue(Jx, pos0 + (unsigned long long)c * cp * K, cnt(c), K, c);
    __syncwarp();
#pragma unroll
    for(int u = 0; u < NU; u++)
    {
      const int e = lane + 32 * u;
      if(e < K)
      {
        const int c = range_ci(e, P, k1, k2, k3);
        const int cl = pick4(c, rp->cls[0], rp->cls[1], rp->cls[2], rp->cls[3]);
        vs[e] = v[S.cls_rows[S.cls_ptr[cl] + e - pick4(c, 0, k1, k2, k3)]];
      }
    }
    __syncwarp();
    for(int c = 0; c < nch; c++)
    {
      const int cn = c + RANGE_NST - 1;
      R.issue(Jx, pos0 + (unsigned long long)cn * cp * K, cnt(cn), K, cn % RANGE_NST);
      const double* tl = R.wait(pos0 + (unsigned long long)c * cp * K, c % RANGE_NST);
      const int nq = cnt(c);
      // lane <-> (period, class) pair = one measurement column: its dot product with v
      for(int pr = lane; pr < nq * P; pr += 32)
      {
        const int qq = pr / P, cc = pr - qq * P;
        const int a0 = pick4(cc, 0, k1, k2, k3), a1 = cc + 1 < P ? pick4(cc + 1, 0, k1, k2, k3) : K;
        const double* col = tl + qq * K;
        // four interleaved partial sums (fixed order): the k-long dependent chain becomes k/4
        double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
        int a = a0;
        for(; a + 4 <= a1; a += 4)
        {
          d0 = fma(col[a], vs[a], d0); d1 = fma(col[a+1], vs[a+1], d1);
          d2 = fma(col[a+2], vs[a+2], d2); d3 = fma(col[a+3], vs[a+3], d3);
        }
        for(; a < a1; a++) d0 = fma(col[a], vs[a], d0);
        const double d = (d0 + d1) + (d2 + d3);
        total = fma(d, d, total);
      }
      __syncwarp();
    }
  }
  double out[5];
  if(grid_reduce5(total, 0.0, 0.0, 0.0, 0.0, part, counter, out))
    *dst = out[0] + (add0 ? *add0 : 0.0) + (add1 ? *add1 : 0.0);
}

// ------------------------------------------- small tasks: a group of lanes each
// Problems with very many pattern classes of a few columns each (bundle adjustment: one class
// per camera-point pair, two measurement columns) would leave a CTA per task idle; their tasks
// are "small" (few member columns, k <= 32) and are handled by a group of G = 8, 16 or 32
// lanes each (G >= the longest small column), two tasks in flight per group. Everything a task
// needs is in one 32-byte record, so the dependent-load chain is record -> member -> values.
template<int G>
__global__ void __launch_bounds__(DLB_NT)
k_sparse_grad_small(DlbSparseDev S, const DlbSmallTask* __restrict__ infos, int ninfo,
                    const double* __restrict__ Jx, const double* __restrict__ x,
                    double* __restrict__ gpart, double* __restrict__ n2part)
{
  __shared__ double sh[32];
  const int a = threadIdx.x & (G - 1);
  const int grp = (blockIdx.x * DLB_NT + threadIdx.x) / G, ngrp = gridDim.x * (DLB_NT / G);
  double n2 = 0.0;
  for(int st = grp; st < ninfo; st += 2 * ngrp)
  {
    const int st2 = st + ngrp;
    const DlbSmallTask t0 = infos[st];
    const bool has2 = st2 < ninfo;
    const DlbSmallTask t1 = infos[has2 ? st2 : st];
    double acc0 = 0.0, acc1 = 0.0;
    const int nm = max(t0.nm, has2 ? t1.nm : 0);
    for(int m = 0; m < nm; m++)
    {
      const bool on0 = m < t0.nm, on1 = has2 && m < t1.nm;
      const double x0 = on0 ? x[S.mem_col[t0.m0 + m]] : 0.0, x1 = on1 ? x[S.mem_col[t1.m0 + m]] : 0.0;
      const double v0 = (on0 && a < t0.k) ? ldg_stream(Jx + S.mem_pos[t0.m0 + m] + a) : 0.0;
      const double v1 = (on1 && a < t1.k) ? ldg_stream(Jx + S.mem_pos[t1.m0 + m] + a) : 0.0;
      acc0 = fma(v0, x0, acc0); acc1 = fma(v1, x1, acc1);
      if(a == 0) { n2 = fma(x0, x0, n2); n2 = fma(x1, x1, n2); }
    }
    if(a < t0.k) gpart[t0.goff + a] = acc0;
    if(has2 && a < t1.k) gpart[t1.goff + a] = acc1;
  }
  n2 = block_sum(n2, sh);
  if(threadIdx.x == 0) n2part[blockIdx.x] = n2;
}

// |J v|^2 over the small tasks (+ *add_or_null, the total of the big tasks) -> *dst
template<int G>
__global__ void __launch_bounds__(DLB_NT)
k_sparse_jv_small(DlbSparseDev S, const DlbSmallTask* __restrict__ infos, int ninfo,
                  const double* __restrict__ Jx, const double* __restrict__ v,
                  double* part, unsigned int* counter, const double* add0, const double* add1, double* dst)
{
  const int a = threadIdx.x & (G - 1);
  const int grp = (blockIdx.x * DLB_NT + threadIdx.x) / G, ngrp = gridDim.x * (DLB_NT / G);
  double total = 0.0;          // accumulated in lane 0 of every group
  // every lane of a warp runs the same number of rounds (full-warp shuffles)
  const int rounds = (ninfo + 2 * ngrp - 1) / (2 * ngrp);
  for(int it = 0; it < rounds; it++)
  {
    const int st = grp + it * 2 * ngrp, st2 = st + ngrp;
    const bool has1 = st < ninfo, has2 = st2 < ninfo;
    const DlbSmallTask t0 = infos[has1 ? st : 0];
    const DlbSmallTask t1 = infos[has2 ? st2 : 0];
    const double va0 = (has1 && a < t0.k) ? v[S.cls_rows[t0.r0 + a]] : 0.0;
    const double va1 = (has2 && a < t1.k) ? v[S.cls_rows[t1.r0 + a]] : 0.0;
    int nm = max(has1 ? t0.nm : 0, has2 ? t1.nm : 0);
    // the longest member list among the warp's groups
#pragma unroll
    for(int o = 16; o >= G; o >>= 1) nm = max(nm, __shfl_xor_sync(0xffffffffu, nm, o));
    for(int m = 0; m < nm; m++)
    {
      const bool on0 = has1 && m < t0.nm && a < t0.k, on1 = has2 && m < t1.nm && a < t1.k;
      double d0 = on0 ? ldg_stream(Jx + S.mem_pos[t0.m0 + m] + a) * va0 : 0.0;
      double d1 = on1 ? ldg_stream(Jx + S.mem_pos[t1.m0 + m] + a) * va1 : 0.0;
#pragma unroll
      for(int o = G / 2; o > 0; o >>= 1) { d0 += __shfl_xor_sync(0xffffffffu, d0, o); d1 += __shfl_xor_sync(0xffffffffu, d1, o); }
      if(a == 0) { total = fma(d0, d0, total); total = fma(d1, d1, total); }
    }
  }
  double out[5];
  if(grid_reduce5(total, 0.0, 0.0, 0.0, 0.0, part, counter, out))
    *dst = out[0] + (add0 ? *add0 : 0.0) + (add1 ? *add1 : 0.0);
}

// ---------------------------------------------------------------- assembly
// Gpart[task_Goff[t] + q], q = a(a+1)/2 + b (a>=b): sum over the task's member
// columns of J(a,col) J(b,col) -- the class-local lower triangle of Jt Jt'.
//
// This is a small SYRK per class: G = V V' with V = (k slots) x (member columns).
// For k <= 32 it runs on the FP64 tensor cores: mma.sync m8n8k4 (SASS DMMA), the
// K dimension being 4 member columns per instruction. The A fragment
// (row slot 8*ti+g, column member m+t for lane 4g+t) and the B fragment of the
// transposed operand are the SAME register, so each lane loads one double per
// 8-row tile straight from HBM (8 consecutive doubles of 4 columns per warp load)
// and all ntile(ntile+1)/2 lower tiles are updated from those registers.
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// one warp: partial G over member columns [m0,m1) into dst[q] (shared or global memory).
// GRAD: the same pass also forms the warp's partial gradient (dstg[a] = sum_col J(a,col) x[col], the four
// lanes that hold one row slot folded by two shuffles) and adds the x^2 of its columns to n2 -- the
// fused evaluation reads every Jacobian value exactly once for Jt*x, |x|^2 AND the class block of Jt*Jt'.
template<int NTILE, bool GRAD>
__device__ __forceinline__ void assemble_task_dmma(const DlbSparseDev& S, const double* __restrict__ Jx,
                                                   const double* __restrict__ x, double* dst, double* dstg, double& n2,
                                                   int k, int m0, int m1, int lane)
{
  const int g = lane >> 2, tt = lane & 3;
  constexpr int NPAIR = NTILE * (NTILE + 1) / 2;
  double acc[NPAIR][2];
#pragma unroll
  for(int i = 0; i < NPAIR; i++) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
  bool on[NTILE];
  double gacc[NTILE];
#pragma unroll
  for(int ti = 0; ti < NTILE; ti++) { on[ti] = 8 * ti + g < k; gacc[ti] = 0.0; }

#pragma unroll 4
  for(int m = m0; m < m1; m += 4)
  {
    const int mm = m + tt;
    const bool valid = mm < m1;
    const unsigned int pos = valid ? S.mem_pos[mm] : 0u;
    double xv = 0.0;
    if(GRAD) { xv = valid ? x[S.mem_col[mm]] : 0.0; if(g == 0) n2 = fma(xv, xv, n2); }
    double v[NTILE];
#pragma unroll
    for(int ti = 0; ti < NTILE; ti++) v[ti] = (valid && on[ti]) ? ldg_stream(Jx + pos + 8 * ti + g) : 0.0;
    if(GRAD)
    {
#pragma unroll
      for(int ti = 0; ti < NTILE; ti++) gacc[ti] = fma(v[ti], xv, gacc[ti]);
    }
    int idx = 0;
#pragma unroll
    for(int ti = 0; ti < NTILE; ti++)
#pragma unroll
      for(int tj = 0; tj <= ti; tj++, idx++) dmma_m8n8k4(acc[idx][0], acc[idx][1], v[ti], v[tj]);
  }
  int idx = 0;
#pragma unroll
  for(int ti = 0; ti < NTILE; ti++)
#pragma unroll
    for(int tj = 0; tj <= ti; tj++, idx++)
    {
      const int a = 8 * ti + g, b0 = 8 * tj + 2 * tt;
      if(a < k)
      {
        const int row = a * (a + 1) / 2;
        if(b0 <= a)     dst[row + b0]     = acc[idx][0];
        if(b0 + 1 <= a) dst[row + b0 + 1] = acc[idx][1];
      }
    }
  if(GRAD)
  {
#pragma unroll
    for(int ti = 0; ti < NTILE; ti++)
    {
      double s0 = gacc[ti];
      s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
      s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
      if(tt == 0 && on[ti]) dstg[8 * ti + g] = s0;
    }
  }
}
template<bool GRAD>
__device__ __forceinline__ void assemble_task_dmma_k(const DlbSparseDev& S, const double* __restrict__ Jx,
                                                     const double* __restrict__ x, double* dst, double* dstg, double& n2,
                                                     int k, int m0, int m1, int lane)
{
  if(k <= 8)       assemble_task_dmma<1, GRAD>(S, Jx, x, dst, dstg, n2, k, m0, m1, lane);
  else if(k <= 16) assemble_task_dmma<2, GRAD>(S, Jx, x, dst, dstg, n2, k, m0, m1, lane);
  else if(k <= 24) assemble_task_dmma<3, GRAD>(S, Jx, x, dst, dstg, n2, k, m0, m1, lane);
  else             assemble_task_dmma<4, GRAD>(S, Jx, x, dst, dstg, n2, k, m0, m1, lane);
}

// scalar FP64 path for long columns (k > 32): lane <-> pair, 8 pairs per lane per sweep
#define ASM_ACC 8
__device__ __forceinline__ void assemble_task_scalar(const DlbSparseDev& S, const double* __restrict__ Jx,
                                                     double* __restrict__ Gpart, int t, int lane, int w)
{
  const int c  = S.task_cls[t];
  const int m0 = S.task_m0[t], m1 = S.task_m1[t];
  const int k  = S.cls_ptr[c+1] - S.cls_ptr[c];
  const int npairs = k * (k + 1) / 2;
  const long long Goff = S.task_Goff[t];
  // the warps of the CTA split the pairs; every warp sweeps all member columns
  for(int q0 = 32 * ASM_ACC * w; q0 < npairs; q0 += 32 * ASM_ACC * TASK_WARPS)
  {
    int pa[ASM_ACC], pb[ASM_ACC];
    double acc[ASM_ACC];
#pragma unroll
    for(int u = 0; u < ASM_ACC; u++)
    {
      const int q = q0 + u * 32 + lane;
      int a = 0, b = 0;
      if(q < npairs)
      {
        a = (int)((sqrt(8.0 * (double)q + 1.0) - 1.0) * 0.5);
        while((a + 1) * (a + 2) / 2 <= q) a++;
        while(a * (a + 1) / 2 > q) a--;
        b = q - a * (a + 1) / 2;
      }
      pa[u] = a; pb[u] = b; acc[u] = 0.0;
    }
    for(int m = m0; m < m1; m++)
    {
      const double* col = Jx + S.mem_pos[m];
#pragma unroll
      for(int u = 0; u < ASM_ACC; u++) acc[u] = fma(col[pa[u]], col[pb[u]], acc[u]);
    }
#pragma unroll
    for(int u = 0; u < ASM_ACC; u++)
    {
      const int q = q0 + u * 32 + lane;
      if(q < npairs) Gpart[Goff + q] = acc[u];
    }
  }
}

// gradient of a long-column task (k > 32), the warps of the CTA over the row slots
__device__ __forceinline__ void grad_task_scalar(const DlbSparseDev& S, const double* __restrict__ Jx, const double* __restrict__ x,
                                                 double* __restrict__ gdst, int t, int lane, int w, double& n2)
{
  const int c  = S.task_cls[t];
  const int m0 = S.task_m0[t], m1 = S.task_m1[t];
  const int k  = S.cls_ptr[c+1] - S.cls_ptr[c];
  for(int a0 = 32 * w; a0 < k; a0 += 32 * TASK_WARPS)
  {
    const int a = a0 + lane;
    double acc = 0.0;
    for(int m = m0; m < m1; m++)
    {
      const double xv = x[S.mem_col[m]];
      if(a < k) acc = fma(ldg_stream(Jx + S.mem_pos[m] + a), xv, acc);
    }
    if(a < k) gdst[a] = acc;
  }
  if(w == 0) for(int m = m0 + lane; m < m1; m += 32) { const double xv = x[S.mem_col[m]]; n2 = fma(xv, xv, n2); }
}

#define ASM_PAIRS_MAX 528        // 32*33/2: class-local lower triangle for k <= 32
// GRAD = false: the class blocks only (Gpart). GRAD = true: the fused evaluation -- class blocks,
// partial gradients (gpart, one block of k entries per task) and the CTA's share of |x|^2 (n2part).
template<bool GRAD>
__global__ void __launch_bounds__(DLB_NT, 3)      // 3 CTAs per SM: the gradient variant must not cost the third one
k_sparse_assemble(DlbSparseDev S, const double* __restrict__ Jx, const double* __restrict__ x,
                  double* __restrict__ Gpart, double* __restrict__ gpart, double* __restrict__ n2part)
{
  __shared__ double shG[TASK_WARPS][ASM_PAIRS_MAX];
  __shared__ double shg[GRAD ? TASK_WARPS : 1][32];
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double n2 = 0.0;
  for(int bt = blockIdx.x; bt < S.nbig; bt += gridDim.x)
  {
    const int t = S.big_tasks[bt];
    const int c = S.task_cls[t];
    const int k = S.cls_ptr[c+1] - S.cls_ptr[c];
    if(k > 32)
    {
      assemble_task_scalar(S, Jx, Gpart, t, lane, w);
      if(GRAD) grad_task_scalar(S, Jx, x, gpart + S.task_goff[t], t, lane, w, n2);
      continue;
    }
    int m0, m1;
    warp_range(S.task_m0[t], S.task_m1[t], w, m0, m1);
    assemble_task_dmma_k<GRAD>(S, Jx, x, shG[w], shg[GRAD ? w : 0], n2, k, m0, m1, lane);
    __syncthreads();
    const int npairs = k * (k + 1) / 2;
    const long long Goff = S.task_Goff[t];
    for(int q = threadIdx.x; q < npairs; q += DLB_NT)
    {
      double s0 = 0.0;
#pragma unroll
      for(int u = 0; u < TASK_WARPS; u++) s0 += shG[u][q];
      Gpart[Goff + q] = s0;
    }
    if(GRAD && threadIdx.x < k)
    {
      double s0 = 0.0;
#pragma unroll
      for(int u = 0; u < TASK_WARPS; u++) s0 += shg[GRAD ? u : 0][threadIdx.x];
      gpart[S.task_goff[t] + threadIdx.x] = s0;
    }
    __syncthreads();
  }
  if(GRAD)
  {
    n2 = block_sum(n2, sh);
    if(threadIdx.x == 0) n2part[blockIdx.x] = n2;
  }
}

// small tasks: the same DMMA SYRK, one warp per task, partial G (and gradient) straight to global memory
template<bool GRAD>
__global__ void __launch_bounds__(DLB_NT)
k_sparse_assemble_small(DlbSparseDev S, const int* __restrict__ tasks, int ntasks,
                        const double* __restrict__ Jx, const double* __restrict__ x,
                        double* __restrict__ Gpart, double* __restrict__ gpart, double* __restrict__ n2part)
{
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31;
  const int wg = blockIdx.x * TASK_WARPS + (threadIdx.x >> 5), nw = gridDim.x * TASK_WARPS;
  double n2 = 0.0;
  for(int st = wg; st < ntasks; st += nw)
  {
    const int t = tasks[st];
    const int c = S.task_cls[t];
    const int k = S.cls_ptr[c+1] - S.cls_ptr[c];
    assemble_task_dmma_k<GRAD>(S, Jx, x, Gpart + S.task_Goff[t], GRAD ? gpart + S.task_goff[t] : (double*)0, n2,
                               k, S.task_m0[t], S.task_m1[t], lane);
  }
  if(GRAD)
  {
    n2 = block_sum(n2, sh);
    if(threadIdx.x == 0) n2part[blockIdx.x] = n2;
  }
}

// ------------------------------------------------------------ host launchers
static inline int grid_for_tasks(int ntasks, int sm_count)
{
  const int cap = sm_count * 8;
  return ntasks < 1 ? 1 : (ntasks > cap ? cap : ntasks);
}
// warp-per-task pipelines: two CTAs of 8 warps fit per SM (shared memory)
static inline int grid_for_warp_tasks(int ntasks, int sm_count)
{
  const int cap = sm_count * 2;
  const int g = (ntasks + TASK_WARPS - 1) / TASK_WARPS;
  return g < 1 ? 1 : (g > cap ? cap : g);
}
static inline int grid_for_small(int nsmall, int sm_count)
{
  const int cap = sm_count * 8;
  const int g = (nsmall + TASK_WARPS - 1) / TASK_WARPS;
  return g < 1 ? 1 : (g > cap ? cap : g);
}
// lane groups of G, two tasks per group and round
static inline int grid_for_groups(int nsmall, int G, int sm_count)
{
  const int cap = sm_count * 8;
  const int per_cta = 2 * (DLB_NT / G);
  const int g = (nsmall + per_cta - 1) / per_cta;
  return g < 1 ? 1 : (g > cap ? cap : g);
}
static inline int grid_for_range(int nrange, int sm_count)
{
  const int cap = sm_count * 3;          // three CTAs of 8 warps per SM (shared-memory rings)
  const int g = (nrange + TASK_WARPS - 1) / TASK_WARPS;
  return g < 1 ? 1 : (g > cap ? cap : g);
}
int dlb_sparse_n2part_size(const DlbSparseDev& S, int sm_count)
{
  const int a = grid_for_range(S.nrange, sm_count) + grid_for_warp_tasks(S.ngj_big, sm_count) + grid_for_groups(S.nsmall, 32, sm_count);
  const int b = grid_for_tasks(S.nbig, sm_count) + grid_for_small(S.nasm_small, sm_count) + grid_for_groups(S.nfused, 32, sm_count);
  return a > b ? a : b;
}

template<int CHUNK, int NST>
static void launch_range_grad(const DlbSparseDev& S, const double* Jx, const double* x, double* gpart, double* n2part,
                              int g0, int nu, cudaStream_t st)
{
  const size_t smem = sizeof(double) * TASK_WARPS * (NST * (CHUNK / 8 + 4) + 128 + 4);
  static DlbPerDeviceOnce attr_once;
  if(attr_once.first())
  {
    cudaFuncSetAttribute(k_range_grad<1, CHUNK, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_range_grad<2, CHUNK, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_range_grad<3, CHUNK, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_range_grad<4, CHUNK, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  if(nu <= 1)      k_range_grad<1, CHUNK, NST><<<g0, DLB_NT, smem, st>>>(S, Jx, x, gpart, n2part);
  else if(nu == 2) k_range_grad<2, CHUNK, NST><<<g0, DLB_NT, smem, st>>>(S, Jx, x, gpart, n2part);
  else if(nu == 3) k_range_grad<3, CHUNK, NST><<<g0, DLB_NT, smem, st>>>(S, Jx, x, gpart, n2part);
  else             k_range_grad<4, CHUNK, NST><<<g0, DLB_NT, smem, st>>>(S, Jx, x, gpart, n2part);
}

void dlb_launch_sparse_grad(const DlbSparseDev& S, const double* Jx, const double* x, double* gpart,
                            double* n2part, double* Jtx, double* part, unsigned int* counter,
                            DlbScalars* sc, int sm_count, cudaStream_t st)
{
  int g1 = 0;
  if(S.nrange > 0)
  {
    const int g0 = grid_for_range(S.nrange, sm_count);
    const int nu = (S.range_kmax + 31) / 32;
    launch_range_grad<RANGE_CHUNK, RANGE_NST>(S, Jx, x, gpart, n2part, g0, nu, st);
    g1 += g0;
  }
  if(S.ngj_big > 0)
  {
    const int gb = grid_for_warp_tasks(S.ngj_big, sm_count);
    const size_t smem = sizeof(double) * TASK_WARPS * PIPE_WARP_DOUBLES;
    static DlbPerDeviceOnce attr_once;
    if(attr_once.first()) { cudaFuncSetAttribute(k_sparse_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); }
    k_sparse_grad<<<gb, DLB_NT, smem, st>>>(S, Jx, x, gpart, n2part + g1);
    g1 += gb;
  }
  if(S.nsmall > 0)
  {
    const int G = S.small_group;
    const int g2 = grid_for_groups(S.nsmall, G, sm_count);
    if(G == 8)       k_sparse_grad_small<8><<<g2, DLB_NT, 0, st>>>(S, S.small_info, S.nsmall, Jx, x, gpart, n2part + g1);
    else if(G == 16) k_sparse_grad_small<16><<<g2, DLB_NT, 0, st>>>(S, S.small_info, S.nsmall, Jx, x, gpart, n2part + g1);
    else             k_sparse_grad_small<32><<<g2, DLB_NT, 0, st>>>(S, S.small_info, S.nsmall, Jx, x, gpart, n2part + g1);
    g1 += g2;
  }
  int g = (S.n + DLB_NT - 1) / DLB_NT; if(g < (S.nmedium + 7) / 8) g = (S.nmedium + 7) / 8; if(g < S.nheavy) g = S.nheavy;
  if(g > sm_count * 8) g = sm_count * 8; if(g < 1) g = 1;
  k_sparse_grad_reduce<<<g, DLB_NT, 0, st>>>(S, gpart, n2part, g1, Jtx, part, counter, sc, (DlbPublished*)0, 0ull, 0);
}

// The fused evaluation: ONE pass over the Jacobian values gives the class blocks of Jt*Jt' (Gpart), the
// partial gradients (one block per task: the plan must have been built without range tasks) and
// |x|^2; then the per-state reduction. Classes assembled by the fused leaf kernel (dlb_leaf.cu) have no
// class block: their gradient comes from the lane-group kernel. pub != NULL: the reduction publishes
// the scalars to the host.
int dlb_launch_sparse_eval_pass(const DlbSparseDev& S, const double* Jx, const double* x, double* Gpart, double* gpart,
                                double* n2part, int sm_count, cudaStream_t st)
{
  int g1 = 0;
  if(S.nbig > 0)
  {
    const int gb = grid_for_tasks(S.nbig, sm_count);
    k_sparse_assemble<true><<<gb, DLB_NT, 0, st>>>(S, Jx, x, Gpart, gpart, n2part);
    g1 += gb;
  }
  if(S.nasm_small > 0)
  {
    const int gs = grid_for_small(S.nasm_small, sm_count);
    k_sparse_assemble_small<true><<<gs, DLB_NT, 0, st>>>(S, S.asm_small_tasks, S.nasm_small, Jx, x, Gpart, gpart, n2part + g1);
    g1 += gs;
  }
  if(S.nfused > 0)
  {
    const int G = S.small_group;
    const int g2 = grid_for_groups(S.nfused, G, sm_count);
    if(G == 8)       k_sparse_grad_small<8><<<g2, DLB_NT, 0, st>>>(S, S.fused_info, S.nfused, Jx, x, gpart, n2part + g1);
    else if(G == 16) k_sparse_grad_small<16><<<g2, DLB_NT, 0, st>>>(S, S.fused_info, S.nfused, Jx, x, gpart, n2part + g1);
    else             k_sparse_grad_small<32><<<g2, DLB_NT, 0, st>>>(S, S.fused_info, S.nfused, Jx, x, gpart, n2part + g1);
    g1 += g2;
  }
  return g1;
}
void dlb_launch_sparse_eval_reduce(const DlbSparseDev& S, const double* gpart, const double* n2part, int n2count, double* Jtx,
                                   double* part, unsigned int* counter, DlbScalars* sc, DlbPublished* pub, unsigned long long seq,
                                   int n2_behind_Jtx, int sm_count, cudaStream_t st)
{
  int g = (S.n + DLB_NT - 1) / DLB_NT; if(g < (S.nmedium + 7) / 8) g = (S.nmedium + 7) / 8; if(g < S.nheavy) g = S.nheavy;
  if(g > sm_count * 8) g = sm_count * 8; if(g < 1) g = 1;
  k_sparse_grad_reduce<<<g, DLB_NT, 0, st>>>(S, gpart, n2part, n2count, Jtx, part, counter, sc, pub, seq, n2_behind_Jtx);
}

void dlb_launch_sparse_jv(const DlbSparseDev& S, const double* Jx, const double* v, double* part,
                          unsigned int* counter, double* dst, int sm_count, cudaStream_t st)
{
  // up to three kernels (range tasks, big class tasks, small class tasks): the earlier ones leave
  // their totals in scratch scalars, the last one adds them and writes *dst
  double* scratch = part + 5 * (size_t)sm_count * 8 + 8;
  const int kinds = (S.nrange > 0) + (S.ngj_big > 0) + (S.nsmall > 0);
  int done = 0;
  const double* adds[2] = {NULL, NULL};
  auto target = [&]() { return done == kinds - 1 ? dst : scratch + done; };
  if(S.nrange > 0)
  {
    const size_t smem = sizeof(double) * TASK_WARPS * RANGE_WARP_DOUBLES;
    static DlbPerDeviceOnce attr_once;
    if(attr_once.first())
    {
      cudaFuncSetAttribute(k_range_jv<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(k_range_jv<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(k_range_jv<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(k_range_jv<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    double* out = target();
    const int gr = grid_for_range(S.nrange, sm_count), nu = (S.range_kmax + 31) / 32;
    if(nu <= 1)      k_range_jv<1><<<gr, DLB_NT, smem, st>>>(S, Jx, v, part, counter, adds[0], adds[1], out);
    else if(nu == 2) k_range_jv<2><<<gr, DLB_NT, smem, st>>>(S, Jx, v, part, counter, adds[0], adds[1], out);
    else if(nu == 3) k_range_jv<3><<<gr, DLB_NT, smem, st>>>(S, Jx, v, part, counter, adds[0], adds[1], out);
    else             k_range_jv<4><<<gr, DLB_NT, smem, st>>>(S, Jx, v, part, counter, adds[0], adds[1], out);
    if(out != dst) adds[done] = out;
    done++;
  }
  if(S.ngj_big > 0)
  {
    const size_t smem = sizeof(double) * TASK_WARPS * PIPE_WARP_DOUBLES;
    static DlbPerDeviceOnce attr_once;
    if(attr_once.first()) { cudaFuncSetAttribute(k_sparse_jv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); }
    double* out = target();
    k_sparse_jv<<<grid_for_warp_tasks(S.ngj_big, sm_count), DLB_NT, smem, st>>>(S, Jx, v, part, counter, adds[0], adds[1], out);
    if(out != dst) adds[done] = out;
    done++;
  }
  if(S.nsmall > 0)
  {
    const int G = S.small_group;
    const int g2 = grid_for_groups(S.nsmall, G, sm_count);
    double* out = target();
    if(G == 8)       k_sparse_jv_small<8><<<g2, DLB_NT, 0, st>>>(S, S.small_info, S.nsmall, Jx, v, part, counter, adds[0], adds[1], out);
    else if(G == 16) k_sparse_jv_small<16><<<g2, DLB_NT, 0, st>>>(S, S.small_info, S.nsmall, Jx, v, part, counter, adds[0], adds[1], out);
    else             k_sparse_jv_small<32><<<g2, DLB_NT, 0, st>>>(S, S.small_info, S.nsmall, Jx, v, part, counter, adds[0], adds[1], out);
    done++;
  }
  if(kinds == 0) cudaMemsetAsync(dst, 0, sizeof(double), st);
}

// |J v|^2 = v' (Jt Jt') v from the class-local blocks the assembly pass has already formed (what the
// reference's dense-products path does, dogleg.c:582-596): sum over the tasks of
// sum_{a>=b} (2 - [a==b]) v[row_a] G_ab v[row_b]. One warp per task, the lanes over the packed pairs;
// ~5 MB instead of another pass over the 180 MB of Jacobian values.
__global__ void __launch_bounds__(DLB_NT)
k_gpart_quadform(DlbSparseDev S, const int* __restrict__ listA, int nA, const int* __restrict__ listB, int nB,
                 const double* __restrict__ Gpart, const double* __restrict__ v,
                 double* part, unsigned int* counter, const double* add0, double* dst)
{
  const double total = quadform_partial(S, listA, nA, listB, nB, Gpart, v, blockIdx.x * TASK_WARPS + (threadIdx.x >> 5),
                                        gridDim.x * TASK_WARPS, threadIdx.x & 31);
  double out[5];
  if(grid_reduce5(total, 0.0, 0.0, 0.0, 0.0, part, counter, out)) *dst = out[0] + (add0 ? *add0 : 0.0);
}

// |J v|^2 with the assembled blocks: quadratic form over every task that has a Gpart block, a pass
// over the Jacobian only for the classes the fused leaf kernel assembles itself
void dlb_launch_sparse_jv_quad(const DlbSparseDev& S, const double* Jx, const double* Gpart, const double* v, double* part,
                               unsigned int* counter, double* dst, int sm_count, cudaStream_t st)
{
  double* scratch = part + 5 * (size_t)sm_count * 8 + 8;
  const double* add = NULL;
  if(S.nfused > 0)
  {
    const int G = S.small_group;
    const int g2 = grid_for_groups(S.nfused, G, sm_count);
    double* out = (S.nbig + S.nasm_small > 0) ? scratch : dst;
    if(G == 8)       k_sparse_jv_small<8><<<g2, DLB_NT, 0, st>>>(S, S.fused_info, S.nfused, Jx, v, part, counter, NULL, NULL, out);
    else if(G == 16) k_sparse_jv_small<16><<<g2, DLB_NT, 0, st>>>(S, S.fused_info, S.nfused, Jx, v, part, counter, NULL, NULL, out);
    else             k_sparse_jv_small<32><<<g2, DLB_NT, 0, st>>>(S, S.fused_info, S.nfused, Jx, v, part, counter, NULL, NULL, out);
    if(out != dst) add = out;
  }
  const int nt = S.nbig + S.nasm_small;
  if(nt > 0)
  {
    int g = (nt + TASK_WARPS - 1) / TASK_WARPS;
    if(g > sm_count * 8) g = sm_count * 8;
    k_gpart_quadform<<<g, DLB_NT, 0, st>>>(S, S.big_tasks, S.nbig, S.asm_small_tasks, S.nasm_small, Gpart, v, part, counter, add, dst);
  }
  else if(S.nfused == 0) cudaMemsetAsync(dst, 0, sizeof(double), st);
}

void dlb_launch_sparse_assemble(const DlbSparseDev& S, const double* Jx, double* Gpart, int all_small,
                                int sm_count, cudaStream_t st)
{
  if(S.nbig > 0)   k_sparse_assemble<false><<<grid_for_tasks(S.nbig, sm_count), DLB_NT, 0, st>>>(S, Jx, (const double*)0, Gpart, (double*)0, (double*)0);
  const int* tasks = all_small ? S.small_tasks : S.asm_small_tasks;
  const int nt = all_small ? S.nsmall : S.nasm_small;
  if(nt > 0) k_sparse_assemble_small<false><<<grid_for_small(nt, sm_count), DLB_NT, 0, st>>>(S, tasks, nt, Jx, (const double*)0, Gpart, (double*)0, (double*)0);
}
