// dlb_sparse.cu -- streaming kernels over the CCS Jacobian Jt (values only).
//
// Layout in HBM: the Jacobian values stay exactly as the callback wrote them
// (CCS order, column j = measurement j). The fixed sparsity pattern is NOT read
// per nonzero: measurement columns with identical row lists form a "pattern
// class" (dlb_symbolic.h); a task is (class, contiguous range of its member
// columns) and is processed by one warp with lane == slot inside the column, so
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

// ---------------------------------------------------------------- gradient
// gpart[task_goff[t] + a] = sum over the task's member columns of J(a,col)*x[col]
// n2part[t]               = sum over the task's member columns of x[col]^2
__global__ void __launch_bounds__(DLB_NT)
k_sparse_grad(DlbSparseDev S, const double* __restrict__ Jx, const double* __restrict__ x,
              double* __restrict__ gpart, double* __restrict__ n2part)
{
  const int warps_per_cta = DLB_NT / 32;
  const int lane = threadIdx.x & 31;
  for(int t = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); t < S.ntasks; t += gridDim.x * warps_per_cta)
  {
    const int c  = S.task_cls[t];
    const int m0 = S.task_m0[t], m1 = S.task_m1[t];
    const int k  = S.cls_ptr[c+1] - S.cls_ptr[c];
    const long long goff = S.task_goff[t];

    for(int a0 = 0; a0 < k; a0 += 32)
    {
      const int a = a0 + lane;
      const bool on = a < k;
      double acc = 0.0;
      int m = m0;
      for(; m + 4 <= m1; m += 4)
      {
        const unsigned int p0 = S.mem_pos[m], p1 = S.mem_pos[m+1], p2 = S.mem_pos[m+2], p3 = S.mem_pos[m+3];
        const double x0 = x[S.mem_col[m]], x1 = x[S.mem_col[m+1]], x2 = x[S.mem_col[m+2]], x3 = x[S.mem_col[m+3]];
        double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
        if(on) { v0 = ldg_stream(Jx + p0 + a); v1 = ldg_stream(Jx + p1 + a); v2 = ldg_stream(Jx + p2 + a); v3 = ldg_stream(Jx + p3 + a); }
        acc = fma(v0, x0, acc); acc = fma(v1, x1, acc); acc = fma(v2, x2, acc); acc = fma(v3, x3, acc);
      }
      for(; m < m1; m++)
      {
        const double xv = x[S.mem_col[m]];
        if(on) acc = fma(ldg_stream(Jx + S.mem_pos[m] + a), xv, acc);
      }
      if(on) gpart[goff + a] = acc;
    }
    // |x|^2 over the member columns of this task
    double n2 = 0.0;
    for(int m = m0 + lane; m < m1; m += 32) { const double xv = x[S.mem_col[m]]; n2 = fma(xv, xv, n2); }
    n2 = warp_sum(n2);
    if(lane == 0) n2part[t] = n2;
  }
}

// Jt_x[i] = sum of its gpart contributions in a fixed order (one warp per state),
// then |x|^2, |Jt x|^2, max|Jt x| for the whole vector
__global__ void __launch_bounds__(DLB_NT)
k_sparse_grad_reduce(DlbSparseDev S, const double* __restrict__ gpart, const double* __restrict__ n2part,
                     double* __restrict__ Jtx, double* part, unsigned int* counter, DlbScalars* sc)
{
  const int warps_per_cta = DLB_NT / 32;
  const int lane = threadIdx.x & 31;
  double g2 = 0.0, gmax = 0.0, n2 = 0.0;
  for(int i = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); i < S.n; i += gridDim.x * warps_per_cta)
  {
    double s = 0.0;
    for(int q = S.ginv_ptr[i] + lane; q < S.ginv_ptr[i+1]; q += 32) s += gpart[S.ginv_idx[q]];
    s = warp_sum(s);
    if(lane == 0) { Jtx[i] = s; g2 = fma(s, s, g2); gmax = fmax(gmax, fabs(s)); }
  }
  for(int t = blockIdx.x * blockDim.x + threadIdx.x; t < S.ntasks; t += gridDim.x * blockDim.x) n2 += n2part[t];
  double out[5];
  if(grid_reduce5(n2, g2, 0.0, 0.0, gmax, part, counter, out))
  {
    sc->norm2_x = out[0]; sc->norm2_Jtx = out[1]; sc->maxabs_Jtx = out[4];
  }
}

// ------------------------------------------------------------------ |J v|^2
// jvpart[t] = sum over the task's member columns of (sum_a J(a,col) v[row_a])^2
__global__ void __launch_bounds__(DLB_NT)
k_sparse_jv(DlbSparseDev S, const double* __restrict__ Jx, const double* __restrict__ v,
            double* __restrict__ jvpart)
{
  const int warps_per_cta = DLB_NT / 32;
  const int lane = threadIdx.x & 31;
  for(int t = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); t < S.ntasks; t += gridDim.x * warps_per_cta)
  {
    const int c  = S.task_cls[t];
    const int m0 = S.task_m0[t], m1 = S.task_m1[t];
    const int r0 = S.cls_ptr[c], k = S.cls_ptr[c+1] - r0;
    double total = 0.0;
    if(k <= 32)
    {
      const bool on = lane < k;
      const double va = on ? v[S.cls_rows[r0 + lane]] : 0.0;
      int m = m0;
      for(; m + 4 <= m1; m += 4)
      {
        const unsigned int p0 = S.mem_pos[m], p1 = S.mem_pos[m+1], p2 = S.mem_pos[m+2], p3 = S.mem_pos[m+3];
        double d0 = 0, d1 = 0, d2 = 0, d3 = 0;
        if(on) { d0 = ldg_stream(Jx + p0 + lane) * va; d1 = ldg_stream(Jx + p1 + lane) * va;
                 d2 = ldg_stream(Jx + p2 + lane) * va; d3 = ldg_stream(Jx + p3 + lane) * va; }
        d0 = warp_sum_all(d0); d1 = warp_sum_all(d1); d2 = warp_sum_all(d2); d3 = warp_sum_all(d3);
        total = fma(d0, d0, total); total = fma(d1, d1, total); total = fma(d2, d2, total); total = fma(d3, d3, total);
      }
      for(; m < m1; m++)
      {
        double d = on ? ldg_stream(Jx + S.mem_pos[m] + lane) * va : 0.0;
        d = warp_sum_all(d);
        total = fma(d, d, total);
      }
    }
    else
    {
      for(int m = m0; m < m1; m++)
      {
        const unsigned int p = S.mem_pos[m];
        double d = 0.0;
        for(int a = lane; a < k; a += 32) d = fma(ldg_stream(Jx + p + a), v[S.cls_rows[r0 + a]], d);
        d = warp_sum_all(d);
        total = fma(d, d, total);
      }
    }
    if(lane == 0) jvpart[t] = total;
  }
}

// dst = sum of jvpart in task order
__global__ void __launch_bounds__(DLB_NT)
k_sum_partials(const double* __restrict__ src, int n, double* dst)
{
  __shared__ double sh[32];
  double s = 0.0;
  for(int i = threadIdx.x; i < n; i += blockDim.x) s += src[i];
  s = block_sum(s, sh);
  if(threadIdx.x == 0) *dst = s;
}

// ---------------------------------------------------------------- assembly
// Gpart[task_Goff[t] + q], q = a(a+1)/2 + b (a>=b): sum over the task's member
// columns of J(a,col) J(b,col) -- the class-local lower triangle of Jt Jt'.
// First (scalar FP64) version: lane <-> pair, 8 pairs per lane per sweep.
#define ASM_ACC 8
__global__ void __launch_bounds__(DLB_NT)
k_sparse_assemble(DlbSparseDev S, const double* __restrict__ Jx, double* __restrict__ Gpart)
{
  const int warps_per_cta = DLB_NT / 32;
  const int lane = threadIdx.x & 31;
  for(int t = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); t < S.ntasks; t += gridDim.x * warps_per_cta)
  {
    const int c  = S.task_cls[t];
    const int m0 = S.task_m0[t], m1 = S.task_m1[t];
    const int k  = S.cls_ptr[c+1] - S.cls_ptr[c];
    const int npairs = k * (k + 1) / 2;
    const long long Goff = S.task_Goff[t];
    for(int q0 = 0; q0 < npairs; q0 += 32 * ASM_ACC)
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
}

// ------------------------------------------------------------ host launchers
static inline int grid_for_tasks(int ntasks, int sm_count)
{
  const int warps_per_cta = DLB_NT / 32;
  int g = (ntasks + warps_per_cta - 1) / warps_per_cta;
  const int cap = sm_count * 8;
  return g < 1 ? 1 : (g > cap ? cap : g);
}

void dlb_launch_sparse_grad(const DlbSparseDev& S, const double* Jx, const double* x, double* gpart,
                            double* n2part, double* Jtx, double* part, unsigned int* counter,
                            DlbScalars* sc, int sm_count, cudaStream_t st)
{
  k_sparse_grad<<<grid_for_tasks(S.ntasks, sm_count), DLB_NT, 0, st>>>(S, Jx, x, gpart, n2part);
  int g = (S.n + 7) / 8; if(g > sm_count * 4) g = sm_count * 4; if(g < 1) g = 1;
  k_sparse_grad_reduce<<<g, DLB_NT, 0, st>>>(S, gpart, n2part, Jtx, part, counter, sc);
}

void dlb_launch_sparse_jv(const DlbSparseDev& S, const double* Jx, const double* v, double* jvpart,
                          double* dst, int sm_count, cudaStream_t st)
{
  k_sparse_jv<<<grid_for_tasks(S.ntasks, sm_count), DLB_NT, 0, st>>>(S, Jx, v, jvpart);
  k_sum_partials<<<1, DLB_NT, 0, st>>>(jvpart, S.ntasks, dst);
}

void dlb_launch_sparse_assemble(const DlbSparseDev& S, const double* Jx, double* Gpart,
                                int sm_count, cudaStream_t st)
{
  k_sparse_assemble<<<grid_for_tasks(S.ntasks, sm_count), DLB_NT, 0, st>>>(S, Jx, Gpart);
}
