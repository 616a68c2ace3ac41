// dlb_leaf.cu -- leaf fronts of the elimination tree, one WARP per front.
//
// A bundle adjustment has a million leaf supernodes (the 3 coordinates of a point) whose fronts
// are ~39 x 39: a CTA per front (dlb_front.cu) and a separate Jt*Jt' pass through Gpart
// (dlb_sparse.cu) would be almost entirely launch/barrier overhead and intermediate traffic.
// Here one warp
//   1. reads the few measurement columns of the front's pattern classes straight from Jt->x,
//   2. forms their Jt*Jt' contributions in the packed lower triangle of the front (shared memory),
//   3. adds lambda, eliminates the pivot columns (right-looking, rsqrt pivots),
//   4. writes the L panel and the update matrix (lower triangle only) to the front's storage,
// i.e. it replaces, for these fronts, k_sparse_assemble_small + k_front_level of the reference's
// cholmod_factorize (dogleg.c:656-665) in one pass: 8 x 12 doubles in, 780 doubles out per
// point. The matching triangular solves (cholmod_solve, dogleg.c:853-856) are warp-per-front too.
// Everything is summed in a fixed order: bit-reproducible.
#include "dlb_common.cuh"
#include "dlb_device.h"

#define LEAF_WARPS 4
#define LEAF_MAX_MEMBERS 4

// packed column-major lower triangle of an r x r matrix: column c starts at c*r - c(c-1)/2
__device__ __forceinline__ int tri_col(int c, int r) { return c * r - (c * (c - 1)) / 2; }

// (row, column) of every entry of the packed lower triangle for the table's r (most leaf fronts
// of a problem have the same size); fronts of another size decode incrementally
__device__ __forceinline__ void tri_next(int& c, int& p, int r) { while(c < r && p >= r - c) { p -= r - c; c++; } }

__global__ void __launch_bounds__(32 * LEAF_WARPS)
k_leaf_fronts(DlbFrontDev F, DlbSparseDev S, int q0, int q1, const double* __restrict__ Jx,
              double* __restrict__ fronts, double lambda, long long* minor, int tri_max, int table_r, int eliminate)
{
  extern __shared__ double sh_leaf[];           // per warp: tri_max doubles of front + 32*LEAF_MAX_MEMBERS of values; then the table
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double* T = sh_leaf + (size_t)w * (tri_max + 32 * LEAF_MAX_MEMBERS);
  double* V = T + tri_max;
  unsigned short* tab = (unsigned short*)(sh_leaf + (size_t)LEAF_WARPS * (tri_max + 32 * LEAF_MAX_MEMBERS));
  {
    const int nt = table_r * (table_r + 1) / 2;
    for(int idx = threadIdx.x; idx < nt; idx += blockDim.x)
    { // column c of the packed triangle starts at c*r - c(c-1)/2
      int c = (int)((2.0 * table_r + 1.0 - sqrt((2.0 * table_r + 1.0) * (2.0 * table_r + 1.0) - 8.0 * idx)) * 0.5);
      while(c > 0 && tri_col(c, table_r) > idx) c--;
      while(tri_col(c + 1, table_r) <= idx) c++;
      tab[idx] = (unsigned short)((c << 8) | (c + idx - tri_col(c, table_r)));
    }
  }
  __syncthreads();
  for(int q = q0 + blockIdx.x * LEAF_WARPS + w; q < q1; q += gridDim.x * LEAF_WARPS)
  {
    const int s  = F.level_sn[q];
    const int c0 = F.sn_first[s], nc = F.sn_first[s+1] - c0;
    const int r  = F.rows_ptr[s+1] - F.rows_ptr[s];
    const int ntri = r * (r + 1) / 2;
    const bool use_tab = r == table_r;
    for(int idx = lane; idx < ntri; idx += 32) T[idx] = 0.0;
    __syncwarp();
    // ---- elements: the front's pattern classes, straight from the Jacobian values ----
    for(int ci = F.fcls_ptr[s]; ci < F.fcls_ptr[s+1]; ci++)
    {
      const int c = F.fcls_list[ci];
      const int r0 = S.cls_ptr[c], k = S.cls_ptr[c+1] - r0;
      const int myloc = lane < k ? S.cls_loc[r0 + lane] : 0;
      const int t = S.cls_task_ptr[c];
      const int m0 = S.task_m0[t], nm = S.task_m1[t] - m0;
      for(int m = 0; m < nm; m++) if(lane < k) V[32 * m + lane] = ldg_stream(Jx + S.mem_pos[m0 + m] + lane);
      __syncwarp();
      const int npairs = k * (k + 1) / 2;
      // lane-strided walk over the pairs (a >= b), a and b advanced incrementally; all lanes run
      // the same number of rounds (the shuffles below need the whole warp)
      int a = 0, b = lane;
      while(b > a) { b -= a + 1; a++; }
      for(int p0 = 0; p0 < npairs; p0 += 32)
      {
        const bool on = p0 + lane < npairs;
        const int aa = on ? a : 0, bb = on ? b : 0;
        const int la = __shfl_sync(0xffffffffu, myloc, aa), lb = __shfl_sync(0xffffffffu, myloc, bb);
        if(on)
        {
          double g = 0.0;
          for(int m = 0; m < nm; m++) g = fma(V[32 * m + a], V[32 * m + b], g);
          const int row = la > lb ? la : lb, col = la > lb ? lb : la;
          T[tri_col(col, r) + row - col] += g;
        }
        b += 32;
        while(b > a) { b -= a + 1; a++; }
      }
      __syncwarp();
    }
    if(eliminate)
    {
      for(int j = lane; j < nc; j += 32) T[tri_col(j, r)] += lambda;
      __syncwarp();
      // ---- the nc pivot columns, left-looking inside the panel ----
      bool failed = false;
      for(int j = 0; j < nc; j++)
      {
        const int oj = tri_col(j, r);
        for(int i = j + lane; i < r; i += 32)
        {
          double v = T[oj + i - j];
          for(int jj = 0; jj < j; jj++) { const int o = tri_col(jj, r) - jj; v = fma(-T[o + i], T[o + j], v); }
          T[oj + i - j] = v;
        }
        __syncwarp();
        const double d = T[oj];
        if(!(d > 0.0) || isinf(d)) { if(lane == 0) atomicMin(minor, (long long)(c0 + j)); failed = true; break; }
        const double rs = rsqrt(d);
        __syncwarp();
        for(int i = j + lane; i < r; i += 32) T[oj + i - j] = (i == j) ? d * rs : T[oj + i - j] * rs;
        __syncwarp();
      }
      // ---- one pass over the trailing block: every entry gets all nc rank-1 updates ----
      if(!failed && nc < r)
      {
        const int o1 = tri_col(nc, r);
        int c = nc, p = lane;
        tri_next(c, p, r);
        for(int idx = o1 + lane; idx < ntri; idx += 32)
        {
          int row, col;
          if(use_tab) { const unsigned int rc = tab[idx]; col = rc >> 8; row = rc & 0xff; }
          else        { col = c; row = c + p; p += 32; tri_next(c, p, r); }
          double v = T[idx];
          for(int jj = 0; jj < nc; jj++) { const int o = tri_col(jj, r) - jj; v = fma(-T[o + row], T[o + col], v); }
          T[idx] = v;
        }
      }
      __syncwarp();
    }
    // ---- write the lower triangle into the front's r x r column-major storage ----
    {
      double* A = fronts + F.front_off[s];
      int c = 0, p = lane;
      tri_next(c, p, r);
      for(int idx = lane; idx < ntri; idx += 32)
      {
        int row, col;
        if(use_tab) { const unsigned int rc = tab[idx]; col = rc >> 8; row = rc & 0xff; }
        else        { col = c; row = c + p; p += 32; tri_next(c, p, r); }
        A[(size_t)col * r + row] = T[idx];
      }
    }
    __syncwarp();
  }
}

void dlb_launch_leaf_fronts(const DlbFrontDev& F, const DlbSparseDev& S, int q0, int q1, const double* Jx,
                            double* fronts, double lambda, long long* minor, int max_rows, int eliminate,
                            int sm_count, cudaStream_t st)
{
  if(q1 <= q0) return;
  const int tri_max = max_rows * (max_rows + 1) / 2;
  const size_t smem = sizeof(double) * LEAF_WARPS * (size_t)(tri_max + 32 * LEAF_MAX_MEMBERS) + sizeof(unsigned short) * tri_max + 16;
  static bool attr_set = false;
  if(!attr_set)
  {
    cudaFuncSetAttribute(k_leaf_fronts, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    attr_set = true;
  }
  long long g = ((long long)(q1 - q0) + LEAF_WARPS - 1) / LEAF_WARPS;
  const long long cap = (long long)sm_count * 16;
  if(g > cap) g = cap;
  // the decode table is built for the largest front size (in a bundle adjustment: the size of nearly all)
  k_leaf_fronts<<<(int)g, 32 * LEAF_WARPS, smem, st>>>(F, S, q0, q1, Jx, fronts, lambda, minor, tri_max, max_rows, eliminate);
}

// ------------------------------------------------------------------ solves
// forward: y = L^-1 P b for leaf fronts (no children): lane i owns rows i and i+32
__global__ void __launch_bounds__(256)
k_leaf_solve_fwd(DlbFrontDev F, int q0, int q1, const double* __restrict__ fronts,
                 const double* __restrict__ rhs, double* __restrict__ ywork, double* __restrict__ zperm, int nrhs)
{
  const int lane = threadIdx.x & 31;
  const int wg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = gridDim.x * (blockDim.x >> 5);
  for(int q = q0 + wg; q < q1; q += nw)
  {
    const int s  = F.level_sn[q];
    const int c0 = F.sn_first[s], nc = F.sn_first[s+1] - c0;
    const int rp = F.rows_ptr[s], r = F.rows_ptr[s+1] - rp;
    const double* A = fronts + F.front_off[s];
    for(int rh = 0; rh < nrhs; rh++)
    {
      double y0 = lane < nc ? rhs[(size_t)rh * F.n + F.perm[c0 + lane]] : 0.0, y1 = 0.0;   // nc <= 32
      for(int j = 0; j < nc; j++)
      {
        const double yj = __shfl_sync(0xffffffffu, y0, j) / A[j + (size_t)j * r];
        if(lane == j) y0 = yj;
        if(lane > j && lane < r)  y0 = fma(-A[lane + (size_t)j * r], yj, y0);
        if(lane + 32 < r)         y1 = fma(-A[lane + 32 + (size_t)j * r], yj, y1);
      }
      double* yg = ywork + (size_t)rh * F.ytot + rp;
      if(lane < r) yg[lane] = y0;
      if(lane + 32 < r) yg[lane + 32] = y1;
      if(lane < nc) zperm[(size_t)rh * F.n + c0 + lane] = y0;
    }
  }
}
// backward: x = L^-T y in place in zperm
__global__ void __launch_bounds__(256)
k_leaf_solve_bwd(DlbFrontDev F, int q0, int q1, const double* __restrict__ fronts, double* __restrict__ zperm, int nrhs)
{
  const int lane = threadIdx.x & 31;
  const int wg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = gridDim.x * (blockDim.x >> 5);
  for(int q = q0 + wg; q < q1; q += nw)
  {
    const int s  = F.level_sn[q];
    const int c0 = F.sn_first[s], nc = F.sn_first[s+1] - c0;
    const int rp = F.rows_ptr[s], r = F.rows_ptr[s+1] - rp;
    const double* A = fronts + F.front_off[s];
    const int* rows = F.rows + rp;
    for(int rh = 0; rh < nrhs; rh++)
    {
      double* z = zperm + (size_t)rh * F.n;
      double x0 = lane < r ? z[rows[lane]] : 0.0;
      const double x1 = lane + 32 < r ? z[rows[lane + 32]] : 0.0;
      for(int j = nc - 1; j >= 0; j--)
      {
        double acc = 0.0;
        if(lane > j && lane < r) acc = A[lane + (size_t)j * r] * x0;
        if(lane + 32 < r)        acc = fma(A[lane + 32 + (size_t)j * r], x1, acc);
        acc = warp_sum_all(acc);
        const double xj = (__shfl_sync(0xffffffffu, x0, j) - acc) / A[j + (size_t)j * r];
        if(lane == j) x0 = xj;
      }
      if(lane < nc) z[c0 + lane] = x0;
    }
  }
}
static inline int leaf_solve_grid(int n, int sm_count)
{
  long long g = ((long long)n + 7) / 8;
  const long long cap = (long long)sm_count * 16;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}
void dlb_launch_leaf_solve_fwd(const DlbFrontDev& F, int q0, int q1, const double* fronts, const double* rhs,
                               double* ywork, double* zperm, int nrhs, int sm_count, cudaStream_t st)
{
  if(q1 <= q0) return;
  k_leaf_solve_fwd<<<leaf_solve_grid(q1 - q0, sm_count), 256, 0, st>>>(F, q0, q1, fronts, rhs, ywork, zperm, nrhs);
}
void dlb_launch_leaf_solve_bwd(const DlbFrontDev& F, int q0, int q1, const double* fronts, double* zperm,
                               int nrhs, int sm_count, cudaStream_t st)
{
  if(q1 <= q0) return;
  k_leaf_solve_bwd<<<leaf_solve_grid(q1 - q0, sm_count), 256, 0, st>>>(F, q0, q1, fronts, zperm, nrhs);
}
