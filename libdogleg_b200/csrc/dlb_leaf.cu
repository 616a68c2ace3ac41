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
// eligibility limits (dlb_engine.cu checks them): per front at most 32 classes, 32 (class, member)
// pairs, LEAF_VALS Jacobian values and LEAF_LOCS class slots
#define LEAF_VALS 256
#define LEAF_LOCS 128

// packed column-major lower triangle of an r x r matrix: column c starts at c*r - c(c-1)/2
__device__ __forceinline__ int tri_col(int c, int r) { return c * r - (c * (c - 1)) / 2; }

// (row, column) of every entry of the packed lower triangle for the table's r (most leaf fronts
// of a problem have the same size); fronts of another size decode incrementally
__device__ __forceinline__ void tri_next(int& c, int& p, int r) { while(c < r && p >= r - c) { p -= r - c; c++; } }

// Index data is read once per front from DRAM, so the dependent-load chain is what a warp waits
// for: one record per leaf (DlbLeaf), one per class (DlbClsInfo), the (class, member) column
// positions fetched by one lane each, then all value loads of the front in flight together.
__global__ void __launch_bounds__(32 * LEAF_WARPS)
k_leaf_fronts(DlbFrontDev F, DlbSparseDev S, int q0, int q1, const double* __restrict__ Jx,
              double* __restrict__ fronts, double lambda, long long* minor, int tri_max, int table_r, int eliminate)
{
  extern __shared__ double sh_leaf[];           // per warp: front (tri_max) + values (LEAF_VALS) + loc (LEAF_LOCS ints); then the table
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int per_warp = tri_max + LEAF_VALS + LEAF_LOCS / 2;
  double* T = sh_leaf + (size_t)w * per_warp;
  double* V = T + tri_max;
  int* LOC = (int*)(V + LEAF_VALS);
  // CTA-wide tables: (row, col) and col*r+row of every packed entry for r == table_r; (a, b) of
  // every class-local pair index; all index arithmetic of the hot loops becomes a 16-bit load
  unsigned short* tab = (unsigned short*)(sh_leaf + (size_t)LEAF_WARPS * per_warp);
  unsigned short* tabOff = tab + tri_max;
  unsigned short* tabAB = tabOff + tri_max;
  {
    const int nt = table_r * (table_r + 1) / 2;
    for(int idx = threadIdx.x; idx < nt; idx += blockDim.x)
    { // column c of the packed triangle starts at c*r - c(c-1)/2
      int c = (int)((2.0 * table_r + 1.0 - sqrt((2.0 * table_r + 1.0) * (2.0 * table_r + 1.0) - 8.0 * idx)) * 0.5);
      while(c > 0 && tri_col(c, table_r) > idx) c--;
      while(tri_col(c + 1, table_r) <= idx) c++;
      const int row = c + idx - tri_col(c, table_r);
      tab[idx] = (unsigned short)((c << 8) | row);
      tabOff[idx] = (unsigned short)(c * table_r + row);
    }
    for(int p = threadIdx.x; p < 528; p += blockDim.x)
    {
      int a = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
      while((a + 1) * (a + 2) / 2 <= p) a++;
      while(a * (a + 1) / 2 > p) a--;
      tabAB[p] = (unsigned short)((a << 8) | (p - a * (a + 1) / 2));
    }
  }
  __syncthreads();
  for(int q = q0 + blockIdx.x * LEAF_WARPS + w; q < q1; q += gridDim.x * LEAF_WARPS)
  {
    const DlbLeaf lf = F.leaf[q - q0];
    const int c0 = lf.c0, nc = lf.nc, r = lf.r;
    const int ntri = r * (r + 1) / 2;
    const bool use_tab = r == table_r;
    // ---- metadata: lane ci <-> class ci of the front ----
    DlbClsInfo info = {0, 0, 0, 0};
    if(lane < lf.ncls) info = S.cls_info[F.fcls_list[lf.fcls0 + lane]];
    // exclusive prefix sums over the classes: member pairs, values, loc entries
    int pm = info.nm, pv = info.nm * info.k, pl = info.k;
#pragma unroll
    for(int o = 1; o < 32; o <<= 1)
    {
      const int am = __shfl_up_sync(0xffffffffu, pm, o), av = __shfl_up_sync(0xffffffffu, pv, o), al = __shfl_up_sync(0xffffffffu, pl, o);
      if(lane >= o) { pm += am; pv += av; pl += al; }
    }
    const int npair = __shfl_sync(0xffffffffu, pm, 31);
    pm -= info.nm; pv -= info.nm * info.k; pl -= info.k;
    // lane <-> (class, member) pair: its column position, k, and where its values go
    int my_ci = 0;
    for(int ci = 1; ci < lf.ncls; ci++) if(__shfl_sync(0xffffffffu, pm, ci) <= lane) my_ci = ci;
    const int p_k  = __shfl_sync(0xffffffffu, info.k, my_ci);
    const int p_m  = lane - __shfl_sync(0xffffffffu, pm, my_ci);
    const int p_vo = __shfl_sync(0xffffffffu, pv, my_ci) + p_m * p_k;
    const int p_m0 = __shfl_sync(0xffffffffu, info.m0, my_ci);
    const unsigned int p_pos = lane < npair ? S.mem_pos[p_m0 + p_m] : 0u;
    for(int idx = lane; idx < ntri; idx += 32) T[idx] = 0.0;
    // all values of the front, and the local rows of the class slots
    for(int pi = 0; pi < npair; pi++)
    {
      const unsigned int pos = __shfl_sync(0xffffffffu, p_pos, pi);
      const int k = __shfl_sync(0xffffffffu, p_k, pi), vo = __shfl_sync(0xffffffffu, p_vo, pi);
      if(lane < k) V[vo + lane] = ldg_stream(Jx + pos + lane);
    }
    for(int ci = 0; ci < lf.ncls; ci++)
    {
      const int k = __shfl_sync(0xffffffffu, info.k, ci), r0 = __shfl_sync(0xffffffffu, info.r0, ci), lo = __shfl_sync(0xffffffffu, pl, ci);
      if(lane < k) { const int l = S.cls_loc[r0 + lane]; LOC[lo + lane] = (tri_col(l, r) << 8) | l; }
    }
    __syncwarp();
    // ---- elements: Jt*Jt' of every class into the packed front ----
    for(int ci = 0; ci < lf.ncls; ci++)
    {
      const int k = __shfl_sync(0xffffffffu, info.k, ci), nm = __shfl_sync(0xffffffffu, info.nm, ci);
      const double* Vc = V + __shfl_sync(0xffffffffu, pv, ci);
      const int* loc = LOC + __shfl_sync(0xffffffffu, pl, ci);
      const int npairs = k * (k + 1) / 2;
      for(int p = lane; p < npairs; p += 32)
      {
        const unsigned int ab = tabAB[p];
        const int a = ab >> 8, b = ab & 0xff;
        double g = 0.0;
        for(int m = 0; m < nm; m++) g = fma(Vc[m * k + a], Vc[m * k + b], g);
        const int la = loc[a], lb = loc[b];          // loc[] holds the packed column start in the high bits
        const int ra = la & 0xff, rb = lb & 0xff;
        const int idx = ra > rb ? (lb >> 8) + ra - rb : (la >> 8) + rb - ra;
        T[idx] += g;
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
        if(use_tab && nc == 3)
        { // the bundle-adjustment case: 3 pivot columns, table decode, everything in registers
          const double* L0 = T, *L1 = T + tri_col(1, r) - 1, *L2 = T + tri_col(2, r) - 2;
          for(int idx = o1 + lane; idx < ntri; idx += 32)
          {
            const unsigned int rc = tab[idx];
            const int col = rc >> 8, row = rc & 0xff;
            double v = T[idx];
            v = fma(-L0[row], L0[col], v); v = fma(-L1[row], L1[col], v); v = fma(-L2[row], L2[col], v);
            T[idx] = v;
          }
        }
        else
        {
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
      }
      __syncwarp();
    }
    // ---- write the lower triangle into the front's r x r column-major storage ----
    {
      double* A = fronts + lf.off;
      if(use_tab)
        for(int idx = lane; idx < ntri; idx += 32) A[tabOff[idx]] = T[idx];
      else
      {
        int c = 0, p = lane;
        tri_next(c, p, r);
        for(int idx = lane; idx < ntri; idx += 32)
        {
          A[(size_t)c * r + c + p] = T[idx];
          p += 32; tri_next(c, p, r);
        }
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
  const size_t smem = sizeof(double) * LEAF_WARPS * (size_t)(tri_max + LEAF_VALS + LEAF_LOCS / 2) + sizeof(unsigned short) * (2 * (size_t)tri_max + 528) + 16;
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
    const DlbLeaf lf = F.leaf[q - q0];
    const int c0 = lf.c0, nc = lf.nc, rp = lf.rp, r = lf.r;
    const double* A = fronts + lf.off;
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
    const DlbLeaf lf = F.leaf[q - q0];
    const int c0 = lf.c0, nc = lf.nc, rp = lf.rp, r = lf.r;
    const double* A = fronts + lf.off;
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
