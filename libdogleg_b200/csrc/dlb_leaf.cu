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
#include <algorithm>

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
  static DlbPerDeviceOnce attr_once;
  if(attr_once.first())
  {
    cudaFuncSetAttribute(k_leaf_fronts, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  }
  long long g = ((long long)(q1 - q0) + LEAF_WARPS - 1) / LEAF_WARPS;
  const long long cap = (long long)sm_count * 16;
  if(g > cap) g = cap;
  // the decode table is built for the largest front size (in a bundle adjustment: the size of nearly all)
  k_leaf_fronts<<<(int)g, 32 * LEAF_WARPS, smem, st>>>(F, S, q0, q1, Jx, fronts, lambda, minor, tri_max, max_rows, eliminate);
}

// ------------------------------------------------------------------ solves
// Leaf fronts have at most 8 pivot columns and 48 rows: lane i owns rows i and i+32, the whole pivot panel sits in
// registers (all its loads are issued at once, from clamped addresses, before the first dependent operation), the
// pivots are inverted up front, and the record of the next front is fetched while this one is solved (two
// alternating register sets, as in k_leaf_fronts_mma).
template<int NC> struct LeafPanel { double a0[NC], a1[NC], inv[NC]; };
template<int NC>
__device__ __forceinline__ void leaf_panel_load(LeafPanel<NC>& P, const double* __restrict__ A, int r, int nc, int lane)
{
  const int i0 = lane < r ? lane : r - 1, i1 = lane + 32 < r ? lane + 32 : r - 1;
  double d[NC];
#pragma unroll
  for(int j = 0; j < NC; j++)
  {
    const size_t jc = (size_t)(j < nc ? j : nc - 1);
    P.a0[j] = A[i0 + jc * r]; P.a1[j] = A[i1 + jc * r]; d[j] = A[jc + jc * r];
  }
#pragma unroll
  for(int j = 0; j < NC; j++) P.inv[j] = 1.0 / d[j];
}
// forward: y = L^-1 P b for leaf fronts (no children); NC = 4 (every leaf has <= 4 pivots) or 8
template<int NC>
__global__ void __launch_bounds__(256, NC == 4 ? 4 : 2)
k_leaf_solve_fwd(DlbFrontDev F, int q0, int q1, const double* __restrict__ fronts,
                 const double* __restrict__ rhs, double* __restrict__ ywork, double* __restrict__ zperm, int nrhs)
{
  const int lane = threadIdx.x & 31;
  const int wg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = gridDim.x * (blockDim.x >> 5);
  auto fetch = [&](DlbLeaf& L, int qq) { L = F.leaf[(qq < q1 ? qq : q1 - 1) - q0]; };
  auto body = [&](const DlbLeaf& lf) {
    const int c0 = lf.c0, nc = lf.nc, rp = lf.rp, r = lf.r;
    const int pc = F.perm[c0 + (lane < nc ? lane : nc - 1)];
    LeafPanel<NC> P;
    leaf_panel_load(P, fronts + lf.off, r, nc, lane);
    for(int rh = 0; rh < nrhs; rh++)
    {
      const double b = rhs[(size_t)rh * F.n + pc];
      double y0 = lane < nc ? b : 0.0, y1 = 0.0;
#pragma unroll
      for(int j = 0; j < NC; j++)
        if(j < nc)
        {
          const double yj = __shfl_sync(0xffffffffu, y0, j) * P.inv[j];
          if(lane == j) y0 = yj;
          if(lane > j && lane < r) y0 = fma(-P.a0[j], yj, y0);
          if(lane + 32 < r)        y1 = fma(-P.a1[j], yj, y1);
        }
      double* yg = ywork + (size_t)rh * F.ytot + rp;
      if(lane < r) yg[lane] = y0;
      if(lane + 32 < r) yg[lane + 32] = y1;
      if(lane < nc) zperm[(size_t)rh * F.n + c0 + lane] = y0;
    }
  };
  int q = q0 + wg;
  DlbLeaf LA, LB;
  if(q < q1) fetch(LA, q);
  while(q < q1)
  {
    fetch(LB, q + nw); body(LA); q += nw;
    if(q >= q1) break;
    fetch(LA, q + nw); body(LB); q += nw;
  }
}
// backward: x = L^-T y in place in zperm. The sums over the rows below the pivots do not depend on the pivots'
// own solution: all nc of them are reduced first (independent shuffle trees), the nc x nc triangle follows.
template<int NC>
__global__ void __launch_bounds__(256, NC == 4 ? 4 : 2)
k_leaf_solve_bwd(DlbFrontDev F, int q0, int q1, const double* __restrict__ fronts, double* __restrict__ zperm, int nrhs)
{
  const int lane = threadIdx.x & 31;
  const int wg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = gridDim.x * (blockDim.x >> 5);
  auto fetch = [&](DlbLeaf& L, int qq) { L = F.leaf[(qq < q1 ? qq : q1 - 1) - q0]; };
  auto body = [&](const DlbLeaf& lf) {
    const int c0 = lf.c0, nc = lf.nc, rp = lf.rp, r = lf.r;
    const int* rows = F.rows + rp;
    const int g0 = rows[lane < r ? lane : r - 1], g1 = rows[lane + 32 < r ? lane + 32 : r - 1];
    LeafPanel<NC> P;
    leaf_panel_load(P, fronts + lf.off, r, nc, lane);
    for(int rh = 0; rh < nrhs; rh++)
    {
      double* z = zperm + (size_t)rh * F.n;
      const double z0 = z[g0], z1 = z[g1];
      double x0 = lane < r ? z0 : 0.0;
      const double x1 = lane + 32 < r ? z1 : 0.0;
      double below[NC];
#pragma unroll
      for(int j = 0; j < NC; j++)
      {
        double acc = (lane >= nc && lane < r) ? P.a0[j] * x0 : 0.0;
        if(lane + 32 < r) acc = fma(P.a1[j], x1, acc);
        below[j] = acc;
      }
#pragma unroll
      for(int o = 16; o > 0; o >>= 1)
#pragma unroll
        for(int j = 0; j < NC; j++) below[j] += __shfl_xor_sync(0xffffffffu, below[j], o);
#pragma unroll
      for(int j = NC - 1; j >= 0; j--)
        if(j < nc)
        {
          // rows j+1 .. nc-1 of column j: lanes inside the triangle, already solved
          double acc = (lane > j && lane < nc) ? P.a0[j] * x0 : 0.0;
#pragma unroll
          for(int o = 4; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);      // nc <= 8: lanes 0..7
          const double xj = (__shfl_sync(0xffffffffu, x0, j) - below[j] - __shfl_sync(0xffffffffu, acc, 0)) * P.inv[j];
          if(lane == j) x0 = xj;
        }
      if(lane < nc) z[c0 + lane] = x0;
    }
  };
  int q = q0 + wg;
  DlbLeaf LA, LB;
  if(q < q1) fetch(LA, q);
  while(q < q1)
  {
    fetch(LB, q + nw); body(LA); q += nw;
    if(q >= q1) break;
    fetch(LA, q + nw); body(LB); q += nw;
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
  if(F.leaf_max_nc <= 4) k_leaf_solve_fwd<4><<<leaf_solve_grid(q1 - q0, sm_count), 256, 0, st>>>(F, q0, q1, fronts, rhs, ywork, zperm, nrhs);
  else                   k_leaf_solve_fwd<8><<<leaf_solve_grid(q1 - q0, sm_count), 256, 0, st>>>(F, q0, q1, fronts, rhs, ywork, zperm, nrhs);
}
void dlb_launch_leaf_solve_bwd(const DlbFrontDev& F, int q0, int q1, const double* fronts, double* zperm,
                               int nrhs, int sm_count, cudaStream_t st)
{
  if(q1 <= q0) return;
  if(F.leaf_max_nc <= 4) k_leaf_solve_bwd<4><<<leaf_solve_grid(q1 - q0, sm_count), 256, 0, st>>>(F, q0, q1, fronts, zperm, nrhs);
  else                   k_leaf_solve_bwd<8><<<leaf_solve_grid(q1 - q0, sm_count), 256, 0, st>>>(F, q0, q1, fronts, zperm, nrhs);
}

// ------------------------------------------------------------------ tensor-core leaf fronts
// The same work as k_leaf_fronts for fronts with at most 4 pivot columns and at most 32
// measurement columns (every point of a bundle adjustment), on the FP64 tensor cores:
//   F = V' V with V = (measurement columns) x (front rows), dense in shared memory (zeros where a
//   column does not touch a row): NT(NT+1)/2 lower 8x8 tiles, one DMMA m8n8k4 per tile and 4
//   columns -- the A fragment of V' and the B fragment of V are the same register;
//   the pivot panel (first tile column) goes to shared memory, is factorized there (3 columns x
//   39 rows: scalar), and the trailing update F22 -= L21 L21' is ONE more DMMA per tile (K = 4
//   covers the <= 4 pivots). ~45 DMMAs per point instead of ~3000 scalar FMA lane-iterations
//   and their index arithmetic (the scalar kernel is issue bound at 4.8 k warp-instructions per front).
// row stride KS of the V' tile: the smallest 8j+4 >= the number of measurement columns (4, 12, 20, 28,
// 36): lanes (g, tt) then read 32 different 8-byte words of 16 double-banks -> two wavefronts, the minimum
__device__ __forceinline__ void leaf_dmma(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

#ifndef LEAF_MMA_MINB
#define LEAF_MMA_MINB 4
#endif
template<int NT>
__global__ void __launch_bounds__(32 * LEAF_WARPS, LEAF_MMA_MINB)
k_leaf_fronts_mma(DlbFrontDev F, DlbSparseDev S, int q0, int q1, const double* __restrict__ Jx,
                  double* __restrict__ fronts, double lambda, long long* minor, int eliminate, int LEAF_KS)
{
  constexpr int R8 = 8 * NT, NPAIR = NT * (NT + 1) / 2;
  extern __shared__ double sh_leaf[];           // per warp: V' (R8 x LEAF_KS), panel (R8 x 4), loc (LEAF_LOCS ints)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int g = lane >> 2, tt = lane & 3;
  const int PER_WARP = R8 * LEAF_KS + R8 * 4 + LEAF_LOCS / 2;
  double* Vt = sh_leaf + (size_t)w * PER_WARP;
  double* Pn = Vt + R8 * LEAF_KS;
  int* LOC = (int*)(Pn + R8 * 4);
  // The per-leaf record (DlbLeaf + flat pair / slot tables, dlb_device.h) of the NEXT front is fetched while this
  // one is computed: of the chain record -> Jacobian values -> compute -> store only the values load stays exposed
  // (the chain leaf -> class list -> class info -> member positions -> values cost five DRAM round trips per front).
  const int ps = F.leaf_ps, lw = F.leaf_lw;
  const int stride = gridDim.x * LEAF_WARPS;
  int q = q0 + blockIdx.x * LEAF_WARPS + w;
  // two record buffers used alternately (a loop-carried copy made the compiler wait for the prefetch right away)
  struct LeafRec { DlbLeaf lf; unsigned int pos, kl, loc; };
  auto fetch = [&](LeafRec& R, int qq) {
    const size_t i = (size_t)((qq < q1 ? qq : q1 - 1) - q0);          // clamped: unconditional loads
    R.lf = F.leaf[i];
    R.pos = F.leaf_pos[i * ps + (lane < ps ? lane : 0)]; R.kl = F.leaf_kl[i * ps + (lane < ps ? lane : 0)];
    R.loc = F.leaf_loc[i * lw + (lane < lw ? lane : 0)];
  };
  unsigned char* LOCB = (unsigned char*)LOC;
  auto body = [&](const LeafRec& R) {
    const DlbLeaf& lf = R.lf;
    const unsigned int p_pos = R.pos, kl = lane < ps ? R.kl : 0u, locw = R.loc;
    const int c0 = lf.c0, nc = lf.nc, r = lf.r;
    const int p_k = (int)(kl & 255u), p_lo = (int)(kl >> 8);
    const int npair = __popc(__ballot_sync(0xffffffffu, p_k > 0));
    const int kcols = (npair + 3) & ~3;                          // measurement columns, padded to the DMMA K
    for(int idx = lane; idx < R8 * LEAF_KS; idx += 32) Vt[idx] = 0.0;
    if(lane < lw) ((unsigned int*)LOCB)[lane] = locw;
    __syncwarp();
    // scatter the Jacobian values: V'[local row of the slot][measurement column]
    // (eight columns at a time: all their loads are issued before the first value is stored -- a load followed by
    // its own store per column made every column wait for the previous one's DRAM round trip)
    for(int pi0 = 0; pi0 < npair; pi0 += 8)
    {
      double v[8]; int dst[8];
#pragma unroll
      for(int u = 0; u < 8; u++)
      {
        const int pi = pi0 + u < npair ? pi0 + u : npair - 1;
        const unsigned int pos = __shfl_sync(0xffffffffu, p_pos, pi);
        const int k = __shfl_sync(0xffffffffu, p_k, pi), lo = __shfl_sync(0xffffffffu, p_lo, pi);
        const bool on = pi0 + u < npair && lane < k;
        v[u] = ldg_stream(Jx + pos + (lane < k ? lane : 0));
        dst[u] = on ? LOCB[lo + lane] * LEAF_KS + pi : -1;
      }
#pragma unroll
      for(int u = 0; u < 8; u++) if(dst[u] >= 0) Vt[dst[u]] = v[u];
    }
    __syncwarp();
    // ---- F = V' V on the tensor cores ----
    double acc[NPAIR][2];
#pragma unroll
    for(int i = 0; i < NPAIR; i++) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
    for(int m = 0; m < kcols; m += 4)
    {
      double v[NT];
#pragma unroll
      for(int ti = 0; ti < NT; ti++) v[ti] = Vt[(8 * ti + g) * LEAF_KS + m + tt];
      int idx = 0;
#pragma unroll
      for(int ti = 0; ti < NT; ti++)
#pragma unroll
        for(int tj = 0; tj <= ti; tj++, idx++) leaf_dmma(acc[idx][0], acc[idx][1], v[ti], v[tj]);
    }
    if(eliminate)
    {
      // ---- pivot panel: first tile column, columns 0..3, to shared memory (+ lambda) ----
      {
        int idx = 0;
#pragma unroll
        for(int ti = 0; ti < NT; ti++)
#pragma unroll
          for(int tj = 0; tj <= ti; tj++, idx++)
            if(tj == 0 && tt < 2)
            {
              const int row = 8 * ti + g, col = 2 * tt;
              Pn[row * 4 + col]     = acc[idx][0] + (row == col && col < nc ? lambda : 0.0);
              Pn[row * 4 + col + 1] = acc[idx][1] + (row == col + 1 && col + 1 < nc ? lambda : 0.0);
            }
      }
      __syncwarp();
      // left-looking factorization of the nc <= 4 panel columns (rows 0..r-1, two rows per lane)
      bool failed = false;
      for(int j = 0; j < nc; j++)
      {
        for(int i = j + lane; i < r; i += 32)
        {
          double vv = Pn[i * 4 + j];
          for(int jj = 0; jj < j; jj++) vv = fma(-Pn[i * 4 + jj], Pn[j * 4 + jj], vv);
          Pn[i * 4 + j] = vv;
        }
        __syncwarp();
        const double d = Pn[j * 4 + j];
        if(!(d > 0.0) || isinf(d)) { if(lane == 0) atomicMin(minor, (long long)(c0 + j)); failed = true; break; }
        const double rs = rsqrt(d);
        __syncwarp();
        for(int i = j + lane; i < r; i += 32) Pn[i * 4 + j] = (i == j) ? d * rs : Pn[i * 4 + j] * rs;
        __syncwarp();
      }
      // columns >= nc of the panel array and rows above the diagonal are zero for the update below
      for(int idx = lane; idx < R8 * 4; idx += 32)
      {
        const int row = idx >> 2, col = idx & 3;
        if(col >= nc || row < col || row >= r) Pn[idx] = 0.0;
      }
      __syncwarp();
      if(!failed)
      { // F22 -= L21 L21': one DMMA per tile (rows of the pivot block themselves give garbage
        // in the pivot columns, which are not written from the accumulators)
        double lv[NT];
#pragma unroll
        for(int ti = 0; ti < NT; ti++) lv[ti] = Pn[(8 * ti + g) * 4 + tt];
        int idx = 0;
#pragma unroll
        for(int ti = 0; ti < NT; ti++)
#pragma unroll
          for(int tj = 0; tj <= ti; tj++, idx++) leaf_dmma(acc[idx][0], acc[idx][1], -lv[ti], lv[tj]);
      }
    }
    // ---- write: pivot columns from the panel, the rest of the lower triangle from the tiles ----
    double* A = fronts + lf.off;
    const int npiv = eliminate ? nc : 0;
    for(int idx = lane; idx < npiv * r; idx += 32)
    {
      const int col = idx / r, row = idx - col * r;
      if(row >= col) A[(size_t)col * r + row] = Pn[row * 4 + col];
    }
    {
      int idx = 0;
#pragma unroll
      for(int ti = 0; ti < NT; ti++)
#pragma unroll
        for(int tj = 0; tj <= ti; tj++, idx++)
        {
          const int row = 8 * ti + g, col = 8 * tj + 2 * tt;
          if(row < r)
          {
            if(col >= npiv && col <= row)         A[(size_t)col * r + row]       = acc[idx][0];
            if(col + 1 >= npiv && col + 1 <= row) A[(size_t)(col + 1) * r + row] = acc[idx][1];
          }
        }
    }
    // Complete the 32-byte sectors the lower triangle only partly covers: the slots above the diagonal of column j that
    // share a sector with its first element (rows j-lead .. j-1) and those of column j+1 behind its last element
    // (rows 0 .. tail-1, above the diagonal while tail <= j+1) are never read by anybody -- zeros there let the L2
    // write whole sectors back instead of fetching each of them from DRAM first to merge (9.5 GB of traffic for
    // 6.9 GB of payload in the bundle-adjustment config). The front's own first and last sector are left alone:
    // their other slots belong to the neighbouring fronts.
    for(int j = lane; j < r; j += 32)
    {
      const size_t first = (size_t)lf.off + (size_t)j * r + j, last = (size_t)lf.off + (size_t)j * r + r - 1;
      const int lead = (int)(first & 3), tail = 3 - (int)(last & 3);
      if(j > 0)      for(int u = 1; u <= lead && u <= j; u++) fronts[first - u] = 0.0;
      if(j + 1 < r)  for(int u = 1; u <= tail && u <= j + 1; u++) fronts[last + u] = 0.0;
    }
    __syncwarp();
  };
  LeafRec RA, RB;
  if(q < q1) fetch(RA, q);
  while(q < q1)
  {
    fetch(RB, q + stride); body(RA); q += stride;
    if(q >= q1) break;
    fetch(RA, q + stride); body(RB); q += stride;
  }
}

template<int NT>
static void launch_leaf_mma(const DlbFrontDev& F, const DlbSparseDev& S, int q0, int q1, const double* Jx,
                            double* fronts, double lambda, long long* minor, int eliminate, int max_pairs, int sm_count, cudaStream_t st)
{
  const int kc = std::max(4, (max_pairs + 3) & ~3);            // measurement columns padded to the DMMA K
  const int LEAF_KS = 8 * ((kc - 4 + 7) / 8) + 4;                // smallest 8j+4 >= kc
  const size_t smem = sizeof(double) * LEAF_WARPS * (size_t)(8 * NT * LEAF_KS + 8 * NT * 4 + LEAF_LOCS / 2);
  static DlbPerDeviceOnce attr_once;
  if(attr_once.first())
  {
    cudaFuncSetAttribute(k_leaf_fronts_mma<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  }
  long long g = ((long long)(q1 - q0) + LEAF_WARPS - 1) / LEAF_WARPS;
  const long long cap = (long long)sm_count * 16;
  if(g > cap) g = cap;
  k_leaf_fronts_mma<NT><<<(int)g, 32 * LEAF_WARPS, smem, st>>>(F, S, q0, q1, Jx, fronts, lambda, minor, eliminate, LEAF_KS);
}
void dlb_launch_leaf_fronts_mma(const DlbFrontDev& F, const DlbSparseDev& S, int q0, int q1, const double* Jx,
                                double* fronts, double lambda, long long* minor, int max_rows, int max_pairs, int eliminate,
                                int sm_count, cudaStream_t st)
{
  if(q1 <= q0) return;
  const int nt = (max_rows + 7) / 8;
  if(nt <= 2)      launch_leaf_mma<2>(F, S, q0, q1, Jx, fronts, lambda, minor, eliminate, max_pairs, sm_count, st);
  else if(nt == 3) launch_leaf_mma<3>(F, S, q0, q1, Jx, fronts, lambda, minor, eliminate, max_pairs, sm_count, st);
  else if(nt == 4) launch_leaf_mma<4>(F, S, q0, q1, Jx, fronts, lambda, minor, eliminate, max_pairs, sm_count, st);
  else if(nt == 5) launch_leaf_mma<5>(F, S, q0, q1, Jx, fronts, lambda, minor, eliminate, max_pairs, sm_count, st);
  else             launch_leaf_mma<6>(F, S, q0, q1, Jx, fronts, lambda, minor, eliminate, max_pairs, sm_count, st);
}
