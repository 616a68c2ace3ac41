// dlb_bigsolve.cu -- triangular solves with the large fronts (more rows than fit in shared memory).
//
// Replaces, for those fronts, the one-CTA-per-front level kernels of dlb_front.cu (round 1: the camera
// fronts of a bundle adjustment, up to 1000 x 360 panels = 2.9 MB each, were streamed by ONE CTA: 4.6 ms
// per Gauss-Newton solve for 0.42 GB of factor, 0.06 of the HBM peak). Part of cholmod_solve(CHOLMOD_A) at
// reference dogleg.c:853-856 (and dpotrs at :872-892 for a large dense JtJ). Per front
//   forward   y1 = L11^-1 (P b + children)            k_bs_fwd_tri   one CTA per front (the nc x nc triangle)
//             y2 = (children) - L21 y1                k_bs_fwd_gemv  one CTA per 128-row chunk of L21
//   backward  t  = L21' x2, in 128-row chunks         k_bs_bwd_gemv  one CTA per chunk, partials per chunk
//             x1 = L11^-T (y1 - sum of the partials)  k_bs_bwd_tri   one CTA per front, chunks in order
// so the (r - nc) x nc block below the triangle -- most of the panel -- is streamed by as many CTAs as it
// has 128-row chunks, every load a run of consecutive rows of one column (coalesced, column-major front).
// The triangle itself is walked in blocks of 64 columns with the INVERTED diagonal blocks the factorization leaves
// behind (DlbBigFront::inv_off): two matrix-vector products per block, no substitution. Fronts with more than 1024
// pivot columns (the dense solve types) go in super-blocks of 512 columns: the triangle of a super-block by one
// CTA, everything behind it -- the rest of the triangle included -- by the chunked kernels.
// All sums run in a fixed order (rows ascending within a lane, lanes folded by a fixed shuffle tree,
// chunks ascending): bit-reproducible.
#include "dlb_common.cuh"
#include "dlb_device.h"

#define BS_NT 256
#define BS_CHUNK 128
#define BS_TRI_NT 512           // the triangle kernels: one CTA per front, as many rows in flight as possible


// Ask the L2 for the nc x nc triangle of a front (rows j..nc of column j) before the blocked solve walks through it:
// every block step otherwise starts with a cold DRAM round trip for its diagonal block and another for its columns.
__device__ __forceinline__ void bs_prefetch_triangle(const double* A, int r, int nc, int tid, int nthreads)
{
  const int lane = tid & 31, w = tid >> 5, nw = nthreads >> 5;
  for(int j = w; j < nc; j += nw)
  {
    const char* p = (const char*)(A + j + (size_t)j * r);
    const int nbytes = (nc - j) * 8;
    for(int o = lane * 128; o < nbytes; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(p + o));
  }
}

// ---- forward, the triangle: y1 <- L11^-1 y1, in place in the front's rows of the work vector ----
// y (all r rows) = [P b on the pivot rows | 0] + the children's gathered contributions (already in ywork).
// Blocked by 64 columns, LEFT-looking, with the inverted diagonal blocks the factorization left behind
// (DlbBigFront::inv_off: X_b = L_bb^-1, row-major): per block
//     t_b = y_b - L[block rows, 0:b0] y[0:b0]        all 512 threads: (row of the block) x (eighth of the columns)
//     y_b = X_b t_b                                   a 64 x 64 matrix-vector product, (row) x (eight columns) per thread
// -- no serial substitution (round 2's first version solved 32 x 32 blocks column by column in one warp and then
// updated all remaining rows: two DRAM round trips and 32 dependent shuffle steps per 32 columns).
__global__ void __launch_bounds__(BS_TRI_NT)
k_bs_fwd_tri(DlbFrontDev F, const DlbBigFront* __restrict__ descs, const double* __restrict__ fronts, const double* __restrict__ inv,
             const double* __restrict__ rhs, double* __restrict__ ywork, double* __restrict__ zperm, int nrhs, int cbeg, int cmax)
{
  extern __shared__ double sy[];                       // the columns [cbeg, cend) of this launch
  __shared__ double red[8][64];
  __shared__ double tb[64];
  const DlbBigFront f = descs[blockIdx.x];
  const int r = f.r, nc = f.nc, c0 = f.col0;
  const int rp = F.rows_ptr[f.sn];
  const double* A = fronts + f.off;
  const double* Xall = inv + f.inv_off;
  const int tid = threadIdx.x;
  if(nc <= 1024) bs_prefetch_triangle(A, r, nc, tid, BS_TRI_NT);       // 8 MB at most: a 4096-wide dense front would only flood the L2
  const int il = tid & 63, pp = tid >> 6;              // the block update: row of the block, eighth of the columns
  const int gi = tid >> 3, gp = tid & 7;               // the product with X: row, eight consecutive columns
  // Super-blocks: a launch solves the columns [cbeg, cend) only; the rows behind them (the rest of the triangle and
  // the rows below it) get their update from k_bs_fwd_gemv, many CTAs wide, before the next super-block's launch --
  // one CTA alone would stream the whole triangle of a 4096-wide dense front at the 53 GB/s a single SM can pull.
  if(cbeg >= nc) return;
  const int cend = nc - cbeg < cmax ? nc : cbeg + cmax;
  for(int rh = 0; rh < nrhs; rh++)
  {
    double* yg = ywork + (size_t)rh * F.ytot + rp;
    const bool gathered = F.sg_flag && F.sg_flag[f.sn];      // a large front with children always has them gathered
    if(cbeg == 0)
      for(int i = tid; i < nc; i += BS_TRI_NT)
      { // the right-hand side of ALL pivot rows: the later super-blocks find theirs in the work vector
        const double v = rhs[(size_t)rh * F.n + F.perm[c0 + i]] + (gathered ? yg[i] : 0.0);
        if(i < cend) sy[i] = v; else yg[i] = v;
      }
    else
      for(int i = cbeg + tid; i < cend; i += BS_TRI_NT) sy[i - cbeg] = yg[i];
    __syncthreads();
    for(int b0 = cbeg; b0 < cend; b0 += 64)
    {
      const int nb = cend - b0 < 64 ? cend - b0 : 64;
      double xr[8];
      {
        const double* Xr = Xall + (size_t)(b0 >> 6) * 8192 + gi * 64 + gp * 8;
#pragma unroll
        for(int u = 0; u < 8; u++) xr[u] = Xr[u];
      }
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      {
        const double* Ai = A + b0 + (il < nb ? il : 0);
        const double* sk = sy - cbeg;
        int k = cbeg + pp;
        for(; k + 56 < b0; k += 64)
        { // eight columns in flight per thread
          double l[8];
#pragma unroll
          for(int u = 0; u < 8; u++) l[u] = Ai[(size_t)(k + 8 * u) * r];
#pragma unroll
          for(int u = 0; u < 8; u += 4)
          { a0 = fma(l[u], sk[k + 8 * u], a0); a1 = fma(l[u+1], sk[k + 8 * (u+1)], a1); a2 = fma(l[u+2], sk[k + 8 * (u+2)], a2); a3 = fma(l[u+3], sk[k + 8 * (u+3)], a3); }
        }
        for(; k < b0; k += 8) a0 = fma(Ai[(size_t)k * r], sk[k], a0);
      }
      red[pp][il] = (a0 + a1) + (a2 + a3);
      __syncthreads();
      if(tid < 64)
      {
        double t = 0.0;
        if(tid < nb)
        {
          t = sy[b0 - cbeg + tid];
#pragma unroll
          for(int q = 0; q < 8; q++) t -= red[q][tid];
        }
        tb[tid] = t;
      }
      __syncthreads();
      double acc = 0.0;
#pragma unroll
      for(int u = 0; u < 8; u++) acc = fma(xr[u], tb[gp * 8 + u], acc);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if(gp == 0 && gi < nb) sy[b0 - cbeg + gi] = acc;
      __syncthreads();
    }
    for(int i = cbeg + tid; i < cend; i += BS_TRI_NT) { yg[i] = sy[i - cbeg]; zperm[(size_t)rh * F.n + c0 + i] = sy[i - cbeg]; }
    __syncthreads();
  }
}

// ---- forward, below the triangle: y2[chunk] <- y2[chunk] - L21[chunk, :] y1 ----
__global__ void __launch_bounds__(BS_NT)
k_bs_fwd_gemv(DlbFrontDev F, const DlbBigFront* __restrict__ descs, const double* __restrict__ fronts,
              double* __restrict__ ywork, int nrhs, int cbeg, int cmax)
{
  extern __shared__ double sy[];                       // y of the columns [cbeg, cend)
  __shared__ double part[BS_NT / BS_CHUNK][BS_CHUNK];
  const DlbBigFront f = descs[blockIdx.y];
  const int r = f.r, nc = f.nc;
  if(cbeg >= nc) return;
  const int cend = nc - cbeg < cmax ? nc : cbeg + cmax, ncols = cend - cbeg;
  const int row0 = cend + (int)blockIdx.x * BS_CHUNK;  // every row behind the columns: rest of the triangle + rows below
  if(row0 >= r) return;
  const int rp = F.rows_ptr[f.sn];
  const double* A = fronts + f.off;
  const int tid = threadIdx.x;
  // thread = (row of the chunk, half of the columns): two threads per row, their sums added in a fixed order
  const int rr = tid & (BS_CHUNK - 1), half = tid / BS_CHUNK;
  const int row = row0 + rr;
  const int ca = half == 0 ? 0 : (ncols + 1) / 2, cb = half == 0 ? (ncols + 1) / 2 : ncols;
  for(int rh = 0; rh < nrhs; rh++)
  {
    double* yg = ywork + (size_t)rh * F.ytot + rp;
    const bool gathered = F.sg_flag && F.sg_flag[f.sn];
    for(int i = tid; i < ncols; i += BS_NT) sy[i] = yg[cbeg + i];
    __syncthreads();
    double acc = 0.0;
    if(row < r)
    {
      const double* Ai = A + row + (size_t)cbeg * r;
      int c = ca;
      for(; c + 8 <= cb; c += 8)
      {
        double l[8];
#pragma unroll
        for(int u = 0; u < 8; u++) l[u] = Ai[(size_t)(c + u) * r];
#pragma unroll
        for(int u = 0; u < 8; u++) acc = fma(l[u], sy[c + u], acc);
      }
      for(; c < cb; c++) acc = fma(Ai[(size_t)c * r], sy[c], acc);
    }
    part[half][rr] = acc;
    __syncthreads();
    // pivot rows and rows already touched by an earlier super-block hold a value; otherwise: the gathered children or 0
    if(half == 0 && row < r) yg[row] = ((row < nc || cbeg > 0 || gathered) ? yg[row] : 0.0) - (part[0][rr] + part[1][rr]);
    __syncthreads();
  }
}

// ---- backward, below the triangle: partial[chunk][c] = sum over the chunk's rows of L21[i, c] x2[i] ----
// warp = column (8 columns in flight per CTA), lanes over the chunk's 128 rows (4 each)
__global__ void __launch_bounds__(BS_NT)
k_bs_bwd_gemv(DlbFrontDev F, const DlbBigFront* __restrict__ descs, const double* __restrict__ fronts,
              const double* __restrict__ zperm, double* __restrict__ partial, const long long* __restrict__ part_off, int nrhs,
              int cbeg, int cmax)
{
  __shared__ double sx[BS_CHUNK];
  const DlbBigFront f = descs[blockIdx.y];
  const int r = f.r;
  if(cbeg >= f.nc) return;
  const int cend = f.nc - cbeg < cmax ? f.nc : cbeg + cmax, nc = cend - cbeg;      // nc: the columns of this launch
  const int row0 = cend + (int)blockIdx.x * BS_CHUNK;
  if(row0 >= r) return;
  const int rp = F.rows_ptr[f.sn];
  const int* rows = F.rows + rp;
  const double* A = fronts + f.off + (size_t)cbeg * r;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int nchunk = (r - cend + BS_CHUNK - 1) / BS_CHUNK;
  for(int rh = 0; rh < nrhs; rh++)
  {
    const double* z = zperm + (size_t)rh * F.n;
    for(int i = tid; i < BS_CHUNK; i += BS_NT) sx[i] = row0 + i < r ? z[rows[row0 + i]] : 0.0;
    __syncthreads();
    double* out = partial + part_off[blockIdx.y] + ((size_t)rh * nchunk + blockIdx.x) * nc;
    for(int c0 = w; c0 < nc; c0 += 2 * (BS_NT / 32))
    { // two columns per warp iteration: 8 independent loads in flight per lane
      const int c1 = c0 + BS_NT / 32;
      const double* A0 = A + row0 + (size_t)c0 * r;
      const double* A1 = A + row0 + (size_t)(c1 < nc ? c1 : c0) * r;
      double l0[4], l1[4];
#pragma unroll
      for(int u = 0; u < 4; u++)
      {
        const int i = lane + 32 * u;
        const int ic = row0 + i < r ? i : 0;
        l0[u] = A0[ic]; l1[u] = A1[ic];
      }
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for(int u = 0; u < 4; u++)
      {
        const int i = lane + 32 * u;
        const double xv = row0 + i < r ? sx[i] : 0.0;
        a0 = fma(l0[u], xv, a0); a1 = fma(l1[u], xv, a1);
      }
      a0 = warp_sum(a0); a1 = warp_sum(a1);
      if(lane == 0) { out[c0] = a0; if(c1 < nc) out[c1] = a1; }
    }
    __syncthreads();
  }
}

// ---- backward, the triangle: x1 <- L11^-T (y1 - sum over the chunks of partial) ----
// Blocks of 64 columns from the last to the first: s_b = t_b - L[rows behind the block, block]' x[behind] (a warp per
// column, lanes over the rows), then x_b = X_b' s_b with the column-major copy of the inverted diagonal block.
__global__ void __launch_bounds__(BS_TRI_NT)
k_bs_bwd_tri(DlbFrontDev F, const DlbBigFront* __restrict__ descs, const double* __restrict__ fronts, const double* __restrict__ inv,
             double* __restrict__ zperm, const double* __restrict__ partial, const long long* __restrict__ part_off, int nrhs,
             int cbeg, int cmax)
{
  extern __shared__ double sx[];                       // the columns [cbeg, cend) of this launch
  __shared__ double tb[64];
  const DlbBigFront f = descs[blockIdx.x];
  const int r = f.r, nc = f.nc, c0 = f.col0;
  if(cbeg >= nc) return;
  const int cend = nc - cbeg < cmax ? nc : cbeg + cmax, ncols = cend - cbeg;
  const double* A = fronts + f.off;
  const double* Xall = inv + f.inv_off;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int nchunk = (r - cend + BS_CHUNK - 1) / BS_CHUNK;    // k_bs_bwd_gemv covered every row behind the columns
  if(nc <= 1024) bs_prefetch_triangle(A, r, nc, tid, BS_TRI_NT);
  const int gi = tid >> 3, gp = tid & 7;
  for(int rh = 0; rh < nrhs; rh++)
  {
    double* z = zperm + (size_t)rh * F.n;
    const double* pin = partial + part_off[blockIdx.x] + (size_t)rh * nchunk * ncols;
    for(int i = tid; i < ncols; i += BS_TRI_NT)
    {
      double t = 0.0;
      for(int ch = 0; ch < nchunk; ch++) t += pin[(size_t)ch * ncols + i];
      sx[i] = z[c0 + cbeg + i] - t;
    }
    __syncthreads();
    const double* sk = sx - cbeg;
    for(int b0 = cbeg + ((ncols - 1) / 64) * 64; b0 >= cbeg; b0 -= 64)
    {
      const int nb = cend - b0 < 64 ? cend - b0 : 64;
      double xc[8];
      {
        const double* Xc = Xall + (size_t)(b0 >> 6) * 8192 + 4096 + gi * 64 + gp * 8;
#pragma unroll
        for(int u = 0; u < 8; u++) xc[u] = Xc[u];
      }
      // the block's columns against the pivots already solved behind the block (inside this launch's columns):
      // warp w takes the columns w, w+16, w+32, w+48, their four sums in flight together
      {
        const int i0 = b0 + nb;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        const double* Ac[4];
#pragma unroll
        for(int q = 0; q < 4; q++) Ac[q] = A + (size_t)(b0 + (w + 16 * q < nb ? w + 16 * q : 0)) * r;
        int i = i0 + lane;
        for(; i + 32 < cend; i += 64)
        { // 8 loads in flight per lane
          double l[8];
#pragma unroll
          for(int q = 0; q < 4; q++) { l[q] = Ac[q][i]; l[4 + q] = Ac[q][i + 32]; }
          const double xv = sk[i], xw = sk[i + 32];
#pragma unroll
          for(int q = 0; q < 4; q++) acc[q] = fma(l[4 + q], xw, fma(l[q], xv, acc[q]));
        }
        for(; i < cend; i += 32)
        {
          const double xv = sk[i];
#pragma unroll
          for(int q = 0; q < 4; q++) acc[q] = fma(Ac[q][i], xv, acc[q]);
        }
#pragma unroll
        for(int o = 16; o > 0; o >>= 1)
#pragma unroll
          for(int q = 0; q < 4; q++) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
        if(lane == 0)
#pragma unroll
          for(int q = 0; q < 4; q++)
          {
            const int c = w + 16 * q;
            tb[c] = c < nb ? sk[b0 + c] - acc[q] : 0.0;
          }
      }
      __syncthreads();
      double a = 0.0;
#pragma unroll
      for(int u = 0; u < 8; u++) a = fma(xc[u], tb[gp * 8 + u], a);
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      a += __shfl_xor_sync(0xffffffffu, a, 4);
      __syncthreads();                                   // every warp has read sx[behind] / tb before the block is overwritten
      if(gp == 0 && gi < nb) sx[b0 - cbeg + gi] = a;
      __syncthreads();
    }
    for(int i = tid; i < ncols; i += BS_TRI_NT) z[c0 + cbeg + i] = sx[i];
    __syncthreads();
  }
}

// ------------------------------------------------------------------ launchers
// descs: the large fronts of one level (device array), max_r / max_nc: maxima over them.
// partial / part_off: scratch for the backward partial sums; front f of the level uses
// partial[part_off[f] ...] with nrhs * nchunk(f) * nc(f) doubles.
#define BS_SUPER 512            // columns per launch for fronts with more than 2 * BS_SUPER pivot columns (a 792-column
                                // root of a bundle adjustment is still faster in one piece: 87 against 130 us)
static inline int bs_super(int max_nc) { return max_nc > 2 * BS_SUPER ? BS_SUPER : max_nc; }
void dlb_launch_bigsolve_fwd(const DlbFrontDev& F, const DlbBigFront* d_descs, int nfronts, int max_r, int max_nc,
                             const double* fronts, const double* inv, const double* rhs, double* ywork, double* zperm, int nrhs,
                             cudaStream_t st)
{
  if(nfronts <= 0) return;
  static DlbPerDeviceOnce attr_once;
  if(attr_once.first())
  {
    cudaFuncSetAttribute(k_bs_fwd_tri, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    cudaFuncSetAttribute(k_bs_fwd_gemv, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  }
  const int cmax = bs_super(max_nc);
  const size_t smem = sizeof(double) * (size_t)cmax;
  for(int s0 = 0; s0 < max_nc; s0 += cmax)
  {
    k_bs_fwd_tri<<<nfronts, BS_TRI_NT, smem, st>>>(F, d_descs, fronts, inv, rhs, ywork, zperm, nrhs, s0, cmax);
    const int nch = (max_r - s0 - 1 + BS_CHUNK - 1) / BS_CHUNK;      // upper bound over the batch: the rows behind column s0
    if(nch > 0)
      for(int f0 = 0; f0 < nfronts; f0 += 65535)
        k_bs_fwd_gemv<<<dim3(nch, nfronts - f0 < 65535 ? nfronts - f0 : 65535), BS_NT, smem, st>>>(F, d_descs + f0, fronts, ywork, nrhs, s0, cmax);
  }
}
void dlb_launch_bigsolve_bwd(const DlbFrontDev& F, const DlbBigFront* d_descs, int nfronts, int max_r, int max_nc,
                             const double* fronts, const double* inv, double* zperm, double* partial, const long long* d_part_off,
                             int nrhs, cudaStream_t st)
{
  if(nfronts <= 0) return;
  static DlbPerDeviceOnce attr_once;
  if(attr_once.first()) cudaFuncSetAttribute(k_bs_bwd_tri, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  const int cmax = bs_super(max_nc);
  for(int s0 = ((max_nc - 1) / cmax) * cmax; s0 >= 0; s0 -= cmax)
  {
    const int nch = (max_r - s0 - 1 + BS_CHUNK - 1) / BS_CHUNK;
    if(nch > 0)
      for(int f0 = 0; f0 < nfronts; f0 += 65535)
        k_bs_bwd_gemv<<<dim3(nch, nfronts - f0 < 65535 ? nfronts - f0 : 65535), BS_NT, 0, st>>>(F, d_descs + f0, fronts, zperm, partial,
                                                                                                d_part_off + f0, nrhs, s0, cmax);
    k_bs_bwd_tri<<<nfronts, BS_TRI_NT, sizeof(double) * (size_t)cmax, st>>>(F, d_descs, fronts, inv, zperm, partial, d_part_off, nrhs, s0, cmax);
  }
}
// doubles of backward-solve scratch a front needs per right-hand side, whatever the super-block size of its level
long long dlb_bigsolve_partial_size(int r, int nc)
{
  long long need = (long long)((r - nc + BS_CHUNK - 1) / BS_CHUNK) * nc;
  for(int s0 = 0; s0 < nc; s0 += BS_SUPER)
  {
    const int cend = nc - s0 < BS_SUPER ? nc : s0 + BS_SUPER;
    need = std::max(need, (long long)((r - cend + BS_CHUNK - 1) / BS_CHUNK) * (cend - s0));
  }
  return need;
}
