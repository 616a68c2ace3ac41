// dlb_bigfront.cu -- FP64 tensor-core kernels for matrices that do not fit in shared memory:
//   * blocked right-looking Cholesky of the pivot columns of a large front (the dense solve
//     types with Nstate > 158 are the single-front case: this is the dpotrf/dpptrf replacement of
//     reference dogleg.c:779-804 at scale, config C5; large supernodes of the sparse
//     factorization use the same code)
//   * J'J for a large dense Jacobian (reference dogleg.c:709-714 does it as Nmeas rank-1 updates)
// All matrix-matrix work is mma.sync.m8n8k4.f64 (SASS DMMA) fed from shared memory tiles whose
// row stride (== 4 mod 16 doubles) makes the fragment loads bank-conflict free. tcgen05 has no
// FP64 kind, so DMMA is the tensor path for this workload on sm_100a.
#include "dlb_common.cuh"
#include "dlb_device.h"

#define BF_NB 64                 // pivot block width
#define BF_LDS 68                // shared-memory row stride of 64-wide tiles (68 mod 16 == 4)

__device__ __forceinline__ void bf_dmma(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// A batch of large fronts (all the large fronts of one elimination-tree level) is processed
// together: blockIdx.y selects the front, every kernel handles pivot block 'step' of each
// front that still has one.
//
// ---- 1. Cholesky of the nb x nb diagonal block at (k0,k0), one CTA, in shared memory ----
// Blocked by 8 columns: an 8x8 diagonal factorization by one warp, an 8-column panel solve with a
// thread per row, a rank-8 trailing update by all threads: 3 barriers per 8 columns.
__global__ void __launch_bounds__(256)
k_bf_potrf(const DlbBigFront* __restrict__ descs, double* __restrict__ fronts, int step, long long* minor)
{
  const DlbBigFront f = descs[blockIdx.y];
  const int k0 = step * BF_NB;
  if(k0 >= f.nc) return;
  const int nb = f.nc - k0 < BF_NB ? f.nc - k0 : BF_NB;
  const int ld = f.r;
  double* A = fronts + f.off;
  __shared__ double T[BF_NB][BF_NB + 1];
  __shared__ int fail_col;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if(tid == 0) fail_col = -1;
  for(int idx = tid; idx < nb * nb; idx += 256)
  {
    const int j = idx / nb, i = idx - j * nb;
    T[i][j] = i >= j ? A[(size_t)(k0 + j) * ld + k0 + i] : 0.0;
  }
  __syncthreads();
  for(int b0 = 0; b0 < nb; b0 += 8)
  {
    const int bw = nb - b0 < 8 ? nb - b0 : 8;
    if(w == 0)
    { // 8x8 diagonal block, right-looking, one warp
      for(int j = b0; j < b0 + bw; j++)
      {
        const double d = T[j][j];
        if(!(d > 0.0) || isinf(d)) { if(lane == 0 && fail_col < 0) fail_col = j; break; }
        // reciprocal square root + multiplications: FP64 sqrt followed by a division is the
        // longest dependent chain of the whole factorization (<= 1.5 ulp instead of 1)
        const double rs = rsqrt(d);
        __syncwarp();
        if(lane == 0) T[j][j] = d * rs;
        const int i = j + 1 + lane;
        if(i < b0 + bw) T[i][j] *= rs;
        __syncwarp();
        // lane <-> (row, col) of the remaining lower triangle inside the 8x8 block (<= 28 pairs)
        const int rem = b0 + bw - j - 1;
        if(lane < rem * (rem + 1) / 2)
        {
          int a = 0; while((a + 1) * (a + 2) / 2 <= lane) a++;
          const int c = lane - a * (a + 1) / 2;
          T[j + 1 + a][j + 1 + c] = fma(-T[j + 1 + a][j], T[j + 1 + c][j], T[j + 1 + a][j + 1 + c]);
        }
        __syncwarp();
      }
    }
    __syncthreads();
    if(fail_col >= 0) break;
    // panel: rows below the 8x8 block, forward substitution with its 8 columns
    {
      const int i = b0 + bw + tid;
      if(i < nb)
      {
        double x[8], rd[8];
#pragma unroll
        for(int c = 0; c < 8; c++) rd[c] = c < bw ? 1.0 / T[b0 + c][b0 + c] : 0.0;
#pragma unroll
        for(int c = 0; c < 8; c++)
          if(c < bw)
          {
            double v = T[i][b0 + c];
#pragma unroll
            for(int cp = 0; cp < c; cp++) v = fma(-x[cp], T[b0 + c][b0 + cp], v);
            x[c] = v * rd[c];
            T[i][b0 + c] = x[c];
          }
      }
    }
    __syncthreads();
    // trailing update of the rest of the diagonal block
    {
      const int t0 = b0 + bw, wd = nb - t0;
      for(int idx = tid; idx < wd * wd; idx += 256)
      {
        const int cc = idx / wd, ii = idx - cc * wd;
        if(ii < cc) continue;
        double acc = T[t0 + ii][t0 + cc];
#pragma unroll
        for(int c = 0; c < 8; c++) if(c < bw) acc = fma(-T[t0 + ii][b0 + c], T[t0 + cc][b0 + c], acc);
        T[t0 + ii][t0 + cc] = acc;
      }
    }
    __syncthreads();
  }
  if(fail_col >= 0)
  {
    if(tid == 0) atomicMin(minor, (long long)(f.col0 + k0 + fail_col));
    return;
  }
  for(int idx = tid; idx < nb * nb; idx += 256)
  {
    const int j = idx / nb, i = idx - j * nb;
    if(i >= j) A[(size_t)(k0 + j) * ld + k0 + i] = T[i][j];
  }
}

// ---- 2. panel solve: rows below the diagonal block, X L_kk' = B, one thread per row ----
// The 8x8 diagonal blocks of L_kk are inverted first (one warp each), so that the 64-step
// dependent chain of a plain substitution becomes 8 steps of 8 independent dot products.
__global__ void __launch_bounds__(64)
k_bf_trsm(const DlbBigFront* __restrict__ descs, double* __restrict__ fronts, int step)
{
  const DlbBigFront f = descs[blockIdx.y];
  const int k0 = step * BF_NB;
  if(k0 >= f.nc) return;
  const int nb = f.nc - k0 < BF_NB ? f.nc - k0 : BF_NB;
  const int ld = f.r, r = f.r;
  if(k0 + nb + (int)blockIdx.x * 64 >= r) return;
  double* A = fronts + f.off;
  __shared__ double L[BF_NB][BF_NB + 1];
  __shared__ double Dinv[8][8][9];
  const int tid = threadIdx.x;
  for(int idx = tid; idx < BF_NB * BF_NB; idx += 64)
  {
    const int j = idx / BF_NB, i = idx - j * BF_NB;
    L[i][j] = (i >= j && i < nb) ? A[(size_t)(k0 + j) * ld + k0 + i] : (i == j ? 1.0 : 0.0);
  }
  __syncthreads();
  { // thread (b, c) computes column c of the inverse of diagonal block b by forward substitution
    const int b = tid >> 3, c = tid & 7;
    double col[8], rd[8];
#pragma unroll
    for(int i = 0; i < 8; i++) rd[i] = 1.0 / L[8 * b + i][8 * b + i];
#pragma unroll
    for(int i = 0; i < 8; i++)
    {
      double v = i == c ? 1.0 : 0.0;
#pragma unroll
      for(int p = 0; p < i; p++) v = fma(-L[8 * b + i][8 * b + p], col[p], v);
      col[i] = v * rd[i];
    }
#pragma unroll
    for(int i = 0; i < 8; i++) Dinv[b][i][c] = col[i];
  }
  __syncthreads();
  const int row = k0 + nb + blockIdx.x * 64 + tid;
  if(row >= r) return;
  double x[BF_NB];
#pragma unroll
  for(int c = 0; c < BF_NB; c++) x[c] = c < nb ? A[(size_t)(k0 + c) * ld + row] : 0.0;
#pragma unroll
  for(int b = 0; b < 8; b++)
  {
    if(8 * b >= nb) break;
    double t[8];
#pragma unroll
    for(int c = 0; c < 8; c++)
    {
      double v = x[8 * b + c];
#pragma unroll
      for(int cp = 0; cp < 8 * b; cp++) v = fma(-x[cp], L[8 * b + c][cp], v);
      t[c] = v;
    }
    // x_b = t * inv(L_bb)' : x[c] = sum_{p <= c} t[p] * Dinv[c][p]
#pragma unroll
    for(int c = 0; c < 8; c++)
    {
      double v = 0.0;
#pragma unroll
      for(int p = 0; p <= c; p++) v = fma(t[p], Dinv[b][c][p], v);
      x[8 * b + c] = v;
    }
  }
#pragma unroll
  for(int c = 0; c < BF_NB; c++) if(c < nb) A[(size_t)(k0 + c) * ld + row] = x[c];
}

// ---- 3. trailing update C -= P P' on the tensor cores: one 64x64 lower tile per CTA ----
// P = A[k0+nb .. r, k0 .. k0+nb) (the panel just solved); tile (ti,tj) covers rows
// t0+64ti.., columns t0+64tj.. with t0 = k0+nb. Warp w owns the 8 rows 8w..8w+7 of the tile.
__global__ void __launch_bounds__(256)
k_bf_syrk_update(const DlbBigFront* __restrict__ descs, double* __restrict__ fronts, int step)
{
  extern __shared__ double sm_p[];
  const DlbBigFront f = descs[blockIdx.y];
  const int k0 = step * BF_NB;
  if(k0 >= f.nc) return;
  const int nb = f.nc - k0 < BF_NB ? f.nc - k0 : BF_NB;
  const int ld = f.r, r = f.r;
  double* A = fronts + f.off;
  double* Pi = sm_p;
  double* Pj = sm_p + 64 * BF_LDS;
  int t = blockIdx.x, ti = 0;
  {
    ti = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while((ti + 1) * (ti + 2) / 2 <= t) ti++;
    while(ti * (ti + 1) / 2 > t) ti--;
  }
  const int tj = t - ti * (ti + 1) / 2;
  const int t0 = k0 + nb;
  const int i0 = t0 + 64 * ti, j0 = t0 + 64 * tj;
  if(i0 >= r) return;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for(int idx = tid; idx < 64 * BF_NB; idx += 256)
  {
    const int kk = idx / 64, rr = idx - kk * 64;
    Pi[rr * BF_LDS + kk] = (kk < nb && i0 + rr < r) ? A[(size_t)(k0 + kk) * ld + i0 + rr] : 0.0;
    Pj[rr * BF_LDS + kk] = (kk < nb && j0 + rr < r) ? A[(size_t)(k0 + kk) * ld + j0 + rr] : 0.0;
  }
  __syncthreads();
  const int g = lane >> 2, tt = lane & 3;
  double acc[8][2];
#pragma unroll
  for(int c = 0; c < 8; c++) { acc[c][0] = 0.0; acc[c][1] = 0.0; }
  for(int k = 0; k < BF_NB; k += 4)
  {
    if(k >= nb) break;
    const double a = Pi[(8 * w + g) * BF_LDS + k + tt];
#pragma unroll
    for(int c = 0; c < 8; c++) bf_dmma(acc[c][0], acc[c][1], a, Pj[(8 * c + g) * BF_LDS + k + tt]);
  }
  const int row = i0 + 8 * w + g;
  if(row < r)
#pragma unroll
    for(int c = 0; c < 8; c++)
    {
      const int col = j0 + 8 * c + 2 * tt;
      if(col <= row && col < r)         A[(size_t)col * ld + row]       -= acc[c][0];
      if(col + 1 <= row && col + 1 < r) A[(size_t)(col + 1) * ld + row] -= acc[c][1];
    }
}

// Partial Cholesky of the first nc columns of every front of a batch (r x r column-major lower,
// ld = r): afterwards the first nc columns hold L, the trailing block holds the update matrix.
// descs: device array; max_nc / max_r: maxima over the batch (host-side copies of the shapes).
void dlb_bigfront_factor_batch(const DlbBigFront* d_descs, int nfronts, int max_r, int max_nc, double* fronts,
                               long long* minor, cudaStream_t st, double* n_launch)
{
  if(nfronts <= 0) return;
  const size_t bf_smem = sizeof(double) * 2 * 64 * BF_LDS;
  static DlbPerDeviceOnce attr_once;
  if(attr_once.first())
  {
    cudaFuncSetAttribute(k_bf_syrk_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bf_smem);
  }
  const int nsteps = (max_nc + BF_NB - 1) / BF_NB;
  for(int step = 0; step < nsteps; step++)
  {
    k_bf_potrf<<<dim3(1, nfronts), 256, 0, st>>>(d_descs, fronts, step, minor);
    if(n_launch) *n_launch += 1;
    const int below = max_r - step * BF_NB - 1;     // an upper bound over the batch (nb >= 1)
    if(below <= 0) continue;
    k_bf_trsm<<<dim3((below + 63) / 64, nfronts), 64, 0, st>>>(d_descs, fronts, step);
    const int nt = (below + 63) / 64;
    k_bf_syrk_update<<<dim3(nt * (nt + 1) / 2, nfronts), 256, bf_smem, st>>>(d_descs, fronts, step);
    if(n_launch) *n_launch += 2;
  }
}

// ---------------------------------------------------------------------------------- J'J
// front (N x N column-major lower) = J'J for a row-first M x N Jacobian: 128x128 output tile per
// CTA (only tiles on or below the diagonal), rows of J streamed through shared memory 16 at a
// time, double buffered; warp w owns tile rows 16w..16w+15 (2 fragments) x all 128 columns.
#define SJ_T 128
#define SJ_K 32
#define SJ_LDS 132               // 132 mod 16 == 4
__global__ void __launch_bounds__(256)
k_dense_syrk_dmma(const double* __restrict__ J, int M, int N, int rows_per_slice,
                  double* __restrict__ out, size_t slice_stride, int direct)
{
  extern __shared__ double sm[];
  double* As[2] = { sm, sm + 2 * SJ_K * SJ_LDS };
  double* Bs[2] = { sm + SJ_K * SJ_LDS, sm + 3 * SJ_K * SJ_LDS };
  int t = blockIdx.x, ti = 0;
  while((ti + 1) * (ti + 2) / 2 <= t) ti++;
  const int tj = t - ti * (ti + 1) / 2;
  const int slice = blockIdx.y;
  const int m0 = slice * rows_per_slice, m1 = min(M, m0 + rows_per_slice);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, tt = lane & 3;
  const int ca0 = ti * SJ_T, cb0 = tj * SJ_T;
  double acc[2][16][2];
#pragma unroll
  for(int a = 0; a < 2; a++)
#pragma unroll
    for(int c = 0; c < 16; c++) { acc[a][c][0] = 0.0; acc[a][c][1] = 0.0; }

  // asynchronous global -> shared copies (cp.async, 8 bytes, zero-filled outside the matrix)
  auto load_stage = [&](int buf, int mb) {
    for(int idx = tid; idx < SJ_K * SJ_T; idx += 256)
    {
      const int kk = idx / SJ_T, cc = idx - kk * SJ_T;
      const int row = mb + kk;
      const bool rv = row < m1;
      const bool va = rv && ca0 + cc < N, vb = rv && cb0 + cc < N;
      const double* srca = J + (va ? (size_t)row * N + ca0 + cc : 0);
      const double* srcb = J + (vb ? (size_t)row * N + cb0 + cc : 0);
      const unsigned da = (unsigned)__cvta_generic_to_shared(&As[buf][kk * SJ_LDS + cc]);
      const unsigned db = (unsigned)__cvta_generic_to_shared(&Bs[buf][kk * SJ_LDS + cc]);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(da), "l"(srca), "r"(va ? 8 : 0));
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(db), "l"(srcb), "r"(vb ? 8 : 0));
    }
    asm volatile("cp.async.commit_group;");
  };
  int buf = 0;
  load_stage(0, m0);
  asm volatile("cp.async.wait_group 0;");
  __syncthreads();
  for(int mb = m0; mb < m1; mb += SJ_K)
  {
    if(mb + SJ_K < m1) load_stage(buf ^ 1, mb + SJ_K);
    const double* Ab = As[buf];
    const double* Bb = Bs[buf];
#pragma unroll
    for(int k = 0; k < SJ_K; k += 4)
    {
      const double a0 = Ab[(k + tt) * SJ_LDS + 16 * w + g];
      const double a1 = Ab[(k + tt) * SJ_LDS + 16 * w + 8 + g];
#pragma unroll
      for(int c = 0; c < 16; c++)
      {
        const double b = Bb[(k + tt) * SJ_LDS + 8 * c + g];
        bf_dmma(acc[0][c][0], acc[0][c][1], a0, b);
        bf_dmma(acc[1][c][0], acc[1][c][1], a1, b);
      }
    }
    asm volatile("cp.async.wait_group 0;");
    __syncthreads();
    buf ^= 1;
  }
  double* dst = out + (direct ? 0 : slice * slice_stride);
#pragma unroll
  for(int a = 0; a < 2; a++)
#pragma unroll
    for(int c = 0; c < 16; c++)
    {
      const int i = ca0 + 16 * w + 8 * a + g;
      const int j = cb0 + 8 * c + 2 * tt;
      if(i < N)
      {
        if(j < N && j <= i)         dst[i + (size_t)j * N]       = acc[a][c][0];
        if(j + 1 < N && j + 1 <= i) dst[i + (size_t)(j + 1) * N] = acc[a][c][1];
      }
    }
}

size_t dlb_dense_syrk_dmma_smem() { return sizeof(double) * 4 * SJ_K * SJ_LDS; }

void dlb_launch_dense_syrk_dmma(const double* J, int M, int N, int ntile, int nslice, int rows_per_slice,
                                double* dst, size_t slice_stride, int direct, cudaStream_t st)
{
  static DlbPerDeviceOnce attr_once;
  if(attr_once.first())
  {
    cudaFuncSetAttribute(k_dense_syrk_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dlb_dense_syrk_dmma_smem());
  }
  k_dense_syrk_dmma<<<dim3(ntile, nslice), 256, dlb_dense_syrk_dmma_smem(), st>>>(J, M, N, rows_per_slice, dst, slice_stride, direct);
}
