// dlb_dense.cu -- dense-Jacobian kernels (DOGLEG_DENSE, DOGLEG_DENSE_PRODUCTS).
//
// J is M x N row-first exactly as the callback wrote it (reference dogleg.h:67).
// Replaces mul_matrix_t_densevector + norm2 (reference dogleg.c:284-292, 1045-1048),
// norm2_mul_matrix_vector (:293-306), the rank-1 JtJ build (:214-220, 709-723),
// mul_xt_Apacked_upper_x / mul_xt_A_x (:309-347) and the memcpy+lambda of the
// products path (:739-770). The factorization itself is the single-front case
// of dlb_front.cu.
#include "dlb_common.cuh"
#include "dlb_device.h"

// ------------------------------------------------------------- J' x and |x|^2
// CTA b owns rows [b*rows_per, ...); thread (tx = column lane, ty = row lane).
// work[b*N + k] = partial J'x, work2[b] = partial |x|^2
__global__ void __launch_bounds__(DLB_NT)
k_dense_grad(const double* __restrict__ J, const double* __restrict__ x, int M, int N, int rows_per,
             double* __restrict__ work, double* __restrict__ work2)
{
  __shared__ double sh[8][33];
  __shared__ double shr[32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r0 = blockIdx.x * rows_per;
  const int r1 = min(M, r0 + rows_per);
  for(int k0 = 0; k0 < N; k0 += 32)
  {
    const int k = k0 + tx;
    double acc = 0.0;
    if(k < N)
      for(int i = r0 + ty; i < r1; i += 8) acc = fma(ldg_stream(J + (size_t)i * N + k), x[i], acc);
    sh[ty][tx] = acc;
    __syncthreads();
    if(ty == 0 && k < N)
    {
      double s = 0.0;
#pragma unroll
      for(int u = 0; u < 8; u++) s += sh[u][tx];
      work[(size_t)blockIdx.x * N + k] = s;
    }
    __syncthreads();
  }
  double n2 = 0.0;
  for(int i = r0 + threadIdx.x; i < r1; i += DLB_NT) n2 = fma(x[i], x[i], n2);
  n2 = block_sum(n2, shr);
  if(threadIdx.x == 0) work2[blockIdx.x] = n2;
}
__global__ void __launch_bounds__(DLB_NT)
k_dense_grad_reduce(const double* __restrict__ work, const double* __restrict__ work2, int nblk, int N,
                    double* __restrict__ Jtx, double* part, unsigned int* counter, DlbScalars* sc)
{
  double g2 = 0.0, gmax = 0.0, n2 = 0.0;
  for(int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x)
  {
    double s = 0.0;
    for(int b = 0; b < nblk; b++) s += work[(size_t)b * N + k];
    Jtx[k] = s; g2 = fma(s, s, g2); gmax = fmax(gmax, fabs(s));
  }
  for(int b = blockIdx.x * blockDim.x + threadIdx.x; b < nblk; b += gridDim.x * blockDim.x) n2 += work2[b];
  double out[5];
  if(grid_reduce5(n2, g2, 0, 0, gmax, part, counter, out))
  { sc->norm2_x = out[0]; sc->norm2_Jtx = out[1]; sc->maxabs_Jtx = out[4]; }
}

static inline int dense_row_blocks(int M, int sm_count)
{
  int g = (M + 63) / 64;
  const int cap = sm_count * 4;
  return g < 1 ? 1 : (g > cap ? cap : g);
}

void dlb_launch_dense_grad(const double* J, const double* x, int M, int N, double* Jtx,
                           double* work, double* part, unsigned int* counter, DlbScalars* sc,
                           int sm_count, cudaStream_t st)
{
  const int nblk = dense_row_blocks(M, sm_count);
  const int rows_per = (M + nblk - 1) / nblk;
  double* work2 = work + (size_t)nblk * N;
  k_dense_grad<<<nblk, DLB_NT, 0, st>>>(J, x, M, N, rows_per, work, work2);
  int g = (N + DLB_NT - 1) / DLB_NT; if(g > sm_count) g = sm_count;
  k_dense_grad_reduce<<<g, DLB_NT, 0, st>>>(work, work2, nblk, N, Jtx, part, counter, sc);
}

// ------------------------------------------------------------------ |J v|^2
__global__ void __launch_bounds__(DLB_NT)
k_dense_jv(const double* __restrict__ J, const double* __restrict__ v, int M, int N, int rows_per,
           double* __restrict__ work)
{
  __shared__ double shr[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r0 = blockIdx.x * rows_per;
  const int r1 = min(M, r0 + rows_per);
  double total = 0.0;
  for(int i = r0 + w; i < r1; i += DLB_NT / 32)
  {
    double d = 0.0;
    for(int k = lane; k < N; k += 32) d = fma(ldg_stream(J + (size_t)i * N + k), v[k], d);
    d = warp_sum_all(d);
    total = fma(d, d, total);
  }
  // every lane of a warp holds the same total: count it once
  total = block_sum(lane == 0 ? total : 0.0, shr);
  if(threadIdx.x == 0) work[blockIdx.x] = total;
}
__global__ void __launch_bounds__(DLB_NT)
k_sum_partials_dense(const double* __restrict__ src, int n, double* dst)
{
  __shared__ double sh[32];
  double s = 0.0;
  for(int i = threadIdx.x; i < n; i += blockDim.x) s += src[i];
  s = block_sum(s, sh);
  if(threadIdx.x == 0) *dst = s;
}
void dlb_launch_dense_jv(const double* J, const double* v, int M, int N, double* work,
                         double* dst, int sm_count, cudaStream_t st)
{
  const int nblk = dense_row_blocks(M, sm_count);
  const int rows_per = (M + nblk - 1) / nblk;
  k_dense_jv<<<nblk, DLB_NT, 0, st>>>(J, v, M, N, rows_per, work);
  k_sum_partials_dense<<<1, DLB_NT, 0, st>>>(work, nblk, dst);
}

// ------------------------------------------------------------------ J' J
// 64x64 output tile per CTA, 4x4 register micro-tile per thread, 16 rows of J
// per shared-memory stage, split over row slices when there are few tiles.
#define SY_T 64
#define SY_K 16
__global__ void __launch_bounds__(256)
k_dense_syrk(const double* __restrict__ J, int M, int N, int ntile, int rows_per_slice,
             double* __restrict__ out, size_t slice_stride, int direct)
{
  __shared__ double As[SY_K][SY_T + 1];
  __shared__ double Bs[SY_K][SY_T + 1];
  // lower-triangular tile index -> (ti, tj), ti >= tj
  int t = blockIdx.x, ti = 0;
  while((ti + 1) * (ti + 2) / 2 <= t) ti++;
  const int tj = t - ti * (ti + 1) / 2;
  (void)ntile;
  const int slice = blockIdx.y;
  const int m0 = slice * rows_per_slice, m1 = min(M, m0 + rows_per_slice);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4];
#pragma unroll
  for(int a = 0; a < 4; a++)
#pragma unroll
    for(int b = 0; b < 4; b++) acc[a][b] = 0.0;

  for(int mb = m0; mb < m1; mb += SY_K)
  {
    for(int idx = threadIdx.x; idx < SY_K * SY_T; idx += 256)
    {
      const int kk = idx / SY_T, cc = idx - kk * SY_T;
      const int row = mb + kk;
      const int ca = ti * SY_T + cc, cb = tj * SY_T + cc;
      As[kk][cc] = (row < m1 && ca < N) ? J[(size_t)row * N + ca] : 0.0;
      Bs[kk][cc] = (row < m1 && cb < N) ? J[(size_t)row * N + cb] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for(int kk = 0; kk < SY_K; kk++)
    {
      double a[4], b[4];
#pragma unroll
      for(int u = 0; u < 4; u++) { a[u] = As[kk][ty * 4 + u]; b[u] = Bs[kk][tx * 4 + u]; }
#pragma unroll
      for(int u = 0; u < 4; u++)
#pragma unroll
        for(int w = 0; w < 4; w++) acc[u][w] = fma(a[u], b[w], acc[u][w]);
    }
    __syncthreads();
  }
  double* dst = out + (direct ? 0 : slice * slice_stride);
#pragma unroll
  for(int u = 0; u < 4; u++)
#pragma unroll
    for(int w = 0; w < 4; w++)
    {
      const int i = ti * SY_T + ty * 4 + u, j = tj * SY_T + tx * 4 + w;
      if(i < N && j < N && i >= j) dst[i + (size_t)j * N] = acc[u][w];
    }
}
__global__ void __launch_bounds__(DLB_NT)
k_syrk_reduce(const double* __restrict__ work, int nslice, size_t slice_stride, int N, double* __restrict__ front)
{
  const size_t total = (size_t)N * N;
  for(size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
  {
    const int j = (int)(idx / N), i = (int)(idx - (size_t)j * N);
    if(i < j) continue;
    double s = 0.0;
    for(int sl = 0; sl < nslice; sl++) s += work[sl * slice_stride + idx];
    front[idx] = s;
  }
}
// big matrices go to the tensor-core kernel of dlb_bigfront.cu (128x128 tiles)
void dlb_launch_dense_syrk_dmma(const double* J, int M, int N, int ntile, int nslice, int rows_per_slice,
                                double* dst, size_t slice_stride, int direct, cudaStream_t st);
#define SY_DMMA_MIN_N 96
static inline void syrk_plan(int M, int N, int sm_count, int& ntile, int& nslice, int& rows_per)
{
  const int T = N >= SY_DMMA_MIN_N ? 128 : SY_T;
  const int nt = (N + T - 1) / T;
  ntile = nt * (nt + 1) / 2;
  nslice = (2 * sm_count) / ntile;
  if(ntile > sm_count && N >= SY_DMMA_MIN_N)
  { // more tiles than SMs (one 128x128 DMMA CTA per SM): cut the rows into 1..8 slices so that the
    // number of CTAs fills whole waves (N = 4096: 528 tiles = 3.57 waves -> 7 slices = 24.97 waves)
    double best = 0.0;
    for(int ns = 1; ns <= 8; ns++)
    {
      const long long total = (long long)ntile * ns, waves = (total + sm_count - 1) / sm_count;
      const double eff = (double)total / (double)(waves * sm_count);
      if(eff > best + 0.02) { best = eff; nslice = ns; }
    }
  }
  const int max_by_rows = (M + 255) / 256;
  if(nslice > max_by_rows) nslice = max_by_rows;
  if(nslice < 1) nslice = 1;
  rows_per = (M + nslice - 1) / nslice;
  rows_per = ((rows_per + SY_K - 1) / SY_K) * SY_K;
  nslice = (M + rows_per - 1) / rows_per;
  if(nslice < 1) nslice = 1;
}
size_t dlb_dense_syrk_work_size(int M, int N, int sm_count)
{
  int ntile, nslice, rows_per;
  syrk_plan(M, N, sm_count, ntile, nslice, rows_per);
  return nslice > 1 ? (size_t)nslice * N * N : 0;
}
void dlb_launch_dense_syrk(const double* J, int M, int N, double* front, double* work,
                           int sm_count, cudaStream_t st)
{
  int ntile, nslice, rows_per;
  syrk_plan(M, N, sm_count, ntile, nslice, rows_per);
  const size_t stride = (size_t)N * N;
  const bool dmma = N >= SY_DMMA_MIN_N;
  if(nslice == 1)
  {
    if(dmma) dlb_launch_dense_syrk_dmma(J, M, N, ntile, 1, rows_per, front, stride, 1, st);
    else     k_dense_syrk<<<dim3(ntile, 1), 256, 0, st>>>(J, M, N, ntile, rows_per, front, stride, 1);
  }
  else
  {
    if(dmma) dlb_launch_dense_syrk_dmma(J, M, N, ntile, nslice, rows_per, work, stride, 0, st);
    else     k_dense_syrk<<<dim3(ntile, nslice), 256, 0, st>>>(J, M, N, ntile, rows_per, work, stride, 0);
    int g = (int)((stride + DLB_NT - 1) / DLB_NT); if(g > sm_count * 8) g = sm_count * 8;
    k_syrk_reduce<<<g, DLB_NT, 0, st>>>(work, nslice, stride, N, front);
  }
}

// -------------------------------------------------- dense-products layouts
// user JtJ layouts (reference dogleg.h:122-128): packed upper row-first,
// packed lower row-first, or full row-first
__device__ __forceinline__ size_t sym_index(int i, int j, int N, int packed, int upper)
{ // i >= j
  if(!packed) return (size_t)i * N + j;
  if(upper)   return (size_t)j * N - (size_t)j * (j - 1) / 2 + (i - j);
  return (size_t)i * (i + 1) / 2 + j;
}
__global__ void k_products_to_front(const double* __restrict__ JtJ, int N, int packed, int upper,
                                    double* __restrict__ front)
{
  const size_t total = (size_t)N * N;
  for(size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
  {
    const int j = (int)(idx / N), i = (int)(idx - (size_t)j * N);
    if(i >= j) front[idx] = JtJ[sym_index(i, j, N, packed, upper)];
  }
}
void dlb_launch_products_to_front(const double* JtJ, int N, int packed, int upper, double* front, cudaStream_t st)
{
  size_t total = (size_t)N * N;
  int g = (int)((total + 255) / 256); if(g > 1184) g = 1184;
  k_products_to_front<<<g, 256, 0, st>>>(JtJ, N, packed, upper, front);
}

// v' A v with A in the user's layout; one warp per row, fixed-order reduction
__global__ void __launch_bounds__(DLB_NT)
k_products_xAx(const double* __restrict__ A, int N, int packed, int upper,
               const double* __restrict__ v, double* __restrict__ work)
{
  __shared__ double shr[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double total = 0.0;
  for(int i = blockIdx.x * (DLB_NT / 32) + w; i < N; i += gridDim.x * (DLB_NT / 32))
  {
    double s = 0.0;
    if(!packed)
      for(int j = lane; j < N; j += 32) s = fma(A[(size_t)i * N + j] * v[i], v[j], s);
    else
      for(int j = i + lane; j < N; j += 32)
      {
        const double a = A[sym_index(j, i, N, packed, upper)];
        s += (j == i) ? a * v[i] * v[i] : 2. * a * v[j] * v[i];
      }
    total += warp_sum_all(s);
  }
  total = block_sum(lane == 0 ? total : 0.0, shr);
  if(threadIdx.x == 0) work[blockIdx.x] = total;
}
void dlb_launch_products_xAx(const double* JtJ, int N, int packed, int upper, const double* v,
                             double* dst, cudaStream_t st)
{
  // dst doubles as scratch: partials live right behind it (engine allocates 1+1024 doubles)
  int g = (N + 7) / 8; if(g > 1024) g = 1024;
  k_products_xAx<<<g, DLB_NT, 0, st>>>(JtJ, N, packed, upper, v, dst + 1);
  k_sum_partials_dense<<<1, DLB_NT, 0, st>>>(dst + 1, g, dst);
}

// front (L, column-major lower) -> what dpptrf/dpotrf would have left in
// ctx->factorization_dense: packed: the same triangle the user layout names;
// full: fortran 'L' of the row-first array, i.e. element (row j, col i>=j)
__global__ void k_front_to_reference(const double* __restrict__ front, int N, int packed, int upper,
                                     double* __restrict__ out)
{
  const size_t total = (size_t)N * N;
  for(size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
  {
    const int j = (int)(idx / N), i = (int)(idx - (size_t)j * N);
    if(i < j) continue;
    const double l = front[idx];
    if(!packed)     out[(size_t)j * N + i] = l;
    else if(upper)  out[(size_t)j * N - (size_t)j * (j - 1) / 2 + (i - j)] = l;
    else            out[(size_t)i * (i + 1) / 2 + j] = l;
  }
}
void dlb_launch_front_to_reference_layout(const double* front, int N, int packed, int upper, double* out, cudaStream_t st)
{
  size_t total = (size_t)N * N;
  int g = (int)((total + 255) / 256); if(g > 1184) g = 1184;
  k_front_to_reference<<<g, 256, 0, st>>>(front, N, packed, upper, out);
}
