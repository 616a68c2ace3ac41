// tests/support/problems_dev.cu -- device-resident version of the synthetic model
// in problems.c (same formulas, same splitmix64-generated data uploaded from the
// host problem), used as the "user callback" of dogleg_gpu_optimize_sparse /
// _dense / _dense_batched in bench.py and the GPU tests. Not product code.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "problems.h"

// device copy of problems.h's splitmix64 counter hash (bit-identical)
__host__ __device__ static inline unsigned long long dev_splitmix64(unsigned long long z)
{
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ static inline double dev_uniform(unsigned long long seed, unsigned long long stream, unsigned long long idx)
{
  unsigned long long h = dev_splitmix64(dev_splitmix64(seed * 0x100000001B3ull + stream) ^ idx);
  return ((double)(h >> 11) + 0.5) * (2.0 / 9007199254740992.0) - 1.0;
}
#include "dogleg_gpu.h"

struct dlb_dev_problem
{
  int N, M; long long nnz;
  int* d_Ap; int* d_Ai; double* d_Ax; double* d_b;     // sparse
  double* d_Adense;                                    // dense (M x N) or batched (B x M x N)
  int B;
  float ms_total; int ncalls;
  cudaEvent_t e0, e1; int timing;
};

__device__ __forceinline__ double phi(double p)  { return p + 0.1 * p * p * p; }
__device__ __forceinline__ double dphi(double p) { return 1.0 + 0.3 * p * p; }

// A CTA takes MODEL_COLS consecutive measurement columns: their nonzeros are one contiguous
// range of the CCS arrays, so values/indices are read and the Jacobian written fully coalesced
// (four independent entries per thread in flight); the products A_q phi(p_k) go through shared
// memory and one thread per column sums them in entry order (the same order as the host callback
// in problems.c).
#define MODEL_COLS 256
#define MODEL_SMEM 6144
__global__ void __launch_bounds__(256)
k_model_sparse(int M, const int* __restrict__ Ap, const int* __restrict__ Ai,
               const double* __restrict__ Ax, const double* __restrict__ b,
               const double* __restrict__ p, double* __restrict__ x, double* __restrict__ Jx)
{
  __shared__ double prod[MODEL_SMEM];
  for(int j0 = blockIdx.x * MODEL_COLS; j0 < M; j0 += gridDim.x * MODEL_COLS)
  {
    const int j1 = min(M, j0 + MODEL_COLS);
    const int q0 = Ap[j0], q1 = Ap[j1];
    const bool fits = q1 - q0 <= MODEL_SMEM;
    for(int qb = q0 + threadIdx.x; qb < q1; qb += 4 * 256)
    {
      int k[4]; double a[4], pk[4];
#pragma unroll
      for(int u = 0; u < 4; u++) { const int q = qb + 256 * u < q1 ? qb + 256 * u : qb; k[u] = Ai[q]; a[u] = Ax[q]; }
#pragma unroll
      for(int u = 0; u < 4; u++) pk[u] = p[k[u]];
#pragma unroll
      for(int u = 0; u < 4; u++)
      {
        const int q = qb + 256 * u;
        if(q < q1)
        {
          Jx[q] = a[u] * dphi(pk[u]);
          if(fits) prod[q - q0] = a[u] * phi(pk[u]);
        }
      }
    }
    __syncthreads();
    for(int j = j0 + threadIdx.x; j < j1; j += 256)
    {
      double s = 0.0;
      const int c0 = Ap[j], c1 = Ap[j+1];
      if(fits) for(int q = c0; q < c1; q++) s += prod[q - q0];
      else     for(int q = c0; q < c1; q++) s += Ax[q] * phi(p[Ai[q]]);
      x[j] = s - b[j];
    }
    __syncthreads();
  }
}

// dense (also the batched layout: rows = B*M, p indexed per problem): one thread per Jacobian
// entry, fully coalesced; the row sum runs over sub-warp groups of `width` lanes (width = the
// power of two >= N, <= 32) with a shuffle tree
__global__ void k_model_dense(long long rows, int M, int N, int width, const double* __restrict__ A,
                              const double* __restrict__ b, const double* __restrict__ p,
                              const int* __restrict__ active,
                              double* __restrict__ x, double* __restrict__ J)
{
  const int per_warp = 32 / width;
  const int lane = threadIdx.x & 31;
  const int sub = lane / width, k = lane % width;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for(long long i0 = warp * per_warp; i0 < rows; i0 += nwarps * per_warp)
  {
    const long long i = i0 + sub;
    const long long prob = i < rows ? i / M : 0;
    const bool on = i < rows && (!active || active[prob]);
    double s = 0.0;
    for(int kk = k; kk < N; kk += width)
      if(on)
      {
        const double a = A[i * N + kk], pk = p[prob * N + kk];
        J[i * N + kk] = a * dphi(pk);
        s += a * phi(pk);
      }
    for(int o = width >> 1; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, width);
    if(on && k == 0) x[i] = s - b[i];
  }
}

// N == 16 (the batched config): four threads per row, four consecutive entries each as two 16-byte accesses,
// two rows per thread in flight -- the generic kernel moves one 8-byte entry per thread and iteration
__global__ void __launch_bounds__(256)
k_model_dense16(long long rows, int M, const double* __restrict__ A, const double* __restrict__ b,
                const double* __restrict__ p, const int* __restrict__ active, double* __restrict__ x, double* __restrict__ J)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
  const int q = (int)(t & 3);
  for(long long i0 = t >> 2; i0 < rows; i0 += 2 * (nt >> 2))
  {
    double2 a[2][2], pv[2][2]; bool on[2]; long long ii[2];
#pragma unroll
    for(int u = 0; u < 2; u++)
    {
      const long long i = i0 + u * (nt >> 2);
      ii[u] = i;
      const long long ic = i < rows ? i : rows - 1;
      const long long prob = ic / M;
      on[u] = i < rows && (!active || active[prob]);
      const double2* Ar = (const double2*)(A + ic * 16 + 4 * q);
      const double2* pr = (const double2*)(p + prob * 16 + 4 * q);
      a[u][0] = Ar[0]; a[u][1] = Ar[1]; pv[u][0] = pr[0]; pv[u][1] = pr[1];
    }
#pragma unroll
    for(int u = 0; u < 2; u++)
    {
      double s = a[u][0].x * phi(pv[u][0].x) + a[u][0].y * phi(pv[u][0].y) + a[u][1].x * phi(pv[u][1].x) + a[u][1].y * phi(pv[u][1].y);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if(on[u])
      {
        double2* Jr = (double2*)(J + ii[u] * 16 + 4 * q);
        Jr[0] = make_double2(a[u][0].x * dphi(pv[u][0].x), a[u][0].y * dphi(pv[u][0].y));
        Jr[1] = make_double2(a[u][1].x * dphi(pv[u][1].x), a[u][1].y * dphi(pv[u][1].y));
        if(q == 0) x[ii[u]] = s - b[ii[u]];
      }
    }
  }
}

static int model_width(int N) { int w = 1; while(w < N && w < 32) w <<= 1; return w; }

extern "C" dlb_dev_problem* dlb_dev_problem_create(const dlb_problem* P)
{
  dlb_dev_problem* D = (dlb_dev_problem*)calloc(1, sizeof(*D));
  D->N = P->N; D->M = P->M; D->nnz = P->nnz; D->B = 1;
  cudaMalloc(&D->d_b, sizeof(double) * (size_t)P->M);
  cudaMemcpy(D->d_b, P->b, sizeof(double) * (size_t)P->M, cudaMemcpyHostToDevice);
  if(P->Adense)
  {
    cudaMalloc(&D->d_Adense, sizeof(double) * (size_t)P->M * P->N);
    cudaMemcpy(D->d_Adense, P->Adense, sizeof(double) * (size_t)P->M * P->N, cudaMemcpyHostToDevice);
  }
  else
  {
    cudaMalloc(&D->d_Ap, sizeof(int) * ((size_t)P->M + 1));
    cudaMalloc(&D->d_Ai, sizeof(int) * (size_t)P->nnz);
    cudaMalloc(&D->d_Ax, sizeof(double) * (size_t)P->nnz);
    cudaMemcpy(D->d_Ap, P->Ap, sizeof(int) * ((size_t)P->M + 1), cudaMemcpyHostToDevice);
    cudaMemcpy(D->d_Ai, P->Ai, sizeof(int) * (size_t)P->nnz, cudaMemcpyHostToDevice);
    cudaMemcpy(D->d_Ax, P->Ax, sizeof(double) * (size_t)P->nnz, cudaMemcpyHostToDevice);
  }
  cudaEventCreate(&D->e0); cudaEventCreate(&D->e1);
  if(cudaDeviceSynchronize() != cudaSuccess) { free(D); return NULL; }
  return D;
}

// B independent dense problems generated on the device: A (B x M x N), b (B x M), p0/p_true (B x N)
__global__ void k_gen_batched(int B, int M, int N, unsigned long long seed,
                              double* A, double* b, double* p0)
{
  const long long total = (long long)B * M;
  for(long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
  {
    const long long prob = i / M;
    const unsigned long long sd = seed + (unsigned long long)prob;
    double s = 0.0;
    for(int k = 0; k < N; k++)
    {
      const double a = dev_uniform(sd, 1, (unsigned long long)((i - prob * M) * N + k));
      A[i * N + k] = a;
      s += a * phi(dev_uniform(sd, 2, k));
    }
    b[i] = s + 0.01 * dev_uniform(sd, 4, (unsigned long long)(i - prob * M));
  }
  for(long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)B * N; i += (long long)gridDim.x * blockDim.x)
  {
    const long long prob = i / N; const int k = (int)(i - prob * N);
    const unsigned long long sd = seed + (unsigned long long)prob;
    p0[i] = dev_uniform(sd, 2, k) + 0.5 * dev_uniform(sd, 3, k);
  }
}
// rows [row0, row0 + Mloc) of the single dense problem of k_gen_batched(B=1): the slice a rank of
// a row-sharded solve holds (identical values to the corresponding rows of the whole problem)
__global__ void k_gen_dense_slice(int row0, int Mloc, int N, unsigned long long seed, double* A, double* b, double* p0)
{
  for(long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < Mloc; i += (long long)gridDim.x * blockDim.x)
  {
    const long long gi = row0 + i;
    double s = 0.0;
    for(int k = 0; k < N; k++)
    {
      const double a = dev_uniform(seed, 1, (unsigned long long)(gi * N + k));
      A[i * N + k] = a;
      s += a * phi(dev_uniform(seed, 2, k));
    }
    b[i] = s + 0.01 * dev_uniform(seed, 4, (unsigned long long)gi);
  }
  for(long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < N; k += (long long)gridDim.x * blockDim.x)
    p0[k] = dev_uniform(seed, 2, k) + 0.5 * dev_uniform(seed, 3, k);
}
extern "C" dlb_dev_problem* dlb_dev_problem_create_dense_slice(int row0, int Mloc, int N, unsigned long long seed, double* p0_host)
{
  dlb_dev_problem* D = (dlb_dev_problem*)calloc(1, sizeof(*D));
  D->N = N; D->M = Mloc; D->B = 1;
  double* d_p0 = 0;
  if(cudaMalloc(&D->d_Adense, sizeof(double) * (size_t)Mloc * N) != cudaSuccess ||
     cudaMalloc(&D->d_b, sizeof(double) * (size_t)Mloc) != cudaSuccess ||
     cudaMalloc(&d_p0, sizeof(double) * (size_t)N) != cudaSuccess) { free(D); return NULL; }
  k_gen_dense_slice<<<1184, 256>>>(row0, Mloc, N, seed, D->d_Adense, D->d_b, d_p0);
  cudaMemcpy(p0_host, d_p0, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost);
  cudaFree(d_p0);
  cudaEventCreate(&D->e0); cudaEventCreate(&D->e1);
  if(cudaDeviceSynchronize() != cudaSuccess) { free(D); return NULL; }
  return D;
}
extern "C" dlb_dev_problem* dlb_dev_problem_create_batched(int B, int M, int N, unsigned long long seed, double* p0_host)
{
  dlb_dev_problem* D = (dlb_dev_problem*)calloc(1, sizeof(*D));
  D->N = N; D->M = M; D->B = B;
  double* d_p0 = 0;
  if(cudaMalloc(&D->d_Adense, sizeof(double) * (size_t)B * M * N) != cudaSuccess ||
     cudaMalloc(&D->d_b, sizeof(double) * (size_t)B * M) != cudaSuccess ||
     cudaMalloc(&d_p0, sizeof(double) * (size_t)B * N) != cudaSuccess) { free(D); return NULL; }
  k_gen_batched<<<1184, 256>>>(B, M, N, seed, D->d_Adense, D->d_b, d_p0);
  cudaMemcpy(p0_host, d_p0, sizeof(double) * (size_t)B * N, cudaMemcpyDeviceToHost);
  cudaFree(d_p0);
  cudaEventCreate(&D->e0); cudaEventCreate(&D->e1);
  if(cudaDeviceSynchronize() != cudaSuccess) { free(D); return NULL; }
  return D;
}
// B copies of the reference's sample.c surface fit (problems.c kind 1), each with its own start
// point: the bilinear model makes Gauss-Newton overshoot, so this exercises rejected steps,
// cauchy and interpolated steps in the batched automaton. d_Adense holds (x_i, y_i) pairs.
__global__ void k_model_sample_batched(int B, int M, const double* __restrict__ xy, const double* __restrict__ meas,
                                       const double* __restrict__ p, const int* __restrict__ active,
                                       double* __restrict__ x, double* __restrict__ J)
{
  const long long total = (long long)B * M;
  for(long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
  {
    const long long prob = idx / M; const int i = (int)(idx - prob * M);
    if(active && !active[prob]) continue;
    const double* pp = p + prob * 6;
    const double X = xy[2*i], Y = xy[2*i+1];
    double* g = J + idx * 6;
    g[0] = pp[1]*X*X; g[1] = pp[0]*X*X + pp[2]*Y*Y; g[2] = pp[1]*Y*Y + X*Y; g[3] = X; g[4] = Y; g[5] = 1.0;
    x[idx] = pp[0]*pp[1]*X*X + pp[1]*pp[2]*Y*Y + pp[2]*X*Y + pp[3]*X + pp[4]*Y + pp[5] - meas[i];
  }
}
extern "C" dlb_dev_problem* dlb_dev_problem_create_sample_batched(const dlb_problem* P, int B)
{
  dlb_dev_problem* D = (dlb_dev_problem*)calloc(1, sizeof(*D));
  D->N = 6; D->M = P->M; D->B = B; D->nnz = -1;      // nnz < 0 marks the sample model
  cudaMalloc(&D->d_Adense, sizeof(double) * 2 * (size_t)P->M);
  cudaMalloc(&D->d_b, sizeof(double) * (size_t)P->M);
  cudaMemcpy(D->d_Adense, P->Ax, sizeof(double) * 2 * (size_t)P->M, cudaMemcpyHostToDevice);
  cudaMemcpy(D->d_b, P->b, sizeof(double) * (size_t)P->M, cudaMemcpyHostToDevice);
  cudaEventCreate(&D->e0); cudaEventCreate(&D->e1);
  if(cudaDeviceSynchronize() != cudaSuccess) { free(D); return NULL; }
  return D;
}

extern "C" void dlb_dev_problem_free(dlb_dev_problem* D)
{
  if(!D) return;
  cudaFree(D->d_Ap); cudaFree(D->d_Ai); cudaFree(D->d_Ax); cudaFree(D->d_b); cudaFree(D->d_Adense);
  cudaEventDestroy(D->e0); cudaEventDestroy(D->e1);
  free(D);
}
extern "C" void dlb_dev_problem_timing(dlb_dev_problem* D, int on) { D->timing = on; D->ms_total = 0; D->ncalls = 0; }
extern "C" double dlb_dev_problem_ms(dlb_dev_problem* D) { return D->ms_total; }
extern "C" int dlb_dev_problem_ncalls(dlb_dev_problem* D) { return D->ncalls; }

extern "C" void dlb_dev_cb_sparse(const double* d_p, double* d_x, double* d_J, void* stream, void* cookie)
{
  dlb_dev_problem* D = (dlb_dev_problem*)cookie;
  cudaStream_t st = (cudaStream_t)stream;
  if(D->timing) cudaEventRecord(D->e0, st);
  k_model_sparse<<<148 * 4, 256, 0, st>>>(D->M, D->d_Ap, D->d_Ai, D->d_Ax, D->d_b, d_p, d_x, d_J);
  if(D->timing) { cudaEventRecord(D->e1, st); cudaEventSynchronize(D->e1); float ms; cudaEventElapsedTime(&ms, D->e0, D->e1); D->ms_total += ms; }
  D->ncalls++;
}
extern "C" void dlb_dev_cb_dense(const double* d_p, double* d_x, double* d_J, void* stream, void* cookie)
{
  dlb_dev_problem* D = (dlb_dev_problem*)cookie;
  cudaStream_t st = (cudaStream_t)stream;
  k_model_dense<<<148 * 16, 256, 0, st>>>((long long)D->M, D->M, D->N, model_width(D->N), D->d_Adense, D->d_b, d_p, NULL, d_x, d_J);
  D->ncalls++;
}
extern "C" void dlb_dev_cb_dense_batched(const double* d_p, double* d_x, double* d_J, const int* d_active, int B,
                                         void* stream, void* cookie)
{
  dlb_dev_problem* D = (dlb_dev_problem*)cookie;
  cudaStream_t st = (cudaStream_t)stream;
  if(D->timing) cudaEventRecord(D->e0, st);
  if(D->nnz < 0) k_model_sample_batched<<<148 * 8, 256, 0, st>>>(B, D->M, D->d_Adense, D->d_b, d_p, d_active, d_x, d_J);
  else if(D->N == 16)
  k_model_dense16<<<148 * 16, 256, 0, st>>>((long long)B * D->M, D->M, D->d_Adense, D->d_b, d_p, d_active, d_x, d_J);
  else
  k_model_dense<<<148 * 16, 256, 0, st>>>((long long)B * D->M, D->M, D->N, model_width(D->N), D->d_Adense, D->d_b, d_p, d_active, d_x, d_J);
  if(D->timing) { cudaEventRecord(D->e1, st); cudaEventSynchronize(D->e1); float ms; cudaEventElapsedTime(&ms, D->e0, D->e1); D->ms_total += ms; }
  D->ncalls++;
}
extern "C" void* dlb_dev_cb_sparse_ptr(void)        { return (void*)&dlb_dev_cb_sparse; }
extern "C" void* dlb_dev_cb_dense_ptr(void)         { return (void*)&dlb_dev_cb_dense; }
extern "C" void* dlb_dev_cb_dense_batched_ptr(void) { return (void*)&dlb_dev_cb_dense_batched; }
