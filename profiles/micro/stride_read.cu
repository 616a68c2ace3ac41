// micro-benchmark: how fast can HBM deliver (a) a contiguous stream, (b) every other 144-byte /
// 192-byte column (what one pattern class of the mrcal layout reads), with plain 8-byte loads,
// one column per warp load? Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 stride_read.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_cols(const double* __restrict__ J, long long ncols, int k, int stride_cols, int phase, int unroll_dummy, double* out)
{
  // warp w handles columns phase + stride_cols*(w + i*nwarps)
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5), w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  double acc = 0;
  const long long per = (ncols / stride_cols + nw - 1) / nw;   // contiguous range of "own" columns per warp
  const long long c0 = w * per, c1 = min(ncols / stride_cols, c0 + per);
  for(long long c = c0; c + 4 <= c1; c += 4)
  {
    double v[4];
#pragma unroll
    for(int u = 0; u < 4; u++) { const long long col = phase + stride_cols * (c + u); v[u] = lane < k ? __ldg(J + col * k + lane) : 0.0; }
#pragma unroll
    for(int u = 0; u < 4; u++) acc += v[u];
  }
  if(acc == 123.456) out[0] = acc;
}
__global__ void k_contig(const double2* __restrict__ J, long long n2, double* out)
{
  double acc = 0;
  for(long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) { double2 v = J[i]; acc += v.x + v.y; }
  if(acc == 123.456) out[0] = acc;
}
int main()
{
  const int k = 24; const long long ncols = 1000000;   // 192 MB
  double* J; double* out; cudaMalloc(&J, sizeof(double) * ncols * k); cudaMalloc(&out, 8);
  cudaMemset(J, 0, sizeof(double) * ncols * k);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for(int rep = 0; rep < 2; rep++)
  {
    cudaEventRecord(e0); k_contig<<<148 * 8, 256>>>((const double2*)J, ncols * k / 2, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("contiguous double2 grid-stride: %.1f us  %.0f GB/s\n", ms * 1e3, ncols * k * 8 / ms / 1e6);
    for(int g = 2; g <= 8; g *= 2)
    {
      cudaEventRecord(e0); k_cols<<<148 * g, 256>>>(J, ncols, k, 1, 0, 0, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1); printf("all columns, warp per column range, %d CTAs/SM: %.1f us  %.0f GB/s\n", g, ms * 1e3, ncols * k * 8 / ms / 1e6);
      cudaEventRecord(e0); k_cols<<<148 * g, 256>>>(J, ncols, k, 2, 0, 0, out); k_cols<<<148 * g, 256>>>(J, ncols, k, 2, 1, 0, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1); printf("even then odd columns (two passes), %d CTAs/SM: %.1f us  %.0f GB/s\n", g, ms * 1e3, ncols * k * 8 / ms / 1e6);
    }
  }
  return 0;
}
