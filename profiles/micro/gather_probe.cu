// profiles/micro/gather_probe.cu -- why does the extend-add gather of the mrcal-shaped tree (442 warps, each
// summing 136-entry blocks of 16 child fronts that another kernel/phase has just written) take 20+ us?
// Variants: who wrote the sources (same launch boundary), load flavour, front stride, batch shape.
#include <cstdio>
#include <cuda_runtime.h>
#define NF 200
#define R 80
__global__ void k_write(double* fronts, size_t stride)
{
  double* A = fronts + (size_t)blockIdx.x * stride;
  for(int i = threadIdx.x; i < R * R; i += blockDim.x) A[i] = 1e-3 * (i + blockIdx.x);
}
template<int MODE>
__device__ __forceinline__ double ld(const double* p)
{
  double v;
  if(MODE == 0) v = *p;
  else if(MODE == 1) asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
// warp per target: 13 chunks x 34 strips; strip s = columns 2s,2s+1 of the 68-row border block, rows >= col
template<int MODE>
__global__ void k_gather(const double* fronts, size_t stride, double* out, long long* cyc)
{
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if(wid >= 442) return;
  const int chunk = wid / 34, strip = wid % 34;
  const int col0 = 12 + 2 * strip, h = 80 - col0;
  long long t0 = clock64();
  double acc[5] = {0, 0, 0, 0, 0};
  for(int q8 = 0; q8 < 16; q8 += 8)
  {
    double v[8][5];
#pragma unroll
    for(int qq = 0; qq < 8; qq++)
    {
      const int child = min(chunk * 16 + q8 + qq, NF - 1);
      const double* src = fronts + (size_t)child * stride + col0 + (size_t)col0 * R;
#pragma unroll
      for(int u = 0; u < 5; u++)
      {
        const int e = lane + 32 * u, j = e / h, i = e - j * h;
        v[qq][u] = (e < 2 * h && i >= j) ? ld<MODE>(src + i + j * R) : 0.0;
      }
    }
#pragma unroll
    for(int qq = 0; qq < 8; qq++)
#pragma unroll
      for(int u = 0; u < 5; u++) acc[u] += v[qq][u];
  }
  long long t1 = clock64();
  for(int u = 0; u < 5; u++) out[(size_t)wid * 160 + lane + 32 * u] = acc[u];
  if(wid == 0 && lane == 0) cyc[0] = t1 - t0;
}
int main()
{
  double *fronts, *out; long long* cyc;
  const size_t maxstride = R * R + 1024;
  cudaMalloc(&fronts, NF * maxstride * 8); cudaMalloc(&out, 442 * 160 * 8); cudaMalloc(&cyc, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for(int pass = 0; pass < 2; pass++)
  for(size_t stride : {(size_t)R * R, (size_t)R * R + 16, (size_t)R * R + 48})
    for(int mode = 0; mode < 3; mode++)
      for(int rewrite = 0; rewrite < 2; rewrite++)
      {
        float best = 1e9; long long hc = 0;
        for(int rep = 0; rep < 5; rep++)
        {
          if(rewrite || rep == 0) k_write<<<NF, 256>>>(fronts, stride);
          cudaEventRecord(e0);
          if(mode == 0) k_gather<0><<<56, 256>>>(fronts, stride, out, cyc);
          else if(mode == 1) k_gather<1><<<56, 256>>>(fronts, stride, out, cyc);
          else k_gather<2><<<56, 256>>>(fronts, stride, out, cyc);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          if(ms < best) { best = ms; cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost); }
        }
        if(pass) printf("stride %zu mode %d (%s) sources %s: kernel %.1f us, warp 0 loads+adds %lld cycles\n", stride, mode,
               mode == 0 ? "ld" : (mode == 1 ? "ld.cg" : "ld.nc"), rewrite ? "rewritten before every gather" : "written once", best * 1e3, hc);
      }
  return 0;
}
