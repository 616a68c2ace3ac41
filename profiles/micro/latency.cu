// profiles/micro/latency.cu -- dependent-operation latencies on B200 that bound the persistent trial
// kernel (one CTA per front, few warps per SM: everything is a latency chain).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/latency profiles/micro/latency.cu && gpurun_out/latency
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_chase(const int* next, int steps, int spinners, volatile unsigned int* flag, long long* out)
{
  if(blockIdx.x > 0)
  { // CTAs > 0 spin on a flag like the waiters of a grid barrier
    if(threadIdx.x == 0 && (int)blockIdx.x <= spinners) while(*flag == 0) { }
    return;
  }
  if(threadIdx.x == 0)
  {
    int p = 0;
    long long t0 = clock64();
    for(int i = 0; i < steps; i++) p = next[p];
    long long t1 = clock64();
    out[0] = t1 - t0; out[1] = p;
    __threadfence();
    *flag = 1;
  }
}
__global__ void k_alu(int steps, double seed, long long* out, double* sink)
{
  __shared__ double sh[1024];
  __shared__ int shn[1024];
  const int lane = threadIdx.x;
  for(int i = lane; i < 1024; i += 32) { sh[i] = 1.0 + 1e-9 * i; shn[i] = (i * 37 + 11) & 1023; }
  __syncwarp();
  double a = seed, b = 1.0000001;
  long long t0 = clock64();
  for(int i = 0; i < steps; i++) a = fma(a, b, 1e-9);
  long long t1 = clock64();
  double r = seed + 2.0;
  for(int i = 0; i < steps; i++) r = rsqrt(r) + 2.0;
  long long t2 = clock64();
  int p = lane;
  for(int i = 0; i < steps; i++) p = shn[p];
  long long t3 = clock64();
  double s = seed;
  for(int i = 0; i < steps; i++) s = __shfl_sync(0xffffffffu, s, (lane + 1) & 31) + 1.0;
  long long t4 = clock64();
  double d = seed + 3.0;
  for(int i = 0; i < steps; i++) d = 1.0 / d + 2.0;
  long long t5 = clock64();
  double q = seed + 3.0;
  for(int i = 0; i < steps; i++) q = sqrt(q) + 2.0;
  long long t6 = clock64();
  // dependent FP64 tensor-core chain (mma.sync m8n8k4, SASS DMMA), and 4 independent chains
  double c0 = seed, c1 = seed;
  for(int i = 0; i < steps; i++)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(b), "d"(b));
  long long t7 = clock64();
  double e0[4] = {seed, seed, seed, seed}, e1[4] = {seed, seed, seed, seed};
  for(int i = 0; i < steps; i++)
#pragma unroll
    for(int u = 0; u < 4; u++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(e0[u]), "+d"(e1[u]) : "d"(b), "d"(b));
  long long t8 = clock64();
  if(lane == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; out[4] = t5 - t4; out[5] = t6 - t5; out[6] = t7 - t6; out[7] = t8 - t7; }
  sink[lane] = a + r + p + s + d + q + c0 + c1 + e0[0] + e0[1] + e0[2] + e0[3] + e1[0] + e1[1] + e1[2] + e1[3];
}
int main()
{
  const int n = 1 << 20;               // 4 MB of ints: L2 resident
  int* h = new int[n];
  for(int i = 0; i < n; i++) h[i] = (int)(((long long)i * 7919 + 104729) % n);
  int* d; cudaMalloc(&d, n * sizeof(int)); cudaMemcpy(d, h, n * sizeof(int), cudaMemcpyHostToDevice);
  long long* out; cudaMalloc(&out, 128); unsigned int* flag; cudaMalloc(&flag, 4);
  double* sink; cudaMalloc(&sink, 32 * 8);
  long long ho[8];
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("SM clock (attr) %d kHz\n", clk);
  for(int spinners : {0, 0, 100, 199, 295})
  {
    cudaMemset(flag, 0, 4);
    k_chase<<<296, 256>>>(d, 4000, spinners, flag, out);
    cudaDeviceSynchronize();
    cudaMemcpy(ho, out, 16, cudaMemcpyDeviceToHost);
    printf("global pointer chase (4 MB, L2 after first pass), %3d spinning CTAs: %.0f cycles/load\n", spinners, ho[0] / 4000.0);
  }
  for(int rep = 0; rep < 2; rep++)
  {
    k_alu<<<1, 32>>>(4000, 1.5, out, sink);
    cudaDeviceSynchronize();
    cudaMemcpy(ho, out, 64, cudaMemcpyDeviceToHost);
    printf("dependent DFMA %.1f | rsqrt(double)+add %.1f | LDS chase %.1f | shfl+add %.1f | 1/x+add %.1f | sqrt+add %.1f | dependent DMMA %.1f | 4 independent DMMA chains %.1f per DMMA cycles\n",
           ho[0] / 4000.0, ho[1] / 4000.0, ho[2] / 4000.0, ho[3] / 4000.0, ho[4] / 4000.0, ho[5] / 4000.0, ho[6] / 4000.0, ho[7] / 16000.0);
  }
  return 0;
}
