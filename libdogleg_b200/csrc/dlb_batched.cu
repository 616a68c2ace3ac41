// dlb_batched.cu -- B independent small dense problems (config C3: Nstate=16, Nmeas=256,
// B=100k), the whole dog-leg automaton device-resident. Additive entry point
// dogleg_gpu_optimize_dense_batched() (include/dogleg_gpu.h); per problem it follows the
// reference's DOGLEG_DENSE path step for step:
//   evaluation           dogleg.c:1034-1053, 1073-1081   (J'x, |x|^2, inf-norm test)
//   trust-region update  dogleg.c:1303-1356
//   loop / termination   dogleg.c:1359-1476
//   Cauchy               dogleg.c:529-617
//   JtJ + lambda, dpptrf dogleg.c:699-816,  dpptrs :867-898
//   step / interpolation dogleg.c:927-998, 1172-1297, expected improvement :1112-1127
//
// One CTA per problem and per trial: the 32 KB Jacobian the callback just produced is loaded
// into shared memory ONCE and everything above is computed from there, so HBM traffic per
// trial is the algorithmic minimum 8(MN+M) bytes. JtJ = J'J runs on the FP64
// tensor cores (mma.sync m8n8k4, SASS DMMA), the 16x16 Cholesky and the vector work on one
// warp with lane == state index. All reductions have a fixed order.
//
// The callback writes trial t into Jacobian buffer t%2. A rejected step needs the Jacobian of
// the *before* point again (for |J step|^2 of the retried step): it is still in the other
// buffer, and is moved to a third buffer only when the next callback would overwrite it.
#include "dlb_common.cuh"
#include "dogleg_gpu.h"
#include <vector>
#include <cstring>
#include <string>
#include <algorithm>

extern "C" void dlb_set_error(const char* msg);

#define BT_NT 128
#define BT_WARPS (BT_NT / 32)
#define BT_NMAX 32

struct BatchState
{
  int B, N, M, max_iterations;
  double tr0, dec_factor, dec_thr, inc_factor, inc_thr, Jtx_thr, upd_thr, tr_thr;
  // per problem
  double *tr, *n2x_before, *n2c, *n2gn, *expected, *lambda;
  int *steps, *flags, *jb, *active;
  double *p_before, *ptrial, *Jtx, *cauchy, *gn;
  // callback buffers
  double *x;            // B x M
  double *J[3];         // B x M x N each
  int *n_active;
};
enum { FL_CAUCHY = 1, FL_GN = 2, FL_EDGE = 4, FL_PENDING = 8 };

__device__ __forceinline__ void bt_dmma(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// |J v|^2 with J (M x N) in shared memory: one thread per row, fixed-order block sum
__device__ __forceinline__ double bt_norm2_Jv(const double* sJ, const double* v, int M, int N, double* red)
{
  double acc = 0.0;
  for(int i = threadIdx.x; i < M; i += BT_NT)
  {
    double d = 0.0;
    for(int k = 0; k < N; k++) d = fma(sJ[i * N + k], v[k], d);
    acc = fma(d, d, acc);
  }
  acc = block_sum(acc, red);
  __shared__ double bcast;
  if(threadIdx.x == 0) bcast = acc;
  __syncthreads();
  return bcast;
}

__global__ void __launch_bounds__(BT_NT)
k_batched_trial(BatchState S, int t)
{
  extern __shared__ double sm[];
  const int b = blockIdx.x;
  if(!S.active[b]) return;
  const int N = S.N, M = S.M;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int NT8 = (N + 7) / 8;
  double* sJ   = sm;                       // M*N
  double* sx   = sJ + (size_t)M * N;       // M
  double* sg   = sx + M;                   // N : Jt_x of the before point (after phase 3)
  double* sgn  = sg + BT_NMAX;             // gradient of the new point
  double* sc   = sgn + BT_NMAX;            // cauchy
  double* sn   = sc + BT_NMAX;             // gauss-newton
  double* ss   = sn + BT_NMAX;             // step
  double* sA   = ss + BT_NMAX;             // BT_NMAX*BT_NMAX JtJ / factor
  double* sW   = sA + BT_NMAX * BT_NMAX;   // BT_WARPS*BT_NMAX*BT_NMAX partial tiles / scratch
  __shared__ double red[32];
  __shared__ double scal[8];
  __shared__ int    ctl[4];                // 0: done, 1: need reload of J_before, 2: accepted

  const int cur = t & 1;
  const double* gJ = S.J[cur] + (size_t)b * M * N;
  const double* gx = S.x + (size_t)b * M;
  for(int i = tid; i < M * N; i += BT_NT) sJ[i] = ldg_stream(gJ + i);
  for(int i = tid; i < M; i += BT_NT) sx[i] = ldg_stream(gx + i);
  __syncthreads();

  // ---- evaluation of the trial point: |x|^2 and J'x (reference dogleg.c:1045-1048) ----
  double n2 = 0.0;
  for(int i = tid; i < M; i += BT_NT) n2 = fma(sx[i], sx[i], n2);
  n2 = block_sum(n2, red);
  if(tid == 0) scal[0] = n2;
  {
    // groups of 32 threads share the rows; lane == state index
    double acc = 0.0;
    if(lane < N) for(int i = w; i < M; i += BT_WARPS) acc = fma(sJ[i * N + lane], sx[i], acc);
    sW[w * BT_NMAX + lane] = acc;
    __syncthreads();
    if(tid < N)
    {
      double s0 = 0.0;
      for(int u = 0; u < BT_WARPS; u++) s0 += sW[u * BT_NMAX + tid];
      sgn[tid] = s0;
    }
    __syncthreads();
  }

  // ---- accept / reject and trust-region update: one thread (dogleg.c:1303-1356, 1359-1470) ----
  if(tid == 0)
  {
    int flags = S.flags[b];
    double tr = S.tr[b];
    const double n2x_new = scal[0];
    double gmax = 0.0;
    for(int k = 0; k < N; k++) gmax = fmax(gmax, fabs(sgn[k]));
    const bool converged = !(gmax > S.Jtx_thr);
    int done = 0, reload = 0, accepted = 0, steps = S.steps[b];
    if(!(flags & FL_PENDING))
    { // the initial operating point
      accepted = 1;
      if(converged || S.max_iterations <= 0) done = 1;
    }
    else
    {
      const double observed = S.n2x_before[b] - n2x_new;
      double rho = observed / S.expected[b];
      if(!isfinite(n2x_new)) rho = -INFINITY;          // see DESIGN.md, divergence 2
      if(rho < S.dec_thr)
      {
        if(!(flags & FL_EDGE)) tr = sqrt(S.n2gn[b]);
        tr *= S.dec_factor;
      }
      else if(rho > S.inc_thr && (flags & FL_EDGE)) tr *= S.inc_factor;
      if(rho > 0.0)
      {
        accepted = 1;
        steps++;
        if(converged || steps >= S.max_iterations) done = 1;
      }
      else
      {
        if(tr < S.tr_thr) done = 1;
        else reload = 1;
      }
    }
    if(accepted)
    {
      S.n2x_before[b] = n2x_new;
      S.jb[b] = cur;
      flags &= ~(FL_CAUCHY | FL_GN | FL_EDGE);
    }
    S.tr[b] = tr; S.steps[b] = steps; S.flags[b] = flags;
    ctl[0] = done; ctl[1] = reload; ctl[2] = accepted;
    scal[1] = tr;
  }
  __syncthreads();
  const bool accepted = ctl[2] != 0;
  if(accepted)
  { // the trial point becomes the current one
    if(tid < N)
    {
      S.p_before[(size_t)b * N + tid] = S.ptrial[(size_t)b * N + tid];
      S.Jtx[(size_t)b * N + tid] = sgn[tid];
      sg[tid] = sgn[tid];
    }
  }
  if(ctl[0])
  {
    if(tid == 0) { S.active[b] = 0; atomicSub(S.n_active, 1); }
    return;
  }
  if(ctl[1])
  { // rejected: bring back the Jacobian and the cached vectors of the before point
    const double* gJb = S.J[S.jb[b]] + (size_t)b * M * N;
    __syncthreads();
    for(int i = tid; i < M * N; i += BT_NT) sJ[i] = gJb[i];
    if(tid < N)
    {
      sg[tid] = S.Jtx[(size_t)b * N + tid];
      sc[tid] = S.cauchy[(size_t)b * N + tid];
      sn[tid] = S.gn[(size_t)b * N + tid];
    }
  }
  __syncthreads();
  int flags = S.flags[b];
  const double tr = scal[1];

  // ---- Cauchy step (dogleg.c:529-617) ----
  double n2c;
  if(!(flags & FL_CAUCHY))
  {
    const double jg2 = bt_norm2_Jv(sJ, sg, M, N, red);
    double g2 = 0.0;
    for(int k = 0; k < N; k++) g2 = fma(sg[k], sg[k], g2);
    const double kc = -g2 / jg2;
    n2c = kc * kc * g2;
    if(tid < N) { sc[tid] = kc * sg[tid]; S.cauchy[(size_t)b * N + tid] = sc[tid]; }
    if(tid == 0) { S.n2c[b] = n2c; }
    flags |= FL_CAUCHY;
    __syncthreads();
  }
  else n2c = S.n2c[b];

  // ---- choose the step (dogleg.c:1192-1255) ----
  int type;
  double n2gn = 0.0;
  if(n2c >= tr * tr) { type = 0; flags |= FL_EDGE; }
  else
  {
    if(!(flags & FL_GN))
    {
      // JtJ = J'J on the tensor cores: every warp takes every BT_WARPS-th group of 4 rows,
      // the warps' partial tiles are added in warp order
      {
        double acc[10][2];
#pragma unroll
        for(int i = 0; i < 10; i++) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
        const int g = lane >> 2, tt = lane & 3;
        for(int r0 = 4 * w; r0 < M; r0 += 4 * BT_WARPS)
        {
          const int row = r0 + tt;
          double v[4];
#pragma unroll
          for(int ti = 0; ti < 4; ti++)
            v[ti] = (ti < NT8 && row < M && 8 * ti + g < N) ? sJ[row * N + 8 * ti + g] : 0.0;
          int idx = 0;
#pragma unroll
          for(int ti = 0; ti < 4; ti++)
#pragma unroll
            for(int tj = 0; tj <= ti; tj++, idx++)
              if(ti < NT8) bt_dmma(acc[idx][0], acc[idx][1], v[ti], v[tj]);
        }
        double* mine = sW + (size_t)w * BT_NMAX * BT_NMAX;
        int idx = 0;
#pragma unroll
        for(int ti = 0; ti < 4; ti++)
#pragma unroll
          for(int tj = 0; tj <= ti; tj++, idx++)
            if(ti < NT8)
            {
              const int a = 8 * ti + g, b0 = 8 * tj + 2 * tt;
              mine[a * BT_NMAX + b0] = acc[idx][0];
              mine[a * BT_NMAX + b0 + 1] = acc[idx][1];
            }
      }
      __syncthreads();
      for(int e = tid; e < N * N; e += BT_NT)
      {
        const int a = e / N, c = e - a * N;
        if(c <= a)
        {
          double s0 = 0.0;
          for(int u = 0; u < BT_WARPS; u++) s0 += sW[(size_t)u * BT_NMAX * BT_NMAX + a * BT_NMAX + c];
          sA[a * BT_NMAX + c] = s0;
        }
      }
      __syncthreads();
      // Cholesky with the lambda ladder, then the two triangular solves: warp 0, lane == row
      if(w == 0)
      {
        double lam = S.lambda[b];
        double* L = sW;                                  // N x N scratch (row a, col c <= a)
        for(;;)
        {
          for(int e = lane; e < N * N; e += 32)
          {
            const int a = e / N, c = e - a * N;
            if(c <= a) L[a * BT_NMAX + c] = sA[a * BT_NMAX + c] + (a == c ? lam : 0.0);
          }
          __syncwarp();
          bool ok = true;
          for(int j = 0; j < N; j++)
          {
            const double d = L[j * BT_NMAX + j];
            if(!(d > 0.0) || isinf(d)) { ok = false; break; }
            const double sd = sqrt(d);
            __syncwarp();
            if(lane == j) L[j * BT_NMAX + j] = sd;
            else if(lane > j && lane < N) L[lane * BT_NMAX + j] /= sd;
            __syncwarp();
            // trailing update: lane owns row a = lane
            if(lane > j && lane < N)
            {
              const double la = L[lane * BT_NMAX + j];
              for(int c = j + 1; c <= lane; c++) L[lane * BT_NMAX + c] = fma(-la, L[c * BT_NMAX + j], L[lane * BT_NMAX + c]);
            }
            __syncwarp();
          }
          if(ok) break;
          lam = lam == 0.0 ? 1e-10 : lam * 10.0;         // dogleg.c:811-813
          if(!isfinite(lam)) break;
          __syncwarp();
        }
        if(lane == 0) S.lambda[b] = lam;
        // forward / backward substitution; u in sn
        if(lane < N) sn[lane] = sg[lane];
        __syncwarp();
        for(int j = 0; j < N; j++)
        {
          const double yj = sn[j] / L[j * BT_NMAX + j];
          __syncwarp();
          if(lane == j) sn[j] = yj;
          else if(lane > j && lane < N) sn[lane] = fma(-L[lane * BT_NMAX + j], yj, sn[lane]);
          __syncwarp();
        }
        for(int j = N - 1; j >= 0; j--)
        {
          const double xj = sn[j] / L[j * BT_NMAX + j];
          __syncwarp();
          if(lane == j) sn[j] = xj;
          else if(lane < j) sn[lane] = fma(-L[j * BT_NMAX + lane], xj, sn[lane]);
          __syncwarp();
        }
        if(lane < N) { sn[lane] = -sn[lane]; S.gn[(size_t)b * N + lane] = sn[lane]; }
        __syncwarp();
        double q = 0.0;
        for(int k = 0; k < N; k++) q = fma(sn[k], sn[k], q);
        if(lane == 0) { S.n2gn[b] = q; scal[2] = q; }
      }
      flags |= FL_GN;
      __syncthreads();
      n2gn = scal[2];
    }
    else n2gn = S.n2gn[b];
    if(n2gn <= tr * tr) { type = 1; flags &= ~FL_EDGE; }
    else                { type = 2; flags |= FL_EDGE; }
  }

  // ---- the step itself, p + step, Jt_x . step, max|step| : every thread redundantly (N <= 32) ----
  double kk = 0.0;
  if(type == 2)
  {
    double l2 = 0.0, negc = 0.0;
    for(int k = 0; k < N; k++) { const double d = sc[k] - sn[k]; l2 = fma(d, d, l2); negc = fma(d, sc[k], negc); }
    double disc = negc * negc - l2 * (n2c - tr * tr);
    if(disc < 0.0) disc = 0.0;
    kk = (negc + sqrt(disc)) / l2;
  }
  __syncthreads();
  if(tid < N)
  {
    double sv;
    if(type == 0)      sv = (tr / sqrt(n2c)) * sc[tid];
    else if(type == 1) sv = sn[tid];
    else               sv = sc[tid] + kk * (sn[tid] - sc[tid]);
    ss[tid] = sv;
    S.ptrial[(size_t)b * N + tid] = S.p_before[(size_t)b * N + tid] + sv;
  }
  __syncthreads();
  double gd = 0.0, smax = 0.0;
  for(int k = 0; k < N; k++) { gd = fma(sg[k], ss[k], gd); smax = fmax(smax, fabs(ss[k])); }
  const double js2 = bt_norm2_Jv(sJ, ss, M, N, red);
  double expected = -2.0 * gd - js2;
  bool finished = false;
  if(!(smax > S.upd_thr)) { expected = -1.0; finished = true; }   // dogleg.c:1289-1296, 1403-1408
  if(tid == 0)
  {
    S.expected[b] = expected;
    S.flags[b] = flags | FL_PENDING;
    if(finished) { S.active[b] = 0; atomicSub(S.n_active, 1); }
  }
  if(finished) return;

  // ---- keep the before-point Jacobian alive across the next callback (which writes buffer (t+1)%2) ----
  const int jb = S.jb[b];
  if(jb == ((t + 1) & 1))
  {
    double* keep = S.J[2] + (size_t)b * M * N;
    for(int i = tid; i < M * N; i += BT_NT) keep[i] = sJ[i];
    if(tid == 0) S.jb[b] = 2;
  }
}

// statistics of the last batched solve of this thread, for bench.py:
// [0] kernel launches, [1] sum over launches of active problems, [2] ms in k_batched_trial (CUDA
// events on the launching stream), [3] ms in the user callback (same events), [4] total accepted steps
static thread_local double g_batched_stats[8];
extern "C" void dogleg_gpu_batched_stats(double out[8]) { memcpy(out, g_batched_stats, sizeof(g_batched_stats)); }

static size_t batched_smem_bytes(int N, int M)
{
  return sizeof(double) * ((size_t)M * N + M + 5 * BT_NMAX + BT_NMAX * BT_NMAX + (size_t)BT_WARPS * BT_NMAX * BT_NMAX);
}

#define CUB(call) do { cudaError_t _e = (call); if(_e != cudaSuccess) { \
  dlb_set_error((std::string(#call) + ": " + cudaGetErrorString(_e)).c_str()); goto fail; } } while(0)

extern "C" int dogleg_gpu_optimize_dense_batched(double* p, unsigned int Nstate, unsigned int Nmeas,
                                                 unsigned int B, dogleg_gpu_callback_dense_batched_t* f,
                                                 void* cookie, const dogleg_parameters2_t* parameters,
                                                 double* norm2x_out, int* iterations_out)
{
  if(dogleg_gpu_device_count() <= 0) { dlb_set_error("no CUDA device available: libdogleg-b200 has no CPU fallback"); return -1; }
  if(!f || !p || B == 0) { dlb_set_error("dense_batched: bad arguments"); return -1; }
  const int N = (int)Nstate, M = (int)Nmeas;
  const size_t smem = batched_smem_bytes(N, M);
  if(N > BT_NMAX || N < 1 || smem > 220 * 1024)
  {
    dlb_set_error("dense_batched: needs Nstate <= 32 and a Jacobian that fits in shared memory; "
                  "use dogleg_optimize_dense2 per problem for bigger ones");
    return -1;
  }
  dogleg_parameters2_t P;
  if(parameters) P = *parameters; else dogleg_getDefaultParameters(&P);
  cudaSetDevice(dogleg_gpu_get_device());

  BatchState S;
  memset(&S, 0, sizeof(S));
  S.B = (int)B; S.N = N; S.M = M; S.max_iterations = P.max_iterations;
  S.tr0 = P.trustregion0; S.dec_factor = P.trustregion_decrease_factor; S.dec_thr = P.trustregion_decrease_threshold;
  S.inc_factor = P.trustregion_increase_factor; S.inc_thr = P.trustregion_increase_threshold;
  S.Jtx_thr = P.Jt_x_threshold; S.upd_thr = P.update_threshold; S.tr_thr = P.trustregion_threshold;
  std::vector<void*> allocs;
  cudaStream_t st = 0;
  int result = -1;
  int* h_active = 0;
  {
    auto dalloc = [&](size_t bytes) -> void* {
      void* q = 0;
      if(cudaMalloc(&q, bytes ? bytes : 8) != cudaSuccess) return (void*)0;
      allocs.push_back(q);
      return q;
    };
    const size_t bN = (size_t)B * N * sizeof(double), bd = (size_t)B * sizeof(double), bi = (size_t)B * sizeof(int);
    S.tr = (double*)dalloc(bd); S.n2x_before = (double*)dalloc(bd); S.n2c = (double*)dalloc(bd);
    S.n2gn = (double*)dalloc(bd); S.expected = (double*)dalloc(bd); S.lambda = (double*)dalloc(bd);
    S.steps = (int*)dalloc(bi); S.flags = (int*)dalloc(bi); S.jb = (int*)dalloc(bi); S.active = (int*)dalloc(bi);
    S.p_before = (double*)dalloc(bN); S.ptrial = (double*)dalloc(bN); S.Jtx = (double*)dalloc(bN);
    S.cauchy = (double*)dalloc(bN); S.gn = (double*)dalloc(bN);
    S.x = (double*)dalloc((size_t)B * M * sizeof(double));
    for(int u = 0; u < 3; u++) S.J[u] = (double*)dalloc((size_t)B * M * N * sizeof(double));
    S.n_active = (int*)dalloc(sizeof(int));
    for(void* q : allocs) if(!q) { dlb_set_error("dense_batched: out of device memory"); goto fail; }
    if(allocs.size() != 20) { dlb_set_error("dense_batched: out of device memory"); goto fail; }
  }
  CUB(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CUB(cudaFuncSetAttribute(k_batched_trial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUB(cudaMemcpyAsync(S.ptrial, p, (size_t)B * N * sizeof(double), cudaMemcpyHostToDevice, st));
  CUB(cudaMemcpyAsync(S.p_before, p, (size_t)B * N * sizeof(double), cudaMemcpyHostToDevice, st));
  CUB(cudaMemsetAsync(S.lambda, 0, (size_t)B * sizeof(double), st));
  CUB(cudaMemsetAsync(S.steps, 0, (size_t)B * sizeof(int), st));
  CUB(cudaMemsetAsync(S.flags, 0, (size_t)B * sizeof(int), st));
  CUB(cudaMemsetAsync(S.jb, 0, (size_t)B * sizeof(int), st));
  {
    std::vector<double> tr(B, S.tr0);
    std::vector<int> ones(B, 1);
    const int nb = (int)B;
    CUB(cudaMemcpyAsync(S.tr, tr.data(), (size_t)B * sizeof(double), cudaMemcpyHostToDevice, st));
    CUB(cudaMemcpyAsync(S.active, ones.data(), (size_t)B * sizeof(int), cudaMemcpyHostToDevice, st));
    CUB(cudaMemcpyAsync(S.n_active, &nb, sizeof(int), cudaMemcpyHostToDevice, st));
    CUB(cudaStreamSynchronize(st));
  }
  CUB(cudaHostAlloc((void**)&h_active, sizeof(int), cudaHostAllocDefault));
  *h_active = (int)B;
  memset(g_batched_stats, 0, sizeof(g_batched_stats));
  {
    cudaEvent_t ev[3];
    for(int u = 0; u < 3; u++) cudaEventCreate(&ev[u]);
    // every problem needs at most (max_iterations accepted + rejected) trials; rejected trials are
    // bounded by the trust region collapsing below its threshold
    const long long trial_cap = 64LL * (long long)std::max(P.max_iterations, 1) + 4096;
    for(long long t = 0; *h_active > 0 && t < trial_cap; t++)
    {
      g_batched_stats[0] += 1; g_batched_stats[1] += *h_active;
      cudaEventRecord(ev[0], st);
      f(S.ptrial, S.x, S.J[t & 1], S.active, (int)B, (void*)st, cookie);
      cudaEventRecord(ev[1], st);
      k_batched_trial<<<B, BT_NT, smem, st>>>(S, (int)(t & 0x7fffffff));
      cudaEventRecord(ev[2], st);
      CUB(cudaGetLastError());
      CUB(cudaMemcpyAsync(h_active, S.n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
      CUB(cudaStreamSynchronize(st));
      float ms_cb = 0, ms_k = 0;
      cudaEventElapsedTime(&ms_cb, ev[0], ev[1]); cudaEventElapsedTime(&ms_k, ev[1], ev[2]);
      g_batched_stats[2] += ms_k; g_batched_stats[3] += ms_cb;
    }
    for(int u = 0; u < 3; u++) cudaEventDestroy(ev[u]);
  }
  CUB(cudaMemcpyAsync(p, S.p_before, (size_t)B * N * sizeof(double), cudaMemcpyDeviceToHost, st));
  if(norm2x_out) CUB(cudaMemcpyAsync(norm2x_out, S.n2x_before, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, st));
  if(iterations_out) CUB(cudaMemcpyAsync(iterations_out, S.steps, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUB(cudaStreamSynchronize(st));
  result = (int)B - *h_active;
fail:
  if(h_active) cudaFreeHost(h_active);
  if(st) cudaStreamDestroy(st);
  for(void* q : allocs) cudaFree(q);
  return result;
}
