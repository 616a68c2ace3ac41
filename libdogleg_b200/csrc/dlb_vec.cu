// dlb_vec.cu -- the O(Nstate) pieces of the dog-leg step, fused so that the
// host only ever branches on the handful of scalars in DlbScalars.
//
// Replaces: compute_updateCauchy's vector part (reference dogleg.c:605-610),
// vec_negate + norm2 after the solve (:862-865, 894-897),
// computeInterpolatedUpdate (:964-987), the step construction and p+step in
// takeStepFrom (:1204-1207, 1231-1234, 1259), inner(Jt_x,step) (:1108) and the
// update-threshold scan (:1289-1291).
#include "dlb_common.cuh"
#include "dlb_device.h"

static inline int vec_grid(int N, int sm_count)
{
  int g = (N + DLB_NT * 4 - 1) / (DLB_NT * 4);
  const int cap = sm_count * 4;
  return g < 1 ? 1 : (g > cap ? cap : g);
}

// k = -|g|^2 / |J g|^2 ; cauchy = k g ; |cauchy|^2 = k^2 |g|^2
__global__ void __launch_bounds__(DLB_NT)
k_cauchy(const double* __restrict__ g, int N, double* __restrict__ cauchy, DlbScalars* sc)
{
  const double g2 = sc->norm2_Jtx, jg2 = sc->norm2_JJtx;
  const double k = -g2 / jg2;
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) cauchy[i] = k * g[i];
  if(blockIdx.x == 0 && threadIdx.x == 0) { sc->k_cauchy = k; sc->norm2_cauchy = k * k * g2; }
}
void dlb_launch_cauchy(const double* Jtx, int N, double* cauchy, DlbScalars* sc, int sm_count, cudaStream_t st)
{
  k_cauchy<<<vec_grid(N, sm_count), DLB_NT, 0, st>>>(Jtx, N, cauchy, sc);
}

// gn[perm[k]] = -z[k] (perm == NULL: identity), |gn|^2
__global__ void __launch_bounds__(DLB_NT)
k_gn_finish(const double* __restrict__ z, const int* __restrict__ perm, int N, double* __restrict__ gn,
            double* part, unsigned int* counter, DlbScalars* sc)
{
  double n2 = 0.0;
  for(int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x)
  {
    const double v = -z[k];
    gn[perm ? perm[k] : k] = v;
    n2 = fma(v, v, n2);
  }
  double out[5];
  if(grid_reduce5(n2, 0, 0, 0, 0, part, counter, out)) sc->norm2_gn = out[0];
}
void dlb_launch_gn_finish(const double* zperm, const int* perm_or_null, int N, double* gn,
                          double* part, unsigned int* counter, DlbScalars* sc, int sm_count, cudaStream_t st)
{
  k_gn_finish<<<vec_grid(N, sm_count), DLB_NT, 0, st>>>(zperm, perm_or_null, N, gn, part, counter, sc);
}

// interpolation coefficients: d = a - b, l2 = |d|^2, neg_c = d.a,
// disc = neg_c^2 - l2 (|a|^2 - delta^2) clamped at 0, k = (neg_c + sqrt(disc)) / l2
__global__ void __launch_bounds__(DLB_NT)
k_interp_dots(const double* __restrict__ a, const double* __restrict__ b, int N, double delta,
              double* part, unsigned int* counter, DlbScalars* sc)
{
  double l2 = 0.0, negc = 0.0;
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
  {
    const double d = a[i] - b[i];
    l2 = fma(d, d, l2);
    negc = fma(d, a[i], negc);
  }
  double out[5];
  if(grid_reduce5(l2, negc, 0, 0, 0, part, counter, out))
  {
    double disc = out[1] * out[1] - out[0] * (sc->norm2_cauchy - delta * delta);
    sc->discriminant = disc;
    if(disc < 0.0) disc = 0.0;
    sc->k_interp = (out[1] + sqrt(disc)) / out[0];
  }
}

// step, p_to = p_from + step, Jt_x . step, |step|^2, max|step|
__global__ void __launch_bounds__(DLB_NT)
k_step_apply(int type, double delta, const double* __restrict__ p_from, const double* __restrict__ g,
             const double* __restrict__ cauchy, const double* __restrict__ gn, int N,
             double* __restrict__ step, double* __restrict__ p_to,
             double* part, unsigned int* counter, DlbScalars* sc)
{
  double scale = 1.0, k = 0.0;
  if(type == 0)      scale = delta / sqrt(sc->norm2_cauchy);
  else if(type == 2) k = sc->k_interp;
  double n2 = 0.0, gd = 0.0, mx = 0.0;
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
  {
    double s;
    if(type == 0)      s = scale * cauchy[i];
    else if(type == 1) s = gn[i];
    else { const double a = cauchy[i]; s = a + k * (gn[i] - a); }
    step[i] = s;
    p_to[i] = p_from[i] + s;
    n2 = fma(s, s, n2);
    gd = fma(g[i], s, gd);
    mx = fmax(mx, fabs(s));
  }
  double out[5];
  if(grid_reduce5(n2, gd, 0, 0, mx, part, counter, out))
  {
    // the reference stores the UNCLIPPED cauchy length and the cached GN length
    // (dogleg.c:1200, 1228); only the interpolated step reports its own norm
    sc->norm2_step   = type == 0 ? sc->norm2_cauchy : (type == 1 ? sc->norm2_gn : out[0]);
    sc->Jtx_dot_step = out[1];
    sc->maxabs_step  = out[4];
  }
}

void dlb_launch_step(int step_type, double delta, const double* p_from, const double* Jtx,
                     const double* cauchy, const double* gn, int N, double* step, double* p_to,
                     double* part, unsigned int* counter, DlbScalars* sc, int sm_count, cudaStream_t st)
{
  const int g = vec_grid(N, sm_count);
  if(step_type == 2) k_interp_dots<<<g, DLB_NT, 0, st>>>(cauchy, gn, N, delta, part, counter, sc);
  k_step_apply<<<g, DLB_NT, 0, st>>>(step_type, delta, p_from, Jtx, cauchy, gn, N, step, p_to, part, counter, sc);
}

// |v|^2 and max|v| into norm2_Jtx / maxabs_Jtx (dense-products: Jt_x comes from the user)
__global__ void __launch_bounds__(DLB_NT)
k_vec_stats_Jtx(const double* __restrict__ v, int N, double* part, unsigned int* counter, DlbScalars* sc)
{
  double n2 = 0.0, mx = 0.0;
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
  { n2 = fma(v[i], v[i], n2); mx = fmax(mx, fabs(v[i])); }
  double out[5];
  if(grid_reduce5(n2, 0, 0, 0, mx, part, counter, out)) { sc->norm2_Jtx = out[0]; sc->maxabs_Jtx = out[4]; }
}
// the same, then the scalars go to the host through mapped pinned memory (row-sharded fused evaluation)
__global__ void __launch_bounds__(DLB_NT)
k_vec_stats_Jtx_pub(const double* __restrict__ v, int N, double* part, unsigned int* counter, DlbScalars* sc,
                    DlbPublished* pub, unsigned long long seq)
{
  double n2 = 0.0, mx = 0.0;
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
  { n2 = fma(v[i], v[i], n2); mx = fmax(mx, fabs(v[i])); }
  double out[5];
  if(grid_reduce5(n2, 0, 0, 0, mx, part, counter, out))
  {
    sc->norm2_Jtx = out[0]; sc->maxabs_Jtx = out[4];
    sc->norm2_x = v[N];                              // the all-reduced |x|^2 sits behind the gradient
    pub->sc = *sc;
    __threadfence_system();
    *(volatile unsigned long long*)&pub->seq = seq;
  }
}
void dlb_launch_vec_stats_Jtx_pub(const double* v, int N, double* part, unsigned int* counter, DlbScalars* sc,
                                  DlbPublished* pub, unsigned long long seq, int sm_count, cudaStream_t st)
{
  k_vec_stats_Jtx_pub<<<vec_grid(N, sm_count), DLB_NT, 0, st>>>(v, N, part, counter, sc, pub, seq);
}
void dlb_launch_vec_stats_Jtx(const double* v, int N, double* part, unsigned int* counter, DlbScalars* sc,
                              int sm_count, cudaStream_t st)
{
  k_vec_stats_Jtx<<<vec_grid(N, sm_count), DLB_NT, 0, st>>>(v, N, part, counter, sc);
}
