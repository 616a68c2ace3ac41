// dlb_front.cu -- multifrontal numeric Cholesky and triangular solves.
//
// Replaces cholmod_factorize[_p] (reference dogleg.c:656-665) and
// cholmod_solve(CHOLMOD_A) (:853-856) -- and, for the dense solve types where
// the whole matrix is a single front, dpptrf/dpotrf + dpptrs/dpotrs
// (:779-804, :872-892).
//
// Every supernode s owns a front: an r x r column-major lower-triangular block.
// One launch handles all fronts of one elimination-tree level, one CTA per
// front:
//   1. zero the front
//   2. add the partial JtJ blocks (Gpart) of the pattern classes assigned to it
//   3. extend-add the update matrices of its children (pull, fixed child order)
//   4. add lambda on the pivot diagonal
//   5. eliminate its ncols pivot columns (right-looking, in shared memory when
//      the front fits); the trailing block is left for the parent
// All sums are performed in a fixed order: results are bit-reproducible.
// A pivot <= 0 or non-finite marks the matrix as not positive definite (the
// LAPACK dpptrf rule; CHOLMOD's simplicial LDL' only flags exact zeros --
// documented divergence): the smallest failing column is kept.
#include "dlb_common.cuh"
#include "dlb_device.h"

#define FRONT_NT 256

__device__ __forceinline__ void pair_from_index(int q, int& a, int& b)
{
  a = (int)((sqrt(8.0 * (double)q + 1.0) - 1.0) * 0.5);
  while((a + 1) * (a + 2) / 2 <= q) a++;
  while(a * (a + 1) / 2 > q) a--;
  b = q - a * (a + 1) / 2;
}

// mode 0: full (assemble + children + lambda + eliminate); mode 1: elements only (tests)
template<bool SMEM, int NT>
__global__ void __launch_bounds__(NT)
k_front_level(DlbFrontDev F, DlbSparseDev S, int l0, double* __restrict__ fronts,
              const double* __restrict__ Gpart, double lambda, long long* minor, int mode)
{
  extern __shared__ double sh_front[];
  const int s   = F.level_sn[l0 + blockIdx.x];
  const int c0  = F.sn_first[s], nc = F.sn_first[s+1] - c0;
  const int rp  = F.rows_ptr[s], r = F.rows_ptr[s+1] - rp;
  double* Ag = fronts + F.front_off[s];
  double* A  = SMEM ? sh_front : Ag;
  const int tid = threadIdx.x;

  // Gpart == NULL: the front arrives pre-filled (dense solve types); else start from zero
  if(Gpart)     for(int idx = tid; idx < r * r; idx += NT) A[idx] = 0.0;
  else if(SMEM) for(int idx = tid; idx < r * r; idx += NT) A[idx] = Ag[idx];
  __syncthreads();

  // ---- elements: classes assigned to this front ----
  if(Gpart)
    for(int ci = F.fcls_ptr[s]; ci < F.fcls_ptr[s+1]; ci++)
    {
      const int c = F.fcls_list[ci];
      const int k = S.cls_ptr[c+1] - S.cls_ptr[c];
      const int* loc = S.cls_loc + S.cls_ptr[c];
      const int npairs = k * (k + 1) / 2;
      const int t0 = F.cls_task_ptr[c], t1 = F.cls_task_ptr[c+1];
      for(int q = tid; q < npairs; q += NT)
      {
        double g = 0.0;
        for(int t = t0; t < t1; t++) g += Gpart[S.task_Goff[t] + q];
        int a, b; pair_from_index(q, a, b);
        const int la = loc[a], lb = loc[b];
        const int row = la > lb ? la : lb, col = la > lb ? lb : la;
        A[row + col * r] += g;
      }
      __syncthreads();
    }

  if(mode == 0)
  {
    // ---- children: extend-add their update matrices ----
    const int g0 = F.grp_ptr ? F.grp_ptr[2*s] : 0, g1 = F.grp_ptr ? F.grp_ptr[2*s+1] : 0;   // [first,last) pairs
    for(int g = g0; g < g1; g++)
    { // pre-summed by k_extend_groups, already in this front's indexing
      const double* T = F.grp_tmp + F.grp_off[g];
      for(int idx = tid; idx < r * r; idx += NT) A[idx] += T[idx];
    }
    if(g1 > g0) __syncthreads();
    for(int ch = (g1 > g0) ? F.child_ptr[s+1] : F.child_ptr[s]; ch < F.child_ptr[s+1]; ch++)
    {
      const int c   = F.child_list[ch];
      const int ncc = F.sn_first[c+1] - F.sn_first[c];
      const int rc  = F.rows_ptr[c+1] - F.rows_ptr[c];
      const int nb  = rc - ncc;
      const double* U = fronts + F.front_off[c];
      const int* rel = F.rel + F.rows_ptr[c] + ncc;
      for(int idx = tid; idx < nb * nb; idx += NT)
      {
        const int j = idx / nb, i = idx - j * nb;
        if(i >= j) A[rel[i] + rel[j] * r] += U[(ncc + i) + (size_t)(ncc + j) * rc];
      }
      __syncthreads();
    }
    for(int j = tid; j < nc; j += NT) A[j + j * r] += lambda;
    __syncthreads();

    // ---- eliminate the pivot columns ----
    bool failed = false;
    for(int j = 0; j < nc; j++)
    {
      const double d = A[j + j * r];
      if(!(d > 0.0) || isinf(d))
      {
        if(tid == 0) atomicMin(minor, (long long)(c0 + j));
        failed = true;
        break;
      }
      const double sd = sqrt(d), inv = 1.0 / sd;
      __syncthreads();
      for(int i = j + tid; i < r; i += NT) A[i + j * r] = (i == j) ? sd : A[i + j * r] * inv;
      __syncthreads();
      const int w = r - j - 1;
      for(int idx = tid; idx < w * w; idx += NT)
      {
        const int cc = idx / w, ii = idx - cc * w;
        if(ii >= cc)
        {
          const int col = j + 1 + cc, row = j + 1 + ii;
          A[row + col * r] = fma(-A[row + j * r], A[col + j * r], A[row + col * r]);
        }
      }
      __syncthreads();
    }
    (void)failed;
  }
  if(SMEM)
  {
    __syncthreads();
    for(int idx = tid; idx < r * r; idx += NT) Ag[idx] = A[idx];
  }
}

// one CTA per group of children of a heavy front: T = sum of their update matrices, scattered
// into the parent's indexing, children in ascending order
__global__ void __launch_bounds__(FRONT_NT)
k_extend_groups(DlbFrontDev F, int g0, const double* __restrict__ fronts)
{
  const int g = g0 + blockIdx.x;
  const int s = F.grp_front[g];
  const int r = F.rows_ptr[s+1] - F.rows_ptr[s];
  double* T = F.grp_tmp + F.grp_off[g];
  const int tid = threadIdx.x;
  for(int idx = tid; idx < r * r; idx += FRONT_NT) T[idx] = 0.0;
  __syncthreads();
  for(int ch = F.grp_child0[g]; ch < F.grp_child1[g]; ch++)
  {
    const int c   = F.child_list[ch];
    const int ncc = F.sn_first[c+1] - F.sn_first[c];
    const int rc  = F.rows_ptr[c+1] - F.rows_ptr[c];
    const int nb  = rc - ncc;
    const double* U = fronts + F.front_off[c];
    const int* rel = F.rel + F.rows_ptr[c] + ncc;
    for(int idx = tid; idx < nb * nb; idx += FRONT_NT)
    {
      const int j = idx / nb, i = idx - j * nb;
      if(i >= j) T[rel[i] + (size_t)rel[j] * r] += U[(ncc + i) + (size_t)(ncc + j) * rc];
    }
    __syncthreads();
  }
}
void dlb_launch_extend_groups(const DlbFrontDev& F, int g0, int g1, const double* fronts, cudaStream_t st)
{
  if(g1 > g0) k_extend_groups<<<g1 - g0, FRONT_NT, 0, st>>>(F, g0, fronts);
}

template<int NT>
static void launch_front_level_nt(const DlbFrontDev& F, const DlbSparseDev& S, int l0, int nf, double* fronts,
                                  const double* Gpart, double lambda, long long* minor, int mode, size_t smem,
                                  cudaStream_t st)
{
  if(smem <= 200 * 1024)
  {
    static bool attr_set = false;
    if(!attr_set)
    {
      cudaFuncSetAttribute(k_front_level<true, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr_set = true;
    }
    k_front_level<true, NT><<<nf, NT, smem, st>>>(F, S, l0, fronts, Gpart, lambda, minor, mode);
  }
  else
    k_front_level<false, NT><<<nf, NT, 0, st>>>(F, S, l0, fronts, Gpart, lambda, minor, mode);
}

void dlb_launch_front_level(const DlbFrontDev& F, const DlbSparseDev& S, int l0, int l1,
                            double* fronts, const double* Gpart, double lambda,
                            long long* minor, int max_rows, cudaStream_t st)
{
  const int nf = l1 - l0;
  if(nf <= 0) return;
  const int mode = lambda < 0.0 ? 1 : 0;       // lambda < 0 selects the elements-only test mode
  const size_t smem = (size_t)max_rows * max_rows * sizeof(double);
  // more threads for bigger fronts: the trailing update has ~r^2/2 independent entries per pivot
  if(max_rows > 96)      launch_front_level_nt<1024>(F, S, l0, nf, fronts, Gpart, lambda, minor, mode, smem, st);
  else if(max_rows > 40) launch_front_level_nt<512>(F, S, l0, nf, fronts, Gpart, lambda, minor, mode, smem, st);
  else                   launch_front_level_nt<256>(F, S, l0, nf, fronts, Gpart, lambda, minor, mode, smem, st);
}

// ------------------------------------------------------------------ solves
// forward: y = L^-1 P b, leaves to root. ywork holds one r-vector per front.
__global__ void __launch_bounds__(FRONT_NT)
k_solve_fwd_level(DlbFrontDev F, int l0, const double* __restrict__ fronts,
                  const double* __restrict__ rhs, double* __restrict__ ywork,
                  double* __restrict__ zperm, int nrhs, long long ytot)
{
  const int s  = F.level_sn[l0 + blockIdx.x];
  const int c0 = F.sn_first[s], nc = F.sn_first[s+1] - c0;
  const int rp = F.rows_ptr[s], r = F.rows_ptr[s+1] - rp;
  const double* A = fronts + F.front_off[s];
  const int tid = threadIdx.x;
  for(int rh = 0; rh < nrhs; rh++)
  {
    double* y = ywork + (size_t)rh * ytot + rp;
    for(int i = tid; i < r; i += FRONT_NT) y[i] = i < nc ? rhs[(size_t)rh * F.n + F.perm[c0 + i]] : 0.0;
    __syncthreads();
    for(int ch = F.child_ptr[s]; ch < F.child_ptr[s+1]; ch++)
    {
      const int c   = F.child_list[ch];
      const int ncc = F.sn_first[c+1] - F.sn_first[c];
      const int rc  = F.rows_ptr[c+1] - F.rows_ptr[c];
      const double* yc = ywork + (size_t)rh * ytot + F.rows_ptr[c];
      const int* rel = F.rel + F.rows_ptr[c];
      for(int i = ncc + tid; i < rc; i += FRONT_NT) y[rel[i]] += yc[i];
      __syncthreads();
    }
    for(int j = 0; j < nc; j++)
    {
      if(tid == 0) y[j] /= A[j + (size_t)j * r];
      __syncthreads();
      const double yj = y[j];
      for(int i = j + 1 + tid; i < r; i += FRONT_NT) y[i] = fma(-A[i + (size_t)j * r], yj, y[i]);
      __syncthreads();
    }
    for(int i = tid; i < nc; i += FRONT_NT) zperm[(size_t)rh * F.n + c0 + i] = y[i];
    __syncthreads();
  }
}

// backward: x = L^-T y, root to leaves, in place in zperm (permuted order)
__global__ void __launch_bounds__(FRONT_NT)
k_solve_bwd_level(DlbFrontDev F, int l0, const double* __restrict__ fronts,
                  double* __restrict__ zperm, int nrhs)
{
  __shared__ double sh[32];
  const int s  = F.level_sn[l0 + blockIdx.x];
  const int c0 = F.sn_first[s], nc = F.sn_first[s+1] - c0;
  const int rp = F.rows_ptr[s], r = F.rows_ptr[s+1] - rp;
  const double* A = fronts + F.front_off[s];
  const int* rows = F.rows + rp;
  const int tid = threadIdx.x;
  for(int rh = 0; rh < nrhs; rh++)
  {
    double* z = zperm + (size_t)rh * F.n;
    for(int j = nc - 1; j >= 0; j--)
    {
      double acc = 0.0;
      for(int i = j + 1 + tid; i < r; i += FRONT_NT) acc = fma(A[i + (size_t)j * r], z[rows[i]], acc);
      acc = block_sum(acc, sh);
      if(tid == 0) z[c0 + j] = (z[c0 + j] - acc) / A[j + (size_t)j * r];
      __syncthreads();
    }
  }
}

void dlb_launch_solve_fwd_level(const DlbFrontDev& F, int l0, int l1, const double* fronts,
                                const double* rhs, double* ywork, double* zperm, int nrhs,
                                int max_rows, cudaStream_t st)
{
  (void)max_rows;
  if(l1 > l0) k_solve_fwd_level<<<l1 - l0, FRONT_NT, 0, st>>>(F, l0, fronts, rhs, ywork, zperm, nrhs, F.ytot);
}
void dlb_launch_solve_bwd_level(const DlbFrontDev& F, int l0, int l1, const double* fronts,
                                double* zperm, int nrhs, int max_rows, cudaStream_t st)
{
  (void)max_rows;
  if(l1 > l0) k_solve_bwd_level<<<l1 - l0, FRONT_NT, 0, st>>>(F, l0, fronts, zperm, nrhs);
}

// tests: scatter the assembled (elements-only) fronts into a dense n x n matrix
__global__ void k_fronts_to_dense(DlbFrontDev F, const double* __restrict__ fronts, double* out)
{
  const int s  = blockIdx.x;
  const int rp = F.rows_ptr[s], r = F.rows_ptr[s+1] - rp;
  const double* A = fronts + F.front_off[s];
  for(int idx = threadIdx.x; idx < r * r; idx += blockDim.x)
  {
    const int j = idx / r, i = idx - j * r;
    if(i < j) continue;
    const double v = A[i + (size_t)j * r];
    if(v == 0.0) continue;
    const int gi = F.perm[F.rows[rp + i]], gj = F.perm[F.rows[rp + j]];
    atomicAdd(&out[(size_t)gi * F.n + gj], v);
    if(gi != gj) atomicAdd(&out[(size_t)gj * F.n + gi], v);
  }
}
void dlb_launch_fronts_to_dense(const DlbFrontDev& F, const double* fronts, double* out, cudaStream_t st)
{
  k_fronts_to_dense<<<F.nsuper, 256, 0, st>>>(F, fronts, out);
}
