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
#include <cstring>
#include <cstdlib>
#include "dlb_device.h"

#define BF_NB 64                 // pivot block width
#define BF_LDS 68                // shared-memory row stride of 64-wide tiles (68 mod 16 == 4)

__device__ __forceinline__ void bf_dmma(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// A batch of large fronts (all the large fronts of one elimination-tree level) is processed together, in panels
// of 64 pivot columns. Three schedules, chosen by how many (front, 64-row tile) pairs the level has:
//   * a few hundred (levels in the middle of a tree): LEFT-looking with look-ahead, ONE launch per panel
//     (k_bf_step): the row tiles of panel k -- every tile refactorizes the 64 x 64 diagonal block itself in shared
//     memory (front_eliminate on a tall 128 x 64 "front") instead of waiting for a one-CTA potrf -- run beside the
//     update of panel k+1 by the panels 0..k-1 (one K loop over hundreds of columns); the update by panel k-1 itself,
//     final only since the previous launch, is folded into the tiles of panel k;
//   * very few (one huge front: the dense solve types, the top of a tree): RIGHT-looking with the update one panel
//     late (k_bf_step, lazy_right): everything behind panel k receives the update of panel k-1 while panel k --
//     which got that update by the fold -- is eliminated;
//   * many hundreds (the levels next to the leaves): k_bf_diag + k_bf_trsm, one diagonal-block CTA per front.
//   once at the end: k_bf_gemm, the Schur complement of the front, C -= L21 L21' with K = all pivots.
// Every schedule leaves the INVERSE of each diagonal block behind (one more tile whose rows below the block are an
// identity matrix): dlb_bigsolve.cu and k_bf_trsm multiply by it.
// (round 1 was right-looking: potrf, trsm and a K = 64 trailing update of the WHOLE remaining front per
// panel: 3 launches per panel, the trailing matrix read and written nc/64 times).
// The GEMM operands are fed by the TMA unit: 1-D bulk copies (cp.async.bulk, SASS UBLKCP) of one
// column segment each into a 3-stage shared-memory ring guarded by mbarriers.
#include "dlb_devfn.cuh"

__device__ __forceinline__ unsigned bf_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bf_mbar_init(void* bar, int count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bf_smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void bf_mbar_expect_tx(void* bar, unsigned bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bf_smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bf_bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(bf_smem_u32(dst)), "l"(src), "r"(bytes), "r"(bf_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bf_mbar_wait(void* bar, unsigned parity)
{
  asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
               :: "r"(bf_smem_u32(bar)), "r"(parity) : "memory");
}

// ---- (b) diagonal block + rows below: one CTA per 64-row tile ----
#define BFP_LD 129               // odd leading dimension of the 128 x 64 shared-memory tile
#define BFW_LD 132               // row stride of the folded panel's K-chunk (132 mod 16 == 4)
#define BFW_KC 16
// fold: the tile first receives the update of the PREVIOUS panel (K = its 64 columns, final since the last
// launch), T -= L[tile rows, k0-64:k0] L[block rows, k0-64:k0]' -- the part of the left-looking panel update
// that the look-ahead launch (k_bf_step) could not do ahead of time. W: BFW_KC x BFW_LD doubles.
// inv_out != NULL: the INVERSE tile of the block -- its "rows below" are an identity matrix, which the elimination
// turns into I * L_bb^-T: the inverse of the diagonal block for the price of one more tile, no extra dependency
// chain. It goes to inv_out[0..4096) row-major (X[i*64+j], X = L_bb^-1) and inv_out[4096..8192) column-major,
// zero-padded; the triangular solves of the right-hand sides (dlb_bigsolve.cu) and k_bf_trsm multiply by it.
__device__ __forceinline__ void bf_panel_tile(const DlbBigFront& f, double* __restrict__ fronts, int step, int bx,
                                              long long* minor, double* T, double* W, bool fold, double* __restrict__ inv_out,
                                              int* cnt)
{
  const int k0 = step * BF_NB;
  if(k0 >= f.nc) return;
  const int nb = f.nc - k0 < BF_NB ? f.nc - k0 : BF_NB;
  const int ld = f.r, r = f.r;
  const bool inverse = inv_out != 0;
  const int row0 = k0 + nb + bx * 64;                             // first row of this tile below the block
  if(!inverse && row0 >= r && bx > 0) return;
  const int mine = inverse ? nb : (row0 < r ? (r - row0 < 64 ? r - row0 : 64) : 0);
  double* A = fronts + f.off;
  const int tid = threadIdx.x;
  // rows 0..nb-1: the diagonal block (lower triangle), rows nb..nb+mine-1: this tile's rows of the panel
  for(int idx = tid; idx < nb * nb; idx += 256)
  {
    const int j = idx / nb, i = idx - j * nb;
    T[i + j * BFP_LD] = i >= j ? A[(size_t)(k0 + j) * ld + k0 + i] : 0.0;
  }
  for(int idx = tid; idx < mine * nb; idx += 256)
  {
    const int j = idx / mine, i = idx - j * mine;
    T[nb + i + j * BFP_LD] = inverse ? (i == j ? 1.0 : 0.0) : A[(size_t)(k0 + j) * ld + row0 + i];
  }
  // Every tile of the front reads the UNFACTORIZED diagonal block, and one of them must overwrite it with L: the
  // tile that loaded it last (cnt: loaders so far; it is reset for the next panel). All tiles hold the same L.
  __shared__ int s_last;
  if(cnt)
  {
    __syncthreads();                                               // all loads of this CTA are done
    if(tid == 0)
    {
      const int below = r - k0 - nb;
      const int ntiles = (below > 0 ? (below + 63) / 64 : 1) + 1;   // row tiles (at least the block's own) + the inverse tile
      const int old = atomicAdd(cnt, 1);
      s_last = old == ntiles - 1;
      if(s_last) *cnt = 0;
    }
    __syncthreads();
  }
  const bool store_diag = cnt ? s_last != 0 : bx == 0;
  if(fold && k0 > 0)
  {
    const int lane = tid & 31, w = tid >> 5, g = lane >> 2, tt = lane & 3;
    double acc[2][8][2];
#pragma unroll
    for(int a = 0; a < 2; a++)
#pragma unroll
      for(int c = 0; c < 8; c++) { acc[a][c][0] = 0.0; acc[a][c][1] = 0.0; }
    for(int kc = 0; kc < BF_NB; kc += BFW_KC)
    {
      __syncthreads();
      // W[k][i] = L[row(i), k0 - 64 + kc + k]; tile rows beyond nb + mine are zero
      for(int idx = tid; idx < BFW_KC * 128; idx += 256)
      {
        const int k = idx >> 7, i = idx & 127;
        const int grow = i < nb ? k0 + i : (inverse ? 0 : row0 + (i - nb));   // the identity rows of an inverse tile get no update
        W[k * BFW_LD + i] = (i < nb + (inverse ? 0 : mine)) ? A[(size_t)(k0 - BF_NB + kc + k) * ld + grow] : 0.0;
      }
      __syncthreads();
#pragma unroll
      for(int k = 0; k < BFW_KC; k += 4)
      {
        const double a0 = W[(k + tt) * BFW_LD + 16 * w + g], a1 = W[(k + tt) * BFW_LD + 16 * w + 8 + g];
#pragma unroll
        for(int c = 0; c < 8; c++)
        {
          const double b = W[(k + tt) * BFW_LD + 8 * c + g];
          bf_dmma(acc[0][c][0], acc[0][c][1], a0, b);
          bf_dmma(acc[1][c][0], acc[1][c][1], a1, b);
        }
      }
    }
#pragma unroll
    for(int a = 0; a < 2; a++)
#pragma unroll
      for(int c = 0; c < 8; c++)
#pragma unroll
        for(int h = 0; h < 2; h++)
        {
          const int i = 16 * w + 8 * a + g, j = 8 * c + 2 * tt + h;
          if(j < nb && i < nb + mine && (i >= nb || i >= j)) T[i + j * BFP_LD] -= acc[a][c][h];
        }
  }
  const int fail = front_eliminate<256>(T, nb, nb, BFP_LD, nb + mine, tid, (double*)0);
  if(fail >= 0)
  {
    if(tid == 0 && bx == 0) atomicMin(minor, (long long)(f.col0 + k0 + fail));
    return;
  }
  if(inverse)
  { // rows nb.. of the tile now hold L_bb^-T: T[nb + i][j] = X[j][i]
    if(store_diag)
      for(int idx = tid; idx < nb * nb; idx += 256)
      {
        const int j = idx / nb, i = idx - j * nb;
        if(i >= j) A[(size_t)(k0 + j) * ld + k0 + i] = T[i + j * BFP_LD];
      }
    for(int idx = tid; idx < 4096; idx += 256)
    {
      const int a = idx >> 6, b = idx & 63;
      const bool in = a < nb && b < nb;
      inv_out[idx]        = in ? T[nb + b + a * BFP_LD] : 0.0;      // row-major:    [a*64 + b] = X[a][b]
      inv_out[4096 + idx] = in ? T[nb + a + b * BFP_LD] : 0.0;      // column-major: [a*64 + b] = X[b][a]
    }
    return;
  }
  if(store_diag)
    for(int idx = tid; idx < nb * nb; idx += 256)
    {
      const int j = idx / nb, i = idx - j * nb;
      if(i >= j) A[(size_t)(k0 + j) * ld + k0 + i] = T[i + j * BFP_LD];
    }
  for(int idx = tid; idx < mine * nb; idx += 256)
  {
    const int j = idx / mine, i = idx - j * mine;
    A[(size_t)(k0 + j) * ld + row0 + i] = T[nb + i + j * BFP_LD];
  }
}

// ---- (a)/(c) C[i-tile, j-tile] -= sum over k in [0, kend) of L[i-tile, k] L[j-tile, k]' ----
// mode 1 (Schur complement): output = the trailing block [nc, r)^2, tiles on or below the diagonal, kend = nc
// mode 2 (the last panel of the right-looking schedule): output = everything behind panel `step`, K = that panel
// (the panel updates of the other schedules call bf_gemm_tile from k_bf_step / k_bf_diag)
// One 64 x 64 output tile per CTA; warp w owns rows 8w..8w+7. The K loop runs over chunks of 32 columns
// through a 3-stage ring: stage = [32 columns][68] for the i rows and the same for the j rows; every
// column segment (64 doubles, contiguous in the column-major front) arrives by ONE bulk copy of its
// 16-byte aligned superset, so a segment may start one double into its slot (parity of its address).
#define BFG_KC 32
#define BFG_LD 68
#define BFG_NST 3
#define BFG_STAGE (2 * BFG_KC * BFG_LD)
// the K loop and the write-back of one 64 x 64 output tile: C[i0.., j0..j0+jw) -= L[i0.., kbeg:kend) L[j0.., kbeg:kend)'
__device__ __forceinline__ void bf_gemm_tile(const DlbBigFront& f, double* __restrict__ fronts, int i0, int j0, int jw,
                                             int kbeg, int kend, double* sm_g)
{
  __shared__ unsigned long long bars[BFG_NST];
  const int ld = f.r, r = f.r;
  double* A = fronts + f.off;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, tt = lane & 3;
  const bool same = i0 == j0;                    // diagonal tile: one operand
  if(tid == 0) for(int s2 = 0; s2 < BFG_NST; s2++) bf_mbar_init(&bars[s2], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int nchunk = (kend - kbeg + BFG_KC - 1) / BFG_KC;
  // absolute element index of (row, column k) = off + k * ld + row; its parity decides the slot offset
  const long long ei0 = f.off + i0, ej0 = f.off + j0;
  auto issue = [&](int c) {
    if(c >= nchunk) return;
    double* stage = sm_g + (size_t)(c % BFG_NST) * BFG_STAGE;
    const int kc = kend - kbeg - c * BFG_KC < BFG_KC ? kend - kbeg - c * BFG_KC : BFG_KC;
    if(w == 0)
    {
      // 66 doubles cover the 64 rows from either parity; all lanes agree on the byte count
      if(lane == 0) bf_mbar_expect_tx(&bars[c % BFG_NST], (unsigned)(kc * 66 * 8 * (same ? 1 : 2)));
      __syncwarp();
      if(lane < kc)
      {
        const long long k = (long long)kbeg + (long long)c * BFG_KC + lane;
        const long long ei = ei0 + k * ld, ej = ej0 + k * ld;
        bf_bulk_g2s(stage + lane * BFG_LD, fronts + (ei & ~1ll), 66 * 8, &bars[c % BFG_NST]);
        if(!same) bf_bulk_g2s(stage + (BFG_KC + lane) * BFG_LD, fronts + (ej & ~1ll), 66 * 8, &bars[c % BFG_NST]);
      }
    }
  };
  for(int c = 0; c < BFG_NST - 1; c++) issue(c);
  double acc[8][2];
#pragma unroll
  for(int c = 0; c < 8; c++) { acc[c][0] = 0.0; acc[c][1] = 0.0; }
  const int ldodd = ld & 1;
  const int pi0 = (int)(ei0 & 1), pj0 = (int)(ej0 & 1);
  for(int c = 0; c < nchunk; c++)
  {
    issue(c + BFG_NST - 1);
    bf_mbar_wait(&bars[c % BFG_NST], (unsigned)((c / BFG_NST) & 1));
    const double* Pi = sm_g + (size_t)(c % BFG_NST) * BFG_STAGE;
    const double* Pj = same ? Pi : Pi + BFG_KC * BFG_LD;
    const int kc = kend - kbeg - c * BFG_KC < BFG_KC ? kend - kbeg - c * BFG_KC : BFG_KC;
    const int kbase = kbeg + c * BFG_KC;
#pragma unroll 2
    for(int k = 0; k < BFG_KC; k += 4)
    {
      if(k >= kc) break;
      const int kk = k + tt;
      const bool kon = kk < kc;
      const int par = ((kbase + kk) & 1) & ldodd;                 // parity flips from column to column iff ld is odd
      const int oi = kk * BFG_LD + (pi0 ^ par), oj = kk * BFG_LD + (pj0 ^ par);
      const double a = kon ? Pi[oi + 8 * w + g] : 0.0;
#pragma unroll
      for(int cc = 0; cc < 8; cc++)
      {
        const double b = kon ? Pj[oj + 8 * cc + g] : 0.0;
        bf_dmma(acc[cc][0], acc[cc][1], a, b);
      }
    }
    __syncthreads();                             // the stage may be refilled
  }
  const int row = i0 + 8 * w + g;
  if(row < r)
#pragma unroll
    for(int cc = 0; cc < 8; cc++)
    {
      const int col = j0 + 8 * cc + 2 * tt;
      if(col <= row && col < j0 + jw)         A[(size_t)col * ld + row]       -= acc[cc][0];
      if(col + 1 <= row && col + 1 < j0 + jw) A[(size_t)(col + 1) * ld + row] -= acc[cc][1];
    }
}

__global__ void __launch_bounds__(256)
k_bf_gemm(const DlbBigFront* __restrict__ descs, double* __restrict__ fronts, int step, int mode)
{
  extern __shared__ __align__(16) double sm_g[];
  const DlbBigFront f = descs[blockIdx.y];
  const int r = f.r;
  int i0, j0, kbeg = 0, kend, jw;                // tile origin, K range, width of the output column range
  { // mode 1: the whole Schur complement at the end (K = all pivots); mode 2: right-looking, everything
    // behind panel `step` gets that panel's update (K = its 64 columns) -- for a batch too small to fill the
    // GPU with panel updates (one huge front: the dense solve types, the top of a tree)
    int t0;
    if(mode == 1) { kend = f.nc; t0 = f.nc; }
    else
    {
      const int k0 = step * BF_NB;
      if(k0 >= f.nc) return;
      kbeg = k0; kend = f.nc - k0 < BF_NB ? f.nc : k0 + BF_NB; t0 = kend;
    }
    int t = blockIdx.x, ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while((ti + 1) * (ti + 2) / 2 <= t) ti++;
    while(ti * (ti + 1) / 2 > t) ti--;
    const int tj = t - ti * (ti + 1) / 2;
    i0 = t0 + 64 * ti; j0 = t0 + 64 * tj; jw = 64;
    if(i0 >= r || kend <= kbeg) return;
  }
  bf_gemm_tile(f, fronts, i0, j0, jw, kbeg, kend, sm_g);
}

// ---- look-ahead step of the left-looking factorization: ONE launch per panel ----
//   CTAs [0, nf * ptiles):  panel `step` -- fold in the update of panel step-1 (final since the previous launch),
//                           then Cholesky of the diagonal block + solve of the rows below (bf_panel_tile)
//   the rest:               panel step+1 receives the updates of the panels 0 .. step-1 (bf_gemm_tile, K = [0, k0))
// so the long K loop of the next panel's update runs beside this panel's latency-bound elimination instead of
// behind it (the dependency chain per panel is max(panel, update) instead of their sum).
__global__ void __launch_bounds__(256)
k_bf_step(const DlbBigFront* __restrict__ descs, double* __restrict__ fronts, double* __restrict__ inv, int* __restrict__ cnt,
          int step, long long* minor, int nf, int ptiles, int gtiles, int lazy_right)
{
  extern __shared__ __align__(16) double sm_g[];
  int bid = (int)blockIdx.x;
  if(bid < nf * ptiles)
  {
    const DlbBigFront f = descs[bid / ptiles];
    const int bx = bid % ptiles;                                 // the last tile of every front: the inverse tile
    bf_panel_tile(f, fronts, step, bx == ptiles - 1 ? -1 : bx, minor, sm_g, sm_g + BFP_LD * BF_NB, true,
                  bx == ptiles - 1 ? inv + f.inv_off + (size_t)step * 8192 : (double*)0, cnt + bid / ptiles);
    return;
  }
  bid -= nf * ptiles;
  const DlbBigFront f = descs[bid / gtiles];
  const int k0 = step * BF_NB, k1 = k0 + BF_NB;          // panel step+1 starts at column k1
  if(lazy_right)
  { // right-looking with the update one panel late: everything BEHIND panel `step` receives the update of panel
    // step-1 (K = its 64 columns) while panel `step` -- which got that update by the fold -- is eliminated
    // (a front whose LAST panel was step-1 gets that panel's update of its update matrix here: t0 = nc)
    if(k0 == 0 || k0 - BF_NB >= f.nc) return;
    const int kend = k0 < f.nc ? k0 : f.nc;
    const int t0 = k0 >= f.nc ? f.nc : (k1 < f.nc ? k1 : f.nc);
    int t = bid % gtiles, ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while((ti + 1) * (ti + 2) / 2 <= t) ti++;
    while(ti * (ti + 1) / 2 > t) ti--;
    const int tj = t - ti * (ti + 1) / 2;
    const int i0 = t0 + 64 * ti, j0 = t0 + 64 * tj;
    if(i0 >= f.r) return;
    bf_gemm_tile(f, fronts, i0, j0, 64, k0 - BF_NB, kend, sm_g);
    return;
  }
  if(k1 >= f.nc || k0 == 0) return;
  const int i0 = k1 + (bid % gtiles) * 64;
  if(i0 >= f.r) return;
  bf_gemm_tile(f, fronts, i0, k1, f.nc - k1 < BF_NB ? f.nc - k1 : BF_NB, 0, k0, sm_g);
}

// ---- throughput form of a panel step (batches with far more row tiles than SM slots) ----
// k_bf_step lets EVERY row tile refactorize the diagonal block itself (no waiting, the right thing when a
// level has a handful of fronts); with hundreds of fronts that is a dozen redundant 64 x 64 factorizations per
// front and per panel, each a latency chain that occupies a quarter of an SM. Here instead:
//   k_bf_diag (one CTA per front, beside the look-ahead update of the next panel): fold + Cholesky of the
//             diagonal block, then its inverse;
//   k_bf_trsm (one CTA per 64-row tile): fold of the previous panel, then tile <- tile * L_bb^-T as a DMMA product.
__global__ void __launch_bounds__(256)
k_bf_diag(const DlbBigFront* __restrict__ descs, double* __restrict__ fronts, double* __restrict__ inv, int step,
          long long* minor, int nf, int gtiles)
{
  extern __shared__ __align__(16) double sm_g[];
  int bid = (int)blockIdx.x;
  if(bid < nf)
  {
    const DlbBigFront f = descs[bid];
    if(step * BF_NB >= f.nc) return;
    // the diagonal block with an identity below it: Cholesky + inverse in one elimination (bf_panel_tile)
    bf_panel_tile(f, fronts, step, 0, minor, sm_g, sm_g + BFP_LD * BF_NB, true, inv + f.inv_off + (size_t)step * 8192, (int*)0);
    return;
  }
  bid -= nf;
  const DlbBigFront f = descs[bid / gtiles];
  const int k0 = step * BF_NB, k1 = k0 + BF_NB;          // panel step+1 starts at column k1
  if(k1 >= f.nc || k0 == 0) return;
  const int i0 = k1 + (bid % gtiles) * 64;
  if(i0 >= f.r) return;
  bf_gemm_tile(f, fronts, i0, k1, f.nc - k1 < BF_NB ? f.nc - k1 : BF_NB, 0, k0, sm_g);
}

#define BFT_LD 68
__global__ void __launch_bounds__(256)
k_bf_trsm(const DlbBigFront* __restrict__ descs, double* __restrict__ fronts, const double* __restrict__ inv, int step, int ptiles)
{
  extern __shared__ __align__(16) double sm_t[];
  double* Bt = sm_t;                                     // the tile, [column][row]: 64 x BFT_LD
  double* Ck[2] = { sm_t + 64 * BFT_LD, sm_t + 64 * BFT_LD + BFW_KC * BFT_LD };   // K chunks of the two operands, [k][row]
  const DlbBigFront f = descs[blockIdx.x / ptiles];
  const int bx = (int)(blockIdx.x % ptiles);
  const int k0 = step * BF_NB;
  if(k0 >= f.nc) return;
  const int nb = f.nc - k0 < BF_NB ? f.nc - k0 : BF_NB;
  const int ld = f.r, r = f.r;
  const int row0 = k0 + nb + bx * 64;
  if(row0 >= r) return;
  const int mine = r - row0 < 64 ? r - row0 : 64;
  double* A = fronts + f.off;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, tt = lane & 3;
  for(int idx = tid; idx < 64 * 64; idx += 256)
  {
    const int j = idx >> 6, i = idx & 63;
    Bt[j * BFT_LD + i] = (j < nb && i < mine) ? A[(size_t)(k0 + j) * ld + row0 + i] : 0.0;
  }
  double acc[8][2];
  if(k0 > 0)
  { // fold: tile -= L[tile rows, previous panel] L[block rows, previous panel]'
#pragma unroll
    for(int c = 0; c < 8; c++) { acc[c][0] = 0.0; acc[c][1] = 0.0; }
    for(int kc = 0; kc < BF_NB; kc += BFW_KC)
    {
      __syncthreads();
      for(int idx = tid; idx < BFW_KC * 64; idx += 256)
      {
        const int k = idx >> 6, i = idx & 63;
        const size_t col = (size_t)(k0 - BF_NB + kc + k) * ld;
        Ck[0][k * BFT_LD + i] = i < mine ? A[col + row0 + i] : 0.0;
        Ck[1][k * BFT_LD + i] = i < nb ? A[col + k0 + i] : 0.0;
      }
      __syncthreads();
#pragma unroll
      for(int k = 0; k < BFW_KC; k += 4)
      {
        const double a0 = Ck[0][(k + tt) * BFT_LD + 8 * w + g];
#pragma unroll
        for(int c = 0; c < 8; c++) bf_dmma(acc[c][0], acc[c][1], a0, Ck[1][(k + tt) * BFT_LD + 8 * c + g]);
      }
    }
    __syncthreads();
#pragma unroll
    for(int c = 0; c < 8; c++)
#pragma unroll
      for(int h = 0; h < 2; h++) Bt[(8 * c + 2 * tt + h) * BFT_LD + 8 * w + g] -= acc[c][h];
  }
  // tile <- tile * X', X = L_bb^-1:  out(i, j) = sum_k tile(i, k) X(j, k); the column-major copy of X gives [k][j] rows
  const double* Xc = inv + f.inv_off + (size_t)step * 8192 + 4096;
#pragma unroll
  for(int c = 0; c < 8; c++) { acc[c][0] = 0.0; acc[c][1] = 0.0; }
  for(int kc = 0; kc < BF_NB; kc += BFW_KC)
  {
    __syncthreads();
    for(int idx = tid; idx < BFW_KC * 64; idx += 256)
    {
      const int k = idx >> 6, j = idx & 63;
      Ck[1][k * BFT_LD + j] = Xc[(size_t)(kc + k) * 64 + j];
    }
    __syncthreads();
#pragma unroll
    for(int k = 0; k < BFW_KC; k += 4)
    {
      const double a0 = Bt[(kc + k + tt) * BFT_LD + 8 * w + g];
#pragma unroll
      for(int c = 0; c < 8; c++) bf_dmma(acc[c][0], acc[c][1], a0, Ck[1][(k + tt) * BFT_LD + 8 * c + g]);
    }
  }
  __syncthreads();
#pragma unroll
  for(int c = 0; c < 8; c++)
#pragma unroll
    for(int h = 0; h < 2; h++) Bt[(8 * c + 2 * tt + h) * BFT_LD + 8 * w + g] = acc[c][h];
  __syncthreads();
  for(int idx = tid; idx < 64 * 64; idx += 256)
  {
    const int j = idx >> 6, i = idx & 63;
    if(j < nb && i < mine) A[(size_t)(k0 + j) * ld + row0 + i] = Bt[j * BFT_LD + i];
  }
}

// Partial Cholesky of the first nc columns of every front of a batch (r x r column-major lower,
// ld = r): afterwards the first nc columns hold L, the trailing block holds the update matrix.
// descs: device array; max_nc / max_r: maxima over the batch (host-side copies of the shapes).
void dlb_bigfront_factor_batch(const DlbBigFront* d_descs, int nfronts, int max_r, int max_nc, double* fronts, double* inv,
                               int* cnt, long long* minor, cudaStream_t st, double* n_launch)
{
  if(nfronts <= 0) return;
  const size_t g_smem = sizeof(double) * BFG_NST * BFG_STAGE, p_smem = sizeof(double) * BFP_LD * BF_NB;
  // k_bf_step / k_bf_diag: the update ring, or the panel tile + the fold chunk
  const size_t f_smem = p_smem + sizeof(double) * BFW_KC * BFW_LD, s_smem = g_smem > f_smem ? g_smem : f_smem;
  const size_t t_smem = sizeof(double) * (64 * BFT_LD + 2 * BFW_KC * BFT_LD);
  static DlbPerDeviceOnce attr_once;
  if(attr_once.first())
  {
    cudaFuncSetAttribute(k_bf_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_smem);
    cudaFuncSetAttribute(k_bf_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s_smem);
    cudaFuncSetAttribute(k_bf_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s_smem);
    cudaFuncSetAttribute(k_bf_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t_smem);
  }
  const int nsteps = (max_nc + BF_NB - 1) / BF_NB;
  // blockIdx.y carries the front: batches beyond the grid limit go in slices
  for(int f0 = 0; f0 < nfronts; f0 += 65535)
  {
    const int nf = nfronts - f0 < 65535 ? nfronts - f0 : 65535;
    const DlbBigFront* d = d_descs + f0;
    // left-looking needs enough (front, row tile) pairs per panel update to fill the GPU; a batch that
    // cannot (one huge front) goes right-looking: after every panel its update of the whole trailing block
    // (front, row tile) pairs of the level decide the schedule (see the top of this file);
    // DOGLEG_GPU_BF_SCHEDULE=left|right|diag forces one (tests)
    const long long pairs = (long long)nf * ((max_r + 63) / 64);
    const char* sched = getenv("DOGLEG_GPU_BF_SCHEDULE");
    const bool force_left = sched && !strcmp(sched, "left"), force_right = sched && !strcmp(sched, "right");
    const bool force_diag = sched && !strcmp(sched, "diag");
    const bool throughput = force_diag || (!force_left && !force_right && pairs > 2 * 296);
    const bool right_looking = !throughput && (force_right || (!force_left && pairs < 2 * 148));
    if(throughput)
    { // (implies left-looking: the Schur complement follows below)
      for(int step = 0; step < nsteps; step++)
      {
        const int below = max_r - step * BF_NB - 1;
        const int ptiles = below > 0 ? (below + 63) / 64 : 0;
        const int rows_next = max_r - (step + 1) * BF_NB;
        const int gtiles = (step >= 1 && step + 1 < nsteps && rows_next > 0) ? (rows_next + 63) / 64 : 0;
        k_bf_diag<<<nf * (1 + gtiles), 256, s_smem, st>>>(d, fronts, inv, step, minor, nf, gtiles);
        if(ptiles > 0) k_bf_trsm<<<nf * ptiles, 256, t_smem, st>>>(d, fronts, inv, step, ptiles);
        if(n_launch) *n_launch += 2;
      }
    }
    else if(!right_looking)
    { // one look-ahead launch per panel: panel `step` (with the fold of panel step-1) beside the update of panel step+1
      for(int step = 0; step < nsteps; step++)
      {
        const int below = max_r - step * BF_NB - 1;               // an upper bound over the batch (nb >= 1)
        const int ptiles = below > 0 ? (below + 63) / 64 : 1;
        const int rows_next = max_r - (step + 1) * BF_NB;         // rows of panel step+1 of the widest front
        const int gtiles = (step >= 1 && step + 1 < nsteps && rows_next > 0) ? (rows_next + 63) / 64 : 0;
        k_bf_step<<<nf * (ptiles + 1 + gtiles), 256, s_smem, st>>>(d, fronts, inv, cnt + f0, step, minor, nf, ptiles + 1, gtiles, 0);
        if(n_launch) *n_launch += 1;
      }
    }
    else
    { // few fronts: right-looking, the trailing update one panel late so that it runs beside the next panel's elimination
      for(int step = 0; step < nsteps; step++)
      {
        const int below = max_r - step * BF_NB - 1;
        const int ptiles = below > 0 ? (below + 63) / 64 : 1;
        const int nt = (below + BF_NB + 63) / 64;                 // tile rows behind the end of panel step-1 (upper bound)
        const int gtiles = step >= 1 ? nt * (nt + 1) / 2 : 0;
        k_bf_step<<<nf * (ptiles + 1 + gtiles), 256, s_smem, st>>>(d, fronts, inv, cnt + f0, step, minor, nf, ptiles + 1, gtiles, 1);
        if(n_launch) *n_launch += 1;
      }
      // the last panel's update of what lies behind the pivots (the update matrix)
      const int below = max_r - (nsteps - 1) * BF_NB - 1;
      if(below > 0)
      {
        const int nt = (below + 63) / 64;
        k_bf_gemm<<<dim3(nt * (nt + 1) / 2, nf), 256, g_smem, st>>>(d, fronts, nsteps - 1, 2);
        if(n_launch) *n_launch += 1;
      }
    }
    const int nt = (max_r - 1 + 63) / 64;                         // trailing tiles of the front with the fewest pivots
    if(nt > 0 && !right_looking)
    {
      k_bf_gemm<<<dim3(nt * (nt + 1) / 2, nf), 256, g_smem, st>>>(d, fronts, 0, 1);
      if(n_launch) *n_launch += 1;
    }

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
