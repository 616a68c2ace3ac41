// dlb_device.h -- device-side data structures (passed to kernels by value) and
// the launch functions each .cu file provides to dlb_engine.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "dlb_taskplan.h"      // DlbRangeTask, DLB_LIGHT_MAX
#include "dlb_common.cuh"       // DlbScalars

// per pattern class (single-task classes): slots, first member, members, offset into cls_rows/cls_loc
struct __align__(16) DlbClsInfo { int k, m0, nm, r0; };
// per fused leaf front (dlb_leaf.cu), in the order of level 0
struct __align__(16) DlbLeaf { long long off; int c0, nc, r, rp, fcls0, ncls; };
// everything a small task needs, one aligned 32-byte load
struct __align__(16) DlbSmallTask { int k, m0, nm, r0; long long goff, Goff; };

// Pattern classes, tasks and the inverse map of the gradient reduction
struct DlbSparseDev
{
  int n, m, ncls, ntasks;
  const int* cls_ptr;          // ncls+1
  const int* cls_rows;         // original state index per class slot
  const int* cls_loc;          // local row in the class's front per class slot
  const int* cls_front;        // ncls
  const int* task_cls;         // ntasks
  const int* task_m0;          // member range [m0,m1) into mem_col/mem_pos
  const int* task_m1;
  const long long* task_goff;  // offset of the task's k partial gradient entries
  const long long* task_Goff;  // offset of the task's k(k+1)/2 partial JtJ entries
  const int* mem_col;          // measurement column of each member (index into x)
  const unsigned int* mem_pos; // position of that column's first value in Jt->x
  const int* ginv_ptr;         // n+1: (class, slot) pairs each state occurs in
  const int* ginv_cls;         //   class, or -1 if it has a single task
  const long long* ginv_off;   //   offset in gpart of that slot in the class's first task
  const int* cls_task_ptr;     // ncls+1: tasks of each class (consecutive, partials contiguous)
  // tasks with few member columns are handled by one warp each ("small"), the others by a CTA
  int nbig, nsmall;
  const int* big_tasks;        // all big tasks (assembly)
  const int* small_tasks;
  // gradient / |Jv|^2: classes covered by range tasks are left out of the class-task lists
  int nrange, range_kmax, ngj_big;
  const DlbRangeTask* rtasks;
  const int* gj_big_tasks;
  const int* gp_count;         // ncls: partial gradient blocks per class (stride k, contiguous)
  const DlbSmallTask* small_info;   // nsmall records, same order as small_tasks
  const DlbClsInfo* cls_info;       // ncls records (first task of each class)
  int small_group;                  // lanes per small task: 8, 16 or 32 (>= the longest small column)
  // the small tasks minus those of the classes that the fused leaf-front kernel (dlb_leaf.cu)
  // assembles itself: what k_sparse_assemble_small has to cover in a normal factorization
  int nasm_small;
  const int* asm_small_tasks;
  // ... and the records of the others (the fused ones): the only classes without a Gpart block
  int nfused;
  const DlbSmallTask* fused_info;
  int nheavy, heavy_threshold; // states occurring in >= heavy_threshold (class, slot) pairs
  const int* heavy_state;      //   are reduced by a whole CTA each,
  int nmedium;                 //   those with DLB_LIGHT_MAX .. heavy_threshold-1 pairs by a warp each,
  const int* medium_state;     //   the rest by one thread each
};

// A precomputed gather: target t is an h x |w| block at pool + dst[t] (leading dimension ld[t];
// w < 0: diagonal block, lower triangle only) that receives the sum of its source blocks
// gs_base[src_ptr[t] .. src_ptr[t+1]) (leading dimensions gs_ld), in list order.
struct DlbGather
{
  const long long* dst;
  const int* ld;
  const int* h;
  const int* w;
  const long long* src_ptr;
  const long long* gs_base;
  const int* gs_ld;
};

// Supernodal / multifrontal structure. Front s is an r x r column-major block
// (ld = r) at fronts + front_off[s]; its first ncols columns become the L panel,
// the trailing (r-ncols)^2 block is the update matrix its parent consumes.
struct DlbFrontDev
{
  int n, nsuper;
  const int* sn_first;         // nsuper+1
  const int* rows_ptr;         // nsuper+1
  const int* rows;             // permuted row indices
  const int* rel;              // position of each below row in the parent's row list
  const int* sn_parent;
  const int* child_ptr;        // nsuper+1
  const int* child_list;
  const long long* front_off;  // nsuper+1
  const int* fcls_ptr;         // classes assembled into each front
  const int* fcls_list;
  const int* cls_task_ptr;     // ncls+1: tasks of each class (consecutive)
  const int* level_sn;         // supernodes sorted by level
  const int* perm;             // n
  const DlbLeaf* leaf;         // the fused leaf fronts: level_sn[level_ptr[0] .. + nleaf)
  // flat per-leaf records of the tensor-core leaf kernel (one load level instead of the chain leaf -> class list ->
  // class info -> member positions): leaf q owns leaf_pos/leaf_kl[q * leaf_ps + pair] (offset of the pair's
  // Jacobian values; k | first slot << 8, k == 0 behind the last pair) and leaf_loc[q * leaf_lw + word]
  // (local front row of every class slot, one byte each)
  const unsigned int* leaf_pos; const unsigned int* leaf_kl; const unsigned int* leaf_loc;
  int leaf_ps, leaf_lw;
  int leaf_max_nc;             // most pivot columns of a fused leaf front
  long long ytot;              // solve work vector per right-hand side: one entry per front row + gather scratch
  // Gathered extend-add (fronts with many children, and all fronts too large for shared memory):
  // k_extend_gather -- one warp per receiving block, walking a precomputed, child-ordered list of
  // source blocks (deterministic, no atomics). All offsets are into one pool:
  // [fronts | temporaries of the small gathered fronts | scratch of the two-pass sums].
  const long long* heavy_tmp_off; // nsuper: >= 0 offset of the front's r x r temporary in heavy_tmp (the front adds
                                  //   it); -2 = children were gathered straight into the (large) front; -1 = the
                                  //   front pulls its children itself
  double* heavy_tmp;
  DlbGather fg;                   // extend-add of the update matrices (pool = the fronts pool)
  // The same for the forward solve: y(parent rows) += y(child rows), targets = row intervals
  // (h x 1 blocks), pool = the solve work vector of one right-hand side [rows | scratch]
  DlbGather sg;
  const char* sg_flag;            // nsuper: 1 = the children's y were gathered into this front's rows of ywork
};

// inv_off: offset (doubles) into the engine's buffer of inverted diagonal blocks: block b (pivots 64b .. 64b+63) of the
// front owns 2 x 4096 doubles there, X = L_bb^-1 row-major (X[i*64+j]) followed by column-major (X[j*64+i]), zero-padded
struct DlbBigFront { long long off; int r, nc, col0, sn; long long inv_off; };

// ---- dlb_trial.cu: one trial step (Cauchy, factorization + solves, step, expected improvement) in
// one persistent cooperative kernel, for trees whose fronts all fit in shared memory ----
#define DLB_TRIAL_PART 8
#define DLB_TRIAL_PROF_MAX 62
enum { DLB_TRIAL_CAUCHY = 0, DLB_TRIAL_GN = 1, DLB_TRIAL_INTERP = 2 };
// what the kernels write into mapped pinned memory for the host: the scalars, then the sequence number
struct DlbPublished { DlbScalars sc; unsigned long long seq; };
struct DlbTrial
{
  // level schedule of the tree (device arrays): fronts level_sn[level_ptr[l] .. level_ptr[l+1]), gather
  // targets [level_gt[2l], level_gt[2l+1]) pass 1 / [.., level_gt[2l+2]) pass 2 (level_sg: forward solve),
  // level_tmp[l] doubles of front temporaries to zero
  int nlev; const int* level_ptr; const long long* level_gt; const long long* level_sg; const long long* level_tmp;
  int max_rows, max_cols, any_solve_gather;
  int small_tail;                 // 2 N doubles fit in the kernel's shared memory: every CTA forms the step by itself
  // the point the step starts from (cached vectors are read when have_* is set, written otherwise)
  const double* Jtx; const double* p_from; const double* Gpart; double* cauchy; double* gn;
  double norm2_Jtx, norm2_cauchy, norm2_gn; int have_cauchy, have_gn;
  // the trial point
  double* step; double* p_to; double* h_p_to;     // h_p_to: mapped host mirror of p_to or NULL
  double* fronts; double* ywork; double* zperm;
  double* part; unsigned int* bar; long long* minor; long long* minor_next;   // failure flag of this / the next launch
  DlbScalars* sc; DlbPublished* pub; unsigned long long seq;
  double delta, lambda;
  // element lists: entry d in [eg_ptr[s], eg_ptr[s+1]) of front s lies at row (eg_dst & 0xffff), column
  // (eg_dst >> 16) and is the sum of Gpart[eg_src[q]], q in [eg_sptr[d], eg_sptr[d+1])
  const int* eg_ptr; const unsigned int* eg_dst; const int* eg_sptr; const int* eg_src;
  unsigned long long* prof;       // NULL, or DLB_TRIAL_PROF_MAX + 1 entries: phase time stamps, then their count
};
size_t dlb_trial_smem_bytes(int max_rows);
int  dlb_trial_threads(int max_rows);
int  dlb_trial_max_grid(int max_rows, int sm_count);      // co-resident CTAs (0: the kernel cannot run)
int  dlb_launch_trial(const DlbSparseDev& S, const DlbFrontDev& F, const DlbTrial& T, int grid, cudaStream_t st);

// ---- dlb_sparse.cu ----
void dlb_launch_sparse_grad(const DlbSparseDev& S, const double* Jx, const double* x, double* gpart,
                            double* n2part, double* Jtx, double* part, unsigned int* counter,
                            DlbScalars* sc, int sm_count, cudaStream_t st);
int  dlb_sparse_n2part_size(const DlbSparseDev& S, int sm_count);
struct DlbPublished;
// fused evaluation (one pass over Jt: class blocks + gradient + |x|^2, then the per-state reduction)
int  dlb_launch_sparse_eval_pass(const DlbSparseDev& S, const double* Jx, const double* x, double* Gpart, double* gpart,
                                 double* n2part, int sm_count, cudaStream_t st);           // returns the number of |x|^2 partials
void dlb_launch_sparse_eval_reduce(const DlbSparseDev& S, const double* gpart, const double* n2part, int n2count, double* Jtx,
                                   double* part, unsigned int* counter, DlbScalars* sc, DlbPublished* pub, unsigned long long seq,
                                   int n2_behind_Jtx, int sm_count, cudaStream_t st);
void dlb_launch_sparse_jv(const DlbSparseDev& S, const double* Jx, const double* v, double* part,
                          unsigned int* counter, double* dst, int sm_count, cudaStream_t st);
// |J v|^2 from the assembled class blocks (Gpart of the same Jacobian) instead of a pass over Jt
void dlb_launch_sparse_jv_quad(const DlbSparseDev& S, const double* Jx, const double* Gpart, const double* v, double* part,
                               unsigned int* counter, double* dst, int sm_count, cudaStream_t st);
// all_small != 0: every small task (tests, partial fronts); else only those no leaf kernel covers
void dlb_launch_sparse_assemble(const DlbSparseDev& S, const double* Jx, double* Gpart, int all_small,
                                int sm_count, cudaStream_t st);

// ---- dlb_leaf.cu ---- leaf fronts level_sn[q0..q1), one warp each, assembled straight from Jt->x
void dlb_launch_leaf_fronts(const DlbFrontDev& F, const DlbSparseDev& S, int q0, int q1, const double* Jx,
                            double* fronts, double lambda, long long* minor, int max_rows, int eliminate,
                            int sm_count, cudaStream_t st);
// the same on the FP64 tensor cores, for fronts with <= 4 pivot columns and <= 32 measurement columns
void dlb_launch_leaf_fronts_mma(const DlbFrontDev& F, const DlbSparseDev& S, int q0, int q1, const double* Jx,
                                double* fronts, double lambda, long long* minor, int max_rows, int max_pairs, int eliminate,
                                int sm_count, cudaStream_t st);
void dlb_launch_leaf_solve_fwd(const DlbFrontDev& F, int q0, int q1, const double* fronts, const double* rhs,
                               double* ywork, double* zperm, int nrhs, int sm_count, cudaStream_t st);
void dlb_launch_leaf_solve_bwd(const DlbFrontDev& F, int q0, int q1, const double* fronts, double* zperm,
                               int nrhs, int sm_count, cudaStream_t st);

// ---- dlb_front.cu ----
// one level of the multifrontal factorization: fronts level_sn[l0..l1)
// lambda < 0: elements only (tests / row-sharded partial fronts). skip_elimination != 0: assemble
// (elements, children, lambda) but leave the pivot columns to dlb_bigfront_factor
void dlb_launch_front_level(const DlbFrontDev& F, const DlbSparseDev& S, int l0, int l1,
                            double* fronts, const double* Gpart, double lambda,
                            long long* minor, int max_rows, int skip_elimination, cudaStream_t st);
// dlb_bigfront.cu: blocked tensor-core partial Cholesky of a batch of large fronts (global memory)
// inv: DlbBigFront::inv_off; cnt: one zero-initialised int per front of the batch (which tile stores a diagonal block)
void dlb_bigfront_factor_batch(const DlbBigFront* d_descs, int nfronts, int max_r, int max_nc, double* fronts, double* inv,
                               int* cnt, long long* minor, cudaStream_t st, double* n_launch);
// gather targets [t0,t1): dst = (accumulate ? dst : 0) + sum of the sources
void dlb_launch_extend_gather(const DlbGather& G, long long t0, long long t1, double* pool, int accumulate, cudaStream_t st);
// zero-fill the large fronts of one level (before their children are gathered into them)
void dlb_launch_zero_bigfronts(const DlbBigFront* d_descs, int nfronts, int max_r, double* fronts, cudaStream_t st);
void dlb_launch_solve_fwd_level(const DlbFrontDev& F, int l0, int l1, const double* fronts,
                                const double* rhs /*original order*/, double* ywork,
                                double* zperm, int nrhs, int max_rows, int max_cols, cudaStream_t st);
void dlb_launch_solve_bwd_level(const DlbFrontDev& F, int l0, int l1, const double* fronts,
                                double* zperm, int nrhs, int max_rows, int max_cols, cudaStream_t st);
// densify the assembled (unfactored) matrix for tests: out is n x n row-first
void dlb_launch_fronts_to_dense(const DlbFrontDev& F, const double* fronts, double* out, cudaStream_t st);

// ---- dlb_bigsolve.cu ---- triangular solves with the large fronts of one level (nrhs right-hand sides)
void dlb_launch_bigsolve_fwd(const DlbFrontDev& F, const DlbBigFront* d_descs, int nfronts, int max_r, int max_nc,
                             const double* fronts, const double* inv, const double* rhs, double* ywork, double* zperm, int nrhs,
                             cudaStream_t st);
long long dlb_bigsolve_partial_size(int r, int nc);   // doubles of backward-solve scratch per front and right-hand side
void dlb_launch_bigsolve_bwd(const DlbFrontDev& F, const DlbBigFront* d_descs, int nfronts, int max_r, int max_nc,
                             const double* fronts, const double* inv, double* zperm, double* partial, const long long* d_part_off,
                             int nrhs, cudaStream_t st);

// ---- dlb_dense.cu ----
void dlb_launch_dense_grad(const double* J, const double* x, int M, int N, double* Jtx,
                           double* work, double* part, unsigned int* counter, DlbScalars* sc,
                           int sm_count, cudaStream_t st);
void dlb_launch_dense_jv(const double* J, const double* v, int M, int N, double* work,
                         double* dst, int sm_count, cudaStream_t st);
// front (N x N column-major lower, ld N) = J'J ; work must hold syrk_work_size doubles
size_t dlb_dense_syrk_work_size(int M, int N, int sm_count);
void dlb_launch_dense_syrk(const double* J, int M, int N, double* front, double* work,
                           int sm_count, cudaStream_t st);
// products: unpack the user's JtJ (packed upper / packed lower / full row-first) into a front
void dlb_launch_products_to_front(const double* JtJ, int N, int packed, int upper, double* front, cudaStream_t st);
void dlb_launch_products_xAx(const double* JtJ, int N, int packed, int upper, const double* v,
                             double* dst, cudaStream_t st);
// front -> the reference's factorization_dense layout
void dlb_launch_front_to_reference_layout(const double* front, int N, int packed, int upper, double* out, cudaStream_t st);
// |v|^2 and max|v| of an N-vector into sc->norm2_Jtx / maxabs_Jtx (products path)
void dlb_launch_vec_stats_Jtx(const double* v, int N, double* part, unsigned int* counter, DlbScalars* sc,
                              int sm_count, cudaStream_t st);
void dlb_launch_vec_stats_Jtx_pub(const double* v, int N, double* part, unsigned int* counter, DlbScalars* sc,
                                  DlbPublished* pub, unsigned long long seq, int sm_count, cudaStream_t st);

// ---- dlb_vec.cu ----
void dlb_launch_cauchy(const double* Jtx, int N, double* cauchy, DlbScalars* sc, int sm_count, cudaStream_t st);
void dlb_launch_gn_finish(const double* zperm, const int* perm_or_null, int N, double* gn,
                          double* part, unsigned int* counter, DlbScalars* sc, int sm_count, cudaStream_t st);
void dlb_launch_step(int step_type, double delta, const double* p_from, const double* Jtx,
                     const double* cauchy, const double* gn, int N, double* step, double* p_to,
                     double* part, unsigned int* counter, DlbScalars* sc, int sm_count, cudaStream_t st);
