// dlb_taskplan.h -- host-side plan of the streaming passes over Jt: which warp / CTA reads which
// measurement columns for the gradient Jt*x (reference dogleg.c:249-261), |J v|^2 (dogleg.c:262-281)
// and the assembly of Jt*Jt' (done inside CHOLMOD behind dogleg.c:666), where the partial results
// land, and the inverse map that sums them per state in a fixed order. Pure integer work on the
// pattern classes of the symbolic analysis: testable without a GPU (tests/test_taskplan.py).
#pragma once
#include "dlb_symbolic.h"
#include <vector>

#define DLB_LIGHT_MAX 8      // states with fewer (class, slot) pairs are summed by one thread each

// a contiguous range of measurement columns whose pattern classes repeat with period P:
// column j0+i has class cls[i % P] and starts at pos0 + (i / P) * Ktot + koff[i % P]
struct alignas(16) DlbRangeTask
{
  int j0, ncols, P, Ktot;
  unsigned int pos0; int pad[3];
  int cls[4], koff[4];
  long long goff[4];            // where this task's partial gradient of each class goes
};

struct DlbTaskPlan
{
  // tasks: (class, chunk [m0, m1) of its member columns); members index mem_col / mem_pos
  std::vector<int> task_cls, task_m0, task_m1;
  std::vector<int> cls_task_ptr;             // ncls+1: tasks of each class
  std::vector<long long> task_goff;          // the task's k partial gradient entries in gpart
  std::vector<long long> task_Goff;          // the task's k(k+1)/2 partial JtJ entries in Gpart
  std::vector<int> mem_col;                  // LOCAL measurement column of each member (index into x)
  std::vector<unsigned int> mem_pos;         // position of that column's first value in the LOCAL Jt->x
  std::vector<int> big_tasks, small_tasks;   // CTA / warp-pipeline tasks vs lane-group tasks
  // range tasks: runs of consecutive columns with periodic classes, read contiguously
  std::vector<DlbRangeTask> rtasks;
  std::vector<char> ranged;                  // per class: covered by range tasks (not by class tasks) in the gradient
  int range_kmax = 1;
  std::vector<int> gj_big_tasks;             // big tasks of the classes that are not ranged
  // partial gradient blocks of each class: gp_count[c] blocks of k entries from gp_first[c]
  std::vector<int> gp_count;
  std::vector<long long> gp_first;
  // inverse map: state i occurs in the (class, slot) pairs ginv_ptr[i] .. ginv_ptr[i+1]
  std::vector<int> ginv_ptr, ginv_cls;       // ginv_cls: -1 = the class has a single block
  std::vector<long long> ginv_off;           // the slot's entry in the class's first block
  int heavy_threshold = 256;
  std::vector<int> heavy_state, medium_state;
  long long goff = 0, Goff = 0;              // doubles of gpart / Gpart
};

// Jp: column pointers of the global pattern; this plan covers the columns [cbk, cbk + Mk) whose
// values start at local position 0 (row-sharded engines hold a slice; everything else: cbk = 0).
// one_task_per_class: every class gets exactly one task (even with no local member column), so that
// the layout of the partial results (gpart, Gpart) is the same on every rank of a row-sharded solve and
// the ranks can sum them with one all-reduce.
void dlb_build_task_plan(const DlbSymbolic& Y, const int* Jp, int cbk, int Mk, int n_state, int sm_count,
                         bool ranges_enabled, DlbTaskPlan& T, bool one_task_per_class = false);
