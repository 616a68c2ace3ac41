// dlb_symbolic.h -- host-side symbolic analysis of Jt*Jt' for the device
// multifrontal Cholesky. Runs once per solve (the Jacobian pattern is fixed,
// reference dogleg.c:648-654) and replaces what cholmod_analyze did there:
// ordering, elimination tree, supernodes, L pattern -- plus the GPU-specific
// products: pattern classes of measurement columns, their assignment to
// fronts, extend-add index maps and the level schedule.
#pragma once
#include <cstdint>
#include <vector>

struct DlbSymbolic
{
  int n = 0, m = 0;
  int64_t nnz = 0;

  // ---- pattern classes: measurement columns with identical row lists ----
  int ncls = 0;
  std::vector<int> cls_ptr;      // ncls+1, into cls_rows / cls_loc
  std::vector<int> cls_rows;     // original state indices, ascending (as in Jt->i)
  std::vector<int> cls_of_col;   // m
  std::vector<int> mem_ptr;      // ncls+1, into mem_col
  std::vector<int> mem_col;      // member columns of each class, ascending

  // ---- ordering ----
  std::vector<int> perm;         // perm[k]  = original state eliminated k-th
  std::vector<int> iperm;        // iperm[i] = k
  bool perm_given = false;

  // ---- column structure of L (for parity tests and the cholmod_factor export) ----
  std::vector<int> parent;       // elimination tree over permuted columns, -1 = root
  std::vector<int> colcount;     // nnz of each column of L incl. diagonal

  // ---- supernodes == fronts ----
  int nsuper = 0;
  int nsuper_fundamental = 0;    // before relaxed amalgamation
  std::vector<int> sn_first;     // nsuper+1: columns [sn_first[s], sn_first[s+1]) in permuted order
  std::vector<int> sn_of_col;    // n
  std::vector<int> rows_ptr;     // nsuper+1, into rows / rel
  std::vector<int> rows;         // sorted permuted row indices of the front; the first ncols are its own columns
  std::vector<int> rel;          // for a below-diagonal row: its position in the PARENT front's row list
  std::vector<int> sn_parent;    // -1 = root
  std::vector<int> child_ptr, child_list;   // children of each supernode, ascending
  std::vector<int64_t> front_off;           // nsuper+1 offsets (doubles) of the r x r fronts
  int max_front_rows = 0;

  // ---- level schedule (leaves are level 0) ----
  int nlevels = 0;
  std::vector<int> level_ptr;    // nlevels+1, into level_sn
  std::vector<int> level_sn;
  std::vector<int> sn_level;

  // ---- element (class) assembly ----
  std::vector<int> cls_front;    // front each class is assembled into
  std::vector<int> cls_loc;      // local row index in that front for each class row (same shape as cls_rows)
  std::vector<int> fcls_ptr, fcls_list;     // classes assembled into each front, ascending

  int64_t nnzL() const { int64_t s = 0; for(int c : colcount) s += c; return s; }
  double  flops() const { double s = 0; for(int c : colcount) s += (double)c * c; return s; }
};

// Ap/Ai: CCS of Jt (n rows = states, m columns = measurements), int32, row
// indices ascending inside a column. user_perm may be NULL (own AMD-style
// ordering) or a permutation of 0..n-1 in the cholmod sense (perm[k] = original
// index of the k-th pivot). postorder: reorder by a postorder of the etree
// (always done for own orderings; optional for injected ones so that tests can
// compare against an oracle that uses the permutation verbatim).
// Returns false on malformed input.
bool dlb_symbolic_analyze(DlbSymbolic& S, int n, int m, const int* Ap, const int* Ai,
                          const int* user_perm, bool postorder_user_perm);

// quotient-graph approximate-minimum-degree ordering on the element (class) graph
void dlb_order_amd(int n, int ncls, const std::vector<int>& cls_ptr,
                   const std::vector<int>& cls_rows, std::vector<int>& perm);
