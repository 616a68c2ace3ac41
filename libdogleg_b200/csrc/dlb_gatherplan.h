// dlb_gatherplan.h -- host-side plan of the extend-add ("assemble the children's update matrices
// into the parent front", what CHOLMOD's supernodal numeric phase does column by column behind
// reference dogleg.c:666) as a precomputed block gather, plus the same plan for the forward
// solve's y(parent) += y(child). Pure integer work on the symbolic structure: testable without a GPU.
#pragma once
#include "dlb_symbolic.h"
#include <vector>

// One gather list: target t is an h x |w| block at pool offset dst[t] with leading dimension
// ld[t] (w < 0: lower-triangular strip, column j starts at row j); it receives the sum of the
// sources gs_base[src_ptr[t] .. src_ptr[t+1]) (same shape, leading dimension gs_ld), in list order.
struct DlbGatherList
{
  std::vector<long long> dst, src_ptr{0}, gs_base;
  std::vector<int> ld, h, w, gs_ld;
};

struct DlbGatherPlan
{
  DlbGatherList fronts;                      // offsets into the pool [fronts | temporaries | scratch]
  DlbGatherList solve;                       // offsets into the solve work vector [rows | scratch]
  std::vector<long long> heavy_tmp_off;      // per supernode: -1 children pulled by the front kernel,
                                             // -2 gathered in place (large front), >= 0 offset of its temporary
  std::vector<char> sg_flag;                 // per supernode: forward solve reads the gathered vector
  // per level l: targets [ptr[2l], ptr[2l+1]) are pass 1 (chunks of long lists into scratch),
  // [ptr[2l+1], ptr[2l+2]) pass 2 (the final blocks); pass 2 may read what pass 1 wrote
  std::vector<long long> level_gt_ptr, level_sg_ptr;
  std::vector<long long> level_tmp_size;     // doubles of temporaries used by each level
  long long pool_tmp = 0, pool_scratch = 0;  // doubles behind the fronts
  long long solve_scratch = 0;               // doubles behind the rows of the solve work vector
};

struct DlbGatherParams
{
  int small_front_max = 158;   // fronts with more rows live in global memory and are gathered in place
  int heavy = 4;               // small fronts with more children than this are gathered too
  int gsplit = 24, gchunk = 16;// source lists longer than gsplit are summed in chunks of gchunk first
  int gtile = 128;             // blocks larger than this are cut into column strips
};

void dlb_build_gather_plan(const DlbSymbolic& Y, const DlbGatherParams& P, DlbGatherPlan& out);
