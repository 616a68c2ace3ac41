// dlb_capi_symbolic.cpp -- C-ABI view of the host-side symbolic analysis
// (include/dogleg_gpu.h) and the cholmod_factor descriptor handed out through
// ctx->factorization (reference dogleg.h:188-195).
#include "dlb_symbolic.h"
#include "dlb_gatherplan.h"
#include "dlb_taskplan.h"
#include "dogleg_internal.h"
#include <cstdlib>
#include <cstring>
#include <algorithm>

struct dlb_symbolic { DlbSymbolic S; };   // layout-compatible handle: the engine stores a DlbSymbolic*

extern "C" dlb_symbolic_t* dlb_symbolic_create(int Nstate, int Nmeas, const int* Jp, const int* Ji,
                                               const int* perm_or_null, int postorder)
{
  dlb_symbolic* h = new dlb_symbolic();
  if(!dlb_symbolic_analyze(h->S, Nstate, Nmeas, Jp, Ji, perm_or_null, postorder != 0))
  {
    delete h;
    dlb_set_error("malformed Jt pattern or permutation");
    return NULL;
  }
  return h;
}
extern "C" void dlb_symbolic_free(dlb_symbolic_t* h) { delete h; }

static const DlbSymbolic& sym(const dlb_symbolic_t* h) { return *reinterpret_cast<const DlbSymbolic*>(h); }

extern "C" void dlb_symbolic_info(const dlb_symbolic_t* h, long long out[8])
{
  const DlbSymbolic& S = sym(h);
  out[0] = S.ncls; out[1] = S.nsuper; out[2] = S.nlevels; out[3] = S.nnzL();
  out[4] = S.max_front_rows; out[5] = S.front_off.empty() ? 0 : S.front_off.back();
  out[6] = (long long)S.flops(); out[7] = (long long)S.rows.size();
}
extern "C" long long dlb_symbolic_get(const dlb_symbolic_t* h, int what, int* out, long long cap)
{
  const DlbSymbolic& S = sym(h);
  const std::vector<int>* v = NULL;
  switch(what)
  {
  case DLB_SYM_PERM: v = &S.perm; break;           case DLB_SYM_PARENT: v = &S.parent; break;
  case DLB_SYM_COLCOUNT: v = &S.colcount; break;   case DLB_SYM_SN_FIRST: v = &S.sn_first; break;
  case DLB_SYM_ROWS_PTR: v = &S.rows_ptr; break;   case DLB_SYM_ROWS: v = &S.rows; break;
  case DLB_SYM_SN_PARENT: v = &S.sn_parent; break; case DLB_SYM_CLS_OF_COL: v = &S.cls_of_col; break;
  case DLB_SYM_CLS_FRONT: v = &S.cls_front; break; case DLB_SYM_SN_LEVEL: v = &S.sn_level; break;
  case DLB_SYM_REL: v = &S.rel; break;             case DLB_SYM_CHILD_PTR: v = &S.child_ptr; break;
  case DLB_SYM_CHILD_LIST: v = &S.child_list; break; case DLB_SYM_LEVEL_PTR: v = &S.level_ptr; break;
  case DLB_SYM_LEVEL_SN: v = &S.level_sn; break;
  case DLB_SYM_CLS_PTR: v = &S.cls_ptr; break;     case DLB_SYM_CLS_ROWS: v = &S.cls_rows; break;
  case DLB_SYM_CLS_LOC: v = &S.cls_loc; break;     case DLB_SYM_MEM_PTR: v = &S.mem_ptr; break;
  case DLB_SYM_MEM_COL: v = &S.mem_col; break;
  default: return -1;
  }
  const long long n = (long long)v->size();
  if(out && cap > 0) std::memcpy(out, v->data(), sizeof(int) * (size_t)std::min(n, cap));
  return n;
}

extern "C" long long dlb_symbolic_front_off(const dlb_symbolic_t* h, long long* out, long long cap)
{
  const DlbSymbolic& S = sym(h);
  const long long n = (long long)S.front_off.size();
  for(long long i = 0; out && i < std::min(n, cap); i++) out[i] = (long long)S.front_off[i];
  return n;
}

// ---- the extend-add / forward-solve gather plan of a symbolic structure (dlb_gatherplan.h) ----
struct dlb_gather_plan { DlbGatherPlan P; long long pool_fronts, solve_rows; int nlevels; };

extern "C" dlb_gather_plan_t* dlb_gather_plan_create(const dlb_symbolic_t* h, int small_front_max, int heavy,
                                                     int gsplit, int gchunk, int gtile)
{
  const DlbSymbolic& S = sym(h);
  DlbGatherParams GP;
  if(small_front_max > 0) GP.small_front_max = small_front_max;
  if(heavy >= 0) GP.heavy = heavy;
  if(gsplit > 0) GP.gsplit = gsplit;
  if(gchunk > 0) GP.gchunk = gchunk;
  if(gtile > 0) GP.gtile = gtile;
  dlb_gather_plan* g = new dlb_gather_plan();
  dlb_build_gather_plan(S, GP, g->P);
  g->pool_fronts = S.front_off.empty() ? 0 : (long long)S.front_off.back();
  g->solve_rows = (long long)S.rows.size();
  g->nlevels = S.nlevels;
  return g;
}
extern "C" void dlb_gather_plan_free(dlb_gather_plan_t* g) { delete g; }
extern "C" void dlb_gather_plan_info(const dlb_gather_plan_t* g, long long out[8])
{
  out[0] = g->pool_fronts; out[1] = g->P.pool_tmp; out[2] = g->P.pool_scratch;
  out[3] = g->solve_rows;  out[4] = g->P.solve_scratch;
  out[5] = (long long)g->P.fronts.dst.size(); out[6] = (long long)g->P.solve.dst.size(); out[7] = g->nlevels;
}
template<class T> static long long copy_wide(const std::vector<T>& v, long long* out, long long cap)
{
  const long long n = (long long)v.size();
  for(long long i = 0; out && i < std::min(n, cap); i++) out[i] = (long long)v[i];
  return n;
}
extern "C" long long dlb_gather_plan_get(const dlb_gather_plan_t* g, int list, int what, long long* out, long long cap)
{
  const DlbGatherList& G = list ? g->P.solve : g->P.fronts;
  switch(what)
  {
  case DLB_GP_DST:       return copy_wide(G.dst, out, cap);
  case DLB_GP_SRC_PTR:   return copy_wide(G.src_ptr, out, cap);
  case DLB_GP_SRC_BASE:  return copy_wide(G.gs_base, out, cap);
  case DLB_GP_LD:        return copy_wide(G.ld, out, cap);
  case DLB_GP_H:         return copy_wide(G.h, out, cap);
  case DLB_GP_W:         return copy_wide(G.w, out, cap);
  case DLB_GP_SRC_LD:    return copy_wide(G.gs_ld, out, cap);
  case DLB_GP_LEVEL_PTR: return copy_wide(list ? g->P.level_sg_ptr : g->P.level_gt_ptr, out, cap);
  case DLB_GP_TMP_OFF:   return copy_wide(g->P.heavy_tmp_off, out, cap);
  case DLB_GP_LEVEL_TMP: return copy_wide(g->P.level_tmp_size, out, cap);
  case DLB_GP_SG_FLAG:   return copy_wide(g->P.sg_flag, out, cap);
  default: return -1;
  }
}

// ---- the streaming-pass plan (dlb_taskplan.h) ----
struct dlb_task_plan { DlbTaskPlan T; };

extern "C" dlb_task_plan_t* dlb_task_plan_create(const dlb_symbolic_t* h, const int* Jp, int col_begin, int ncols,
                                                 int sm_count, int ranges_enabled)
{
  const DlbSymbolic& S = sym(h);
  if(!Jp || col_begin < 0 || ncols < 0 || col_begin + ncols > S.m) { dlb_set_error("task plan: bad column range"); return NULL; }
  dlb_task_plan* t = new dlb_task_plan();
  dlb_build_task_plan(S, Jp, col_begin, ncols, S.n, sm_count > 0 ? sm_count : 148, ranges_enabled != 0, t->T);
  return t;
}
extern "C" void dlb_task_plan_free(dlb_task_plan_t* t) { delete t; }
extern "C" long long dlb_task_plan_get(const dlb_task_plan_t* t, int what, long long* out, long long cap)
{
  const DlbTaskPlan& T = t->T;
  switch(what)
  {
  case DLB_TP_TASK_CLS:     return copy_wide(T.task_cls, out, cap);
  case DLB_TP_TASK_M0:      return copy_wide(T.task_m0, out, cap);
  case DLB_TP_TASK_M1:      return copy_wide(T.task_m1, out, cap);
  case DLB_TP_CLS_TASK_PTR: return copy_wide(T.cls_task_ptr, out, cap);
  case DLB_TP_TASK_GOFF:    return copy_wide(T.task_goff, out, cap);
  case DLB_TP_TASK_GGOFF:   return copy_wide(T.task_Goff, out, cap);
  case DLB_TP_MEM_COL:      return copy_wide(T.mem_col, out, cap);
  case DLB_TP_MEM_POS:      return copy_wide(T.mem_pos, out, cap);
  case DLB_TP_BIG_TASKS:    return copy_wide(T.big_tasks, out, cap);
  case DLB_TP_SMALL_TASKS:  return copy_wide(T.small_tasks, out, cap);
  case DLB_TP_GJ_BIG_TASKS: return copy_wide(T.gj_big_tasks, out, cap);
  case DLB_TP_RANGED:       return copy_wide(T.ranged, out, cap);
  case DLB_TP_GP_COUNT:     return copy_wide(T.gp_count, out, cap);
  case DLB_TP_GP_FIRST:     return copy_wide(T.gp_first, out, cap);
  case DLB_TP_GINV_PTR:     return copy_wide(T.ginv_ptr, out, cap);
  case DLB_TP_GINV_CLS:     return copy_wide(T.ginv_cls, out, cap);
  case DLB_TP_GINV_OFF:     return copy_wide(T.ginv_off, out, cap);
  case DLB_TP_HEAVY:        return copy_wide(T.heavy_state, out, cap);
  case DLB_TP_MEDIUM:       return copy_wide(T.medium_state, out, cap);
  case DLB_TP_SIZES:
  {
    const long long v[4] = {T.goff, T.Goff, (long long)T.range_kmax, (long long)T.heavy_threshold};
    for(long long i = 0; out && i < std::min<long long>(4, cap); i++) out[i] = v[i];
    return 4;
  }
  case DLB_TP_RANGE_TASKS:
  { // 17 values per range task: j0 ncols P Ktot pos0 | cls[4] | koff[4] | goff[4]
    const long long n = (long long)T.rtasks.size() * 17;
    for(size_t r = 0; out && r < T.rtasks.size() && (long long)(r + 1) * 17 <= cap; r++)
    {
      const DlbRangeTask& R = T.rtasks[r];
      long long* o = out + r * 17;
      o[0] = R.j0; o[1] = R.ncols; o[2] = R.P; o[3] = R.Ktot; o[4] = R.pos0;
      for(int i = 0; i < 4; i++) { o[5 + i] = R.cls[i]; o[9 + i] = R.koff[i]; o[13 + i] = R.goff[i]; }
    }
    return n;
  }
  default: return -1;
  }
}

// Supernodal cholmod_factor header: integer structure on the host, numeric
// values in HBM (x == NULL). super/pi/s follow CHOLMOD's supernodal convention:
// supernode k owns columns super[k]..super[k+1]-1 and rows s[pi[k]..pi[k+1]).
extern "C" cholmod_factor* dlb_factor_descriptor_new(const dlb_symbolic_t* h, int n)
{
  cholmod_factor* L = (cholmod_factor*)std::calloc(1, sizeof(cholmod_factor));
  if(!L) return NULL;
  L->n = (size_t)n; L->minor = (size_t)n;
  L->is_ll = 1; L->is_super = 1; L->is_monotonic = 1;
  L->itype = CHOLMOD_INT; L->xtype = CHOLMOD_REAL; L->dtype = CHOLMOD_DOUBLE;
  if(!h) return L;
  const DlbSymbolic& S = sym(h);
  auto dup = [](const std::vector<int>& v) {
    int* p = (int*)std::malloc(sizeof(int) * std::max<size_t>(v.size(), 1));
    if(p && !v.empty()) std::memcpy(p, v.data(), sizeof(int) * v.size());
    return p;
  };
  L->Perm = dup(S.perm); L->IPerm = dup(S.iperm); L->ColCount = dup(S.colcount);
  L->nsuper = (size_t)S.nsuper; L->ssize = S.rows.size();
  L->super = dup(S.sn_first); L->pi = dup(S.rows_ptr); L->s = dup(S.rows);
  L->ordering = S.perm_given ? CHOLMOD_GIVEN : CHOLMOD_AMD;
  size_t xs = 0; int maxc = 0;
  std::vector<int> px(S.nsuper + 1, 0);
  for(int s = 0; s < S.nsuper; s++)
  {
    const int r = S.rows_ptr[s+1] - S.rows_ptr[s], c = S.sn_first[s+1] - S.sn_first[s];
    px[s] = (int)xs; xs += (size_t)r * c; maxc = std::max(maxc, c);
  }
  px[S.nsuper] = (int)xs;
  L->px = dup(px); L->xsize = xs; L->maxcsize = (size_t)maxc;
  return L;
}
extern "C" void dlb_factor_descriptor_free(cholmod_factor* L)
{
  if(!L) return;
  std::free(L->Perm); std::free(L->IPerm); std::free(L->ColCount);
  std::free(L->super); std::free(L->pi); std::free(L->px); std::free(L->s); std::free(L->x);
  std::free(L);
}
