// dlb_capi_symbolic.cpp -- C-ABI view of the host-side symbolic analysis
// (include/dogleg_gpu.h) and the cholmod_factor descriptor handed out through
// ctx->factorization (reference dogleg.h:188-195).
#include "dlb_symbolic.h"
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
  default: return -1;
  }
  const long long n = (long long)v->size();
  if(out && cap > 0) std::memcpy(out, v->data(), sizeof(int) * (size_t)std::min(n, cap));
  return n;
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
