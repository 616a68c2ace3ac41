// dlb_taskplan.cpp -- see dlb_taskplan.h.
#include "dlb_taskplan.h"
#include <algorithm>
#include <climits>
#include <cstring>

void dlb_build_task_plan(const DlbSymbolic& Y, const int* Jp, int cbk, int Mk, int n_state, int sm_count,
                         bool ranges_enabled, DlbTaskPlan& T, bool one_task_per_class)
{
  T = DlbTaskPlan();
  // tasks: (class, chunk of member columns). The gradient / |Jv|^2 kernels give a task to one warp
  // (cp.async pipeline, ~16 resident warps per SM), the assembly kernel to a CTA: chunks of at
  // least 512 columns, about 8 tasks per SM when the classes are large enough
  const int target = sm_count * 8;
  const int chunk = one_task_per_class ? INT_MAX : std::max(512, (Mk + target - 1) / target);
  std::vector<int>& task_cls = T.task_cls; std::vector<int>& task_m0 = T.task_m0; std::vector<int>& task_m1 = T.task_m1;
  std::vector<int>& cls_task_ptr = T.cls_task_ptr; cls_task_ptr.assign(Y.ncls + 1, 0);
  std::vector<long long>& task_goff = T.task_goff; std::vector<long long>& task_Goff = T.task_Goff;
  long long& goff = T.goff; long long& Goff = T.Goff;
  // member columns of each class that live on this rank (all of them unless row-sharded);
  // mem_col / mem_pos index the LOCAL x and value buffers
  std::vector<int>& lmem_col = T.mem_col; lmem_col.reserve(Mk);
  std::vector<unsigned int>& mem_pos = T.mem_pos; mem_pos.reserve(Mk);
  for(int c = 0; c < Y.ncls; c++)
  {
    const int* mb = Y.mem_col.data() + Y.mem_ptr[c];
    const int* me = Y.mem_col.data() + Y.mem_ptr[c+1];
    const int* lo = std::lower_bound(mb, me, cbk);
    const int* hi = std::lower_bound(mb, me, cbk + Mk);
    const int first = (int)lmem_col.size();
    for(const int* q = lo; q < hi; q++) { lmem_col.push_back(*q - cbk); mem_pos.push_back((unsigned int)(Jp[*q] - Jp[cbk])); }
    const int nmem = (int)lmem_col.size() - first;
    const int k = Y.cls_ptr[c+1] - Y.cls_ptr[c];
    const int nt = std::max(1, (nmem + chunk - 1) / chunk);
    const int per = (nmem + nt - 1) / nt;
    cls_task_ptr[c] = (int)task_cls.size();
    for(int t = 0; t < nt; t++)
    {
      const int m0 = first + t * per, m1 = std::min(first + nmem, m0 + per);
      if(m0 >= m1 && t > 0) break;
      task_cls.push_back(c); task_m0.push_back(m0); task_m1.push_back(std::max(m0, m1));
      task_goff.push_back(goff); task_Goff.push_back(Goff);
      goff += k; Goff += (long long)k * (k + 1) / 2;
    }
  }
  cls_task_ptr[Y.ncls] = (int)task_cls.size();
  const int ntasks = (int)task_cls.size();
  // tasks with few member columns (and short columns) get one warp each instead of a CTA
  const int SMALL_MEMBERS = 32;
  std::vector<int>& big_tasks = T.big_tasks; std::vector<int>& small_tasks = T.small_tasks;
  for(int t = 0; t < ntasks; t++)
  {
    const int c = task_cls[t];
    const bool small = task_m1[t] - task_m0[t] <= SMALL_MEMBERS && Y.cls_ptr[c+1] - Y.cls_ptr[c] <= 32;
    (small ? small_tasks : big_tasks).push_back(t);
  }
  // ---- range tasks for the gradient / |Jv|^2 kernels: maximal runs of consecutive local columns
  // whose classes repeat with a period P <= 4 (calibration: x-row, y-row, x-row, ...), cut into
  // pieces of a few hundred columns; a class is covered ("ranged") only if ALL its local member
  // columns lie in such runs, and a run is only used if all its classes are ranged
  std::vector<DlbRangeTask>& rtasks = T.rtasks;
  std::vector<char>& ranged = T.ranged; ranged.assign(Y.ncls, 0);
  std::vector<int>& gp_count = T.gp_count; gp_count.assign(Y.ncls, 0);
  std::vector<long long>& gp_first = T.gp_first; gp_first.assign(Y.ncls, 0);
  int& range_kmax = T.range_kmax; range_kmax = 1;
  {
    const bool enabled = ranges_enabled;
    struct Run { int j0, n, P; bool ok; };
    std::vector<Run> runs;
    const int Ml = Mk;
    auto cls = [&](int j) { return Y.cls_of_col[cbk + j]; };
    auto klen = [&](int j) { return Jp[cbk + j + 1] - Jp[cbk + j]; };
    for(int j = 0; enabled && j < Ml; )
    {
      int bestP = 0, bestLen = 0;
      for(int P = 1; P <= 4 && j + P <= Ml; P++)
      {
        bool distinct = true;
        for(int a = 0; a < P; a++) for(int b = a + 1; b < P; b++) if(cls(j + a) == cls(j + b)) distinct = false;
        if(!distinct) continue;
        int len = P;
        while(j + len < Ml && cls(j + len) == cls(j + len - P)) len++;
        len -= len % P;
        if(len > bestLen) { bestLen = len; bestP = P; }
      }
      int Ktot = 0;
      for(int i = 0; i < bestP; i++) Ktot += klen(j + i);
      if(bestP > 0 && bestLen >= 64 * bestP && Ktot <= 128 && Ktot > 0) { runs.push_back({j, bestLen, bestP, true}); j += bestLen; }
      else j++;
    }
    std::vector<int> nlocal(Y.ncls, 0), inrun(Y.ncls, 0);
    for(int t = 0; t < ntasks; t++) nlocal[task_cls[t]] += task_m1[t] - task_m0[t];
    for(bool changed = true; changed; )
    {
      changed = false;
      std::fill(inrun.begin(), inrun.end(), 0);
      for(const Run& r : runs) if(r.ok) for(int i = 0; i < r.P; i++) inrun[cls(r.j0 + i)] += r.n / r.P;
      for(Run& r : runs)
        if(r.ok)
          for(int i = 0; i < r.P; i++)
            if(inrun[cls(r.j0 + i)] != nlocal[cls(r.j0 + i)]) { r.ok = false; changed = true; break; }
    }
    const int want = std::max(64, Ml / std::max(1, sm_count * 32));
    for(const Run& r : runs)
    {
      if(!r.ok) continue;
      const int per_periods = std::max(1, want / r.P);
      const int nper = r.n / r.P;
      const int nt = (nper + per_periods - 1) / per_periods, each = (nper + nt - 1) / nt;
      for(int t = 0; t < nt; t++)
      {
        const int q0 = t * each, q1 = std::min(nper, q0 + each);
        if(q0 >= q1) break;
        DlbRangeTask rt; memset(&rt, 0, sizeof(rt));
        rt.j0 = r.j0 + q0 * r.P; rt.ncols = (q1 - q0) * r.P; rt.P = r.P;
        rt.pos0 = (unsigned int)(Jp[cbk + rt.j0] - Jp[cbk]);
        int off = 0;
        for(int i = 0; i < r.P; i++) { rt.cls[i] = cls(r.j0 + i); rt.koff[i] = off; off += klen(r.j0 + i); ranged[rt.cls[i]] = 1; gp_count[rt.cls[i]]++; }
        rt.Ktot = off;
        range_kmax = std::max(range_kmax, off);
        rtasks.push_back(rt);
      }
    }
    // partial gradient blocks: class tasks for the other classes, one block per range task for the ranged ones
    for(int c = 0; c < Y.ncls; c++)
    {
      const int k = Y.cls_ptr[c+1] - Y.cls_ptr[c];
      if(ranged[c]) { gp_first[c] = goff; goff += (long long)k * gp_count[c]; gp_count[c] = 0; }
      else { gp_first[c] = task_goff[cls_task_ptr[c]]; gp_count[c] = cls_task_ptr[c+1] - cls_task_ptr[c]; }
    }
    for(DlbRangeTask& rt : rtasks)
      for(int i = 0; i < rt.P; i++)
      {
        const int c = rt.cls[i], k = Y.cls_ptr[c+1] - Y.cls_ptr[c];
        rt.goff[i] = gp_first[c] + (long long)k * gp_count[c]++;
      }
  }
  std::vector<int>& gj_big_tasks = T.gj_big_tasks;
  for(int t : big_tasks) if(!ranged[task_cls[t]]) gj_big_tasks.push_back(t);

  // inverse map of the gradient: the (class, slot) pairs each state occurs in
  std::vector<int>& ginv_ptr = T.ginv_ptr; ginv_ptr.assign(n_state + 1, 0);
  for(int c = 0; c < Y.ncls; c++)
    for(int q = Y.cls_ptr[c]; q < Y.cls_ptr[c+1]; q++) ginv_ptr[Y.cls_rows[q] + 1]++;
  for(int i = 0; i < n_state; i++) ginv_ptr[i+1] += ginv_ptr[i];
  std::vector<int>& ginv_cls = T.ginv_cls; ginv_cls.assign(ginv_ptr[n_state], 0);
  std::vector<long long>& ginv_off = T.ginv_off; ginv_off.assign(ginv_ptr[n_state], 0);
  {
    std::vector<int> fill(ginv_ptr.begin(), ginv_ptr.end() - 1);
    for(int c = 0; c < Y.ncls; c++)
      for(int q = Y.cls_ptr[c]; q < Y.cls_ptr[c+1]; q++)
      {
        const int at = fill[Y.cls_rows[q]]++;
        ginv_cls[at] = gp_count[c] == 1 ? -1 : c;
        ginv_off[at] = gp_first[c] + (q - Y.cls_ptr[c]);
      }
  }

  // states that occur in many (class, slot) pairs get a whole CTA in the gradient reduction
  const int heavy_threshold = T.heavy_threshold;
  std::vector<int>& heavy_state = T.heavy_state; std::vector<int>& medium_state = T.medium_state;
  for(int i = 0; i < n_state; i++)
  {
    const int cnt = ginv_ptr[i+1] - ginv_ptr[i];
    if(cnt >= heavy_threshold) heavy_state.push_back(i);
    else if(cnt >= DLB_LIGHT_MAX) medium_state.push_back(i);
  }
}
