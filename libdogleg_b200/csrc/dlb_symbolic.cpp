// dlb_symbolic.cpp -- see dlb_symbolic.h.
//
// Reference behaviour being replaced: cholmod_analyze(Jt) at dogleg.c:650-654
// (ordering + etree + column counts of Jt*Jt'), done once per solve. CHOLMOD's
// ordering cannot be reproduced bit for bit; an injected
// permutation gives bit-exact etree / column counts / L pattern instead
// (tests/test_symbolic.py checks them against a brute-force elimination).
//
// Key observation used throughout: JtJ's graph is a union of cliques, one per
// measurement column, and columns with the same row list ("pattern class")
// give the same clique. All symbolic work is done on the classes, never on the
// O(sum nnz_col^2) explicit pattern of JtJ.
#include "dlb_symbolic.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <numeric>
#include <unordered_map>
#include <cstdio>
#include <cstdlib>

namespace {

inline uint64_t mix64(uint64_t z)
{
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// open-addressing table of representatives: find-or-insert by hash with a caller-supplied
// equality test (std::unordered_map<hash, vector> cost 5 s of the 18 s analysis of the
// bundle-adjustment config)
struct RepTable
{
  std::vector<int> slot; uint64_t mask;
  explicit RepTable(size_t expected)
  {
    size_t cap = 64; while(cap < 2 * expected + 16) cap <<= 1;
    slot.assign(cap, -1); mask = cap - 1;
  }
  // returns the stored id equal to the candidate, or stores 'fresh' and returns it
  template<class Eq> int find_or_insert(uint64_t h, int fresh, Eq eq)
  {
    for(uint64_t at = h & mask;; at = (at + 1) & mask)
    {
      const int id = slot[at];
      if(id < 0) { slot[at] = fresh; return fresh; }
      if(eq(id)) return id;
    }
  }
};

// ------------------------------------------------------------------ classes
bool build_classes(DlbSymbolic& S, const int* Ap, const int* Ai)
{
  const int m = S.m, n = S.n;
  S.cls_of_col.assign(m, -1);
  S.cls_ptr.assign(1, 0);
  S.cls_rows.clear();
  RepTable table((size_t)m / 2 + 1024);
  std::vector<int> cls_first_col;           // a representative column per class
  auto same = [&](int ca, int cb) {
    const int la = Ap[ca+1] - Ap[ca], lb = Ap[cb+1] - Ap[cb];
    return la == lb && std::memcmp(Ai + Ap[ca], Ai + Ap[cb], sizeof(int) * la) == 0;
  };
  for(int j = 0; j < m; j++)
  {
    if(Ap[j+1] < Ap[j]) return false;
    for(int q = Ap[j]; q < Ap[j+1]; q++)
    {
      if(Ai[q] < 0 || Ai[q] >= n) return false;
      if(q > Ap[j] && Ai[q] <= Ai[q-1]) return false;   // must be strictly ascending
    }
    // cheap look-behind: the same pattern usually recurs 1..4 columns back
    int found = -1;
    for(int back = 1; back <= 4 && j - back >= 0 && found < 0; back++)
      if(same(j, j - back)) found = S.cls_of_col[j - back];
    if(found < 0)
    {
      uint64_t h = mix64((uint64_t)(Ap[j+1] - Ap[j]));
      for(int q = Ap[j]; q < Ap[j+1]; q++) h = mix64(h ^ (uint64_t)Ai[q]);
      const int fresh = (int)cls_first_col.size();
      if(2 * (size_t)fresh + 16 > table.slot.size())
      { // grow: re-insert the representatives
        RepTable bigger(4 * (size_t)fresh + 1024);
        for(int c = 0; c < fresh; c++)
        {
          const int jc = cls_first_col[c];
          uint64_t hc = mix64((uint64_t)(Ap[jc+1] - Ap[jc]));
          for(int q = Ap[jc]; q < Ap[jc+1]; q++) hc = mix64(hc ^ (uint64_t)Ai[q]);
          bigger.find_or_insert(hc, c, [](int) { return false; });
        }
        table = std::move(bigger);
      }
      found = table.find_or_insert(h, fresh, [&](int c) { return same(j, cls_first_col[c]); });
      if(found == fresh)
      {
        cls_first_col.push_back(j);
        S.cls_rows.insert(S.cls_rows.end(), Ai + Ap[j], Ai + Ap[j+1]);
        S.cls_ptr.push_back((int)S.cls_rows.size());
      }
    }
    S.cls_of_col[j] = found;
  }
  S.ncls = (int)cls_first_col.size();
  S.mem_ptr.assign(S.ncls + 1, 0);
  for(int j = 0; j < m; j++) S.mem_ptr[S.cls_of_col[j] + 1]++;
  for(int c = 0; c < S.ncls; c++) S.mem_ptr[c+1] += S.mem_ptr[c];
  S.mem_col.resize(m);
  std::vector<int> fill(S.mem_ptr.begin(), S.mem_ptr.end() - 1);
  for(int j = 0; j < m; j++) S.mem_col[fill[S.cls_of_col[j]]++] = j;
  return true;
}

} // namespace

// ================================================================ nested dissection
// Used by dlb_order_amd on the quotient graph that is left once the minimum degree has risen
// above a threshold. Level-structure dissection (George): breadth-first levels from a
// pseudo-peripheral variable, the lightest reasonably balanced level is the separator, recurse
// on both sides. The result is a constraint set per variable -- leaves first, every separator
// after the two sub-trees it separates -- that the minimum-degree loop then honours (the
// CAMD idea), so the order inside every set is still minimum-degree.
struct DlbNdOptions { bool enabled; int min_degree, min_vertices, leaf_vertices; };
static DlbNdOptions dlb_nd_options()
{
  // DOGLEG_GPU_ND="min_degree,min_vertices,leaf_vertices" (or "0" to disable)
  DlbNdOptions o{true, 96, 256, 40};
  if(const char* env = getenv("DOGLEG_GPU_ND"))
  {
    int a = 0, b = 0, c = 0;
    const int k = sscanf(env, "%d,%d,%d", &a, &b, &c);
    if(k == 1 && a == 0) o.enabled = false;
    if(k >= 1 && a > 0) o.min_degree = a;
    if(k >= 2 && b > 0) o.min_vertices = b;
    if(k >= 3 && c > 0) o.leaf_vertices = c;
  }
  return o;
}

// variables: principal, alive. Adjacency of v: the alive elements in vpool[v_start[v] .. + v_len[v]),
// whose variable lists are pool[e_start[e] .. + e_len[e]) (entries with nv <= 0 are dead).
// Writes cset[v] in 1..K for every v of 'verts' and returns K.
static int dlb_nested_dissection(const std::vector<int>& verts, const std::vector<int>& nv,
                                 const std::vector<int64_t>& v_start, const std::vector<int>& v_len,
                                 const std::vector<int>& vpool, const std::vector<int64_t>& e_start,
                                 const std::vector<int>& e_len, const std::vector<char>& e_alive,
                                 const std::vector<int>& pool, int leaf_vertices, std::vector<int>& cset)
{
  const int n = (int)nv.size();
  std::vector<int> part(n, -1);             // id of the (sub)graph a variable currently belongs to
  std::vector<int> vstamp(n, -1), estamp(e_alive.size(), -1), level(n, 0);
  int stamp = 0, nparts = 0, nsets = 0;
  // breadth-first search inside part 'pid' from 'root': order + level boundaries
  std::vector<int> order, lstart;
  auto bfs = [&](int root, int pid) {
    stamp++;
    order.clear(); lstart.clear();
    order.push_back(root); vstamp[root] = stamp; level[root] = 0;
    lstart.push_back(0);
    size_t head = 0;
    while(head < order.size())
    {
      const int v = order[head++];
      for(int q = 0; q < v_len[v]; q++)
      {
        const int e = vpool[v_start[v] + q];
        if(!e_alive[e] || estamp[e] == stamp) continue;
        estamp[e] = stamp;
        for(int t = 0; t < e_len[e]; t++)
        {
          const int u = pool[e_start[e] + t];
          if(nv[u] <= 0 || part[u] != pid || vstamp[u] == stamp) continue;
          vstamp[u] = stamp; level[u] = level[v] + 1;
          if(level[u] == (int)lstart.size()) lstart.push_back((int)order.size());
          order.push_back(u);
        }
      }
    }
    lstart.push_back((int)order.size());
  };
  // explicit recursion: a work item is a vertex set; a finished subtree emits its sets in
  // postorder (children before their separator) through 'emit'
  struct Frame { std::vector<int> verts; std::vector<int> sep; int state; std::vector<int> A, B; };
  std::vector<Frame> stack;
  stack.push_back(Frame{verts, {}, 0, {}, {}});
  for(int v : verts) part[v] = 0;
  nparts = 1;
  auto assign = [&](const std::vector<int>& vs) { if(vs.empty()) return; nsets++; for(int v : vs) cset[v] = nsets; };
  while(!stack.empty())
  {
    Frame& f = stack.back();
    if(f.state == 1)
    { // first side done: now the other side, then the separator
      f.state = 2;
      if(!f.B.empty()) { Frame g{std::move(f.B), {}, 0, {}, {}}; stack.push_back(std::move(g)); }
      continue;
    }
    if(f.state == 2) { assign(f.sep); stack.pop_back(); continue; }
    // state 0: decide
    std::vector<int>& V = f.verts;
    if((int)V.size() <= leaf_vertices) { assign(V); stack.pop_back(); continue; }
    const int pid = nparts++;
    for(int v : V) part[v] = pid;
    // component of V[0]; anything else is split off as a sibling
    bfs(V[0], pid);
    if(order.size() < V.size())
    {
      std::vector<int> comp(order), rest;
      for(int v : comp) part[v] = -2;
      for(int v : V) if(part[v] == pid) rest.push_back(v);
      f.sep.clear(); f.state = 1; f.B = std::move(rest);
      stack.push_back(Frame{std::move(comp), {}, 0, {}, {}});
      continue;
    }
    // pseudo-peripheral root: restart from a farthest variable while the depth grows
    int depth = (int)lstart.size() - 1;
    for(int it = 0; it < 4; it++)
    {
      const int far = order.back();
      std::vector<int> keep_order(order), keep_lstart(lstart);
      bfs(far, pid);
      const int d2 = (int)lstart.size() - 1;
      if(d2 <= depth) { if(d2 < depth) { order.swap(keep_order); lstart.swap(keep_lstart); } break; }
      depth = d2;
    }
    const int nl = (int)lstart.size() - 1;
    if(nl < 3) { assign(V); stack.pop_back(); continue; }
    std::vector<long long> w(nl, 0);
    long long total = 0;
    for(int l = 0; l < nl; l++) { for(int q = lstart[l]; q < lstart[l+1]; q++) w[l] += nv[order[q]]; total += w[l]; }
    int best = -1; double best_cost = 0;
    long long before = w[0];
    for(int l = 1; l <= nl - 2; l++)
    {
      const long long after = total - before - w[l];
      const long long small = std::min(before, after);
      if(small > 0)
      {
        // lightest separator, penalised when the two sides are badly balanced
        const double bal = (double)small / (double)(before + after);        // 0 .. 0.5
        const double cost = (double)w[l] * (bal >= 0.25 ? 1.0 : 0.25 / std::max(bal, 1e-3));
        if(best < 0 || cost < best_cost) { best = l; best_cost = cost; }
      }
      before += w[l];
    }
    if(best < 0) { assign(V); stack.pop_back(); continue; }
    std::vector<int> A(order.begin(), order.begin() + lstart[best]);
    std::vector<int> sep(order.begin() + lstart[best], order.begin() + lstart[best+1]);
    std::vector<int> B(order.begin() + lstart[best+1], order.end());
    f.sep = std::move(sep); f.B = std::move(B); f.state = 1;
    std::vector<int>().swap(f.verts);
    stack.push_back(Frame{std::move(A), {}, 0, {}, {}});
  }
  return nsets;
}

// ====================================================================== AMD
// Approximate minimum degree on a quotient graph whose initial elements are the
// pattern classes (each a clique of JtJ). Follows the published
// Amestoy/Davis/Duff scheme: element absorption, approximate external degree,
// aggressive absorption, mass elimination, supervariable detection by hashing;
// plus an up-front compression of states that occur in exactly the same
// classes (e.g. the 6 pose parameters of a frame, the 3 coordinates of a point).
void dlb_order_amd(int n, int ncls, const std::vector<int>& cls_ptr,
                   const std::vector<int>& cls_rows, std::vector<int>& perm)
{
  perm.clear(); perm.reserve(n);
  if(n == 0) return;

  // ---- membership lists (transpose of the class table) and pre-compression ----
  std::vector<int> vptr(n + 1, 0);
  for(int c = 0; c < ncls; c++) for(int q = cls_ptr[c]; q < cls_ptr[c+1]; q++) vptr[cls_rows[q] + 1]++;
  for(int i = 0; i < n; i++) vptr[i+1] += vptr[i];
  std::vector<int> vcls(vptr[n]);
  {
    std::vector<int> fill(vptr.begin(), vptr.end() - 1);
    for(int c = 0; c < ncls; c++) for(int q = cls_ptr[c]; q < cls_ptr[c+1]; q++) vcls[fill[cls_rows[q]]++] = c;
  }
  std::vector<int> rep(n), nv(n, 0);
  std::vector<int> memb_next(n, -1), memb_tail(n);   // chain of states compressed into a representative
  {
    RepTable tab((size_t)n);
    for(int i = 0; i < n; i++)
    {
      uint64_t h = mix64((uint64_t)(vptr[i+1] - vptr[i]) + 12345);
      for(int q = vptr[i]; q < vptr[i+1]; q++) h = mix64(h ^ (uint64_t)vcls[q]);
      const int la = vptr[i+1] - vptr[i];
      const int r = tab.find_or_insert(h, i, [&](int cand2) {
        return la == vptr[cand2+1] - vptr[cand2] &&
               std::memcmp(&vcls[vptr[i]], &vcls[vptr[cand2]], sizeof(int) * la) == 0; });
      if(r == i) memb_tail[i] = i;
      else       { memb_next[memb_tail[r]] = i; memb_tail[r] = i; }
      rep[i] = r; nv[r]++;
    }
  }

  // ---- dense rows go last (they would only slow the graph down) ----
  // initial exact weighted degree of each representative
  std::vector<int> tag(n, -1);
  std::vector<long long> deg0(n, 0);
  for(int i = 0; i < n; i++)
  {
    if(rep[i] != i) continue;
    long long d = 0;
    tag[i] = i;
    for(int q = vptr[i]; q < vptr[i+1]; q++)
    {
      const int c = vcls[q];
      for(int t = cls_ptr[c]; t < cls_ptr[c+1]; t++)
      {
        const int r = rep[cls_rows[t]];
        if(tag[r] != i) { tag[r] = i; d += nv[r]; }
      }
    }
    deg0[i] = d;
  }
  const double dense_thresh = std::max(16.0, 10.0 * std::sqrt((double)n));
  std::vector<char> is_dense(n, 0);
  std::vector<int> dense_list;
  for(int i = 0; i < n; i++)
    if(rep[i] == i && (double)deg0[i] > dense_thresh) { is_dense[i] = 1; dense_list.push_back(i); }
  std::stable_sort(dense_list.begin(), dense_list.end(), [&](int a, int b) { return deg0[a] < deg0[b]; });

  // ---- quotient graph ----
  // element lists live in 'pool'; variable->element lists live in 'vpool' (in place, they only shrink)
  const int max_elems = ncls + n;
  std::vector<int> pool;  pool.reserve((size_t)cls_rows.size() + 4 * (size_t)n);
  std::vector<int64_t> e_start(max_elems, 0);
  std::vector<int> e_len(max_elems, 0);
  std::vector<long long> e_deg(max_elems, 0);
  std::vector<char> e_alive(max_elems, 0);
  std::vector<int> vpool(vptr[n] + n);            // room for one extra entry per variable
  std::vector<int64_t> v_start(n, 0);
  std::vector<int> v_len(n, 0);
  std::fill(tag.begin(), tag.end(), -1);
  {
    int64_t at = 0;
    for(int i = 0; i < n; i++)
    {
      if(rep[i] != i || is_dense[i]) continue;
      v_start[i] = at;
      at += (vptr[i+1] - vptr[i]) + 1;
    }
  }
  int nelems = 0;
  std::vector<int> stampv(n, -1);
  for(int c = 0; c < ncls; c++)
  {
    const int e = nelems++;
    e_start[e] = (int64_t)pool.size();
    long long w = 0; int len = 0;
    for(int t = cls_ptr[c]; t < cls_ptr[c+1]; t++)
    {
      const int r = rep[cls_rows[t]];
      if(is_dense[r] || stampv[r] == c) continue;
      stampv[r] = c;
      pool.push_back(r); len++; w += nv[r];
    }
    e_len[e] = len; e_deg[e] = w; e_alive[e] = len > 0;
    if(len > 0)
      for(int t = 0; t < len; t++)
      {
        const int r = pool[e_start[e] + t];
        vpool[v_start[r] + v_len[r]++] = e;
      }
  }

  // degrees and bucket lists
  long long ntotal = 0;
  for(int i = 0; i < n; i++) if(rep[i] == i && !is_dense[i]) ntotal += nv[i];
  std::vector<long long> deg(n, 0);
  std::fill(tag.begin(), tag.end(), -1);
  for(int i = 0; i < n; i++)
  {
    if(rep[i] != i || is_dense[i]) continue;
    long long d = 0;
    tag[i] = i;
    for(int q = 0; q < v_len[i]; q++)
    {
      const int e = vpool[v_start[i] + q];
      for(int t = 0; t < e_len[e]; t++)
      {
        const int r = pool[e_start[e] + t];
        if(tag[r] != i) { tag[r] = i; d += nv[r]; }
      }
    }
    deg[i] = d;
  }
  const int nbuckets = n + 1;
  std::vector<int> head(nbuckets, -1), nxt(n, -1), prv(n, -1);
  auto bucket_of = [&](long long d) { return (int)std::min<long long>(std::max<long long>(d, 0), n); };
  // constraint sets (nested dissection, below): only variables of the current set are candidates
  std::vector<int> cset(n, 0);
  std::vector<char> inlist(n, 0);
  std::vector<std::vector<int>> set_members;
  int cur_set = 0, nsets = 1;
  auto list_insert = [&](int i) {
    if(cset[i] != cur_set) return;
    const int b = bucket_of(deg[i]);
    nxt[i] = head[b]; prv[i] = -1;
    if(head[b] >= 0) prv[head[b]] = i;
    head[b] = i;
    inlist[i] = 1;
  };
  auto list_remove = [&](int i) {
    if(!inlist[i]) return;
    const int b = bucket_of(deg[i]);
    if(prv[i] >= 0) nxt[prv[i]] = nxt[i]; else head[b] = nxt[i];
    if(nxt[i] >= 0) prv[nxt[i]] = prv[i];
    nxt[i] = prv[i] = -1;
    inlist[i] = 0;
  };
  for(int i = n - 1; i >= 0; i--) if(rep[i] == i && !is_dense[i]) list_insert(i);

  std::vector<int> order; order.reserve(n);       // principal variables in elimination order
  std::vector<int> merged_next(n, -1), merged_tail(n);   // supervariables found during elimination
  for(int i = 0; i < n; i++) merged_tail[i] = i;
  std::vector<int> wstamp(max_elems, -1);
  std::vector<long long> wval(max_elems, 0);
  std::vector<int> Lp_tag(n, -1);
  std::vector<uint64_t> hsh(n, 0);
  std::vector<int> emark(max_elems, -1);
  long long nel = 0;
  int mindeg = 0, stamp = 0, mstamp = 0;
  std::vector<int> survivors;

  // supervariable detection among the survivors of an elimination step (or round), final clamp
  // of their degrees, back into the degree lists
  auto finish_survivors = [&]() {
    // ---- supervariable detection among the survivors ----
  if(survivors.size() > 1)
  {
    std::sort(survivors.begin(), survivors.end(), [&](int a, int b) {
      return hsh[a] != hsh[b] ? hsh[a] < hsh[b] : (v_len[a] != v_len[b] ? v_len[a] < v_len[b] : a < b); });
    size_t a = 0;
    while(a < survivors.size())
    {
      size_t b = a + 1;
      while(b < survivors.size() && hsh[survivors[b]] == hsh[survivors[a]] &&
            v_len[survivors[b]] == v_len[survivors[a]]) b++;
      for(size_t s = a; s < b; s++)
      {
        const int i = survivors[s];
        if(nv[i] <= 0) continue;
        bool marked = false;
        for(size_t u = s + 1; u < b; u++)
        {
          const int j = survivors[u];
          if(nv[j] <= 0 || cset[j] != cset[i]) continue;
          if(!marked)
          {
            mstamp++;
            for(int q = 0; q < v_len[i]; q++) emark[vpool[v_start[i] + q]] = mstamp;
            marked = true;
          }
          bool eq = true;
          for(int q = 0; q < v_len[j] && eq; q++) eq = emark[vpool[v_start[j] + q]] == mstamp;
          if(!eq) continue;
          // j is indistinguishable from i
          nv[i] += nv[j];
          deg[i] -= nv[j];
          nv[j] = 0; v_len[j] = 0;
          merged_next[merged_tail[i]] = j;
          merged_tail[i] = merged_tail[j];
        }
      }
      a = b;
    }
  }

    // ---- finalise ----
  const long long left = ntotal - nel;
  for(int i : survivors)
  {
    if(nv[i] <= 0) continue;
    deg[i] = std::max<long long>(0, std::min<long long>(deg[i], left - nv[i]));
    list_insert(i);
    if(bucket_of(deg[i]) < mindeg) mindeg = bucket_of(deg[i]);
  }
  };
  DlbNdOptions nd = dlb_nd_options();
  bool nd_done = !nd.enabled;
  // multiple elimination (Liu): while the minimum degree is small, ALL variables of the minimum
  // degree bucket that are not adjacent to a pivot of the same round are eliminated before any
  // degree is updated; the touched variables then get their element lists rebuilt and their exact
  // external degree computed once. A bundle adjustment's 1 M points (degree 36, pairwise
  // non-adjacent) go in one round instead of 1 M degree updates of their ~400-element cameras.
  const char* me_env = getenv("DOGLEG_GPU_MULTI_ELIM");
  const long long multi_max_deg = me_env ? atoll(me_env) : 64;
  std::vector<int> dirty_stamp(n, -1), dirty_list, cand, pend_head(n, -1), pend_next, pend_elem, e_round(max_elems, -1);
  int round_stamp = 0;
  while(nel < ntotal)
  {
    while(mindeg < nbuckets && head[mindeg] < 0) mindeg++;
    if(mindeg >= nbuckets)
    { // the current constraint set is exhausted: open the next one
      if(cur_set + 1 >= nsets) break;
      cur_set++;
      for(int v : set_members[cur_set]) if(nv[v] > 0 && !inlist[v]) list_insert(v);
      mindeg = 0;
      continue;
    }
    if(!nd_done && mindeg > nd.min_degree)
    { // Everything cheap is gone (e.g. all the points of a bundle adjustment): what is left is the
      // expensive core. A minimum-degree order of a band/mesh-like core gives a chain-like
      // elimination tree (thousands of sequential levels for the device factorization); nested
      // dissection of the core gives a tree of logarithmic depth instead.
      nd_done = true;
      std::vector<int> alive;
      for(int i = 0; i < n; i++) if(rep[i] == i && !is_dense[i] && nv[i] > 0) alive.push_back(i);
      if((int)alive.size() >= nd.min_vertices)
      {
        for(int v : alive) list_remove(v);
        nsets = 1 + dlb_nested_dissection(alive, nv, v_start, v_len, vpool, e_start, e_len, e_alive, pool,
                                          nd.leaf_vertices, cset);
        set_members.assign(nsets, {});
        for(int v : alive) set_members[cset[v]].push_back(v);
        continue;                         // set 0 is empty now: the branch above opens set 1
      }
    }
    if(mindeg <= multi_max_deg)
    {
      round_stamp++;
      dirty_list.clear(); cand.clear(); pend_next.clear(); pend_elem.clear();
      for(int v = head[mindeg]; v >= 0; v = nxt[v]) cand.push_back(v);
      for(int pv : cand)
      {
        if(dirty_stamp[pv] == round_stamp || nv[pv] <= 0) continue;       // adjacent to a pivot of this round
        list_remove(pv);
        stamp++;
        const int pe2 = nelems++;
        e_start[pe2] = (int64_t)pool.size();
        long long dL = 0;
        Lp_tag[pv] = stamp;
        for(int q = 0; q < v_len[pv]; q++)
        {
          const int e = vpool[v_start[pv] + q];
          if(!e_alive[e]) continue;
          for(int t = 0; t < e_len[e]; t++)
          {
            const int v = pool[e_start[e] + t];
            if(nv[v] <= 0 || Lp_tag[v] == stamp) continue;
            Lp_tag[v] = stamp;
            pool.push_back(v);
            dL += nv[v];
            if(dirty_stamp[v] != round_stamp) { dirty_stamp[v] = round_stamp; dirty_list.push_back(v); list_remove(v); }
            pend_next.push_back(pend_head[v]); pend_elem.push_back(pe2); pend_head[v] = (int)pend_elem.size() - 1;
          }
          e_alive[e] = 0;
        }
        e_len[pe2] = (int)((int64_t)pool.size() - e_start[pe2]);
        e_deg[pe2] = dL; e_alive[pe2] = dL > 0; e_round[pe2] = round_stamp;
        order.push_back(pv);
        nel += nv[pv];
        nv[pv] = 0; v_len[pv] = 0;
      }
      // rebuild the element lists of the touched variables (dead elements out, this round's in);
      // a variable left with a single, new element is eliminated with that pivot at no extra fill
      survivors.clear();
      for(int i : dirty_list)
      {
        int keep = 0;
        for(int q = 0; q < v_len[i]; q++)
        {
          const int e = vpool[v_start[i] + q];
          if(e_alive[e]) vpool[v_start[i] + keep++] = e;
        }
        for(int k = pend_head[i]; k >= 0; k = pend_next[k])
          if(e_alive[pend_elem[k]]) vpool[v_start[i] + keep++] = pend_elem[k];
        pend_head[i] = -1;
        v_len[i] = keep;
        if(keep == 1 && e_round[vpool[v_start[i]]] == round_stamp)
        {
          const int e = vpool[v_start[i]];
          order.push_back(i);
          nel += nv[i];
          e_deg[e] -= nv[i]; if(e_deg[e] <= 0) e_alive[e] = 0;
          nv[i] = 0; v_len[i] = 0;
          continue;
        }
        survivors.push_back(i);
      }
      // exact external degrees and hashes of the survivors
      for(int i : survivors)
      {
        stamp++;
        Lp_tag[i] = stamp;
        long long d = 0; uint64_t h = 0;
        int keep = 0;
        for(int q = 0; q < v_len[i]; q++)
        {
          const int e = vpool[v_start[i] + q];
          if(!e_alive[e]) continue;                      // emptied by a mass elimination above
          vpool[v_start[i] + keep++] = e;
          h += (uint64_t)e * 0x9E3779B97F4A7C15ull;
          for(int t = 0; t < e_len[e]; t++)
          {
            const int v = pool[e_start[e] + t];
            if(nv[v] > 0 && Lp_tag[v] != stamp) { Lp_tag[v] = stamp; d += nv[v]; }
          }
        }
        v_len[i] = keep;
        deg[i] = d; hsh[i] = h;
      }
      finish_survivors();
      continue;
    }
    const int p = head[mindeg];
    list_remove(p);
    stamp++;

    // ---- new element pe = union of p's elements, minus p ----
    const int pe = nelems++;
    e_start[pe] = (int64_t)pool.size();
    long long degLp = 0;
    Lp_tag[p] = stamp;
    for(int q = 0; q < v_len[p]; q++)
    {
      const int e = vpool[v_start[p] + q];
      if(!e_alive[e]) continue;
      for(int t = 0; t < e_len[e]; t++)
      {
        const int v = pool[e_start[e] + t];
        if(nv[v] <= 0 || Lp_tag[v] == stamp) continue;
        Lp_tag[v] = stamp;
        pool.push_back(v);
        degLp += nv[v];
        list_remove(v);
      }
      e_alive[e] = 0;
    }
    const int Lp_len = (int)((int64_t)pool.size() - e_start[pe]);
    e_len[pe] = Lp_len;
    order.push_back(p);
    nel += nv[p];
    const int nvp = nv[p];
    nv[p] = 0;
    (void)nvp;

    // ---- pass 1: w[e] = |Le \ Lp| for every element touching Lp ----
    for(int t = 0; t < Lp_len; t++)
    {
      const int i = pool[e_start[pe] + t];
      for(int q = 0; q < v_len[i]; q++)
      {
        const int e = vpool[v_start[i] + q];
        if(!e_alive[e]) continue;
        if(wstamp[e] != stamp) { wstamp[e] = stamp; wval[e] = e_deg[e]; }
        wval[e] -= nv[i];
      }
    }

    // ---- pass 2: compact element lists, new approximate degrees, mass elimination ----
    survivors.clear();
    const long long nleft = ntotal - nel;
    for(int t = 0; t < Lp_len; t++)
    {
      const int i = pool[e_start[pe] + t];
      int keep = 0; long long d = 0; uint64_t h = 0;
      for(int q = 0; q < v_len[i]; q++)
      {
        const int e = vpool[v_start[i] + q];
        if(!e_alive[e]) continue;
        if(wval[e] <= 0) { e_alive[e] = 0; continue; }          // aggressive absorption
        vpool[v_start[i] + keep++] = e;
        d += wval[e];
        h += (uint64_t)e * 0x9E3779B97F4A7C15ull;
      }
      if(keep == 0)
      { // only the new element is left: eliminate together with p, no extra fill
        order.push_back(i);
        nel += nv[i]; degLp -= nv[i];
        nv[i] = 0; v_len[i] = 0;
        continue;
      }
      vpool[v_start[i] + keep++] = pe;
      v_len[i] = keep;
      h += (uint64_t)pe * 0x9E3779B97F4A7C15ull;
      hsh[i] = h;
      long long dnew = std::min<long long>(deg[i] + degLp - nv[i], d + degLp - nv[i]);
      deg[i] = dnew;        // clamped against nleft once degLp is final
      survivors.push_back(i);
    }
    (void)nleft;

    finish_survivors();
    e_deg[pe] = degLp;
    e_alive[pe] = degLp > 0;
  }

  // ---- expand: principal -> merged supervariable members -> pre-compressed states ----
  std::vector<char> emitted(n, 0);
  auto emit_rep = [&](int r) {
    for(int s = r; s >= 0; s = memb_next[s]) if(!emitted[s]) { emitted[s] = 1; perm.push_back(s); }
  };
  for(int pvt : order)
    for(int v = pvt; v >= 0; v = merged_next[v]) emit_rep(v);
  for(int r : dense_list) emit_rep(r);
  for(int i = 0; i < n; i++) if(!emitted[i]) { emitted[i] = 1; perm.push_back(i); }   // safety net
}

// ===================================================== symbolic factorization
namespace {

// star graph of the permuted class cliques: edge (min row) -> every other row
void build_star(const DlbSymbolic& S, const std::vector<int>& iperm,
                std::vector<int>& col_ptr, std::vector<int>& col_rows,
                std::vector<int>& row_ptr, std::vector<int>& row_cols)
{
  const int n = S.n;
  col_ptr.assign(n + 1, 0); row_ptr.assign(n + 1, 0);
  std::vector<int> cmin(S.ncls, -1);
  for(int c = 0; c < S.ncls; c++)
  {
    int mn = n;
    for(int q = S.cls_ptr[c]; q < S.cls_ptr[c+1]; q++) mn = std::min(mn, iperm[S.cls_rows[q]]);
    cmin[c] = mn;
    if(mn == n) continue;
    const int k = S.cls_ptr[c+1] - S.cls_ptr[c] - 1;
    col_ptr[mn + 1] += k;
    for(int q = S.cls_ptr[c]; q < S.cls_ptr[c+1]; q++)
    {
      const int r = iperm[S.cls_rows[q]];
      if(r != mn) row_ptr[r + 1]++;
    }
  }
  for(int i = 0; i < n; i++) { col_ptr[i+1] += col_ptr[i]; row_ptr[i+1] += row_ptr[i]; }
  col_rows.resize(col_ptr[n]); row_cols.resize(row_ptr[n]);
  std::vector<int> cf(col_ptr.begin(), col_ptr.end() - 1), rf(row_ptr.begin(), row_ptr.end() - 1);
  for(int c = 0; c < S.ncls; c++)
  {
    const int mn = cmin[c];
    if(mn == n) continue;
    for(int q = S.cls_ptr[c]; q < S.cls_ptr[c+1]; q++)
    {
      const int r = iperm[S.cls_rows[q]];
      if(r == mn) continue;
      col_rows[cf[mn]++] = r;
      row_cols[rf[r]++] = mn;
    }
  }
}

void etree_liu(int n, const std::vector<int>& row_ptr, const std::vector<int>& row_cols,
               std::vector<int>& parent)
{
  parent.assign(n, -1);
  std::vector<int> anc(n, -1);
  for(int k = 0; k < n; k++)
    for(int q = row_ptr[k]; q < row_ptr[k+1]; q++)
    {
      int i = row_cols[q];
      while(i != -1 && i < k)
      {
        const int inext = anc[i];
        anc[i] = k;
        if(inext == -1) parent[i] = k;
        i = inext;
      }
    }
}

// Children are visited in ascending order, except that the child with the largest subtree comes
// last: its columns then directly precede the parent's, which is what lets a chain of
// supernodes be amalgamated (below) and keeps the live update matrices small.
void postorder(int n, const std::vector<int>& parent, std::vector<int>& post)
{
  std::vector<int> head(n, -1), next(n, -1), stack, size(n, 1), best(n, -1);
  for(int j = 0; j < n; j++) if(parent[j] >= 0) size[parent[j]] += size[j];       // parent[j] > j
  for(int j = 0; j < n; j++)
    if(parent[j] >= 0 && (best[parent[j]] < 0 || size[j] >= size[best[parent[j]]])) best[parent[j]] = j;
  for(int p = 0; p < n; p++) if(best[p] >= 0) head[p] = best[p];
  for(int j = n - 1; j >= 0; j--) if(parent[j] >= 0 && best[parent[j]] != j) { next[j] = head[parent[j]]; head[parent[j]] = j; }
  post.clear(); post.reserve(n);
  for(int r = 0; r < n; r++)
  {
    if(parent[r] >= 0) continue;
    stack.push_back(r);
    while(!stack.empty())
    {
      const int v = stack.back();
      const int c = head[v];
      if(c < 0) { post.push_back(v); stack.pop_back(); }
      else      { head[v] = next[c]; stack.push_back(c); }
    }
  }
}

} // namespace

bool dlb_symbolic_analyze(DlbSymbolic& S, int n, int m, const int* Ap, const int* Ai,
                          const int* user_perm, bool postorder_user_perm)
{
  S = DlbSymbolic();
  S.n = n; S.m = m; S.nnz = Ap[m];
  if(n <= 0 || m < 0) return false;
  // DOGLEG_GPU_VERBOSE >= 2: wall time of every phase on stderr
  const bool timing = getenv("DOGLEG_GPU_VERBOSE") && atoi(getenv("DOGLEG_GPU_VERBOSE")) >= 2;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if(!timing) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "libdogleg-b200: symbolic %-28s %.3f s\n", what, std::chrono::duration<double>(now - t_last).count());
    t_last = now;
  };
  if(!build_classes(S, Ap, Ai)) return false;
  lap("pattern classes");

  // ---- ordering ----
  if(user_perm)
  {
    S.perm.assign(user_perm, user_perm + n);
    std::vector<char> seen(n, 0);
    for(int k = 0; k < n; k++)
    {
      if(S.perm[k] < 0 || S.perm[k] >= n || seen[S.perm[k]]) return false;
      seen[S.perm[k]] = 1;
    }
    S.perm_given = true;
  }
  else dlb_order_amd(n, S.ncls, S.cls_ptr, S.cls_rows, S.perm);
  S.iperm.assign(n, 0);
  for(int k = 0; k < n; k++) S.iperm[S.perm[k]] = k;
  lap("ordering");

  std::vector<int> col_ptr, col_rows, row_ptr, row_cols;
  if(!user_perm || postorder_user_perm)
  {
    build_star(S, S.iperm, col_ptr, col_rows, row_ptr, row_cols);
    std::vector<int> par, post;
    etree_liu(n, row_ptr, row_cols, par);
    postorder(n, par, post);
    std::vector<int> p2(n);
    for(int j = 0; j < n; j++) p2[j] = S.perm[post[j]];
    S.perm.swap(p2);
    for(int k = 0; k < n; k++) S.iperm[S.perm[k]] = k;
    lap("etree + postorder");
  }
  build_star(S, S.iperm, col_ptr, col_rows, row_ptr, row_cols);
  lap("star graph");

  // ---- one sweep: column etree, column counts, maximal supernodes, row lists ----
  S.parent.assign(n, -1);
  S.colcount.assign(n, 0);
  S.sn_of_col.assign(n, -1);
  S.sn_first.clear(); S.rows_ptr.assign(1, 0); S.rows.clear(); S.sn_parent.clear();
  std::vector<int> mark(n, -1);
  std::vector<int> child_head(n, -1), child_next;     // closed supernodes waiting at their parent column
  std::vector<int> sn_pcol;                            // parent column of each closed supernode
  std::vector<int> extra;
  int open = -1;                                       // index of the open supernode

  auto below_begin = [&](int s, int upto_col) {        // first row-list entry beyond column upto_col
    return S.rows_ptr[s] + (upto_col - S.sn_first[s] + 1);
  };
  auto close_open = [&](int last_col) {
    const int b0 = below_begin(open, last_col);
    const int pcol = b0 < S.rows_ptr[open + 1] ? S.rows[b0] : -1;
    sn_pcol[open] = pcol;
    if(pcol >= 0) { child_next[open] = child_head[pcol]; child_head[pcol] = open; }
    S.parent[last_col] = pcol;
  };

  for(int j = 0; j < n; j++)
  {
    mark[j] = j;
    extra.clear();
    int nb = 0;
    bool chained = false;        // column j-1 (open supernode) has parent j
    if(open >= 0)
    {
      const int b0 = below_begin(open, j - 1);
      if(b0 < S.rows_ptr[open + 1] && S.rows[b0] == j)
      {
        chained = true;
        for(int q = b0 + 1; q < S.rows_ptr[open + 1]; q++) { mark[S.rows[q]] = j; nb++; }
      }
    }
    for(int q = col_ptr[j]; q < col_ptr[j+1]; q++)
    {
      const int r = col_rows[q];
      if(mark[r] != j) { mark[r] = j; extra.push_back(r); }
    }
    for(int t = child_head[j]; t >= 0; t = child_next[t])
    {
      const int last = S.sn_first[t + 1] - 1;
      for(int q = below_begin(t, last); q < S.rows_ptr[t + 1]; q++)
      {
        const int r = S.rows[q];
        if(mark[r] != j) { mark[r] = j; extra.push_back(r); }
      }
    }
    if(chained && extra.empty())
    { // struct(L_j) == struct(L_{j-1}) \ {j}: same supernode
      S.sn_of_col[j] = open;
      S.colcount[j] = S.colcount[j-1] - 1;
      S.parent[j-1] = j;
      S.sn_first[open + 1] = j + 1;
      continue;
    }
    // start a new supernode at j
    int prev_open = open;
    if(open >= 0) close_open(j - 1);
    const int s = (int)S.sn_first.size() - (S.sn_first.empty() ? 0 : 1);
    if(S.sn_first.empty()) S.sn_first.push_back(j); else S.sn_first.back() = j;
    S.sn_first.push_back(j + 1);
    sn_pcol.push_back(-1); child_next.push_back(-1);
    S.rows.push_back(j);
    const size_t base = S.rows.size();
    if(chained)
      for(int q = below_begin(prev_open, j - 1) + 1; q < S.rows_ptr[prev_open + 1]; q++)
      { const int r = S.rows[q]; S.rows.push_back(r); }
    S.rows.insert(S.rows.end(), extra.begin(), extra.end());
    std::sort(S.rows.begin() + base, S.rows.end());
    S.rows_ptr.push_back((int)S.rows.size());
    S.sn_of_col[j] = s;
    S.colcount[j] = 1 + nb + (int)extra.size();
    open = s;
  }
  if(open >= 0) close_open(n - 1);
  S.nsuper = (int)S.sn_first.size() - 1;
  S.nsuper_fundamental = S.nsuper;
  lap("supernodes + row lists");

  // ---- relaxed amalgamation: merge a supernode into its parent when its columns directly
  // precede the parent's (it is the last child in the postorder) and the explicit zeros this
  // adds stay below a fraction of the merged panel. Long chains of narrow supernodes with big
  // fronts (the camera system of a bundle adjustment) become a few wide ones: fewer levels,
  // K >= 64 for the tensor-core updates, far less read-modify-write of the update matrices.
  // colcount / parent keep describing the exact L; the fronts carry the explicit zeros.
  {
    double f_small = 0.25, f_mid = 0.2, f_big = 0.1;
    bool relax = true;
    if(const char* env = getenv("DOGLEG_GPU_RELAX"))
    {
      double a = 0, b = 0, c = 0;
      const int k = sscanf(env, "%lf,%lf,%lf", &a, &b, &c);
      if(k == 1 && a == 0) relax = false;
      if(k == 3) { f_small = a; f_mid = b; f_big = c; }
    }
    if(relax && S.nsuper > 1)
    {
      std::vector<int> g_first, g_last;            // fundamental supernodes [g_first, g_last] of every group
      long long ncg = 0, rg = 0; double zg = 0;
      for(int s2 = 0; s2 < S.nsuper; s2++)
      {
        const long long ncp = S.sn_first[s2+1] - S.sn_first[s2], rp = S.rows_ptr[s2+1] - S.rows_ptr[s2];
        bool merge = false;
        if(s2 > 0 && sn_pcol[s2-1] >= 0 && S.sn_of_col[sn_pcol[s2-1]] == s2)
        {
          const long long below = rg - ncg;
          const double z = zg + (double)ncg * (double)(rp - below);
          const long long nc2 = ncg + ncp, r2 = ncg + rp;
          const double total = (double)nc2 * (double)r2 - 0.5 * (double)nc2 * (double)(nc2 - 1);
          const double lim = nc2 <= 64 ? f_small : (nc2 <= 512 ? f_mid : f_big);
          if(z <= lim * total) { merge = true; zg = z; ncg = nc2; rg = r2; }
        }
        if(merge) g_last.back() = s2;
        else { g_first.push_back(s2); g_last.push_back(s2); ncg = ncp; rg = rp; zg = 0; }
      }
      if((int)g_first.size() < S.nsuper)
      {
        std::vector<int> nf, nrp(1, 0), nrows, npcol;
        nrows.reserve(S.rows.size());
        for(size_t g = 0; g < g_first.size(); g++)
        {
          const int a = g_first[g], b = g_last[g];
          nf.push_back(S.sn_first[a]);
          for(int c = S.sn_first[a]; c < S.sn_first[b+1]; c++) { nrows.push_back(c); S.sn_of_col[c] = (int)g; }
          const int ncb = S.sn_first[b+1] - S.sn_first[b];
          nrows.insert(nrows.end(), S.rows.begin() + S.rows_ptr[b] + ncb, S.rows.begin() + S.rows_ptr[b+1]);
          nrp.push_back((int)nrows.size());
          npcol.push_back(sn_pcol[b]);
        }
        nf.push_back(n);
        S.sn_first.swap(nf); S.rows_ptr.swap(nrp); S.rows.swap(nrows); sn_pcol.swap(npcol);
        S.nsuper = (int)g_first.size();
      }
    }
  }

  lap("relaxed amalgamation");
  // ---- supernode tree, relative indices, levels, front offsets ----
  S.sn_parent.assign(S.nsuper, -1);
  for(int s = 0; s < S.nsuper; s++) if(sn_pcol[s] >= 0) S.sn_parent[s] = S.sn_of_col[sn_pcol[s]];
  S.child_ptr.assign(S.nsuper + 1, 0);
  for(int s = 0; s < S.nsuper; s++) if(S.sn_parent[s] >= 0) S.child_ptr[S.sn_parent[s] + 1]++;
  for(int s = 0; s < S.nsuper; s++) S.child_ptr[s+1] += S.child_ptr[s];
  S.child_list.resize(S.child_ptr[S.nsuper]);
  {
    std::vector<int> fill(S.child_ptr.begin(), S.child_ptr.end() - 1);
    for(int s = 0; s < S.nsuper; s++) if(S.sn_parent[s] >= 0) S.child_list[fill[S.sn_parent[s]]++] = s;
  }
  S.rel.assign(S.rows.size(), -1);
  S.front_off.assign(S.nsuper + 1, 0);
  S.sn_level.assign(S.nsuper, 0);
  S.max_front_rows = 0;
  for(int s = 0; s < S.nsuper; s++)
  {
    const int r = S.rows_ptr[s+1] - S.rows_ptr[s];
    const int c = S.sn_first[s+1] - S.sn_first[s];
    S.max_front_rows = std::max(S.max_front_rows, r);
    S.front_off[s+1] = S.front_off[s] + (int64_t)r * r;
    const int ps = S.sn_parent[s];
    if(ps >= 0)
    {
      int at = S.rows_ptr[ps];
      for(int q = S.rows_ptr[s] + c; q < S.rows_ptr[s+1]; q++)
      {
        while(S.rows[at] != S.rows[q]) at++;
        S.rel[q] = at - S.rows_ptr[ps];
      }
      S.sn_level[ps] = std::max(S.sn_level[ps], S.sn_level[s] + 1);
    }
  }
  S.nlevels = 0;
  for(int s = 0; s < S.nsuper; s++) S.nlevels = std::max(S.nlevels, S.sn_level[s] + 1);
  S.level_ptr.assign(S.nlevels + 1, 0);
  for(int s = 0; s < S.nsuper; s++) S.level_ptr[S.sn_level[s] + 1]++;
  for(int l = 0; l < S.nlevels; l++) S.level_ptr[l+1] += S.level_ptr[l];
  S.level_sn.resize(S.nsuper);
  {
    std::vector<int> fill(S.level_ptr.begin(), S.level_ptr.end() - 1);
    for(int s = 0; s < S.nsuper; s++) S.level_sn[fill[S.sn_level[s]]++] = s;
  }

  lap("tree, rel, levels");
  // ---- assign every class to the front of its first-eliminated state ----
  S.cls_front.assign(S.ncls, -1);
  S.cls_loc.assign(S.cls_rows.size(), -1);
  S.fcls_ptr.assign(S.nsuper + 1, 0);
  for(int c = 0; c < S.ncls; c++)
  {
    int mn = n;
    for(int q = S.cls_ptr[c]; q < S.cls_ptr[c+1]; q++) mn = std::min(mn, S.iperm[S.cls_rows[q]]);
    if(mn == n) continue;                      // empty column: contributes nothing
    const int s = S.sn_of_col[mn];
    S.cls_front[c] = s;
    S.fcls_ptr[s + 1]++;
    const int* rb = &S.rows[S.rows_ptr[s]];
    const int* re = &S.rows[S.rows_ptr[s+1] - 1] + 1;
    for(int q = S.cls_ptr[c]; q < S.cls_ptr[c+1]; q++)
    {
      const int* it = std::lower_bound(rb, re, S.iperm[S.cls_rows[q]]);
      if(it == re || *it != S.iperm[S.cls_rows[q]]) return false;    // cannot happen for a correct symbolic phase
      S.cls_loc[q] = (int)(it - rb);
    }
  }
  for(int s = 0; s < S.nsuper; s++) S.fcls_ptr[s+1] += S.fcls_ptr[s];
  S.fcls_list.resize(S.fcls_ptr[S.nsuper]);
  {
    std::vector<int> fill(S.fcls_ptr.begin(), S.fcls_ptr.end() - 1);
    for(int c = 0; c < S.ncls; c++) if(S.cls_front[c] >= 0) S.fcls_list[fill[S.cls_front[c]]++] = c;
  }
  lap("class -> front");
  return true;
}
