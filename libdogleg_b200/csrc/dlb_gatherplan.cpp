// dlb_gatherplan.cpp -- see dlb_gatherplan.h.
//
// Fronts with more than P.heavy children, and all fronts too large for shared memory: instead of
// pulling the children one after the other (a barrier per child, one CTA per front), the
// extend-add is a precomputed gather. The rows of the receiving front are cut into intervals
// such that every run of consecutive rows of every child is a union of whole intervals (the 9
// parameters of a camera, the 6 of a frame ...); a target is a pair of intervals = a
// rectangular block of the front, with the list of its source blocks, children in ascending
// order. Irregular fronts degenerate to 1x1 blocks. Long source lists (the diagonal block of a
// camera receives from every point it sees) are summed in two passes: chunks of P.gchunk sources
// into scratch blocks (pass 1), then the scratch blocks in chunk order (pass 2).
// Small fronts receive into a temporary that k_front_level adds (reused from level to level);
// large fronts (zero-filled beforehand) receive straight into their own storage.
// The forward solve's y(parent) += y(child) uses the same intervals (h x 1 targets).
#include "dlb_gatherplan.h"
#include <algorithm>

namespace {

const long long TAG_TMP = 1ll << 60, TAG_SCR = 1ll << 61;   // relative offsets, fixed up at the end

struct Tgt { long long dst; int ld, h, w; size_t s0, s1; };

// the targets of one level, collected front by front, then emitted as pass 1 + pass 2
struct LevelBuilder
{
  DlbGatherList& G;
  std::vector<Tgt> finals;
  std::vector<long long> fs_base; std::vector<int> fs_ld;
  long long scratch_max = 0;
  explicit LevelBuilder(DlbGatherList& g) : G(g) {}

  void emit(long long dst, int ld, int h, int w, size_t s0, size_t s1)
  {
    G.dst.push_back(dst); G.ld.push_back(ld); G.h.push_back(h); G.w.push_back(w);
    for(size_t k = s0; k < s1; k++) { G.gs_base.push_back(fs_base[k]); G.gs_ld.push_back(fs_ld[k]); }
    G.src_ptr.push_back((long long)G.gs_base.size());
  }
  void flush_level(long long& p0, long long& p1, long long& p2, const DlbGatherParams& P)
  {
    long long scr = 0;
    p0 = (long long)G.dst.size();
    for(Tgt& t : finals)
    { // pass 1: chunks of the long source lists into scratch blocks
      if(t.s1 - t.s0 <= (size_t)P.gsplit) continue;
      const int ww = t.w < 0 ? -t.w : t.w;
      const size_t first_new = fs_base.size();
      for(size_t c0 = t.s0; c0 < t.s1; c0 += P.gchunk)
      {
        emit(TAG_SCR + scr, t.h, t.h, t.w, c0, std::min(t.s1, c0 + (size_t)P.gchunk));
        fs_base.push_back(TAG_SCR + scr); fs_ld.push_back(t.h);
        scr += (long long)t.h * ww;
      }
      t.s0 = first_new; t.s1 = fs_base.size();
    }
    p1 = (long long)G.dst.size();
    for(const Tgt& t : finals) emit(t.dst, t.ld, t.h, t.w, t.s0, t.s1);   // pass 2: the final targets
    p2 = (long long)G.dst.size();
    scratch_max = std::max(scratch_max, scr);
    finals.clear(); fs_base.clear(); fs_ld.clear();
  }
};

} // namespace

void dlb_build_gather_plan(const DlbSymbolic& Y, const DlbGatherParams& P, DlbGatherPlan& out)
{
  out = DlbGatherPlan();
  LevelBuilder FB(out.fronts), SB(out.solve);
  const long long pool_fronts = (long long)Y.front_off[Y.nsuper];
  const long long yrows = (long long)Y.rows.size();
  out.heavy_tmp_off.assign(Y.nsuper, -1);
  out.sg_flag.assign(Y.nsuper, 0);
  out.level_gt_ptr.assign(2 * (size_t)Y.nlevels + 1, 0);
  out.level_sg_ptr.assign(2 * (size_t)Y.nlevels + 1, 0);
  out.level_tmp_size.assign(Y.nlevels, 0);

  struct Src { long long key; long long base; int ld; };
  std::vector<Src> srcs;
  struct YSrc { int iv; long long base; };
  std::vector<YSrc> ysrcs;
  std::vector<int> interval_of, interval_start, seg_iv, seg_off;
  std::vector<char> cut;
  for(int l = 0; l < Y.nlevels; l++)
  {
    long long tmp_level = 0;
    for(int q = Y.level_ptr[l]; q < Y.level_ptr[l+1]; q++)
    {
      const int s = Y.level_sn[q];
      const int r = Y.rows_ptr[s+1] - Y.rows_ptr[s];
      const int nch = Y.child_ptr[s+1] - Y.child_ptr[s];
      const bool large = r > P.small_front_max;
      if(nch == 0 || (nch <= P.heavy && !large)) continue;
      long long dst0;
      if(large) { out.heavy_tmp_off[s] = -2; dst0 = (long long)Y.front_off[s]; }
      else      { out.heavy_tmp_off[s] = tmp_level; dst0 = TAG_TMP + tmp_level; tmp_level += (long long)r * r; }
      out.sg_flag[s] = 1;
      // interval boundaries: wherever a run of some child starts or ends
      cut.assign((size_t)r + 1, 0); cut[0] = cut[r] = 1;
      for(int ch = Y.child_ptr[s]; ch < Y.child_ptr[s+1]; ch++)
      {
        const int c = Y.child_list[ch];
        const int ncc = Y.sn_first[c+1] - Y.sn_first[c], nb = Y.rows_ptr[c+1] - Y.rows_ptr[c] - ncc;
        const int* rel = &Y.rel[Y.rows_ptr[c] + ncc];
        for(int i = 0; i < nb; i++)
        {
          if(i == 0 || rel[i] != rel[i-1] + 1) cut[rel[i]] = 1;
          if(i == nb - 1 || rel[i+1] != rel[i] + 1) cut[rel[i] + 1] = 1;
        }
      }
      interval_of.assign(r, 0); interval_start.clear();
      for(int i = 0; i < r; i++) { if(cut[i]) interval_start.push_back(i); interval_of[i] = (int)interval_start.size() - 1; }
      const long long niv = (long long)interval_start.size();
      interval_start.push_back(r);
      srcs.clear(); ysrcs.clear();
      for(int ch = Y.child_ptr[s]; ch < Y.child_ptr[s+1]; ch++)
      {
        const int c = Y.child_list[ch];
        const int ncc = Y.sn_first[c+1] - Y.sn_first[c], rcc = Y.rows_ptr[c+1] - Y.rows_ptr[c], nb = rcc - ncc;
        const int* rel = &Y.rel[Y.rows_ptr[c] + ncc];
        seg_iv.clear(); seg_off.clear();
        for(int i = 0; i < nb; i++)
          if(i == 0 || interval_of[rel[i]] != interval_of[rel[i-1]]) { seg_iv.push_back(interval_of[rel[i]]); seg_off.push_back(ncc + i); }
        for(size_t a = 0; a < seg_iv.size(); a++)
        {
          ysrcs.push_back({seg_iv[a], (long long)Y.rows_ptr[c] + seg_off[a]});
          for(size_t b = 0; b <= a; b++)          // row interval a >= column interval b (rel is ascending)
            srcs.push_back({(long long)seg_iv[a] * niv + seg_iv[b],
                            (long long)Y.front_off[c] + seg_off[a] + (long long)seg_off[b] * rcc, rcc});
        }
      }
      // stable sort by target block: the children stay in ascending order inside every target
      std::stable_sort(srcs.begin(), srcs.end(), [](const Src& x, const Src& y) { return x.key < y.key; });
      // one target per block; blocks of more than gtile entries are cut into column strips so that
      // a block of a big child (hundreds of rows) is spread over many warps
      for(size_t k0 = 0; k0 < srcs.size(); )
      {
        size_t k1 = k0 + 1;
        while(k1 < srcs.size() && srcs[k1].key == srcs[k0].key) k1++;
        const int ia = (int)(srcs[k0].key / niv), ib = (int)(srcs[k0].key % niv);
        const int h = interval_start[ia+1] - interval_start[ia], w = interval_start[ib+1] - interval_start[ib];
        const bool tri = ia == ib;
        const int nstrips = (int)std::min<long long>(w, ((long long)h * w + P.gtile - 1) / P.gtile);
        const int cw = (w + nstrips - 1) / nstrips;
        for(int j0 = 0; j0 < w; j0 += cw)
        {
          const int ww = std::min(cw, w - j0);
          const int i0 = tri ? j0 : 0;            // a strip of a diagonal block starts at its own diagonal
          FB.finals.push_back({dst0 + interval_start[ia] + i0 + (long long)(interval_start[ib] + j0) * r, r, h - i0,
                               tri ? -ww : ww, FB.fs_base.size(), 0});
          for(size_t k = k0; k < k1; k++)
          { FB.fs_base.push_back(srcs[k].base + i0 + (long long)j0 * srcs[k].ld); FB.fs_ld.push_back(srcs[k].ld); }
          FB.finals.back().s1 = FB.fs_base.size();
        }
        k0 = k1;
      }
      // the forward solve: one h x 1 target per interval of the front's rows
      std::stable_sort(ysrcs.begin(), ysrcs.end(), [](const YSrc& x, const YSrc& y) { return x.iv < y.iv; });
      for(size_t k0 = 0; k0 < ysrcs.size(); )
      {
        size_t k1 = k0 + 1;
        while(k1 < ysrcs.size() && ysrcs[k1].iv == ysrcs[k0].iv) k1++;
        const int ia = ysrcs[k0].iv;
        SB.finals.push_back({(long long)Y.rows_ptr[s] + interval_start[ia], 1, interval_start[ia+1] - interval_start[ia], 1,
                             SB.fs_base.size(), 0});
        for(size_t k = k0; k < k1; k++) { SB.fs_base.push_back(ysrcs[k].base); SB.fs_ld.push_back(1); }
        SB.finals.back().s1 = SB.fs_base.size();
        k0 = k1;
      }
    }
    FB.flush_level(out.level_gt_ptr[2*l], out.level_gt_ptr[2*l+1], out.level_gt_ptr[2*l+2], P);
    SB.flush_level(out.level_sg_ptr[2*l], out.level_sg_ptr[2*l+1], out.level_sg_ptr[2*l+2], P);
    out.level_tmp_size[l] = tmp_level;
    out.pool_tmp = std::max(out.pool_tmp, tmp_level);
  }
  out.pool_scratch = FB.scratch_max; out.solve_scratch = SB.scratch_max;
  auto fix = [&](long long& v) {
    if(v & TAG_SCR)      v = pool_fronts + out.pool_tmp + (v & ~TAG_SCR);
    else if(v & TAG_TMP) v = pool_fronts + (v & ~TAG_TMP);
  };
  for(long long& v : out.fronts.dst) fix(v);
  for(long long& v : out.fronts.gs_base) fix(v);
  auto yfix = [&](long long& v) { if(v & TAG_SCR) v = yrows + (v & ~TAG_SCR); };
  for(long long& v : out.solve.dst) yfix(v);
  for(long long& v : out.solve.gs_base) yfix(v);
}
