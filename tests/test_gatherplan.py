"""CPU tests of the extend-add / forward-solve gather plan (libdogleg_b200/csrc/dlb_gatherplan.cpp).

The plan is integer work (bit-exact bar): executed here with numpy on integer-valued doubles, level
by level exactly as the engine launches k_extend_gather (pass 1 overwrites scratch chunks, pass 2
accumulates into the targets), it must reproduce the textbook extend-add
    F_parent[rel[i], rel[j]] += U_child[i, j]   (i >= j over the child's below-diagonal rows)
-- what CHOLMOD's supernodal assembly does behind reference dogleg.c:666 -- and, for the forward
solve, y_parent[rel[i]] += y_child[i]. Also checked: no two targets of one pass overlap (they are
written by different warps without atomics), a pass never reads what it writes, every offset stays
inside its pool, nothing outside the gathered fronts is touched, and only the lower triangles of the
children are read (the upper triangles are poisoned with NaN)."""
import ctypes as C

import numpy as np
import pytest

from libdogleg_b200 import ffi


class Plan:
    def __init__(self, H, prob, small_front_max=0, heavy=-1, gsplit=0, gchunk=0, gtile=0):
        L = ffi.load()
        self.L = L
        Jp, Ji = prob.pattern()
        self.h = L.dlb_symbolic_create(prob.N, prob.M, H.as_ip(Jp), H.as_ip(Ji), None, 1)
        assert self.h
        info = (C.c_longlong * 8)()
        L.dlb_symbolic_info(self.h, info)
        self.nsuper, self.nlevels = info[1], info[2]

        def sget(name):
            ln = L.dlb_symbolic_get(self.h, ffi.SYM[name], None, 0)
            out = np.zeros(max(ln, 1), np.int32)
            L.dlb_symbolic_get(self.h, ffi.SYM[name], H.as_ip(out), ln)
            return out[:ln].astype(np.int64)
        for name in ("sn_first", "rows_ptr", "rel", "child_ptr", "child_list", "level_ptr", "level_sn"):
            setattr(self, name, sget(name))
        fo = np.zeros(self.nsuper + 1, np.int64)
        assert L.dlb_symbolic_front_off(self.h, fo.ctypes.data_as(C.POINTER(C.c_longlong)), len(fo)) == self.nsuper + 1
        self.front_off = fo
        self.g = L.dlb_gather_plan_create(self.h, small_front_max, heavy, gsplit, gchunk, gtile)
        assert self.g
        ginfo = (C.c_longlong * 8)()
        L.dlb_gather_plan_info(self.g, ginfo)
        (self.pool_fronts, self.pool_tmp, self.pool_scratch, self.solve_rows, self.solve_scratch,
         self.n_front_targets, self.n_solve_targets, nl) = list(ginfo)
        assert nl == self.nlevels and self.pool_fronts == fo[-1]

        def gget(lst, name):
            ln = L.dlb_gather_plan_get(self.g, lst, ffi.GP[name], None, 0)
            out = np.zeros(max(ln, 1), np.int64)
            L.dlb_gather_plan_get(self.g, lst, ffi.GP[name], out.ctypes.data_as(C.POINTER(C.c_longlong)), ln)
            return out[:ln]
        self.lists = [{k: gget(lst, k) for k in ("dst", "src_ptr", "src_base", "ld", "h", "w", "src_ld", "level_ptr")}
                      for lst in (0, 1)]
        self.tmp_off = gget(0, "tmp_off")
        self.level_tmp = gget(0, "level_tmp")
        self.sg_flag = gget(0, "sg_flag")

    def close(self):
        self.L.dlb_gather_plan_free(self.g)
        self.L.dlb_symbolic_free(self.h)

    def rows_of(self, s):
        return int(self.rows_ptr[s + 1] - self.rows_ptr[s])

    def cols_of(self, s):
        return int(self.sn_first[s + 1] - self.sn_first[s])

    def children(self, s):
        return [int(c) for c in self.child_list[self.child_ptr[s]:self.child_ptr[s + 1]]]


def block_index(h, w):
    """(i, j) of the entries of an h x |w| target; w < 0: lower-triangular strip (i >= j)."""
    tri = w < 0
    w = abs(w)
    J, I = np.meshgrid(np.arange(w), np.arange(h))
    keep = (I >= J) if tri else np.ones_like(I, bool)
    return I[keep].astype(np.int64), J[keep].astype(np.int64)


def run_pass(G, t0, t1, pool, accumulate, limit):
    """Execute targets [t0, t1) the way k_extend_gather does; returns (written, read) index arrays."""
    written, read = [], []
    for t in range(t0, t1):
        I, J = block_index(int(G["h"][t]), int(G["w"][t]))
        acc = np.zeros(len(I))
        for q in range(int(G["src_ptr"][t]), int(G["src_ptr"][t + 1])):
            idx = G["src_base"][q] + I + J * G["src_ld"][q]
            assert idx.min() >= 0 and idx.max() < limit
            acc += pool[idx]
            read.append(idx)
        d = G["dst"][t] + I + J * G["ld"][t]
        assert d.min() >= 0 and d.max() < limit
        written.append(d)
        if accumulate:
            pool[d] += acc
        else:
            pool[d] = acc
    cat = lambda v: np.concatenate(v) if v else np.zeros(0, np.int64)   # noqa: E731
    return cat(written), cat(read)


def check_plan(P, rng):
    total = P.pool_fronts + P.pool_tmp + P.pool_scratch
    pool = np.zeros(total)
    G = P.lists[0]
    assert len(G["dst"]) == P.n_front_targets and len(G["src_ptr"]) == P.n_front_targets + 1
    assert G["level_ptr"][-1] == P.n_front_targets

    def refill(s):
        r = P.rows_of(s)
        F = np.tril(rng.integers(-50, 50, (r, r)).astype(float)) + np.triu(np.full((r, r), np.nan), 1)
        pool[P.front_off[s]:P.front_off[s] + r * r] = F.flatten(order="F")

    n_gathered = n_two_pass = 0
    for l in range(P.nlevels):
        fronts = [int(s) for s in P.level_sn[P.level_ptr[l]:P.level_ptr[l + 1]]]
        # what the engine does before the gathers of a level: zero the temporaries and the large fronts
        pool[P.pool_fronts:P.pool_fronts + P.level_tmp[l]] = 0.0
        for s in fronts:
            if P.tmp_off[s] == -2:
                pool[P.front_off[s]:P.front_off[s + 1]] = 0.0
        expect = pool[:P.pool_fronts + P.pool_tmp].copy()
        tmp_used = 0
        for s in fronts:
            r = P.rows_of(s)
            if P.tmp_off[s] == -1:
                assert P.sg_flag[s] == 0
                continue
            n_gathered += 1
            assert P.sg_flag[s] == 1
            E = np.zeros((r, r))
            for c in P.children(s):
                rc, ncc = P.rows_of(c), P.cols_of(c)
                U = pool[P.front_off[c]:P.front_off[c] + rc * rc].reshape((rc, rc), order="F")
                rel = P.rel[P.rows_ptr[c] + ncc:P.rows_ptr[c] + rc]
                assert (np.diff(rel) > 0).all() and (len(rel) == 0 or rel[-1] < r)
                for j in range(rc - ncc):
                    E[rel[j:], rel[j]] += U[ncc + j:, ncc + j]
            if P.tmp_off[s] >= 0:
                base = P.pool_fronts + P.tmp_off[s]
                tmp_used = max(tmp_used, int(P.tmp_off[s]) + r * r)
            else:
                base = P.front_off[s]
            expect[base:base + r * r] += E.flatten(order="F")
        assert tmp_used <= P.level_tmp[l] <= P.pool_tmp
        p0, p1, p2 = (int(v) for v in G["level_ptr"][2 * l:2 * l + 3])
        n_two_pass += p1 - p0
        w1, r1 = run_pass(G, p0, p1, pool, 0, total)
        assert len(np.unique(w1)) == len(w1), "pass-1 targets overlap"
        assert len(w1) == 0 or w1.min() >= P.pool_fronts + P.pool_tmp, "pass 1 writes outside the scratch"
        assert len(np.intersect1d(w1, r1)) == 0
        w2, r2 = run_pass(G, p1, p2, pool, 1, total)
        assert len(np.unique(w2)) == len(w2), "pass-2 targets overlap"
        assert len(w2) == 0 or w2.max() < P.pool_fronts + P.pool_tmp, "pass 2 writes into the scratch"
        assert len(np.intersect1d(w2, r2)) == 0, "pass 2 reads what it writes"
        got = pool[:P.pool_fronts + P.pool_tmp]
        assert np.array_equal(got, expect, equal_nan=True), f"level {l}: extend-add differs"
        for s in fronts:          # "factor" the fronts of this level: new contents for the parents
            refill(s)
    return n_gathered, n_two_pass


def check_solve_plan(P, rng):
    total = P.solve_rows + P.solve_scratch
    y = np.zeros(total)
    G = P.lists[1]
    assert len(G["dst"]) == P.n_solve_targets
    for l in range(P.nlevels):
        fronts = [int(s) for s in P.level_sn[P.level_ptr[l]:P.level_ptr[l + 1]]]
        expect = y[:P.solve_rows].copy()
        for s in fronts:
            if not P.sg_flag[s]:
                continue
            for c in P.children(s):
                rc, ncc = P.rows_of(c), P.cols_of(c)
                rel = P.rel[P.rows_ptr[c] + ncc:P.rows_ptr[c] + rc]
                np.add.at(expect, P.rows_ptr[s] + rel, y[P.rows_ptr[c] + ncc:P.rows_ptr[c] + rc])
        p0, p1, p2 = (int(v) for v in G["level_ptr"][2 * l:2 * l + 3])
        w1, _ = run_pass(G, p0, p1, y, 0, total)
        assert len(np.unique(w1)) == len(w1) and (len(w1) == 0 or w1.min() >= P.solve_rows)
        w2, r2 = run_pass(G, p1, p2, y, 1, total)
        assert len(np.unique(w2)) == len(w2) and (len(w2) == 0 or w2.max() < P.solve_rows)
        assert len(np.intersect1d(w2, r2)) == 0
        assert np.array_equal(y[:P.solve_rows], expect), f"level {l}: forward-solve gather differs"
        for s in fronts:          # the fronts of this level are "solved": new values for the parents
            y[P.rows_ptr[s]:P.rows_ptr[s + 1]] = rng.integers(-50, 50, P.rows_of(s))


PROBLEMS = {
    "mrcal_small": lambda H: H.Problem.mrcal(3, 8, 6, seed=11),
    "mrcal_frames": lambda H: H.Problem.mrcal(4, 40, 25, seed=2),
    "ba": lambda H: H.Problem.ba(40, 600, 4, 16, 0, seed=4),
    "ba_longrange": lambda H: H.Problem.ba(24, 300, 4, 8, 30, seed=5),
    "random": lambda H: H.Problem.random_sparse(60, 400, 4, seed=7),
}
# engine defaults; everything "large" + two passes + strips forced on small problems; no strips; the trial-kernel strips
SETTINGS = {
    "default": dict(),
    "forced": dict(small_front_max=6, heavy=0, gsplit=3, gchunk=2, gtile=8),
    "all_small": dict(small_front_max=100000, heavy=0, gsplit=2, gchunk=2, gtile=100000),
    # what the engine uses for trees that run in the persistent trial kernel: strips of 32 entries
    "small_tree": dict(gtile=32),
}


@pytest.mark.parametrize("setting", list(SETTINGS))
@pytest.mark.parametrize("name", list(PROBLEMS))
def test_gather_plan_is_the_extend_add(H, name, setting):
    P = Plan(H, PROBLEMS[name](H), **SETTINGS[setting])
    try:
        rng = np.random.default_rng(1)
        n_gathered, n_two_pass = check_plan(P, rng)
        check_solve_plan(P, rng)
        if setting == "forced":
            # every front with children is gathered, and some list was long enough for two passes
            with_children = sum(1 for s in range(P.nsuper) if P.child_ptr[s + 1] > P.child_ptr[s])
            assert n_gathered == with_children
            if name in ("mrcal_frames", "ba"):
                assert n_two_pass > 0
    finally:
        P.close()


@pytest.mark.parametrize("name", ["ba", "random", "ba_longrange"])
def test_gather_plan_deep_tree(H, name, monkeypatch):
    """Fundamental supernodes (no relaxed amalgamation, no multiple elimination): many levels,
    chains of small fronts, irregular row lists that degenerate to 1x1 blocks."""
    monkeypatch.setenv("DOGLEG_GPU_RELAX", "0")
    monkeypatch.setenv("DOGLEG_GPU_MULTI_ELIM", "-1")
    P = Plan(H, PROBLEMS[name](H), **SETTINGS["forced"])
    try:
        rng = np.random.default_rng(2)
        n_gathered, _ = check_plan(P, rng)
        check_solve_plan(P, rng)
        assert P.nlevels >= 4 and n_gathered >= P.nlevels - 1
    finally:
        P.close()


@pytest.mark.parametrize("seed", range(12))
def test_gather_plan_random_small_patterns(H, seed, monkeypatch):
    """Random tiny patterns (empty / repeated / dense columns, untouched states), every front with
    children gathered, fundamental and relaxed supernodes."""
    from test_symbolic import _Pattern
    rng = np.random.default_rng(300 + seed)
    n, m = int(rng.integers(2, 40)), int(rng.integers(1, 80))
    cols = [np.sort(rng.choice(n, size=int(rng.integers(0, min(n, 4) + 1)), replace=False)) for _ in range(m)]
    prob = _Pattern(n, cols)
    for relax in ("0", None):
        if relax is None:
            monkeypatch.delenv("DOGLEG_GPU_RELAX", raising=False)
        else:
            monkeypatch.setenv("DOGLEG_GPU_RELAX", relax)
        P = Plan(H, prob, **SETTINGS["forced"])
        try:
            check_plan(P, rng)
            check_solve_plan(P, rng)
        finally:
            P.close()
