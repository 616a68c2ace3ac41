"""CPU tests of the streaming-pass plan (libdogleg_b200/csrc/dlb_taskplan.cpp): integer work, bit-exact.

The plan decides which task reads which measurement columns of Jt and where the partial results go.
Executed here with numpy on integer-valued doubles, exactly as the kernels consume it
(k_sparse_grad / k_range_grad -> k_sparse_grad_reduce; k_sparse_assemble -> k_front_level), it must
reproduce the reference's own loops: Jt*x (dogleg.c:249-261) and the entries of Jt*Jt' that CHOLMOD
forms behind dogleg.c:666 -- for the whole problem and for a rank's slice of the columns (8e)."""
import ctypes as C

import numpy as np
import pytest

from libdogleg_b200 import ffi

LLP = C.POINTER(C.c_longlong)


class TaskPlan:
    def __init__(self, H, prob, col_begin=0, ncols=None, sm_count=148, ranges=1):
        L = ffi.load()
        self.L = L
        self.Jp, self.Ji = prob.pattern()
        self.N, self.M = prob.N, prob.M
        self.cb = col_begin
        self.nc = prob.M - col_begin if ncols is None else ncols
        self.h = L.dlb_symbolic_create(prob.N, prob.M, H.as_ip(self.Jp), H.as_ip(self.Ji), None, 1)
        assert self.h

        def sget(name):
            ln = L.dlb_symbolic_get(self.h, ffi.SYM[name], None, 0)
            out = np.zeros(max(ln, 1), np.int32)
            L.dlb_symbolic_get(self.h, ffi.SYM[name], H.as_ip(out), ln)
            return out[:ln].astype(np.int64)
        for name in ("perm", "rows_ptr", "rows", "cls_of_col", "cls_front", "cls_ptr", "cls_rows", "cls_loc"):
            setattr(self, name, sget(name))
        self.ncls = len(self.cls_ptr) - 1
        self.t = L.dlb_task_plan_create(self.h, H.as_ip(self.Jp), col_begin, self.nc, sm_count, ranges)
        assert self.t

        def tget(name):
            ln = L.dlb_task_plan_get(self.t, ffi.TP[name], None, 0)
            out = np.zeros(max(ln, 1), np.int64)
            L.dlb_task_plan_get(self.t, ffi.TP[name], out.ctypes.data_as(LLP), ln)
            return out[:ln]
        for name in ffi.TP:
            setattr(self, name, tget(name))
        self.gsize, self.Gsize, self.range_kmax, self.heavy_threshold = (int(v) for v in self.sizes)
        self.rt = self.range_tasks.reshape(-1, 17)

    def close(self):
        self.L.dlb_task_plan_free(self.t)
        self.L.dlb_symbolic_free(self.h)

    def k(self, c):
        return int(self.cls_ptr[c + 1] - self.cls_ptr[c])


def integer_values(P, rng):
    """Integer-valued local Jacobian values and residuals: every sum below is exact in double."""
    nnz_local = int(P.Jp[P.cb + P.nc] - P.Jp[P.cb])
    return rng.integers(-9, 10, nnz_local).astype(float), rng.integers(-9, 10, P.nc).astype(float)


def check_structure(P):
    """Every local column is read exactly once by the class tasks (assembly) and exactly once by the
    gradient's class tasks of the non-ranged classes + the range tasks."""
    ntasks = len(P.task_cls)
    assert sorted(np.concatenate([P.big_tasks, P.small_tasks]).tolist()) == list(range(ntasks))
    seen = np.zeros(P.nc, int)
    for t in range(ntasks):
        c = int(P.task_cls[t])
        assert P.cls_task_ptr[c] <= t < P.cls_task_ptr[c + 1]
        cols = P.mem_col[P.task_m0[t]:P.task_m1[t]]
        assert (np.diff(cols) > 0).all()                          # ascending: fixed summation order
        assert (P.cls_of_col[P.cb + cols] == c).all()
        assert (P.mem_pos[P.task_m0[t]:P.task_m1[t]] == P.Jp[P.cb + cols] - P.Jp[P.cb]).all()
        seen[cols] += 1
    assert (seen == 1).all()
    gseen = np.zeros(P.nc, int)
    for t in range(ntasks):
        if not P.ranged[P.task_cls[t]]:
            gseen[P.mem_col[P.task_m0[t]:P.task_m1[t]]] += 1
    for r in P.rt:
        j0, ncols, per, Ktot, pos0 = (int(v) for v in r[:5])
        assert 1 <= per <= 4 and ncols % per == 0 and 0 < Ktot <= 128
        cls, koff = r[5:9], r[9:13]
        assert pos0 == P.Jp[P.cb + j0] - P.Jp[P.cb]
        off = 0
        for i in range(per):
            assert P.ranged[cls[i]] and koff[i] == off
            off += P.k(int(cls[i]))
        assert off == Ktot <= P.range_kmax
        q = np.arange(ncols)
        assert (P.cls_of_col[P.cb + j0 + q] == cls[q % per]).all()
        assert (P.Jp[P.cb + j0 + q] - P.Jp[P.cb] == pos0 + (q // per) * Ktot + koff[q % per]).all()
        gseen[j0 + q] += 1
    assert (gseen == 1).all()
    # gj_big_tasks: the big tasks of the non-ranged classes
    assert P.gj_big_tasks.tolist() == [int(t) for t in P.big_tasks if not P.ranged[P.task_cls[t]]]
    # the partial blocks of the different classes do not overlap and fit
    spans = sorted((int(P.gp_first[c]), int(P.gp_first[c]) + P.k(c) * int(P.gp_count[c])) for c in range(P.ncls))
    assert all(a1 <= b0 for (_, a1), (b0, _) in zip(spans, spans[1:])) and (not spans or spans[-1][1] <= P.gsize)


def emulate_gradient(P, Jx, x):
    gpart = np.full(P.gsize, np.nan)
    for t in range(len(P.task_cls)):
        c = int(P.task_cls[t])
        k = P.k(c)
        if P.ranged[c]:
            continue
        acc = np.zeros(k)
        for m in range(int(P.task_m0[t]), int(P.task_m1[t])):
            acc += Jx[P.mem_pos[m]:P.mem_pos[m] + k] * x[P.mem_col[m]]
        g0 = int(P.task_goff[t])
        assert g0 == P.gp_first[c] + k * (t - P.cls_task_ptr[c])
        gpart[g0:g0 + k] = acc
    nblocks = np.zeros(P.ncls, int)
    for r in P.rt:
        j0, ncols, per, Ktot, pos0 = (int(v) for v in r[:5])
        for i in range(per):
            c, k = int(r[5 + i]), P.k(int(r[5 + i]))
            acc = np.zeros(k)
            for q in range(i, ncols, per):
                pos = pos0 + (q // per) * Ktot + int(r[9 + i])
                acc += Jx[pos:pos + k] * x[j0 + q]
            g0 = int(r[13 + i])
            assert (g0 - P.gp_first[c]) % max(k, 1) == 0 and P.gp_first[c] <= g0 <= P.gp_first[c] + k * (P.gp_count[c] - 1)
            assert k == 0 or np.isnan(gpart[g0:g0 + k]).all(), "two range tasks share a block"
            gpart[g0:g0 + k] = acc
            nblocks[c] += 1
    for c in range(P.ncls):
        if P.ranged[c]:
            assert nblocks[c] == P.gp_count[c]
        else:
            assert P.gp_count[c] == P.cls_task_ptr[c + 1] - P.cls_task_ptr[c]
    # k_sparse_grad_reduce: per state, its (class, slot) pairs in list order
    Jtx = np.zeros(P.N)
    cnt = np.diff(P.ginv_ptr)
    assert P.heavy.tolist() == [i for i in range(P.N) if cnt[i] >= P.heavy_threshold]
    assert P.medium.tolist() == [i for i in range(P.N) if 8 <= cnt[i] < P.heavy_threshold]
    for i in range(P.N):
        for q in range(int(P.ginv_ptr[i]), int(P.ginv_ptr[i + 1])):
            c = int(P.ginv_cls[q])
            if c < 0:
                Jtx[i] += gpart[P.ginv_off[q]]
            else:
                assert P.gp_count[c] != 1
                k = P.k(c)
                assert P.cls_rows[P.cls_ptr[c] + (P.ginv_off[q] - P.gp_first[c])] == i
                for t in range(int(P.gp_count[c])):
                    Jtx[i] += gpart[P.ginv_off[q] + t * k]
    return Jtx


def emulate_assembly(P, Jx):
    """Gpart per task (k_sparse_assemble), then the element assembly of k_front_level into the
    permuted lower triangle."""
    A = np.zeros((P.N, P.N))
    for t in range(len(P.task_cls)):
        c = int(P.task_cls[t])
        k = P.k(c)
        G = np.zeros((k, k))
        for m in range(int(P.task_m0[t]), int(P.task_m1[t])):
            v = Jx[P.mem_pos[m]:P.mem_pos[m] + k]
            G += np.outer(v, v)
        assert P.task_Goff[t] + k * (k + 1) // 2 <= P.Gsize
        s = int(P.cls_front[c])
        if k == 0:
            assert s == -1
            continue
        rows = P.rows[P.rows_ptr[s]:P.rows_ptr[s + 1]]
        loc = P.cls_loc[P.cls_ptr[c]:P.cls_ptr[c + 1]]
        gr = rows[loc]                                  # permuted global index of every class slot
        for a in range(k):
            for b in range(a + 1):
                A[max(gr[a], gr[b]), min(gr[a], gr[b])] += G[a, b]
    # the task_Goff blocks are disjoint
    spans = sorted((int(P.task_Goff[t]), int(P.task_Goff[t]) + P.k(int(P.task_cls[t])) * (P.k(int(P.task_cls[t])) + 1) // 2)
                   for t in range(len(P.task_cls)))
    assert all(a1 <= b0 for (_, a1), (b0, _) in zip(spans, spans[1:]))
    return A


def direct(P, Jx, x):
    """The reference's loops on the same slice: Jt*x and the lower triangle of Jt*Jt' (permuted)."""
    Jtx = np.zeros(P.N)
    D = np.zeros((P.nc, P.N))
    base = P.Jp[P.cb]
    for j in range(P.nc):
        p0, p1 = P.Jp[P.cb + j] - base, P.Jp[P.cb + j + 1] - base
        rows = P.Ji[P.Jp[P.cb + j]:P.Jp[P.cb + j + 1]]
        Jtx[rows] += Jx[p0:p1] * x[j]
        D[j, rows] = Jx[p0:p1]
    JtJ = D.T @ D
    return Jtx, np.tril(JtJ[np.ix_(P.perm, P.perm)])


PROBLEMS = {
    "mrcal_runs": lambda H: H.Problem.mrcal(3, 5, 150, seed=9),
    "mrcal_split_classes": lambda H: H.Problem.mrcal(2, 2, 700, seed=5),     # > 512 members: several tasks per class
    "mrcal_frames": lambda H: H.Problem.mrcal(4, 40, 25, seed=2),
    "ba": lambda H: H.Problem.ba(30, 400, 4, 12, 20, seed=8),
    "ragged": lambda H: H.Problem.ragged(40, 3000, 7),
    "random": lambda H: H.Problem.random_sparse(60, 400, 4, seed=7),
}


@pytest.mark.parametrize("ranges", [1, 0])
@pytest.mark.parametrize("name", list(PROBLEMS))
def test_task_plan_reproduces_gradient_and_JtJ(H, name, ranges):
    prob = PROBLEMS[name](H)
    for sm_count in (148, 2):
        P = TaskPlan(H, prob, sm_count=sm_count, ranges=ranges)
        try:
            rng = np.random.default_rng(3)
            Jx, x = integer_values(P, rng)
            check_structure(P)
            Jtx_ref, A_ref = direct(P, Jx, x)
            assert np.array_equal(emulate_gradient(P, Jx, x), Jtx_ref)
            assert np.array_equal(emulate_assembly(P, Jx), A_ref)
            if name == "mrcal_runs":
                assert (len(P.rt) > 0) == bool(ranges)          # the x/y runs are found when enabled
            if not ranges:
                assert len(P.rt) == 0 and not P.ranged.any()
                if name == "mrcal_split_classes":
                    assert (P.gp_count > 1).any() and (P.ginv_cls >= 0).any()
        finally:
            P.close()


@pytest.mark.parametrize("name", ["mrcal_runs", "ba", "ragged"])
def test_task_plan_of_a_column_slice(H, name):
    """Row-sharded engines: the plan of a rank's slice produces that slice's partial sums."""
    prob = PROBLEMS[name](H)
    for b, e in H.shard_columns(prob.M, 3, 2):
        P = TaskPlan(H, prob, col_begin=b, ncols=e - b)
        try:
            rng = np.random.default_rng(4)
            Jx, x = integer_values(P, rng)
            check_structure(P)
            Jtx_ref, A_ref = direct(P, Jx, x)
            assert np.array_equal(emulate_gradient(P, Jx, x), Jtx_ref)
            assert np.array_equal(emulate_assembly(P, Jx), A_ref)
        finally:
            P.close()


@pytest.mark.parametrize("seed", range(12))
def test_task_plan_random_small_patterns(H, seed):
    """Random tiny patterns with empty, repeated and dense columns and untouched states."""
    from test_symbolic import _Pattern
    rng = np.random.default_rng(200 + seed)
    n, m = int(rng.integers(1, 20)), int(rng.integers(1, 300))
    base = [np.sort(rng.choice(n, size=int(rng.integers(0, min(n, 5) + 1)), replace=False)) for _ in range(4)]
    cols = []
    for j in range(m):
        r = rng.integers(0, 10)
        cols.append(base[j % 2] if r < 6 else (base[int(rng.integers(0, 4))] if r < 9 else np.arange(n)))
    prob = _Pattern(n, cols)
    for ranges in (1, 0):
        P = TaskPlan(H, prob, ranges=ranges)
        try:
            Jx, x = integer_values(P, rng)
            check_structure(P)
            Jtx_ref, A_ref = direct(P, Jx, x)
            assert np.array_equal(emulate_gradient(P, Jx, x), Jtx_ref)
            assert np.array_equal(emulate_assembly(P, Jx), A_ref)
        finally:
            P.close()
