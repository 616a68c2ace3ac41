"""CPU run of the whole device factorization schedule, in numpy, from nothing but the host-side
structures the engine uploads: pattern classes -> fronts (cls_front / cls_loc), supernodes and row
lists, relative indices, the level schedule. Each front is assembled from its classes' J J' blocks,
receives its children's update matrices through `rel` (what k_front_level / k_extend_gather do),
loses its pivot columns by a dense partial Cholesky, and the triangular solves walk the same
levels. The result must agree with LAPACK on the permuted Jt*Jt' + lambda*I: L to 1e-12, the
solution of (JtJ + lambda I) u = Jt x (reference dogleg.c:839-866) to 1e-10 -- so a wrong index
anywhere in the symbolic phase fails here, without a GPU, and not only as a wrong answer of a kernel."""
import ctypes as C

import numpy as np
import pytest

from libdogleg_b200 import ffi


def structures(H, prob, perm=None):
    L = ffi.load()
    Jp, Ji = prob.pattern()
    h = L.dlb_symbolic_create(prob.N, prob.M, H.as_ip(Jp), H.as_ip(Ji),
                              H.as_ip(perm) if perm is not None else None, 1)
    assert h

    def sget(name):
        ln = L.dlb_symbolic_get(h, ffi.SYM[name], None, 0)
        out = np.zeros(max(ln, 1), np.int32)
        L.dlb_symbolic_get(h, ffi.SYM[name], H.as_ip(out), ln)
        return out[:ln].astype(np.int64)
    S = {name: sget(name) for name in ffi.SYM}
    L.dlb_symbolic_free(h)
    S["Jp"], S["Ji"] = Jp, Ji
    return S


def multifrontal(S, N, Jx, lam):
    """Returns (L dense in the permuted order, list of factored fronts)."""
    nsuper = len(S["sn_first"]) - 1
    rows_of = lambda s: S["rows"][S["rows_ptr"][s]:S["rows_ptr"][s + 1]]          # noqa: E731
    ncols = lambda s: int(S["sn_first"][s + 1] - S["sn_first"][s])               # noqa: E731
    fronts = [None] * nsuper
    ncls = len(S["cls_ptr"]) - 1
    # classes of every front, ascending (fcls_list)
    by_front = [[] for _ in range(nsuper)]
    for c in range(ncls):
        if S["cls_front"][c] >= 0:
            by_front[S["cls_front"][c]].append(c)
    Ld = np.zeros((N, N))
    nlevels = len(S["level_ptr"]) - 1
    done = np.zeros(nsuper, bool)
    for l in range(nlevels):
        for s in S["level_sn"][S["level_ptr"][l]:S["level_ptr"][l + 1]]:
            s = int(s)
            r, nc = len(rows_of(s)), ncols(s)
            A = np.zeros((r, r))
            for c in by_front[s]:                                   # element assembly
                k = int(S["cls_ptr"][c + 1] - S["cls_ptr"][c])
                loc = S["cls_loc"][S["cls_ptr"][c]:S["cls_ptr"][c + 1]]
                for j in S["mem_col"][S["mem_ptr"][c]:S["mem_ptr"][c + 1]]:
                    v = Jx[S["Jp"][j]:S["Jp"][j] + k]
                    A[np.ix_(loc, loc)] += np.outer(v, v)
            for c in S["child_list"][S["child_ptr"][s]:S["child_ptr"][s + 1]]:   # extend-add
                c = int(c)
                assert done[c], "a child is scheduled after its parent"
                rc, ncc = len(rows_of(c)), ncols(c)
                rel = S["rel"][S["rows_ptr"][c] + ncc:S["rows_ptr"][c] + rc]
                assert (rows_of(s)[rel] == rows_of(c)[ncc:]).all()
                A[np.ix_(rel, rel)] += fronts[c][ncc:, ncc:]
            A[np.arange(nc), np.arange(nc)] += lam
            # partial Cholesky of the nc pivot columns
            L11 = np.linalg.cholesky(A[:nc, :nc])
            L21 = np.linalg.solve(L11, A[nc:, :nc].T).T
            F = np.zeros((r, r))
            F[:nc, :nc] = L11
            F[nc:, :nc] = L21
            F[nc:, nc:] = A[nc:, nc:] - L21 @ L21.T
            fronts[s] = F
            done[s] = True
            c0 = int(S["sn_first"][s])
            Ld[np.ix_(rows_of(s), np.arange(c0, c0 + nc))] = F[:, :nc]
    assert done.all()
    return np.tril(Ld), fronts


def solve_with_fronts(S, N, fronts, rhs):
    """Forward substitution up the levels (y of a front = its rows of the permuted rhs plus the
    children's contributions through rel), backward down the levels."""
    nsuper = len(fronts)
    rows_of = lambda s: S["rows"][S["rows_ptr"][s]:S["rows_ptr"][s + 1]]          # noqa: E731
    ncols = lambda s: int(S["sn_first"][s + 1] - S["sn_first"][s])               # noqa: E731
    b = rhs[S["perm"]]
    y = [None] * nsuper
    z = np.zeros(N)
    nlevels = len(S["level_ptr"]) - 1
    for l in range(nlevels):
        for s in S["level_sn"][S["level_ptr"][l]:S["level_ptr"][l + 1]]:
            s = int(s)
            r, nc = len(rows_of(s)), ncols(s)
            v = np.zeros(r)
            v[:nc] = b[rows_of(s)[:nc]]
            for c in S["child_list"][S["child_ptr"][s]:S["child_ptr"][s + 1]]:
                c = int(c)
                ncc = ncols(c)
                rel = S["rel"][S["rows_ptr"][c] + ncc:S["rows_ptr"][c + 1]]
                v[rel] += y[c][ncc:]
            F = fronts[s]
            v[:nc] = np.linalg.solve(F[:nc, :nc], v[:nc])
            v[nc:] -= F[nc:, :nc] @ v[:nc]
            y[s] = v
            z[rows_of(s)[:nc]] = v[:nc]
    for l in range(nlevels - 1, -1, -1):
        for s in S["level_sn"][S["level_ptr"][l]:S["level_ptr"][l + 1]]:
            s = int(s)
            rws, nc = rows_of(s), ncols(s)
            F = fronts[s]
            t = z[rws[:nc]] - F[nc:, :nc].T @ z[rws[nc:]]
            z[rws[:nc]] = np.linalg.solve(F[:nc, :nc].T, t)
    u = np.zeros(N)
    u[S["perm"]] = z
    return u


CASES = {
    "mrcal": (lambda H: H.Problem.mrcal(3, 8, 6, seed=11), {}),
    "mrcal_frames": (lambda H: H.Problem.mrcal(4, 30, 10, seed=2), {}),
    "ba": (lambda H: H.Problem.ba(30, 400, 4, 12, 20, seed=8), {}),
    "ba_nd": (lambda H: H.Problem.ba(60, 800, 4, 24, 0, seed=4), {"DOGLEG_GPU_ND": "30,16,6"}),
    "ba_fundamental": (lambda H: H.Problem.ba(30, 400, 4, 12, 20, seed=8),
                       {"DOGLEG_GPU_RELAX": "0", "DOGLEG_GPU_MULTI_ELIM": "-1"}),
    "ragged": (lambda H: H.Problem.ragged(40, 3000, 7), {}),
    "random": (lambda H: H.Problem.random_sparse(60, 400, 4, seed=7), {}),
}


@pytest.mark.parametrize("lam", [0.0, 1e-3])
@pytest.mark.parametrize("name", list(CASES))
def test_level_schedule_factorizes_and_solves(H, monkeypatch, name, lam):
    mk, env = CASES[name]
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    prob = mk(H)
    S = structures(H, prob)
    N = prob.N
    rng = np.random.default_rng(5)
    nnz = int(S["Jp"][-1])
    Jx = rng.uniform(-1, 1, nnz)
    x = rng.uniform(-1, 1, prob.M)
    D = np.zeros((prob.M, N))
    for j in range(prob.M):
        D[j, S["Ji"][S["Jp"][j]:S["Jp"][j + 1]]] = Jx[S["Jp"][j]:S["Jp"][j + 1]]
    A = D.T @ D + lam * np.eye(N)
    Lref = np.linalg.cholesky(A[np.ix_(S["perm"], S["perm"])])
    Lgot, fronts = multifrontal(S, N, Jx, lam)
    scale = np.abs(Lref).max()
    assert np.abs(Lgot - Lref).max() <= 1e-12 * scale * max(1.0, np.linalg.cond(A) ** 0.5)
    # the stored pattern covers L: exact zeros outside the supernodal row lists
    g = D.T @ x
    u = solve_with_fronts(S, N, fronts, g)
    uref = np.linalg.solve(A, g)
    assert np.abs(u - uref).max() <= 1e-10 * max(1.0, np.abs(uref).max()) * max(1.0, np.linalg.cond(A) ** 0.5)


@pytest.mark.parametrize("seed", range(12))
def test_level_schedule_random_small_patterns(H, seed):
    """Random tiny patterns incl. empty columns and states no measurement touches (lambda > 0 makes
    those pivots lambda): factor and solve through the level schedule against LAPACK."""
    from test_symbolic import _Pattern
    rng = np.random.default_rng(400 + seed)
    n, m = int(rng.integers(1, 30)), int(rng.integers(1, 120))
    cols = [np.sort(rng.choice(n, size=int(rng.integers(0, min(n, 5) + 1)), replace=False)) for _ in range(m)]
    prob = _Pattern(n, cols)
    S = structures(H, prob)
    lam = 0.5
    nnz = int(S["Jp"][-1])
    Jx = rng.uniform(-1, 1, max(nnz, 1))
    D = np.zeros((m, n))
    for j in range(m):
        D[j, S["Ji"][S["Jp"][j]:S["Jp"][j + 1]]] = Jx[S["Jp"][j]:S["Jp"][j + 1]]
    A = D.T @ D + lam * np.eye(n)
    Lref = np.linalg.cholesky(A[np.ix_(S["perm"], S["perm"])])
    Lgot, fronts = multifrontal(S, n, Jx, lam)
    assert np.abs(Lgot - Lref).max() <= 1e-12 * max(1.0, np.abs(Lref).max())
    g = rng.uniform(-1, 1, n)
    u = solve_with_fronts(S, n, fronts, g)
    assert np.abs(u - np.linalg.solve(A, g)).max() <= 1e-11 * max(1.0, np.abs(g).max())
