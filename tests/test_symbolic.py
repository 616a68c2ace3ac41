"""CPU tests of the host-side symbolic analysis (integer work: bit-exact).
Permutation validity, elimination tree, column counts, supernode partition and
the L pattern are compared with a brute-force boolean elimination of Jt*Jt'
under the SAME ordering -- the contract BASELINE.json states for integer
structures -- and with the oracle's etree/colcounts (CHOLMOD restatement)."""
import ctypes as C

import os

import numpy as np
import pytest

from libdogleg_b200 import ffi


def brute(n, Jp, Ji, perm):
    iperm = np.empty(n, int)
    iperm[perm] = np.arange(n)
    A = np.zeros((n, n), bool)
    for j in range(len(Jp) - 1):
        r = iperm[Ji[Jp[j]:Jp[j + 1]]]
        A[np.ix_(r, r)] = True
    Lm = np.tril(A)
    for k in range(n):
        rows = np.nonzero(Lm[k + 1:, k])[0] + k + 1
        Lm[np.ix_(rows, rows)] |= True
    Lm = np.tril(Lm)
    np.fill_diagonal(Lm, True)
    parent = np.full(n, -1)
    for k in range(n):
        rows = np.nonzero(Lm[k + 1:, k])[0]
        if len(rows):
            parent[k] = rows[0] + k + 1
    return Lm, parent


def analyze(H, prob, perm=None, post=0):
    L = ffi.load()
    Jp, Ji = prob.pattern()
    n = prob.N
    h = L.dlb_symbolic_create(n, prob.M, H.as_ip(Jp), H.as_ip(Ji),
                              H.as_ip(perm) if perm is not None else None, post)
    assert h
    info = (C.c_longlong * 8)()
    L.dlb_symbolic_info(h, info)
    info = list(info)

    def get(name, cap):
        out = np.zeros(max(cap, 1), np.int32)
        ln = L.dlb_symbolic_get(h, ffi.SYM[name], H.as_ip(out), cap)
        return out[:ln]
    res = dict(info=info, perm=get("perm", n), parent=get("parent", n), colcount=get("colcount", n),
               sn_first=get("sn_first", n + 1), rows_ptr=get("rows_ptr", n + 1), rows=get("rows", info[7]),
               sn_parent=get("sn_parent", n), cls_of_col=get("cls_of_col", prob.M), sn_level=get("sn_level", n))
    L.dlb_symbolic_free(h)
    return res


def check_exact(H, prob, perm=None, post=0, relaxed=False):
    """relaxed=False: fundamental supernodes (DOGLEG_GPU_RELAX=0), the supernodal row lists must
    reproduce the brute-force pattern of L exactly. relaxed=True: the default amalgamation, the
    fronts may carry explicit zeros (a superset of L), everything else stays exact."""
    Jp, Ji = prob.pattern()
    n = prob.N
    old = os.environ.get("DOGLEG_GPU_RELAX")
    if not relaxed:
        os.environ["DOGLEG_GPU_RELAX"] = "0"
    try:
        S = analyze(H, prob, perm, post)
    finally:
        if not relaxed:
            if old is None:
                del os.environ["DOGLEG_GPU_RELAX"]
            else:
                os.environ["DOGLEG_GPU_RELAX"] = old
    p = S["perm"]
    assert sorted(p) == list(range(n))
    if perm is not None and not post:
        assert (p == perm).all()
    Lm, par = brute(n, Jp, Ji, p)
    assert (S["parent"] == par).all()
    assert (S["colcount"] == Lm.sum(0)).all()
    L2 = np.zeros((n, n), bool)
    nsuper = S["info"][1]
    for s in range(nsuper):
        r = S["rows"][S["rows_ptr"][s]:S["rows_ptr"][s + 1]]
        ncol = S["sn_first"][s + 1] - S["sn_first"][s]
        assert (r[:ncol] == np.arange(S["sn_first"][s], S["sn_first"][s + 1])).all()
        assert (np.diff(r) > 0).all()
        for c in range(S["sn_first"][s], S["sn_first"][s + 1]):
            L2[r[r >= c], c] = True
        # supernode parent = supernode of the first below-diagonal row; levels increase towards the root
        if len(r) > ncol:
            ps = S["sn_parent"][s]
            assert S["sn_first"][ps] <= r[ncol] < S["sn_first"][ps + 1]
            assert S["sn_level"][ps] > S["sn_level"][s]
        else:
            assert S["sn_parent"][s] == -1
    if relaxed:
        assert (L2 | ~Lm).all()                          # every entry of L is stored
        assert L2.sum() <= 1.6 * Lm.sum()                # bounded explicit zeros
    else:
        assert (L2 == Lm).all()
    assert S["info"][3] == Lm.sum()
    return S


PROBLEMS = [lambda H: H.Problem.sample(),
            lambda H: H.Problem.mrcal(2, 6, 12),
            lambda H: H.Problem.mrcal(4, 20, 5),
            lambda H: H.Problem.random_sparse(60, 300, 5),
            lambda H: H.Problem.random_sparse(200, 900, 4, seed=5),
            lambda H: H.Problem.ba(10, 60, 3, 5),
            lambda H: H.Problem.ba(20, 200, 4, 8, 50)]


@pytest.mark.parametrize("mk", PROBLEMS)
def test_symbolic_bit_exact_own_ordering(H, mk):
    check_exact(H, mk(H))


@pytest.mark.parametrize("mk", PROBLEMS)
def test_symbolic_relaxed_amalgamation(H, mk):
    exact = check_exact(H, mk(H))
    relaxed = check_exact(H, mk(H), relaxed=True)
    assert relaxed["info"][1] <= exact["info"][1]        # supernodes
    assert relaxed["info"][2] <= exact["info"][2]        # levels
    assert (relaxed["perm"] == exact["perm"]).all() and (relaxed["colcount"] == exact["colcount"]).all()


def test_nested_dissection_ordering(H, monkeypatch):
    """Banded camera graph (bundle adjustment): once the points are gone the ordering switches to
    nested dissection. The structures stay bit-exact against the brute-force elimination, the
    elimination tree gets much shallower than with plain minimum degree, the fill stays comparable."""
    prob = H.Problem.ba(60, 1500, 4, 8, 0)
    # one pivot at a time (no multiple elimination): plain minimum degree gives a chain-like tree
    monkeypatch.setenv("DOGLEG_GPU_MULTI_ELIM", "-1")
    monkeypatch.setenv("DOGLEG_GPU_ND", "0")
    amd = check_exact(H, prob)
    monkeypatch.setenv("DOGLEG_GPU_ND", "30,16,6")
    nd = check_exact(H, prob)
    check_exact(H, prob, relaxed=True)
    # the default (multiple elimination below degree 64) with and without dissection: valid too
    monkeypatch.delenv("DOGLEG_GPU_MULTI_ELIM")
    check_exact(H, prob)
    monkeypatch.setenv("DOGLEG_GPU_ND", "0")
    multi = check_exact(H, prob)
    assert multi["info"][3] <= 1.1 * amd["info"][3]        # same fill as one pivot at a time
    assert nd["info"][2] < amd["info"][2]                  # levels
    assert nd["info"][3] <= 1.5 * amd["info"][3]           # nnz(L)
    # with long-range observations (irregular separators) it must still be a valid ordering
    check_exact(H, H.Problem.ba(40, 600, 4, 8, 30))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_symbolic_bit_exact_injected_ordering(H, seed):
    prob = H.Problem.random_sparse(80, 300, 4, seed=9 + seed)
    pm = np.random.default_rng(seed).permutation(80).astype(np.int32)
    check_exact(H, prob, pm, 0)
    check_exact(H, prob, pm, 1)
    # identity ordering and reversed ordering are legal too
    check_exact(H, prob, np.arange(80, dtype=np.int32), 0)
    check_exact(H, prob, np.arange(79, -1, -1, dtype=np.int32), 0)


def test_symbolic_matches_oracle_etree_and_counts(H):
    """Same ordering into the CHOLMOD restatement: etree and column counts agree bit for bit."""
    O = H.oracle_lib()
    prob = H.Problem.mrcal(3, 10, 4)
    Jp, Ji = prob.pattern()
    n = prob.N
    pm = np.zeros(n, np.int32)
    O.orc_min_degree(n, prob.M, H.as_ip(Jp), H.as_ip(Ji), H.as_ip(pm))
    S = analyze(H, prob, pm, 0)

    class OrcFactor(C.Structure):
        _fields_ = [("n", C.c_int), ("perm", C.POINTER(C.c_int)), ("iperm", C.POINTER(C.c_int)),
                    ("parent", C.POINTER(C.c_int)), ("colcount", C.POINTER(C.c_int))]
    F = O.orc_analyze(n, prob.M, H.as_ip(Jp), H.as_ip(Ji), H.as_ip(pm))
    f = C.cast(F, C.POINTER(OrcFactor)).contents
    assert (np.ctypeslib.as_array(f.parent, shape=(n,)) == S["parent"]).all()
    assert (np.ctypeslib.as_array(f.colcount, shape=(n,)) == S["colcount"]).all()
    O.orc_free(F)


def test_amd_fill_is_close_to_exact_minimum_degree(H):
    O = H.oracle_lib()
    for prob in [H.Problem.mrcal(4, 20, 5), H.Problem.random_sparse(200, 900, 4, seed=5),
                 H.Problem.ba(20, 200, 4, 8, 50)]:
        Jp, Ji = prob.pattern()
        pm = np.zeros(prob.N, np.int32)
        O.orc_min_degree(prob.N, prob.M, H.as_ip(Jp), H.as_ip(Ji), H.as_ip(pm))
        exact = analyze(H, prob, pm, 1)["info"][3]
        ours = analyze(H, prob)["info"][3]
        assert ours <= 1.10 * exact


def test_pattern_classes(H):
    prob = H.Problem.mrcal(4, 20, 5)
    S = analyze(H, prob)
    Jp, Ji = prob.pattern()
    cls = S["cls_of_col"]
    assert S["info"][0] == 4 * 20 * 2          # (frame, camera, xy) patterns
    seen = {}
    for j in range(prob.M):
        key = tuple(Ji[Jp[j]:Jp[j + 1]])
        assert seen.setdefault(key, cls[j]) == cls[j]
    assert len(set(seen.values())) == len(seen)


def test_malformed_input_is_rejected(H):
    L = ffi.load()
    Jp = np.array([0, 2, 4], np.int32)
    Ji = np.array([1, 0, 0, 5], np.int32)       # descending, out of range
    assert not L.dlb_symbolic_create(3, 2, H.as_ip(Jp), H.as_ip(Ji), None, 1)
    Ji = np.array([0, 1, 0, 2], np.int32)
    bad = np.array([0, 0, 1], np.int32)         # not a permutation
    assert not L.dlb_symbolic_create(3, 2, H.as_ip(Jp), H.as_ip(Ji), H.as_ip(bad), 0)


class _Pattern:
    """A bare CCS pattern with the interface check_exact needs."""

    def __init__(self, n, cols):
        self.N, self.M = n, len(cols)
        self._Jp = np.concatenate([[0], np.cumsum([len(c) for c in cols])]).astype(np.int32)
        self._Ji = (np.concatenate(cols) if sum(len(c) for c in cols) else np.zeros(0)).astype(np.int32)

    def pattern(self):
        Ji = self._Ji if len(self._Ji) else np.zeros(1, np.int32)      # never hand out a NULL pointer
        return self._Jp, Ji


@pytest.mark.parametrize("seed", range(40))
def test_symbolic_random_small_patterns(H, seed):
    """Random tiny patterns: empty columns, repeated columns, states no measurement touches, a single
    state, dense columns -- exact against the brute-force elimination, own and injected orderings,
    fundamental and relaxed supernodes."""
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(1, 26))
    m = int(rng.integers(1, 60))
    cols = []
    for _ in range(m):
        kind = rng.integers(0, 5)
        if kind == 0:
            c = np.zeros(0, int)                                        # empty column
        elif kind == 1 and cols:
            c = cols[int(rng.integers(0, len(cols)))]                   # repeat an earlier pattern
        elif kind == 2:
            c = np.arange(n)                                            # dense column
        else:
            c = np.sort(rng.choice(n, size=int(rng.integers(1, min(n, 6) + 1)), replace=False))
        cols.append(np.asarray(c, int))
    prob = _Pattern(n, cols)
    check_exact(H, prob)
    check_exact(H, prob, relaxed=True)
    pm = rng.permutation(n).astype(np.int32)
    check_exact(H, prob, pm, 0)
    check_exact(H, prob, pm, 1, relaxed=True)
