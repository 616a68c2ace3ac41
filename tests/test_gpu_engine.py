"""GPU parity tests, one engine call at a time (include/dogleg_gpu.h), against the
oracle restatement of the reference functions on identical inputs.
Floating point: the kernels sum in a different (fixed) order than the scalar
reference loops, so results agree to round-off: rtol 1e-11 on sums of >=100 terms
(BASELINE.json's bar is 1e-9 on cost, 1e-7 on p)."""
import numpy as np
import pytest

from libdogleg_b200 import ffi

pytestmark = pytest.mark.gpu
RT = 1e-11


def dense_J(prob, Jp, Ji, Jx):
    J = np.zeros((prob.M, prob.N))
    for j in range(prob.M):
        J[j, Ji[Jp[j]:Jp[j + 1]]] = Jx[Jp[j]:Jp[j + 1]]
    return J


SPARSE = [lambda H: H.Problem.sample(),
          lambda H: H.Problem.mrcal(2, 6, 12, seed=7),
          lambda H: H.Problem.mrcal(4, 20, 5),
          lambda H: H.Problem.random_sparse(60, 300, 5),
          lambda H: H.Problem.random_sparse(200, 900, 40, seed=5),     # columns longer than a warp
          lambda H: H.Problem.ba(10, 60, 3, 5),
          lambda H: H.Problem.ba(20, 200, 4, 8, 50),
          lambda H: H.Problem.mrcal(10, 12, 4, seed=3),        # 182-row fronts: blocked tensor-core path
          lambda H: H.Problem.mrcal(3, 4, 150, seed=5)]        # runs of 300 columns with period 2: range tasks


@pytest.mark.parametrize("mk", SPARSE)
def test_sparse_engine_ops_match_oracle(H, mk):
    O = H.oracle_lib()
    prob = mk(H)
    Jp, Ji = prob.pattern()
    p = prob.p0()
    x, Jx = prob.evaluate(p)
    J = dense_J(prob, Jp, Ji, Jx)
    N, M = prob.N, prob.M
    E = H.Engine(ffi.SOLVE_SPARSE, N, M, len(Ji))
    E.load_sparse(0, p, x, Jp, Ji, Jx)

    # a2/a3/a4: gradient pass
    sc = E.evaluate(0)
    g_ref = np.zeros(N)
    O.orc_Jt_times_x(H.as_dp(g_ref), N, M, H.as_ip(Jp), H.as_ip(Ji), H.as_dp(Jx), H.as_dp(x))
    g = E.download(0)["Jtx"]
    scale = np.max(np.abs(g_ref))
    assert np.max(np.abs(g - g_ref)) <= RT * scale
    assert np.isclose(sc.norm2_x, np.dot(x, x), rtol=RT)
    assert np.isclose(sc.norm2_Jtx, np.dot(g_ref, g_ref), rtol=1e-10)
    assert np.isclose(sc.maxabs_Jtx, scale, rtol=RT)

    # a5/a6: cauchy
    sc = E.cauchy(0)
    jg2 = O.orc_norm2_J_times_v(M, H.as_ip(Jp), H.as_ip(Ji), H.as_dp(Jx), H.as_dp(g_ref))
    assert np.isclose(sc.norm2_JJtx, jg2, rtol=1e-10)
    k = -np.dot(g_ref, g_ref) / jg2
    assert np.isclose(sc.k_cauchy, k, rtol=1e-10)
    assert np.isclose(sc.norm2_cauchy, k * k * np.dot(g_ref, g_ref), rtol=1e-10)
    cauchy = E.download(0)["cauchy"]
    assert np.allclose(cauchy, k * g_ref, rtol=1e-10, atol=1e-300)

    # a7: assembly of JtJ + lambda I
    A_ref = np.zeros((N, N))
    O.orc_sparse_JtJ_dense(H.as_dp(A_ref), N, M, H.as_ip(Jp), H.as_ip(Ji), H.as_dp(Jx), 0.25)
    A = E.JtJ(0, 0.25)
    assert np.max(np.abs(A - A_ref)) <= RT * np.max(np.abs(A_ref))
    assert np.array_equal(A == 0, A_ref == 0)            # same sparsity

    # a7/a8: factorize + solve
    assert E.factorize(0, 0.0) == -1
    sc = E.gauss_newton(0)
    gn = E.download(0)["gn"]
    gn_ref = -np.linalg.solve(J.T @ J, g_ref)
    tol = 1e-9 * np.linalg.cond(J.T @ J) ** 0.5
    assert np.max(np.abs(gn - gn_ref)) <= max(tol, 1e-9) * np.max(np.abs(gn_ref))
    assert np.isclose(sc.norm2_gn, np.dot(gn, gn), rtol=1e-12)
    # multi-RHS solve (what the outlier helpers use, dogleg.c:1914-1918)
    B = np.random.default_rng(0).standard_normal((N, 4))
    X = E.solve(B)
    assert np.max(np.abs((J.T @ J) @ X - B)) <= 1e-8 * np.max(np.abs(B)) * max(1.0, np.linalg.cond(J.T @ J) ** 0.5 * 1e-3)

    # lambda > 0
    assert E.factorize(0, 0.5) == -1
    E.gauss_newton(0)
    gn2 = E.download(0)["gn"]
    assert np.allclose(gn2, -np.linalg.solve(J.T @ J + 0.5 * np.eye(N), g_ref), rtol=1e-8, atol=1e-12)
    assert E.factorize(0, 0.0) == -1
    E.gauss_newton(0)

    # a10-a12: the three step types
    n2c, n2g = E.scalars().norm2_cauchy, E.scalars().norm2_gn
    for kind, delta in ((ffi.STEP_CAUCHY, 0.5 * np.sqrt(n2c)), (ffi.STEP_GAUSSNEWTON, 2 * np.sqrt(n2g)),
                        (ffi.STEP_INTERPOLATED, 0.5 * (np.sqrt(n2c) + np.sqrt(n2g)))):
        sc = E.step(0, 1, kind, delta)
        d1 = E.download(1)
        if kind == ffi.STEP_CAUCHY:
            step_ref = cauchy * (delta / np.sqrt(n2c))
            assert np.isclose(sc.norm2_step, n2c, rtol=1e-14)          # the unclipped length, dogleg.c:1200
        elif kind == ffi.STEP_GAUSSNEWTON:
            step_ref = gn
            assert np.isclose(sc.norm2_step, n2g, rtol=1e-14)
        else:
            d = cauchy - gn
            l2, negc = d @ d, d @ cauchy
            kk = (negc + np.sqrt(negc * negc - l2 * (n2c - delta * delta))) / l2
            step_ref = cauchy + kk * (gn - cauchy)
            assert np.isclose(sc.k_interp, kk, rtol=1e-9)
            assert np.isclose(np.sqrt(sc.norm2_step), delta, rtol=1e-9)    # lands on the trust-region boundary
        assert np.allclose(d1["step"], step_ref, rtol=1e-9, atol=1e-14 * np.max(np.abs(step_ref)))
        assert np.allclose(d1["p"], p + d1["step"], rtol=0, atol=1e-15 * max(1, np.max(np.abs(p))))
        assert np.isclose(sc.Jtx_dot_step, g_ref @ d1["step"], rtol=1e-10)
        assert np.isclose(sc.maxabs_step, np.max(np.abs(d1["step"])), rtol=1e-15)
        js = O.orc_norm2_J_times_v(M, H.as_ip(Jp), H.as_ip(Ji), H.as_dp(Jx), H.as_dp(d1["step"]))
        assert np.isclose(sc.norm2_Jstep, js, rtol=1e-10)
    E.close()


def test_sparse_results_are_bit_reproducible(H):
    """Fixed summation order everywhere: two runs give identical bits."""
    prob = H.Problem.mrcal(4, 20, 5)
    Jp, Ji = prob.pattern()
    p = prob.p0()
    x, Jx = prob.evaluate(p)
    outs = []
    for _ in range(2):
        E = H.Engine(ffi.SOLVE_SPARSE, prob.N, prob.M, len(Ji))
        E.load_sparse(0, p, x, Jp, Ji, Jx)
        E.evaluate(0)
        E.cauchy(0)
        E.factorize(0, 0.0)
        E.gauss_newton(0)
        d = E.download(0)
        outs.append(np.concatenate([d["Jtx"], d["cauchy"], d["gn"]]))
        E.close()
    assert np.array_equal(outs[0], outs[1])


def test_sparse_injected_permutation_gives_same_solution(H):
    prob = H.Problem.random_sparse(60, 300, 5)
    Jp, Ji = prob.pattern()
    p = prob.p0()
    x, Jx = prob.evaluate(p)
    sols = []
    for perm in (None, np.arange(prob.N)[::-1].copy(), np.random.default_rng(1).permutation(prob.N)):
        E = H.Engine(ffi.SOLVE_SPARSE, prob.N, prob.M, len(Ji))
        E.load_sparse(0, p, x, Jp, Ji, Jx, perm=perm, postorder=0)
        E.evaluate(0)
        assert E.factorize(0, 0.0) == -1
        E.gauss_newton(0)
        sols.append(E.download(0)["gn"])
        E.close()
    for s in sols[1:]:
        assert np.allclose(s, sols[0], rtol=1e-9, atol=1e-13)


def test_singular_JtJ_is_reported_and_lambda_fixes_it(H):
    """A state no measurement depends on: pivot <= 0 must be flagged at that column (the
    signal that drives the lambda ladder, dogleg.c:667-673)."""
    prob = H.Problem.random_sparse(30, 120, 4, seed=2)
    Jp, Ji = prob.pattern()
    Ji = Ji.copy()
    N = prob.N + 1                       # one extra, untouched state
    p = np.append(prob.p0(), 0.0)
    x, Jx = prob.evaluate(prob.p0())
    E = H.Engine(ffi.SOLVE_SPARSE, N, prob.M, len(Ji))
    E.load_sparse(0, p, x, Jp, Ji, Jx)
    E.evaluate(0)
    assert E.factorize(0, 0.0) >= 0
    assert E.factorize(0, 1e-10) == -1
    E.close()


DENSE = [(6, 100, 1), (16, 256, 3), (40, 1000, 4), (130, 700, 5), (200, 300, 6), (515, 2100, 7)]


@pytest.mark.parametrize("N,M,seed", DENSE)
def test_dense_engine_ops_match_oracle(H, N, M, seed):
    O = H.oracle_lib()
    rng = np.random.default_rng(seed)
    J = rng.standard_normal((M, N))
    x = rng.standard_normal(M)
    p = rng.standard_normal(N)
    E = H.Engine(ffi.SOLVE_DENSE, N, M)
    E.load_dense(0, p, x, J)
    sc = E.evaluate(0)
    g = E.download(0)["Jtx"]
    g_ref = J.T @ x
    assert np.max(np.abs(g - g_ref)) <= RT * np.max(np.abs(g_ref))
    assert np.isclose(sc.norm2_x, x @ x, rtol=RT)
    sc = E.cauchy(0)
    assert np.isclose(sc.norm2_JJtx, np.sum((J @ g_ref) ** 2), rtol=1e-10)
    # a17: packed JtJ as the reference builds it
    ap = np.zeros(N * (N + 1) // 2)
    O.orc_dense_JtJ_packed_upper(H.as_dp(ap), H.as_dp(np.ascontiguousarray(J)), M, N, 0.0)
    A = E.JtJ(0, 0.0)
    iu = np.triu_indices(N)
    assert np.max(np.abs(A[iu] - ap)) <= RT * np.max(np.abs(ap))
    assert E.factorize(0, 0.0) == -1
    # the factor in the reference's layout equals dpptrf's (orc_pptrf_lower restates it)
    fac = np.zeros(N * (N + 1) // 2)
    E.check(E.L.dlb_engine_dense_factor_to_host(E.h, H.as_dp(fac)))
    assert O.orc_pptrf_lower(H.as_dp(ap), N) == 0
    assert np.max(np.abs(fac - ap)) <= 1e-9 * np.max(np.abs(ap))
    sc = E.gauss_newton(0)
    gn = E.download(0)["gn"]
    assert np.allclose(gn, -np.linalg.solve(J.T @ J, g_ref), rtol=1e-8, atol=1e-12)
    sc = E.step(0, 1, ffi.STEP_GAUSSNEWTON, 1e9)
    assert np.isclose(sc.norm2_Jstep, np.sum((J @ gn) ** 2), rtol=1e-9)
    E.close()


@pytest.mark.parametrize("packed,upper", [(1, 1), (0, 0), (1, 0)])
def test_products_engine_ops(H, packed, upper):
    N, M = 24, 300
    rng = np.random.default_rng(8)
    J = rng.standard_normal((M, N))
    x = rng.standard_normal(M)
    A = J.T @ J
    xtJ = J.T @ x
    if packed and upper:
        lay = A[np.triu_indices(N)]
    elif packed:
        lay = A[np.tril_indices(N)]
    else:
        lay = A.ravel()
    E = H.Engine(ffi.SOLVE_DENSE_PRODUCTS, N, 0, 0, packed, upper)
    E.load_products(0, np.zeros(N), xtJ, lay)
    sc = E.evaluate(0, float(x @ x))
    assert sc.norm2_x == float(x @ x)
    assert np.isclose(sc.norm2_Jtx, xtJ @ xtJ, rtol=RT)
    assert E.factorize(0, 0.0) == -1
    E.gauss_newton(0)
    gn = E.download(0)["gn"]
    assert np.allclose(gn, -np.linalg.solve(A, xtJ), rtol=1e-8, atol=1e-12)
    if packed and not upper:
        with pytest.raises(RuntimeError):      # as in the reference (dogleg.c:597-601)
            E.cauchy(0)
    else:
        sc = E.cauchy(0)
        assert np.isclose(sc.norm2_JJtx, xtJ @ A @ xtJ, rtol=1e-10)
    E.close()
