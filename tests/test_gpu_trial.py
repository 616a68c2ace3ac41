"""GPU tests of the fused evaluation (gradient + class blocks in one pass, dlb_sparse.cu) and of the
persistent trial kernel (dlb_trial.cu: Cauchy, factorization + solves, step selection, expected
improvement in one cooperative launch) against numpy restatements of reference dogleg.c:529-617,
839-866, 927-998, 1085-1165, 1192-1296 on identical inputs, and against the per-operation engine
calls of the round-1 schedule (DOGLEG_GPU_FUSED=0)."""
import numpy as np
import pytest

from libdogleg_b200 import ffi

pytestmark = pytest.mark.gpu

PROBLEMS = {
    "sample": lambda H: H.Problem.sample(),
    "mrcal_small": lambda H: H.Problem.mrcal(2, 6, 12, seed=7),
    "mrcal_frames": lambda H: H.Problem.mrcal(4, 20, 5),
    "mrcal_runs": lambda H: H.Problem.mrcal(3, 4, 150, seed=5),
    "mrcal_many_fronts": lambda H: H.Problem.mrcal(2, 330, 3, seed=8),      # more fronts than CTAs of the grid
    "random": lambda H: H.Problem.random_sparse(60, 300, 5),
    "random_long_columns": lambda H: H.Problem.random_sparse(120, 900, 40, seed=5),
    "ba_tiny": lambda H: H.Problem.ba(10, 60, 3, 5),
    "ragged": lambda H: H.Problem.ragged(100, 3000, 20),
}


def dense_J(prob, Jp, Ji, Jx):
    J = np.zeros((prob.M, prob.N))
    for j in range(prob.M):
        J[j, Ji[Jp[j]:Jp[j + 1]]] = Jx[Jp[j]:Jp[j + 1]]
    return J


def setup(H, name):
    prob = PROBLEMS[name](H)
    Jp, Ji = prob.pattern()
    p = prob.p0()
    x, Jx = prob.evaluate(p)
    E = H.Engine(ffi.SOLVE_SPARSE, prob.N, prob.M, len(Ji))
    E.load_sparse(0, p, x, Jp, Ji, Jx)
    return prob, E, p, x, dense_J(prob, Jp, Ji, Jx)


@pytest.mark.parametrize("name", list(PROBLEMS))
def test_fused_evaluation_matches_numpy(H, name):
    prob, E, p, x, J = setup(H, name)
    sc = E.evaluate(0)
    g_ref = J.T @ x
    g = E.download(0)["Jtx"]
    assert np.max(np.abs(g - g_ref)) <= 1e-11 * np.max(np.abs(g_ref))
    assert np.isclose(sc.norm2_x, x @ x, rtol=1e-12)
    assert np.isclose(sc.norm2_Jtx, g_ref @ g_ref, rtol=1e-10)
    assert np.isclose(sc.maxabs_Jtx, np.max(np.abs(g_ref)), rtol=1e-11)
    # the class blocks formed in the same pass: |J g|^2 as g'(JtJ)g and the assembled matrix
    sc = E.cauchy(0)
    assert np.isclose(sc.norm2_JJtx, np.sum((J @ g_ref) ** 2), rtol=1e-10)
    A = E.JtJ(0, 0.0)
    assert np.max(np.abs(A - J.T @ J)) <= 1e-11 * np.max(np.abs(J.T @ J))
    E.close()


@pytest.mark.parametrize("name", list(PROBLEMS))
def test_trial_kernel_matches_numpy(H, name):
    prob, E, p, x, J = setup(H, name)
    if not E.has_trial():
        E.close()
        pytest.skip("tree not eligible for the trial kernel")
    N = prob.N
    sc = E.evaluate(0)
    g = J.T @ x
    JtJ = J.T @ J
    k = -(g @ g) / np.sum((J @ g) ** 2)
    cauchy = k * g
    n2c = k * k * (g @ g)
    gn = -np.linalg.solve(JtJ, g)
    n2g = gn @ gn
    lc, lg = np.sqrt(n2c), np.sqrt(n2g)
    assert lc < lg, "fixture: the Cauchy step should be the shorter one"
    tol = max(1e-9, 1e-10 * np.sqrt(np.linalg.cond(JtJ)))
    # Cauchy-clipped, then interpolated, then Gauss-Newton: the second and third launch reuse the
    # cached Cauchy step, the third the cached factorization as well
    for kind, delta in ((ffi.STEP_CAUCHY, 0.5 * lc), (ffi.STEP_INTERPOLATED, 0.5 * (lc + lg)),
                        (ffi.STEP_GAUSSNEWTON, 2.0 * lg)):
        sc = E.trial(0, 1, delta)
        assert sc.minor == -1
        assert int(sc.step_type) == kind
        assert sc.trial_flags == (1.0 if kind == ffi.STEP_INTERPOLATED else 0.0)
        assert np.isclose(sc.norm2_cauchy, n2c, rtol=1e-10)
        d0, d1 = E.download(0), E.download(1)
        assert np.allclose(d0["cauchy"], cauchy, rtol=1e-10, atol=1e-300)
        if kind == ffi.STEP_CAUCHY:
            step_ref = cauchy * (delta / lc)
            assert np.isclose(sc.norm2_step, n2c, rtol=1e-10)            # the unclipped length, dogleg.c:1200
        elif kind == ffi.STEP_GAUSSNEWTON:
            step_ref = gn
            assert np.isclose(sc.norm2_step, n2g, rtol=10 * tol)
        else:
            assert np.max(np.abs(d0["gn"] - gn)) <= tol * np.max(np.abs(gn))
            assert np.isclose(sc.norm2_gn, d0["gn"] @ d0["gn"], rtol=1e-12)
            d = cauchy - gn
            l2, negc = d @ d, d @ cauchy
            kk = (negc + np.sqrt(negc * negc - l2 * (n2c - delta * delta))) / l2
            step_ref = cauchy + kk * (gn - cauchy)
            assert np.isclose(sc.k_interp, kk, rtol=1e3 * tol)
            assert np.isclose(np.sqrt(sc.norm2_step), delta, rtol=1e-9)
        assert np.max(np.abs(d1["step"] - step_ref)) <= 1e2 * tol * np.max(np.abs(step_ref))
        assert np.allclose(d1["p"], p + d1["step"], rtol=0, atol=1e-15 * max(1, np.max(np.abs(p))))
        # p of the trial point also arrives in its pinned host mirror (host callbacks read it there)
        assert np.array_equal(E.host(1, ffi.BUF_P, N), d1["p"])
        assert np.isclose(sc.Jtx_dot_step, g @ d1["step"], rtol=1e-9)
        assert np.isclose(sc.maxabs_step, np.max(np.abs(d1["step"])), rtol=1e-15)
        assert np.isclose(sc.norm2_Jstep, np.sum((J @ d1["step"]) ** 2), rtol=1e-9)
    # the factor the kernel left behind serves multi-RHS solves (outlier helpers)
    B = np.random.default_rng(1).standard_normal((N, 3))
    X = E.solve(B)
    assert np.max(np.abs(JtJ @ X - B)) <= 1e-7 * np.max(np.abs(B)) * max(1.0, np.sqrt(np.linalg.cond(JtJ)) * 1e-3)
    E.close()


def test_trial_kernel_reports_singular_matrix_and_takes_lambda(H):
    """Two identical columns in J make JtJ singular: minor >= 0, nothing after the factorization is
    done; with lambda > 0 the same call succeeds (dogleg.c:668-677)."""
    prob = H.Problem.random_sparse(40, 200, 4, seed=11)
    Jp, Ji = prob.pattern()
    p = prob.p0()
    x, Jx = prob.evaluate(p)
    J = dense_J(prob, Jp, Ji, Jx)
    # make state 7 a copy of state 3 wherever both appear... simplest: zero a state's column entirely
    for j in range(prob.M):
        for q in range(Jp[j], Jp[j + 1]):
            if Ji[q] == 5:
                Jx[q] = 0.0
    J[:, 5] = 0.0
    E = H.Engine(ffi.SOLVE_SPARSE, prob.N, prob.M, len(Ji))
    E.load_sparse(0, p, x, Jp, Ji, Jx)
    if not E.has_trial():
        E.close()
        pytest.skip("tree not eligible for the trial kernel")
    E.evaluate(0)
    sc = E.trial(0, 1, 1e6)
    assert sc.minor >= 0
    lam = 1e-3
    sc = E.trial(0, 1, 1e6, lam)
    assert sc.minor == -1 and int(sc.step_type) == ffi.STEP_GAUSSNEWTON and sc.trial_flags == 1.0
    g = J.T @ x
    gn = -np.linalg.solve(J.T @ J + lam * np.eye(prob.N), g)
    assert np.max(np.abs(E.download(0)["gn"] - gn)) <= 1e-8 * np.max(np.abs(gn))
    E.close()


@pytest.mark.parametrize("name", ["mrcal_frames", "mrcal_runs", "random", "ragged", "sample"])
def test_fused_schedule_and_round1_schedule_agree(H, monkeypatch, name):
    """The same solve through the fused schedule (default) and through the per-operation schedule of
    round 1 (DOGLEG_GPU_FUSED=0): identical callback counts, cost to 1e-12, p to 1e-9."""
    monkeypatch.setenv("DOGLEG_GPU_ENGINE_CACHE", "0")
    res = []
    for v in ("1", "0"):
        monkeypatch.setenv("DOGLEG_GPU_FUSED", v)
        res.append(H.solve_product(PROBLEMS[name](H), "sparse", max_iterations=30))
    a, b = res
    assert a.ncalls == b.ncalls and a.accepted == b.accepted
    assert abs(a.norm2x - b.norm2x) <= 1e-12 * abs(b.norm2x)
    assert np.max(np.abs(a.p - b.p)) <= 1e-9 * max(1.0, np.max(np.abs(b.p)))
