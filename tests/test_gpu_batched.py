"""GPU parity tests of the batched dense mode (config C3, additive entry point
dogleg_gpu_optimize_dense_batched): every problem of the batch must end where the
reference's dogleg_optimize_dense2 ends for the same problem -- same number of accepted
steps, cost within 1e-9 relative, p within 1e-7."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def ref_solve(H, prob, p0, **pk):
    if H.reference_lib() is not None:
        r = H.solve_reference(prob, "dense", p0=p0, **pk)
        o = H.solve_oracle(prob, "dense", p0=p0, **pk)
        r.accepted = o.accepted
        return r
    return H.solve_oracle(prob, "dense", p0=p0, **pk)


@pytest.mark.parametrize("tr0", [1e3, 0.3])
def test_batched_synthetic_matches_reference_per_problem(H, tr0):
    B, N, M, seed = 40, 16, 256, 1000
    DL = H.dev_problems_lib()
    p0 = np.zeros((B, N))
    dev = DL.dlb_dev_problem_create_batched(B, M, N, seed, H.as_dp(p0))
    assert dev
    rc, p, n2, it = H.solve_batched(dev, p0, N, M, max_iterations=30, trustregion0=tr0)
    DL.dlb_dev_problem_free(dev)
    assert rc == B
    for b in range(0, B, 3):
        prob = H.Problem.dense(N, M, seed=seed + b)
        assert np.allclose(prob.p0(), p0[b], rtol=0, atol=1e-14)      # host and device generators agree
        ref = ref_solve(H, prob, p0[b], max_iterations=30, trustregion0=tr0)
        assert it[b] == ref.accepted
        assert abs(n2[b] - ref.norm2x) <= 1e-9 * ref.norm2x
        assert np.max(np.abs(p[b] - ref.p)) <= 1e-7 * max(1.0, np.max(np.abs(ref.p)))


def test_batched_sample_surface_with_rejections(H):
    """The reference's own sample problem (6 states: not a multiple of the 8-wide tensor-core tile)
    from many start points: rejected steps, cauchy, interpolated and Gauss-Newton steps, lambda = 0."""
    prob = H.Problem.sample()
    DL = H.dev_problems_lib()
    rng = np.random.default_rng(5)
    starts = [[5, -3, 2, 1, 0, 10], [-2, 4, -1, 3, 3, 3], [0.1, 0.1, 0.1, 0, 0, 0], list(prob.p0())]
    starts += [list(rng.uniform(-3, 6, 6)) for _ in range(28)]
    p0 = np.array(starts, dtype=float)
    B = len(p0)
    dev = DL.dlb_dev_problem_create_sample_batched(C.cast(prob.ptr, C.c_void_p), B)
    assert dev
    rc, p, n2, it = H.solve_batched(dev, p0, 6, prob.M, max_iterations=60)
    DL.dlb_dev_problem_free(dev)
    assert rc == B
    nrej = 0
    for b in range(B):
        ref = H.solve_oracle(prob, "dense", p0=p0[b], max_iterations=60)
        nrej += sum(1 for t in ref.trials if not t.accepted)
        assert it[b] == ref.accepted, (b, it[b], ref.accepted)
        assert abs(n2[b] - ref.norm2x) <= 1e-9 * ref.norm2x
        assert np.max(np.abs(p[b] - ref.p)) <= 1e-7 * max(1.0, np.max(np.abs(ref.p)))
    assert nrej > 10        # the fixture really exercises the rejection path


def test_batched_rejects_oversized_problems(H):
    lib = H.dlb.load()
    p = np.zeros((2, 64))
    assert lib.dogleg_gpu_optimize_dense_batched(H.as_dp(p), 64, 256, 2, H.dev_problems_lib().dlb_dev_cb_dense_batched_ptr(),
                                                 None, None, None, None) < 0
    assert b"Nstate <= 32" in lib.dogleg_gpu_last_error()
