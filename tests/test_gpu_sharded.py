"""Row-sharded solves. On one GPU the sharded entry point is exercised with a
single rank (the partial-fronts path forced on, no collective); with >= 2 GPUs visible the real
NCCL path runs under torchrun (tests/multi_gpu/sharded_check.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("force", ["0", "1"])
def test_sharded_entry_point_single_rank(H, force, monkeypatch):
    monkeypatch.setenv("DOGLEG_GPU_FORCE_REDUCE_PATH", force)
    prob = H.Problem.mrcal(3, 8, 6, seed=11)
    ref = H.solve_oracle(prob, "sparse", max_iterations=30, trustregion0=0.3)
    for devcb in (True, False):
        got = H.solve_product_sharded(prob, prob.slice(0, prob.M), 0, device_callbacks=devcb,
                                      max_iterations=30, trustregion0=0.3)
        assert got.accepted == ref.accepted
        assert abs(got.norm2x - ref.norm2x) <= 1e-9 * ref.norm2x
        assert np.max(np.abs(got.p - ref.p)) <= 1e-7 * max(1.0, np.max(np.abs(ref.p)))


def test_gather_mode_single_rank(H, monkeypatch):
    """The gather flavour of a sharded solve (full-size device arrays, this rank's slice written in
    place) with one rank holding everything: must equal the plain solve."""
    monkeypatch.setenv("DOGLEG_GPU_SHARD_MODE", "gather")
    monkeypatch.setenv("DOGLEG_GPU_ENGINE_CACHE", "0")
    prob = H.Problem.ba(30, 600, 4, 12, 0, seed=8)
    ref = H.solve_oracle(prob, "sparse", max_iterations=20)
    for devcb in (True, False):
        got = H.solve_product_sharded(prob, prob.slice(0, prob.M), 0, device_callbacks=devcb, max_iterations=20)
        assert got.accepted == ref.accepted
        assert abs(got.norm2x - ref.norm2x) <= 1e-9 * ref.norm2x
        assert np.max(np.abs(got.p - ref.p)) <= 1e-7 * max(1.0, np.max(np.abs(ref.p)))


def test_a_slice_alone_equals_the_sliced_problem(H, monkeypatch):
    """A rank that holds only some of the columns (and nobody to sum with) must solve exactly the
    problem made of those columns: checks the member filtering / re-basing of the sharded layout."""
    monkeypatch.setenv("DOGLEG_GPU_FORCE_REDUCE_PATH", "1")
    prob = H.Problem.random_sparse(30, 400, 12, seed=6)
    b, e = 120, 330
    part = prob.slice(b, e - b)
    ref = H.solve_product(part, "sparse", max_iterations=30)
    got = H.solve_product_sharded(prob, part, b, device_callbacks=True, max_iterations=30)
    assert got.accepted == ref.accepted
    assert abs(got.norm2x - ref.norm2x) <= 1e-9 * ref.norm2x
    assert np.max(np.abs(got.p - ref.p)) <= 1e-7 * max(1.0, np.max(np.abs(ref.p)))


@pytest.mark.parametrize("force", ["0", "1"])
def test_dense_sharded_entry_point_single_rank(H, force, monkeypatch):
    """dogleg_gpu_optimize_dense_sharded with one rank holding all rows (and, forced, the path that
    would all-reduce the N x N J'J): same walk as the reference's dense solve."""
    import ctypes as C
    from libdogleg_b200 import ffi
    monkeypatch.setenv("DOGLEG_GPU_FORCE_REDUCE_PATH", force)
    prob = H.Problem.dense(16, 256, seed=3)
    ref = H.solve_oracle(prob, "dense", max_iterations=30)
    lib = ffi.load()
    lib.dogleg_gpu_optimize_dense_sharded.restype = C.c_double
    lib.dogleg_gpu_optimize_dense_sharded.argtypes = [H.dp, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p,
                                                      C.c_void_p, C.c_void_p, C.c_void_p]
    P = H.make_params(lib, max_iterations=30)
    p = prob.p0().copy()
    prob.reset()
    r = lib.dogleg_gpu_optimize_dense_sharded(H.as_dp(p), prob.N, prob.M, 0, prob.M, H.problems_lib().dlb_cb_dense_ptr(), None,
                                              C.cast(prob.ptr, C.c_void_p), C.cast(C.byref(P), C.c_void_p), None)
    assert r >= 0
    assert abs(r - ref.norm2x) <= 1e-9 * ref.norm2x
    assert np.max(np.abs(p - ref.p)) <= 1e-7 * max(1.0, np.max(np.abs(ref.p)))


def test_two_gpus_nccl(H):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(ROOT, "tests", "multi_gpu", "sharded_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert "SHARDED_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_solves_on_the_second_device_of_one_process(H):
    """dogleg_gpu_set_device(1) after solves on device 0 in the same process: engines, the batched
    workspace and the kernels' shared-memory opt-ins (cudaFuncSetAttribute is per device) must all
    follow; then back to device 0."""
    import torch
    from libdogleg_b200 import ffi
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    L = ffi.load()
    cases = [(H.Problem.mrcal(3, 5, 150, seed=9), "sparse"), (H.Problem.ba(60, 1500, 4, 24, 0, seed=4), "sparse"),
             (H.Problem.dense(200, 1024, seed=3), "dense")]
    refs = [H.solve_oracle(p, m, max_iterations=20) for p, m in cases]
    try:
        for dev in (0, 1, 0):
            assert L.dogleg_gpu_set_device(dev) == 0
            assert L.dogleg_gpu_get_device() == dev
            for (prob, mode), ref in zip(cases, refs):
                got = H.solve_product(prob, mode, max_iterations=20)
                assert got.norm2x >= 0, L.dogleg_gpu_last_error()
                assert got.ncalls == ref.ncalls
                assert abs(got.norm2x - ref.norm2x) <= 1e-9 * ref.norm2x
                assert np.max(np.abs(got.p - ref.p)) <= 1e-7 * max(1.0, np.max(np.abs(ref.p)))
    finally:
        L.dogleg_gpu_set_device(0)
