"""GPU parity tests of the outlier / confidence helpers against the UNMODIFIED
reference run on the same problem: dogleg_getOutliernessFactors (sparse and dense),
dogleg_getOutliernessTrace_newFeature_sparse and dogleg_markOutliers."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def before_offset():
    src = ('#include <stdio.h>\n#include "dogleg.h"\nint main(void){printf("%zu\\n", '
           'offsetof(dogleg_solverContext_t, beforeStep));return 0;}')
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "o.c"), "w").write(src)
        subprocess.run(["gcc", "-std=gnu11", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "compat"),
                        os.path.join(td, "o.c"), "-o", os.path.join(td, "o")], check=True)
        return int(subprocess.run([os.path.join(td, "o")], capture_output=True, text=True).stdout)


def solve_keep_context(H, lib, prob, mode):
    PL = H.problems_lib()
    P = H.make_params(H.dlb.load(), max_iterations=30)
    p = prob.p0()
    ctx = C.c_void_p()
    prob.reset()
    prob.trace(False)
    cookie = C.cast(prob.ptr, C.c_void_p)
    if mode == "sparse":
        r = lib.dogleg_optimize2(H.as_dp(p), prob.N, prob.M, prob.nnz, PL.dlb_cb_sparse_ptr(), cookie, C.byref(P), C.byref(ctx))
    else:
        r = lib.dogleg_optimize_dense2(H.as_dp(p), prob.N, prob.M, PL.dlb_cb_dense_ptr(), cookie, C.byref(P), C.byref(ctx))
    assert r >= 0 and ctx.value
    point = C.c_void_p.from_address(ctx.value + before_offset()).value
    return ctx, point, P


def factors_of(lib, ctx, point, featureSize, Nfeatures, nout=0):
    lib.dogleg_getOutliernessFactors.restype = C.c_bool
    lib.dogleg_getOutliernessFactors.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int,
                                                 C.c_void_p, C.c_void_p]
    f = np.zeros(Nfeatures)
    scale = C.c_double(-1.0)
    ok = lib.dogleg_getOutliernessFactors(f.ctypes.data_as(C.POINTER(C.c_double)), C.byref(scale), featureSize, Nfeatures,
                                          nout, point, ctx)
    assert ok
    return f, scale.value


@pytest.mark.parametrize("featureSize", [1, 2])
def test_outlierness_factors_sparse_match_reference(H, featureSize):
    if H.reference_lib() is None:
        pytest.skip("oracle/_ref not built")
    prob = H.Problem.mrcal(2, 6, 12, seed=7)
    nfeat = prob.M // featureSize
    res = []
    for lib in (H.reference_lib(), H.dlb.load()):
        ctx, point, P = solve_keep_context(H, lib, prob, "sparse")
        res.append(factors_of(lib, ctx, point, featureSize, nfeat))
        lib.dogleg_freeContext.argtypes = [C.POINTER(C.c_void_p)]
        lib.dogleg_freeContext(C.byref(ctx))
    (fr, sr), (fg, sg) = res
    assert np.isclose(sr, sg, rtol=1e-9)
    assert np.allclose(fg, fr, rtol=1e-7, atol=1e-12 * np.max(np.abs(fr)))
    assert np.max(fr) > 0


def test_outlierness_factors_dense_match_reference(H):
    if H.reference_lib() is None:
        pytest.skip("oracle/_ref not built")
    prob = H.Problem.dense(16, 256, seed=3)
    res = []
    for lib in (H.reference_lib(), H.dlb.load()):
        ctx, point, P = solve_keep_context(H, lib, prob, "dense")
        res.append(factors_of(lib, ctx, point, 1, prob.M))
        lib.dogleg_freeContext.argtypes = [C.POINTER(C.c_void_p)]
        lib.dogleg_freeContext(C.byref(ctx))
    (fr, sr), (fg, sg) = res
    assert np.isclose(sr, sg, rtol=1e-9)
    assert np.allclose(fg, fr, rtol=1e-7, atol=1e-12 * np.max(np.abs(fr)))


def test_dense_and_sparse_agree_for_feature_size_two(H):
    """featureSize == 2: the reference's dense variant mis-indexes the second row (dogleg.c:2490);
    ours is consistent between the dense and the sparse route on the same (densified) problem."""
    prob = H.Problem.mrcal(2, 6, 12, seed=7)
    lib = H.dlb.load()
    out = []
    for mode in ("sparse", "dense"):
        ctx, point, P = solve_keep_context(H, lib, prob, mode)
        out.append(factors_of(lib, ctx, point, 2, prob.M // 2)[0])
        lib.dogleg_freeContext.argtypes = [C.POINTER(C.c_void_p)]
        lib.dogleg_freeContext(C.byref(ctx))
    assert np.allclose(out[0], out[1], rtol=1e-6, atol=1e-12 * np.max(np.abs(out[0])))


def test_new_feature_trace_matches_reference(H):
    if H.reference_lib() is None:
        pytest.skip("oracle/_ref not built")
    prob = H.Problem.mrcal(2, 6, 12, seed=7)
    rng = np.random.default_rng(4)
    istate, nact = 20, 12
    Jq = rng.standard_normal((2, nact))
    vals = []
    for lib in (H.reference_lib(), H.dlb.load()):
        ctx, point, P = solve_keep_context(H, lib, prob, "sparse")
        fn = lib.dogleg_getOutliernessTrace_newFeature_sparse
        fn.restype = C.c_double
        fn.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        vals.append(fn(Jq.ctypes.data_as(C.POINTER(C.c_double)), istate, nact, 2, 0, point, ctx))
        lib.dogleg_freeContext.argtypes = [C.POINTER(C.c_void_p)]
        lib.dogleg_freeContext(C.byref(ctx))
    assert vals[0] > 0 and np.isclose(vals[0], vals[1], rtol=1e-8)


def test_mark_outliers_with_a_planted_outlier(H):
    """Corrupt one measurement strongly: its feature must get factor >= 1 and be marked when the
    confidence callback says removing it costs nothing."""
    prob = H.Problem.mrcal(2, 6, 12, seed=7)
    b = np.ctypeslib.as_array(prob.c.b, shape=(prob.M,))
    b[77] += 5.0
    lib = H.dlb.load()
    ctx, point, P = solve_keep_context(H, lib, prob, "sparse")
    f, _ = factors_of(lib, ctx, point, 1, prob.M)
    assert np.argmax(f) == 77 and f[77] >= 1.0

    class Outl(C.Structure):
        _fields_ = [("marked", C.c_ubyte, 1)]
    marked = (Outl * prob.M)()
    CONF = C.CFUNCTYPE(C.c_double, C.c_int)
    conf = CONF(lambda i: 1.0)
    lib.dogleg_markOutliers.restype = C.c_bool
    lib.dogleg_markOutliers.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), CONF, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p]
    scale, nout = C.c_double(-1.0), C.c_int(0)
    assert lib.dogleg_markOutliers(marked, C.byref(scale), C.byref(nout), conf, 1, prob.M, point, ctx)
    assert marked[77].marked == 1 and nout.value >= 1
    lib.dogleg_freeContext.argtypes = [C.POINTER(C.c_void_p)]
    lib.dogleg_freeContext(C.byref(ctx))
