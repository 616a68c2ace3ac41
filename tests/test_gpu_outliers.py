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


@pytest.mark.parametrize("chunked", ["0", "1"])
@pytest.mark.parametrize("featureSize", [1, 2])
def test_outlierness_factors_sparse_match_reference(H, monkeypatch, featureSize, chunked):
    """chunked=0: inv(JtJ) once + one kernel over all features (dlb_engine_outlier_products);
    chunked=1: the multi-right-hand-side solves per 64 measurements (the path for > 16384 states)."""
    if H.reference_lib() is None:
        pytest.skip("oracle/_ref not built")
    monkeypatch.setenv("DOGLEG_GPU_OUTLIER_CHUNKED", chunked)
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


def _dense_JtJ(H, prob, p):
    Jp, Ji = prob.pattern()
    x, Jx = prob.evaluate(p)
    J = np.zeros((prob.M, prob.N))
    for j in range(prob.M):
        J[j, Ji[Jp[j]:Jp[j + 1]]] = Jx[Jp[j]:Jp[j + 1]]
    return J.T @ J


@pytest.mark.parametrize("mk", [lambda H: H.Problem.mrcal(3, 8, 6, seed=11), lambda H: H.Problem.random_sparse(60, 300, 5),
                                lambda H: H.Problem.ba(20, 200, 4, 8, 50)])
def test_returned_context_solve_and_factor_export(H, mk):
    """SURVEY.md 8f2 / reference dogleg.h:188-195, README.pod:105-111: a caller that keeps the context can
    solve with the factorization (dogleg_gpu_solve) and can get the numeric factor itself
    (dogleg_gpu_export_factor: CHOLMOD's supernodal layout); L L' must equal P (JtJ + lambda I) P'."""
    lib = H.dlb.load()
    prob = mk(H)
    ctx, point, P = solve_keep_context(H, lib, prob, "sparse")
    N = prob.N
    lib.dogleg_gpu_solve.argtypes = [C.c_void_p, H.dp, H.dp, C.c_int]
    lib.dogleg_gpu_export_factor.argtypes = [C.c_void_p]
    p_final = np.ctypeslib.as_array(C.cast(C.c_void_p.from_address(point).value, H.dp), shape=(N,)).copy()
    A = _dense_JtJ(H, prob, p_final)
    B = np.asfortranarray(np.random.default_rng(3).standard_normal((N, 3)))
    X = np.zeros_like(B, order="F")
    assert lib.dogleg_gpu_solve(ctx, H.as_dp(B), H.as_dp(X), 3) == 0
    assert np.max(np.abs(A @ X - B)) <= 1e-7 * np.max(np.abs(B)) * max(1.0, np.sqrt(np.linalg.cond(A)) * 1e-3)

    assert lib.dogleg_gpu_export_factor(ctx) == 0

    class Factor(C.Structure):          # compat/cholmod.h cholmod_factor, the members used here
        _fields_ = [("n", C.c_size_t), ("minor", C.c_size_t), ("Perm", C.c_void_p), ("ColCount", C.c_void_p), ("IPerm", C.c_void_p),
                    ("nzmax", C.c_size_t), ("p", C.c_void_p), ("i", C.c_void_p), ("x", C.c_void_p), ("z", C.c_void_p),
                    ("nz", C.c_void_p), ("next", C.c_void_p), ("prev", C.c_void_p),
                    ("nsuper", C.c_size_t), ("ssize", C.c_size_t), ("xsize", C.c_size_t), ("maxcsize", C.c_size_t),
                    ("maxesize", C.c_size_t), ("super", C.c_void_p), ("pi", C.c_void_p), ("px", C.c_void_p), ("s", C.c_void_p)]
    import subprocess, tempfile
    src = ('#include <stdio.h>\n#include <stddef.h>\n#include "dogleg.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", '
           'offsetof(dogleg_solverContext_t, factorization), offsetof(cholmod_factor, Perm), offsetof(cholmod_factor, x), '
           'offsetof(cholmod_factor, nsuper), offsetof(cholmod_factor, super), offsetof(cholmod_factor, pi), '
           'offsetof(cholmod_factor, px), offsetof(cholmod_factor, s));return 0;}')
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "o.c"), "w").write(src)
        subprocess.run(["gcc", "-std=gnu11", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "compat"),
                        os.path.join(td, "o.c"), "-o", os.path.join(td, "o")], check=True)
        offs = [int(v) for v in subprocess.run([os.path.join(td, "o")], capture_output=True, text=True).stdout.split()]
    Lp = C.c_void_p.from_address(ctx.value + offs[0]).value
    rd = lambda o: C.c_void_p.from_address(Lp + o).value
    nsuper = C.c_size_t.from_address(Lp + offs[3]).value
    ints = lambda ptr, n: np.ctypeslib.as_array(C.cast(ptr, H.ip), shape=(n,)).copy()
    perm, sup, pi, px = ints(rd(offs[1]), N), ints(rd(offs[4]), nsuper + 1), ints(rd(offs[5]), nsuper + 1), ints(rd(offs[6]), nsuper + 1)
    srows = ints(rd(offs[7]), pi[nsuper])
    xs = np.ctypeslib.as_array(C.cast(rd(offs[2]), H.dp), shape=(px[nsuper],)).copy()
    Lmat = np.zeros((N, N))
    for k in range(nsuper):
        rows = srows[pi[k]:pi[k + 1]]
        ncol = sup[k + 1] - sup[k]
        panel = xs[px[k]:px[k] + len(rows) * ncol].reshape((ncol, len(rows))).T       # column-major, ld = nsrow
        for c in range(ncol):
            Lmat[rows[c:], sup[k] + c] = panel[c:, c]
    PAP = A[np.ix_(perm, perm)]
    assert np.max(np.abs(Lmat @ Lmat.T - PAP)) <= 1e-10 * np.max(np.abs(PAP))
    lib.dogleg_freeContext.argtypes = [C.POINTER(C.c_void_p)]
    lib.dogleg_freeContext(C.byref(ctx))


def test_outlierness_factors_at_full_c2_scale(H, capsys):
    """f1 at the size of BASELINE.json configs[1] (1 M measurements, 1268 states): every feature's
    J* inv(JtJ) J*' on the device, timed; a random sample of the factors is checked against numpy on the dense
    JtJ of the same point (reference dogleg.c:2401-2791; the reference itself needs 250 000 cholmod_solve calls
    of 4 columns for this and is not run)."""
    import time
    prob = H.Problem.mrcal(4, 200, 625)
    lib = H.dlb.load()
    ctx, point, P = solve_keep_context(H, lib, prob, "sparse")
    featureSize, nfeat = 2, prob.M // 2
    factors_of(lib, ctx, point, featureSize, 1000)                     # warm-up (allocations)
    t0 = time.perf_counter()
    f, scale = factors_of(lib, ctx, point, featureSize, nfeat)
    dt = time.perf_counter() - t0
    with capsys.disabled():
        print(f"\n[f1] dogleg_getOutliernessFactors, {nfeat} features of size 2, Nstate {prob.N}: {dt * 1e3:.1f} ms")
    assert dt < 5.0
    # numpy on a sample: A = J* inv(JtJ) J*' from the Jacobian at the final point
    PT = C.c_void_p(point)
    pbuf = np.ctypeslib.as_array(C.cast(C.c_void_p.from_address(point).value, C.POINTER(C.c_double)), shape=(prob.N,)).copy()
    x, Jx = prob.evaluate(pbuf)
    Jp, Ji = prob.pattern()
    JtJ = np.zeros((prob.N, prob.N))
    rows = np.repeat(np.arange(prob.M), np.diff(Jp))
    import scipy.sparse as sp
    J = sp.csr_matrix((Jx, Ji, Jp), shape=(prob.M, prob.N))
    JtJ = (J.T @ J).toarray()
    Binv = np.linalg.inv(JtJ)
    rng = np.random.default_rng(0)
    k = scale
    for fi in rng.integers(0, nfeat, 40):
        Jf = J[2 * fi:2 * fi + 2].toarray()
        A = Jf @ Binv @ Jf.T
        xf = x[2 * fi:2 * fi + 2]
        det = (1.0 - A[0, 0]) * (1.0 - A[1, 1]) - A[0, 1] ** 2
        B = np.array([[A[1, 1] - 1.0, -A[0, 1]], [-A[0, 1], A[0, 0] - 1.0]])
        ref = (xf @ B @ xf) / det + np.sum((B @ xf) ** 2) / det ** 2
        assert np.isclose(f[fi], ref * k / 8.0, rtol=1e-6, atol=1e-12)
    lib.dogleg_freeContext.argtypes = [C.POINTER(C.c_void_p)]
    lib.dogleg_freeContext(C.byref(ctx))
