"""Test harness: loads the product (libdogleg.so), the unmodified reference
(oracle/_ref/libdogleg_ref.so), the oracle restatement (oracle/_build/liboracle.so)
and the shared synthetic problems (tests/support/libdlb_problems.so), and runs
the same C callbacks through each of them.

Only tests/, __graft_entry__.smoke() and bench.py import this; the product
package never touches oracle/.
"""
import ctypes as C
import os
import sys
from dataclasses import dataclass, field

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import libdogleg_b200 as dlb            # noqa: E402
from libdogleg_b200 import ffi           # noqa: E402

dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p


class CProblem(C.Structure):
    _fields_ = [("kind", C.c_int), ("N", C.c_int), ("M", C.c_int), ("nnz", C.c_int64),
                ("Ap", ip), ("Ai", ip), ("Ax", dp), ("Adense", dp),
                ("b", dp), ("p_true", dp), ("p0", dp),
                ("trace_on", C.c_int), ("ncalls", C.c_int), ("trace_cap", C.c_int),
                ("trace_p", dp), ("trace_norm2x", dp), ("cb_seconds", C.c_double),
                ("packed", C.c_int), ("upper", C.c_int), ("nthreads", C.c_int), ("progress", C.c_void_p)]


class OrcTrial(C.Structure):
    _fields_ = [("iteration", C.c_int), ("accepted", C.c_int), ("step_type", C.c_int)] + \
               [(n, C.c_double) for n in
                ("norm2x_before", "norm2x_after", "step_len_cauchy", "step_len_gn", "step_len_interpolated",
                 "k_cauchy_to_gn", "norm2_step", "expected_improvement", "observed_improvement", "rho",
                 "trustregion_before", "trustregion_after")]


class OrcResult(C.Structure):
    _fields_ = [("norm2_x", C.c_double), ("accepted_steps", C.c_int), ("Ntrials", C.c_int),
                ("Ntrials_max", C.c_int), ("trials", C.POINTER(OrcTrial)), ("lambda_", C.c_double),
                ("use_ll", C.c_int), ("perm", ip)]


_cache = {}


def problems_lib():
    if "prob" not in _cache:
        path = os.path.join(ROOT, "tests", "support", "libdlb_problems.so")
        if not os.path.exists(path):
            import __graft_entry__ as g
            g.build_support()
        L = C.CDLL(path, mode=C.RTLD_GLOBAL)
        PP = C.POINTER(CProblem)
        L.dlb_problem_sample.restype = PP
        L.dlb_problem_random_sparse.restype = PP
        L.dlb_problem_random_sparse.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64]
        L.dlb_problem_ragged.restype = PP
        L.dlb_problem_ragged.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64]
        L.dlb_problem_mrcal.restype = PP
        L.dlb_problem_mrcal.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64]
        L.dlb_problem_ba.restype = PP
        L.dlb_problem_ba.argtypes = [C.c_int] * 5 + [C.c_uint64]
        L.dlb_problem_dense.restype = PP
        L.dlb_problem_dense.argtypes = [C.c_int, C.c_int, C.c_uint64]
        L.dlb_problem_slice.restype = PP
        L.dlb_problem_slice.argtypes = [PP, C.c_int, C.c_int]
        L.dlb_problem_free.argtypes = [PP]
        L.dlb_problem_trace.argtypes = [PP, C.c_int, C.c_int]
        L.dlb_problem_reset.argtypes = [PP]
        for n in ("dlb_cb_sparse_ptr", "dlb_cb_dense_ptr", "dlb_cb_products_ptr"):
            getattr(L, n).restype = vp
        _cache["prob"] = L
    return _cache["prob"]


def reference_lib():
    """The UNMODIFIED reference compiled by oracle/Makefile (None if it was not built)."""
    if "ref" not in _cache:
        path = os.path.join(ROOT, "oracle", "_ref", "libdogleg_ref.so")
        if not os.path.exists(path):
            _cache["ref"] = None
        else:
            L = C.CDLL(path, mode=C.RTLD_LOCAL)
            PP = C.POINTER(ffi.Parameters)
            L.dogleg_getDefaultParameters.argtypes = [PP]
            L.dogleg_optimize2.argtypes = [dp, C.c_uint, C.c_uint, C.c_uint, vp, vp, PP, C.POINTER(vp)]
            L.dogleg_optimize2.restype = C.c_double
            L.dogleg_optimize_dense2.argtypes = [dp, C.c_uint, C.c_uint, vp, vp, PP, C.POINTER(vp)]
            L.dogleg_optimize_dense2.restype = C.c_double
            L.dogleg_optimize_dense_products.argtypes = [dp, C.c_uint, vp, vp, PP, C.POINTER(vp)]
            L.dogleg_optimize_dense_products.restype = C.c_double
            L.dogleg_freeContext.argtypes = [C.POINTER(vp)]
            L.orc_shim_set_permutation.argtypes = [ip, C.c_int]
            L.orc_shim_set_ll.argtypes = [C.c_int]
            _cache["ref"] = L
    return _cache["ref"]


def oracle_lib():
    if "orc" not in _cache:
        path = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
        if not os.path.exists(path):
            import __graft_entry__ as g
            g.build_oracle()
        L = C.CDLL(path, mode=C.RTLD_LOCAL)
        PP = C.POINTER(ffi.Parameters)
        RP = C.POINTER(OrcResult)
        L.orc_optimize_sparse.argtypes = [dp, C.c_uint, C.c_uint, C.c_uint, vp, vp, PP, RP]
        L.orc_optimize_sparse.restype = C.c_double
        L.orc_optimize_dense.argtypes = [dp, C.c_uint, C.c_uint, vp, vp, PP, RP]
        L.orc_optimize_dense.restype = C.c_double
        L.orc_optimize_dense_products.argtypes = [dp, C.c_uint, vp, vp, PP, RP]
        L.orc_optimize_dense_products.restype = C.c_double
        L.orc_Jt_times_x.argtypes = [dp, C.c_int, C.c_int, ip, ip, dp, dp]
        L.orc_Jt_times_x.restype = None
        L.orc_norm2_J_times_v.argtypes = [C.c_int, ip, ip, dp, dp]
        L.orc_norm2_J_times_v.restype = C.c_double
        L.orc_sparse_JtJ_dense.argtypes = [dp, C.c_int, C.c_int, ip, ip, dp, C.c_double]
        L.orc_sparse_JtJ_dense.restype = None
        L.orc_dense_JtJ_packed_upper.argtypes = [dp, dp, C.c_int, C.c_int, C.c_double]
        L.orc_dense_JtJ_packed_upper.restype = None
        L.orc_pptrf_lower.argtypes = [dp, C.c_int]
        L.orc_pptrs_lower.argtypes = [dp, C.c_int, dp]
        L.orc_pptrs_lower.restype = None
        L.orc_min_degree.argtypes = [C.c_int, C.c_int, ip, ip, ip]
        L.orc_min_degree.restype = None
        L.orc_analyze.argtypes = [C.c_int, C.c_int, ip, ip, ip]
        L.orc_analyze.restype = vp
        L.orc_free.argtypes = [vp]
        L.orc_free.restype = None
        _cache["orc"] = L
    return _cache["orc"]


def as_dp(a):
    return a.ctypes.data_as(dp)


def as_ip(a):
    return a.ctypes.data_as(ip)


class Problem:
    """A synthetic problem living in libdlb_problems.so."""

    def __init__(self, ptr):
        self.ptr = ptr
        self.c = ptr.contents
        self.N, self.M, self.nnz = self.c.N, self.c.M, int(self.c.nnz)

    @classmethod
    def sample(cls):
        return cls(problems_lib().dlb_problem_sample())

    @classmethod
    def random_sparse(cls, N, M, nnz_per_meas, seed=1):
        return cls(problems_lib().dlb_problem_random_sparse(N, M, nnz_per_meas, seed))

    @classmethod
    def ragged(cls, N, M, kmax, seed=6):
        return cls(problems_lib().dlb_problem_ragged(N, M, kmax, seed))

    @classmethod
    def mrcal(cls, ncam, nframes, npts, seed=2):
        return cls(problems_lib().dlb_problem_mrcal(ncam, nframes, npts, seed))

    @classmethod
    def ba(cls, ncams, npoints, obs=4, window=8, longrange_permille=0, seed=4):
        return cls(problems_lib().dlb_problem_ba(ncams, npoints, obs, window, longrange_permille, seed))

    @classmethod
    def dense(cls, N, M, seed=3):
        return cls(problems_lib().dlb_problem_dense(N, M, seed))

    def slice(self, col_begin, ncols):
        q = problems_lib().dlb_problem_slice(self.ptr, col_begin, ncols)
        if not q:
            raise ValueError("bad slice")
        return Problem(q)

    def __del__(self):
        try:
            problems_lib().dlb_problem_free(self.ptr)
        except Exception:
            pass

    def p0(self):
        return np.ctypeslib.as_array(self.c.p0, shape=(self.N,)).copy()

    def pattern(self):
        """(Jp, Ji) of Jt as int32 arrays (sparse kinds; the sample problem is fully dense)."""
        if self.c.kind == 1:
            Jp = np.arange(0, (self.M + 1) * self.N, self.N, dtype=np.int32)
            Ji = np.tile(np.arange(self.N, dtype=np.int32), self.M)
            return Jp, Ji
        Jp = np.ctypeslib.as_array(self.c.Ap, shape=(self.M + 1,)).copy()
        Ji = np.ctypeslib.as_array(self.c.Ai, shape=(self.nnz,)).copy()
        return Jp, Ji

    def trace(self, on=True, cap=4096):
        problems_lib().dlb_problem_trace(self.ptr, 1 if on else 0, cap)

    def reset(self):
        problems_lib().dlb_problem_reset(self.ptr)

    def get_trace(self):
        n = min(self.c.ncalls, self.c.trace_cap)
        P = np.ctypeslib.as_array(self.c.trace_p, shape=(max(self.c.trace_cap, 1), self.N))[:n].copy()
        X = np.ctypeslib.as_array(self.c.trace_norm2x, shape=(max(self.c.trace_cap, 1),))[:n].copy()
        return P, X

    def set_layout(self, packed, upper):
        self.c.packed, self.c.upper = int(packed), int(upper)

    def evaluate(self, p):
        """x and Jt values at p through the sparse callback (numpy, for kernel-level tests)."""
        from libdogleg_b200.ffi import Parameters  # noqa: F401
        Jp, Ji = self.pattern()
        x = np.zeros(self.M)
        Jx = np.zeros(len(Ji))
        Jp2 = np.zeros(self.M + 1, dtype=np.int32)
        Ji2 = np.zeros(len(Ji), dtype=np.int32)

        class CS(C.Structure):
            _fields_ = [("nrow", C.c_size_t), ("ncol", C.c_size_t), ("nzmax", C.c_size_t),
                        ("p", vp), ("i", vp), ("nz", vp), ("x", vp), ("z", vp),
                        ("stype", C.c_int), ("itype", C.c_int), ("xtype", C.c_int), ("dtype", C.c_int),
                        ("sorted", C.c_int), ("packed", C.c_int)]
        cs = CS(self.N, self.M, len(Ji), Jp2.ctypes.data, Ji2.ctypes.data, None, Jx.ctypes.data, None,
                0, 0, 1, 0, 1, 1)
        L = problems_lib()
        L.dlb_cb_sparse.argtypes = [dp, dp, vp, vp]
        L.dlb_cb_sparse.restype = None
        pp = np.ascontiguousarray(p, dtype=np.float64)
        was = self.c.trace_on
        self.c.trace_on = 0
        L.dlb_cb_sparse(as_dp(pp), as_dp(x), C.byref(cs), C.cast(self.ptr, vp))
        self.c.trace_on = was
        return x, Jx


@dataclass
class Result:
    norm2x: float
    p: np.ndarray
    accepted: int = -1
    ncalls: int = 0
    trace_p: np.ndarray = None
    trace_norm2x: np.ndarray = None
    trials: list = field(default_factory=list)
    lam: float = 0.0
    stats: np.ndarray = None
    cb_seconds: float = 0.0


def make_params(lib, max_iterations=None, packed=False, upper=False, vnlog=False, debug=False, **kw):
    P = ffi.Parameters()
    lib.dogleg_getDefaultParameters(C.byref(P))
    if max_iterations is not None:
        P.max_iterations = max_iterations
    P.set_flags(debug=debug, packed=packed, upper=upper, vnlog=vnlog)
    for k, v in kw.items():
        setattr(P, k, v)
    return P


def _solve_c_api(lib, prob, mode, p0=None, perm=None, is_product=False, **pk):
    PL = problems_lib()
    packed = mode == "products-packed-upper"
    upper = packed
    prob.set_layout(packed, upper)
    P = make_params(lib, packed=packed, upper=upper, **pk)
    p = (prob.p0() if p0 is None else np.array(p0, dtype=np.float64)).copy()
    prob.reset()
    prob.trace(True)
    cookie = C.cast(prob.ptr, vp)
    if mode == "sparse":
        if perm is not None:
            perm = np.ascontiguousarray(perm, dtype=np.int32)
            if is_product:
                lib.dogleg_gpu_set_permutation(as_ip(perm), len(perm), 0)
            else:
                lib.orc_shim_set_permutation(as_ip(perm), len(perm))
        r = lib.dogleg_optimize2(as_dp(p), prob.N, prob.M, prob.nnz, PL.dlb_cb_sparse_ptr(), cookie,
                                 C.byref(P), None)
        if perm is not None and not is_product:
            lib.orc_shim_set_permutation(None, 0)
    elif mode == "dense":
        r = lib.dogleg_optimize_dense2(as_dp(p), prob.N, prob.M, PL.dlb_cb_dense_ptr(), cookie, C.byref(P), None)
    else:
        r = lib.dogleg_optimize_dense_products(as_dp(p), prob.N, PL.dlb_cb_products_ptr(), cookie, C.byref(P), None)
    tp, tx = prob.get_trace()
    res = Result(norm2x=r, p=p, ncalls=prob.c.ncalls, trace_p=tp, trace_norm2x=tx, cb_seconds=prob.c.cb_seconds)
    if is_product:
        st = np.zeros(8)
        lib.dogleg_gpu_get_stats(None, as_dp(st))
        res.stats = st
        res.accepted = int(st[0])
    return res


def solve_product(prob, mode, **kw):
    return _solve_c_api(dlb.load(), prob, mode, is_product=True, **kw)


def solve_reference(prob, mode, **kw):
    lib = reference_lib()
    if lib is None:
        raise RuntimeError("oracle/_ref/libdogleg_ref.so is not built")
    return _solve_c_api(lib, prob, mode, is_product=False, **kw)


def solve_oracle(prob, mode, p0=None, perm=None, use_ll=0, **pk):
    O = oracle_lib()
    PL = problems_lib()
    packed = mode == "products-packed-upper"
    prob.set_layout(packed, packed)
    P = make_params(dlb.load(), packed=packed, upper=packed, **pk)
    p = (prob.p0() if p0 is None else np.array(p0, dtype=np.float64)).copy()
    prob.reset()
    prob.trace(True)
    cookie = C.cast(prob.ptr, vp)
    ntr = 4096
    trials = (OrcTrial * ntr)()
    R = OrcResult()
    R.trials = C.cast(trials, C.POINTER(OrcTrial))
    R.Ntrials_max = ntr
    R.use_ll = use_ll
    keep = None
    if perm is not None:
        keep = np.ascontiguousarray(perm, dtype=np.int32)
        R.perm = as_ip(keep)
    if mode == "sparse":
        r = O.orc_optimize_sparse(as_dp(p), prob.N, prob.M, prob.nnz, PL.dlb_cb_sparse_ptr(), cookie,
                                  C.byref(P), C.byref(R))
    elif mode == "dense":
        r = O.orc_optimize_dense(as_dp(p), prob.N, prob.M, PL.dlb_cb_dense_ptr(), cookie, C.byref(P), C.byref(R))
    else:
        r = O.orc_optimize_dense_products(as_dp(p), prob.N, PL.dlb_cb_products_ptr(), cookie, C.byref(P), C.byref(R))
    tp, tx = prob.get_trace()
    tl = [trials[i] for i in range(R.Ntrials)]
    return Result(norm2x=r, p=p, accepted=R.accepted_steps, ncalls=prob.c.ncalls, trace_p=tp, trace_norm2x=tx,
                  trials=tl, lam=R.lambda_, cb_seconds=prob.c.cb_seconds)


def has_gpu():
    try:
        return dlb.load().dogleg_gpu_device_count() > 0
    except Exception:
        return False


class Engine:
    """Thin wrapper over the dlb_engine_* C-ABI (include/dogleg_gpu.h) for kernel-level parity tests."""

    def __init__(self, solve_type, N, M, nnz=0, packed=0, upper=0):
        self.L = dlb.load()
        self.N, self.M, self.nnz, self.type = N, M, nnz, solve_type
        self.packed, self.upper = packed, upper
        self.h = self.L.dlb_engine_create(solve_type, N, M, nnz, packed, upper)
        if not self.h:
            raise RuntimeError(self.L.dogleg_gpu_last_error().decode())

    def close(self):
        if self.h:
            self.L.dlb_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def host(self, slot, which, count, dtype=np.float64):
        ptr = self.L.dlb_engine_host_buffer(self.h, slot, which)
        ct = C.c_double if dtype == np.float64 else C.c_int
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(count,))

    def jcount(self):
        if self.type == ffi.SOLVE_SPARSE:
            return self.nnz
        if self.type == ffi.SOLVE_DENSE:
            return self.M * self.N
        return self.N * (self.N + 1) // 2 if self.packed else self.N * self.N

    def check(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.dogleg_gpu_last_error().decode())

    def scalars(self):
        return self.L.dlb_engine_scalars(self.h).contents

    def load_sparse(self, slot, p, x, Jp, Ji, Jx, perm=None, postorder=1):
        self.host(slot, ffi.BUF_P, self.N)[:] = p
        self.host(slot, ffi.BUF_X, self.M)[:] = x
        self.host(slot, ffi.BUF_JP, self.M + 1, np.int32)[:] = Jp
        self.host(slot, ffi.BUF_JI, self.nnz, np.int32)[:] = Ji
        self.host(slot, ffi.BUF_JVALUES, self.nnz)[:] = Jx
        pm = None if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
        self.check(self.L.dlb_engine_set_pattern(self.h, as_ip(np.ascontiguousarray(Jp)), as_ip(np.ascontiguousarray(Ji)),
                                                 as_ip(pm) if pm is not None else None, postorder))
        self.check(self.L.dlb_engine_upload_p(self.h, slot))

    def load_dense(self, slot, p, x, J):
        self.host(slot, ffi.BUF_P, self.N)[:] = p
        self.host(slot, ffi.BUF_X, self.M)[:] = x
        self.host(slot, ffi.BUF_JVALUES, self.M * self.N)[:] = np.asarray(J).ravel()
        self.check(self.L.dlb_engine_upload_p(self.h, slot))

    def load_products(self, slot, p, xtJ, JtJ_layout):
        self.host(slot, ffi.BUF_P, self.N)[:] = p
        self.host(slot, ffi.BUF_JTX, self.N)[:] = xtJ
        self.host(slot, ffi.BUF_JVALUES, self.jcount())[:] = JtJ_layout
        self.check(self.L.dlb_engine_upload_p(self.h, slot))

    def evaluate(self, slot, norm2x_products=0.0):
        self.check(self.L.dlb_engine_evaluate(self.h, slot, 1, norm2x_products))
        return self.scalars()

    def cauchy(self, slot):
        self.check(self.L.dlb_engine_cauchy(self.h, slot))
        return self.scalars()

    def factorize(self, slot, lam=0.0):
        self.check(self.L.dlb_engine_factorize(self.h, slot, lam))
        return self.scalars().minor

    def gauss_newton(self, slot):
        self.check(self.L.dlb_engine_gauss_newton(self.h, slot))
        return self.scalars()

    def step(self, frm, to, kind, delta):
        self.check(self.L.dlb_engine_step(self.h, frm, to, kind, delta))
        return self.scalars()

    def has_trial(self):
        return bool(self.L.dlb_engine_has_trial(self.h))

    def trial(self, frm, to, delta, lam=0.0):
        """the whole trial step in one launch (dlb_trial.cu)"""
        self.check(self.L.dlb_engine_trial(self.h, frm, to, delta, lam))
        return self.scalars()

    def download(self, slot):
        self.check(self.L.dlb_engine_download(self.h, slot))
        return {k: self.host(slot, w, self.N).copy() for k, w in
                (("p", ffi.BUF_P), ("Jtx", ffi.BUF_JTX), ("cauchy", ffi.BUF_CAUCHY), ("gn", ffi.BUF_GN),
                 ("step", ffi.BUF_STEP))}

    def JtJ(self, slot, lam=0.0):
        out = np.zeros((self.N, self.N))
        self.check(self.L.dlb_engine_debug_JtJ(self.h, slot, lam, as_dp(out)))
        return out

    def solve(self, B):
        B = np.asfortranarray(B, dtype=np.float64)
        nrhs = 1 if B.ndim == 1 else B.shape[1]
        X = np.zeros_like(B, order="F")
        self.check(self.L.dlb_engine_solve(self.h, as_dp(B), as_dp(X), nrhs))
        return X


def dev_problems_lib():
    if "dev" not in _cache:
        path = os.path.join(ROOT, "tests", "support", "libdlb_problems_dev.so")
        L = C.CDLL(path, mode=C.RTLD_LOCAL)
        L.dlb_dev_problem_create.restype = vp
        L.dlb_dev_problem_create.argtypes = [vp]
        L.dlb_dev_problem_create_batched.restype = vp
        L.dlb_dev_problem_create_batched.argtypes = [C.c_int, C.c_int, C.c_int, C.c_ulonglong, dp]
        L.dlb_dev_problem_create_sample_batched.restype = vp
        L.dlb_dev_problem_create_sample_batched.argtypes = [vp, C.c_int]
        L.dlb_dev_problem_free.argtypes = [vp]
        L.dlb_dev_problem_timing.argtypes = [vp, C.c_int]
        L.dlb_dev_problem_ms.argtypes = [vp]
        L.dlb_dev_problem_ms.restype = C.c_double
        L.dlb_dev_problem_ncalls.argtypes = [vp]
        for n in ("dlb_dev_cb_sparse_ptr", "dlb_dev_cb_dense_ptr", "dlb_dev_cb_dense_batched_ptr"):
            getattr(L, n).restype = vp
        _cache["dev"] = L
    return _cache["dev"]


def solve_product_device(prob, mode, p0=None, **pk):
    """dogleg_gpu_optimize_sparse / _dense with the device-resident model as callback."""
    lib = dlb.load()
    DL = dev_problems_lib()
    dev = DL.dlb_dev_problem_create(C.cast(prob.ptr, vp))
    assert dev
    P = make_params(lib, **pk)
    p = (prob.p0() if p0 is None else np.array(p0, dtype=np.float64)).copy()
    if mode == "sparse":
        Jp, Ji = prob.pattern()
        r = lib.dogleg_gpu_optimize_sparse(as_dp(p), prob.N, prob.M, prob.nnz, as_ip(Jp), as_ip(Ji),
                                           DL.dlb_dev_cb_sparse_ptr(), C.c_void_p(dev), C.byref(P), None)
    else:
        r = lib.dogleg_gpu_optimize_dense(as_dp(p), prob.N, prob.M, DL.dlb_dev_cb_dense_ptr(), C.c_void_p(dev),
                                          C.byref(P), None)
    st = np.zeros(8)
    lib.dogleg_gpu_get_stats(None, as_dp(st))
    ncalls = DL.dlb_dev_problem_ncalls(dev)
    DL.dlb_dev_problem_free(dev)
    return Result(norm2x=r, p=p, accepted=int(st[0]), ncalls=ncalls, stats=st)


def solve_batched(dev, p0, N, M, **pk):
    """dogleg_gpu_optimize_dense_batched over the device problem set `dev`; p0 is B x N."""
    lib = dlb.load()
    DL = dev_problems_lib()
    P = make_params(lib, **pk)
    p = np.ascontiguousarray(p0, dtype=np.float64).copy()
    B = p.shape[0]
    n2 = np.zeros(B)
    it = np.zeros(B, dtype=np.int32)
    rc = lib.dogleg_gpu_optimize_dense_batched(as_dp(p), N, M, B, DL.dlb_dev_cb_dense_batched_ptr(), C.c_void_p(dev),
                                               C.byref(P), as_dp(n2), as_ip(it))
    if rc < 0:
        raise RuntimeError(lib.dogleg_gpu_last_error().decode())
    return rc, p, n2, it


def shard_columns(M, world, align=1):
    """Contiguous, nearly equal column ranges [begin, end) per rank, boundaries on multiples of `align`
    (e.g. the measurements of one frame) -- the row partition of SURVEY.md 8e."""
    units = M // align
    assert units * align == M
    cuts = [(units * r) // world * align for r in range(world + 1)]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def solve_product_sharded(prob_global, local_prob, col_begin, device_callbacks=True, p0=None, **pk):
    """dogleg_gpu_optimize_sparse_sharded on this rank's slice; NCCL must be initialised
    (dogleg_gpu_nccl_init) when world > 1."""
    lib = dlb.load()
    lib.dogleg_gpu_optimize_sparse_sharded.restype = C.c_double
    lib.dogleg_gpu_optimize_sparse_sharded.argtypes = [dp, C.c_uint, C.c_uint, ip, ip, C.c_uint, C.c_uint, vp, vp, vp,
                                                       C.POINTER(ffi.Parameters), C.POINTER(vp)]
    P = make_params(lib, **pk)
    p = (prob_global.p0() if p0 is None else np.array(p0, dtype=np.float64)).copy()
    Jp, Ji = prob_global.pattern()
    dev = None
    if device_callbacks:
        DL = dev_problems_lib()
        dev = DL.dlb_dev_problem_create(C.cast(local_prob.ptr, vp))
        assert dev
        r = lib.dogleg_gpu_optimize_sparse_sharded(as_dp(p), prob_global.N, prob_global.M, as_ip(Jp), as_ip(Ji),
                                                   col_begin, local_prob.M, None, DL.dlb_dev_cb_sparse_ptr(),
                                                   C.c_void_p(dev), C.byref(P), None)
        DL.dlb_dev_problem_free(dev)
        cbs = 0.0
    else:
        local_prob.reset()
        local_prob.trace(False)
        r = lib.dogleg_gpu_optimize_sparse_sharded(as_dp(p), prob_global.N, prob_global.M, as_ip(Jp), as_ip(Ji),
                                                   col_begin, local_prob.M, problems_lib().dlb_cb_sparse_ptr(), None,
                                                   C.cast(local_prob.ptr, vp), C.byref(P), None)
        cbs = local_prob.c.cb_seconds
    st = np.zeros(8)
    lib.dogleg_gpu_get_stats(None, as_dp(st))
    return Result(norm2x=r, p=p, accepted=int(st[0]), stats=st, cb_seconds=cbs)
