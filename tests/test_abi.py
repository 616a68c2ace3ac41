"""CPU tests of the drop-in boundary: the library loads, exports every symbol the
headers declare, keeps the reference's struct layouts, the host-only parameter
API behaves like the reference's, and -- without a GPU -- solves fail loudly
instead of falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from libdogleg_b200 import ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:dogleg|dlb)_[A-Za-z0-9_]+)\s*\(", text)) -
                  {"dogleg_callback_t", "dogleg_callback_dense_t", "dogleg_callback_dense_products_t"})


def test_every_declared_symbol_is_exported():
    L = ffi.load()
    for hdr in ("dogleg.h", "dogleg_gpu.h"):
        names = declared_functions(hdr)
        assert len(names) >= 15
        for n in names:
            if n.endswith("_t"):
                continue
            assert hasattr(L, n), f"{n} declared in {hdr} but not exported"


def test_reference_public_symbols_all_present():
    """The 20 entry points of libdogleg ABI 2."""
    L = ffi.load()
    for n in """dogleg_getDefaultParameters dogleg_setMaxIterations dogleg_setTrustregionUpdateParameters
                dogleg_setDebug dogleg_setInitialTrustregion dogleg_setThresholds dogleg_optimize dogleg_optimize2
                dogleg_optimize_dense dogleg_optimize_dense2 dogleg_optimize_dense_products
                dogleg_computeJtJfactorization dogleg_testGradient dogleg_testGradient_dense
                dogleg_testGradient_dense_products dogleg_freeContext dogleg_getOutliernessFactors
                dogleg_markOutliers dogleg_reportOutliers dogleg_getOutliernessTrace_newFeature_sparse""".split():
        assert hasattr(L, n)


def test_struct_layouts_match_reference(tmp_path):
    """sizeof/offsetof of the public structs, compiled from OUR header, against the
    numbers SURVEY.md records for the reference ([probe] 72 / 104 bytes, bit 30)."""
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stdio.h>
#include "dogleg.h"
int main(void) {
  dogleg_parameters2_t p = {.debug_vnlog = true};
  dogleg_parameters2_t q = {.debug = true}; dogleg_parameters2_t r = {.JtJ_packed = true}; dogleg_parameters2_t s = {.JtJ_upper = true};
  dogleg_operatingPoint_t o; memset(&o, 0, sizeof(o)); o.didStepToEdgeOfTrustRegion = true;
  printf("%zu %zu %zu %zu %zu %zu %zu %d %d %d %d %d %zu %zu %zu\n",
    sizeof(dogleg_parameters2_t), offsetof(dogleg_parameters2_t, trustregion0), offsetof(dogleg_parameters2_t, trustregion_threshold),
    sizeof(dogleg_operatingPoint_t), offsetof(dogleg_operatingPoint_t, Jt_x), offsetof(dogleg_operatingPoint_t, dummy_bits),
    offsetof(dogleg_operatingPoint_t, step_to_here),
    p.dogleg_debug == DOGLEG_DEBUG_VNLOG, q.dogleg_debug, r.dogleg_debug, s.dogleg_debug, o.dummy_bits[0],
    sizeof(struct dogleg_outliers_t), offsetof(dogleg_solverContext_t, beforeStep) - offsetof(dogleg_solverContext_t, f),
    offsetof(dogleg_solverContext_t, lambda) - offsetof(dogleg_solverContext_t, factorization));
  return 0; }''')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=gnu11", "-include", "string.h", "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(ROOT, "compat"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert out == ["72", "8", "64", "104", "32", "72", "88", "1", "1", "2", "4", "256", "1", "16", "8"]


def test_default_parameters_and_global_setters():
    L = ffi.load()
    P = ffi.default_parameters()
    assert (P.max_iterations, P.dogleg_debug, P.trustregion0) == (100, 0, 1e3)
    assert (P.trustregion_decrease_factor, P.trustregion_decrease_threshold) == (0.1, 0.25)
    assert (P.trustregion_increase_factor, P.trustregion_increase_threshold) == (2.0, 0.75)
    assert (P.Jt_x_threshold, P.update_threshold, P.trustregion_threshold) == (1e-8, 1e-8, 1e-8)
    assert C.sizeof(ffi.Parameters) == 72
    # the setters only touch the process-global copy, never the defaults
    L.dogleg_setMaxIterations(7)
    L.dogleg_setThresholds(1e-3, -1.0, 0.0)
    L.dogleg_setDebug(ffi.DEBUG_VNLOG)
    L.dogleg_setDebug(0)
    assert ffi.default_parameters().max_iterations == 100
    L.dogleg_setMaxIterations(100)
    L.dogleg_setThresholds(1e-8, 1e-8, 1e-8)


def test_reference_programs_compile_against_our_header():
    """The reference's sample.c and test-misc.c build unchanged against include/dogleg.h
    and link against libdogleg.so (only where /root/reference exists)."""
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference sources are not on this machine")
    out = os.path.join(ROOT, "oracle", "_ref")
    os.makedirs(out, exist_ok=True)
    lib = ffi.lib_path()
    for src, exe in (("sample.c", "sample_product"), ("test-misc.c", "test_misc_product")):
        subprocess.run(["gcc", "-O2", "-std=gnu11", "-w", "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(ROOT, "compat"), os.path.join("/root/reference", src),
                        "-o", os.path.join(out, exe), lib,
                        "-Wl,-rpath,$ORIGIN/../../libdogleg_b200", "-lm"], check=True)
    r = subprocess.run([os.path.join(out, "test_misc_product")], capture_output=True, text=True)
    assert r.returncode == 0 and "DOES match" in r.stdout


def test_no_gpu_means_loud_failure(H):
    """Without a CUDA device there is no fallback: the solve returns <0 and says why."""
    L = ffi.load()
    if L.dogleg_gpu_device_count() > 0:
        pytest.skip("a GPU is present")
    r = H.solve_product(H.Problem.sample(), "dense", max_iterations=8)
    assert r.norm2x < 0
    assert b"no CUDA device" in L.dogleg_gpu_last_error()
    assert not L.dlb_engine_create(1, 6, 100, 600, 0, 0)


def test_argument_validation_matches_reference(H):
    L = ffi.load()
    p = np.zeros(3)
    PL = H.problems_lib()
    # sparse with NJnnz == 0 is refused before anything else (reference dogleg.c:1762-1766)
    assert L.dogleg_optimize2(H.as_dp(p), 3, 5, 0, PL.dlb_cb_sparse_ptr(), None, None, None) == -1.0
    assert L.dogleg_optimize_dense2(H.as_dp(p), 3, 5, None, None, None, None) == -1.0
