"""Regenerates tests/golden/sample_*.json from the UNMODIFIED reference
(oracle/_ref, built from /root/reference by oracle/Makefile). Run in the build
container: `python tests/golden/make_golden.py`. The vnlog text is the stdout of
the reference's own `sample --diag vnlog <mode>`; the trace is every p the
reference handed to the callback with the cost it got back."""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from support import harness as H  # noqa: E402

out = {}
for mode, arg in [("sparse", "sparse"), ("dense", "dense"),
                  ("products-packed-upper", "dense-products-packed-upper"),
                  ("products-unpacked", "dense-products-unpacked")]:
    log = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "sample_ref"), "--diag", "vnlog", arg],
                         capture_output=True, text=True, check=True).stdout
    prob = H.Problem.sample()
    r = H.solve_reference(prob, mode, max_iterations=8)
    out[mode] = dict(vnlog=log, norm2x=r.norm2x, p=r.p.tolist(), ncalls=r.ncalls,
                     trace_p=r.trace_p.tolist(), trace_norm2x=r.trace_norm2x.tolist())
with open(os.path.join(HERE, "sample_reference.json"), "w") as f:
    json.dump(out, f, indent=1)

# medium problems through the reference's DENSE path (real LAPACK): the cross-check of SURVEY.md 8c
cases = {}
for name, mk in [("mrcal_2x6x12", lambda: H.Problem.mrcal(2, 6, 12, seed=7)),
                 ("mrcal_4x20x5", lambda: H.Problem.mrcal(4, 20, 5, seed=2)),
                 ("random_60x300", lambda: H.Problem.random_sparse(60, 300, 5, seed=1)),
                 ("ba_10x60", lambda: H.Problem.ba(10, 60, 3, 5, 0, seed=4)),
                 ("dense_16x256", lambda: H.Problem.dense(16, 256, seed=3))]:
    prob = mk()
    r = H.solve_reference(prob, "dense", max_iterations=20)
    cases[name] = dict(norm2x=r.norm2x, p=r.p.tolist(), ncalls=r.ncalls,
                       trace_norm2x=r.trace_norm2x.tolist(), trace_p=r.trace_p.tolist())
with open(os.path.join(HERE, "dense_reference_cases.json"), "w") as f:
    json.dump(cases, f)
print("golden fixtures written")
