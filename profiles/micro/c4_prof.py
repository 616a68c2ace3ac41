"""Per-level CUDA-event times of one factorization and one solve of a bundle-adjustment config (DOGLEG_GPU_SOLVE_PROF=1).
usage (GPU box): python profiles/micro/c4_prof.py [c4|c4m|c4s]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from libdogleg_b200 import ffi
from support import harness as H
import bench
cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
prob = bench.make_problem(H, cfg)
Jp, Ji = prob.pattern()
p = prob.p0()
x, Jx = prob.evaluate(p)
E = H.Engine(ffi.SOLVE_SPARSE, prob.N, prob.M, len(Ji))
E.load_sparse(0, p, x, Jp, Ji, Jx)
E.evaluate(0)
for rep in range(3):
    if rep == 2:
        os.environ["DOGLEG_GPU_SOLVE_PROF"] = "1"
    E.factorize(0, 0.0)
    E.gauss_newton(0)
E.close()
