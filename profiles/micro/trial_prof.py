"""Phase stamps of the persistent trial kernel on the C2 problem (DOGLEG_GPU_TRIAL_PROF=1; with a library built
with `make NVFLAGS_EXTRA=-DDLB_TRIAL_DEBUG` also the SM-cycle stamps inside front_eliminate of CTA 0).
usage (GPU box): DOGLEG_GPU_TRIAL_PROF=1 python profiles/micro/trial_prof.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("DOGLEG_GPU_TRIAL_PROF", "1")
from libdogleg_b200 import ffi
from support import harness as H
prob = H.Problem.mrcal(4, 200, 625)
Jp, Ji = prob.pattern()
p = prob.p0()
x, Jx = prob.evaluate(p)
E = H.Engine(ffi.SOLVE_SPARSE, prob.N, prob.M, len(Ji))
E.load_sparse(0, p, x, Jp, Ji, Jx)
for _ in range(4):
    E.evaluate(0)          # a fresh point: nothing cached
    E.trial(0, 1, 1e9)
E.close()
