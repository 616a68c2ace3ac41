import sys, os, time, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
os.environ["DOGLEG_GPU_PHASE_TIMING"]="1"
from support import harness as H
import libdogleg_b200 as dlb
import bench
L=dlb.load()
prob=bench.make_problem(H,"c4")
Jp,Ji=prob.pattern()
DL=H.dev_problems_lib()
dev=DL.dlb_dev_problem_create(C.cast(prob.ptr,C.c_void_p))
P=H.make_params(L,max_iterations=100)
st=np.zeros(8); ph=np.zeros(8)
for rep in range(3):
    p=prob.p0()
    t0=time.perf_counter()
    r=L.dogleg_gpu_optimize_sparse(H.as_dp(p),prob.N,prob.M,prob.nnz,H.as_ip(Jp),H.as_ip(Ji),DL.dlb_dev_cb_sparse_ptr(),C.c_void_p(dev),C.byref(P),None)
    dt=time.perf_counter()-t0
    L.dogleg_gpu_get_stats(None,H.as_dp(st)); L.dogleg_gpu_get_phase_ms(H.as_dp(ph))
    print("solve %.1f ms cost %.6g stats"%(dt*1e3,r), st.tolist(), "phases", [round(float(v),2) for v in ph], "sum", round(float(ph.sum()),1), flush=True)
