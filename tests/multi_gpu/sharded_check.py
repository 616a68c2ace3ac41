"""Run under torchrun on >= 2 GPUs (gpurun --gpus 2): every rank solves its row slice of one sparse
problem with dogleg_gpu_optimize_sparse_sharded (NCCL all-reduce of the partial gradient, |x|^2,
|J v|^2 and partial fronts); all ranks must agree bit for bit, and with the single-GPU solve and
the oracle within the parity tolerances. Prints SHARDED_OK on success."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from support import harness as H  # noqa: E402
import libdogleg_b200 as dlb  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    L = dlb.load()
    L.dogleg_gpu_set_device(local)
    L.dogleg_gpu_nccl_get_unique_id.argtypes = [C.c_void_p]
    L.dogleg_gpu_nccl_init.argtypes = [C.c_int, C.c_int, C.c_void_p]
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_ubyte * 128)()
        assert L.dogleg_gpu_nccl_get_unique_id(buf) == 0, L.dogleg_gpu_last_error()
        uid = torch.tensor(list(buf), dtype=torch.uint8)
    uid = uid.cuda()
    dist.broadcast(uid, 0)
    idb = (C.c_ubyte * 128)(*uid.cpu().tolist())
    assert L.dogleg_gpu_nccl_init(rank, world, idb) == 0, L.dogleg_gpu_last_error()

    ok = True
    # "reduce": partial fronts summed (small states); "gather": Jacobian slices exchanged, everything
    # replicated (what large bundle-adjustment problems use); both must give the single-GPU answer
    for mk, align, tr0, mode in [(lambda: H.Problem.mrcal(3, 8, 6, seed=11), 2 * 3 * 6, 1e3, "reduce"),
                                 (lambda: H.Problem.mrcal(3, 8, 6, seed=11), 2 * 3 * 6, 0.3, "reduce"),
                                 (lambda: H.Problem.mrcal(4, 40, 125, seed=2), 2 * 4 * 125, 1e3, "reduce"),
                                 (lambda: H.Problem.mrcal(3, 8, 6, seed=11), 2 * 3 * 6, 0.3, "gather"),
                                 (lambda: H.Problem.ba(60, 1500, 4, 24, 0, seed=4), 8, 1e3, "gather")]:
        os.environ["DOGLEG_GPU_SHARD_MODE"] = mode
        os.environ["DOGLEG_GPU_ENGINE_CACHE"] = "0"
        prob = mk()
        b, e = H.shard_columns(prob.M, world, align)[rank]
        local_prob = prob.slice(b, e - b)
        for devcb in (True, False):
            got = H.solve_product_sharded(prob, local_prob, b, device_callbacks=devcb, max_iterations=30, trustregion0=tr0)
            assert got.norm2x >= 0, L.dogleg_gpu_last_error()
            # all ranks identical, bit for bit
            t = torch.tensor(np.append(got.p, got.norm2x), device="cuda")
            ref = t.clone()
            dist.broadcast(ref, 0)
            same = bool(torch.equal(t, ref))
            if rank == 0:
                single = H.solve_product(prob, "sparse", max_iterations=30, trustregion0=tr0)
                orc = H.solve_oracle(prob, "sparse", max_iterations=30, trustregion0=tr0)
                good = (got.accepted == single.accepted == orc.accepted and
                        abs(got.norm2x - orc.norm2x) <= 1e-9 * orc.norm2x and
                        np.max(np.abs(got.p - orc.p)) <= 1e-7 * max(1.0, np.max(np.abs(orc.p))))
                print(f"rank0: {mode} N={prob.N} M={prob.M} tr0={tr0} devcb={devcb} accepted={got.accepted} "
                      f"cost={got.norm2x:.12g} vs oracle {orc.norm2x:.12g} parity={'ok' if good else 'FAIL'}", flush=True)
                ok = ok and good
            ok = ok and same
    # dense, rows of J split over the ranks (dogleg_gpu_optimize_dense_sharded, host callbacks)
    os.environ.pop("DOGLEG_GPU_SHARD_MODE", None)
    L.dogleg_gpu_optimize_dense_sharded.restype = C.c_double
    L.dogleg_gpu_optimize_dense_sharded.argtypes = [H.dp, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p,
                                                    C.c_void_p, C.c_void_p, C.c_void_p]
    for N, M in [(16, 256), (200, 1024)]:
        prob = H.Problem.dense(N, M, seed=3)
        b, e = H.shard_columns(M, world, 1)[rank]
        local_prob = prob.slice(b, e - b)
        P = H.make_params(L, max_iterations=30)
        p = prob.p0().copy()
        local_prob.reset()
        r = L.dogleg_gpu_optimize_dense_sharded(H.as_dp(p), N, M, b, e - b, H.problems_lib().dlb_cb_dense_ptr(), None,
                                                C.cast(local_prob.ptr, C.c_void_p), C.cast(C.byref(P), C.c_void_p), None)
        assert r >= 0, L.dogleg_gpu_last_error()
        t = torch.tensor(np.append(p, r), device="cuda")
        ref = t.clone()
        dist.broadcast(ref, 0)
        ok = ok and bool(torch.equal(t, ref))
        if rank == 0:
            orc = H.solve_oracle(prob, "dense", max_iterations=30)
            good = abs(r - orc.norm2x) <= 1e-9 * orc.norm2x and np.max(np.abs(p - orc.p)) <= 1e-7 * max(1.0, np.max(np.abs(orc.p)))
            print(f"rank0: dense sharded N={N} M={M} cost={r:.12g} vs oracle {orc.norm2x:.12g} parity={'ok' if good else 'FAIL'}", flush=True)
            ok = ok and good
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED_OK" if int(flag.item()) == 1 else "SHARDED_FAIL", flush=True)
    L.dogleg_gpu_nccl_finalize()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
