"""CPU (gloo, world_size 2) tests of the multi-process plumbing used by bench.py and the sharded
solves: the row partition and the max/sum-over-ranks reductions of the timing contract."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_columns_partition(H):
    for M, world, align in [(1000000, 8, 5000), (1000000, 3, 5000), (3600, 2, 36), (10, 4, 1), (7, 8, 1)]:
        parts = H.shard_columns(M, world, align)
        assert parts[0][0] == 0 and parts[-1][1] == M
        for (a, b), (c, d) in zip(parts, parts[1:]):
            assert b == c and a <= b
        assert all(a % align == 0 for a, _ in parts)
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= align


def test_gloo_world2_reductions(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r}); sys.path.insert(0, os.path.join({ROOT!r}, "tests"))
        import torch.distributed as dist
        import bench
        from support import harness as H
        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        dist.init_process_group("gloo")
        t = bench.barrier_max(dist, 1.0 + rank)          # timing = max over ranks
        s = bench.barrier_sum(dist, 10.0 * (rank + 1))   # work = sum over ranks
        b, e = H.shard_columns(3600, world, 36)[rank]
        cover = bench.barrier_sum(dist, float(e - b))
        assert t == 2.0 and s == 30.0 and cover == 3600.0, (t, s, cover)
        dist.barrier(); dist.destroy_process_group()
        print("GLOO_OK", rank)
    """))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                         capture_output=True, text=True, timeout=300)
    assert out.stdout.count("GLOO_OK") == 2, out.stdout[-2000:] + out.stderr[-2000:]


def test_gloo_world2_row_sharded_sums(tmp_path):
    """The arithmetic the row-sharded solves rely on, on CPU with the oracle: every rank forms
    Jt*x, |x|^2, |J v|^2 and the dense JtJ of ITS measurement columns only; the all-reduced sums
    must equal the whole problem's. gloo, world_size 2."""
    script = tmp_path / "w2.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r}); sys.path.insert(0, os.path.join({ROOT!r}, "tests"))
        import numpy as np, torch, torch.distributed as dist
        from support import harness as H
        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        dist.init_process_group("gloo")
        O = H.oracle_lib()
        prob = H.Problem.mrcal(3, 8, 6, seed=11)
        N, M = prob.N, prob.M
        Jp, Ji = prob.pattern()
        p = prob.p0()
        x, Jx = prob.evaluate(p)
        b, e = H.shard_columns(M, world, 2 * 3 * 6)[rank]
        lp = (Jp[b:e + 1] - Jp[b]).astype(np.int32); li = Ji[Jp[b]:Jp[e]].copy(); lx = Jx[Jp[b]:Jp[e]].copy(); xs = x[b:e].copy()
        g = np.zeros(N); O.orc_Jt_times_x(H.as_dp(g), N, e - b, H.as_ip(lp), H.as_ip(li), H.as_dp(lx), H.as_dp(xs))
        v = np.cos(np.arange(N))
        jv2 = O.orc_norm2_J_times_v(e - b, H.as_ip(lp), H.as_ip(li), H.as_dp(lx), H.as_dp(v))
        J = np.zeros((e - b, N))
        for j in range(e - b):
            J[j, li[lp[j]:lp[j + 1]]] = lx[lp[j]:lp[j + 1]]
        part = torch.tensor(np.concatenate([g, [xs @ xs, jv2], (J.T @ J).ravel()]))
        dist.all_reduce(part)
        gf = np.zeros(N); O.orc_Jt_times_x(H.as_dp(gf), N, M, H.as_ip(Jp), H.as_ip(Ji), H.as_dp(Jx), H.as_dp(x))
        jf = O.orc_norm2_J_times_v(M, H.as_ip(Jp), H.as_ip(Ji), H.as_dp(Jx), H.as_dp(v))
        Jf = np.zeros((M, N))
        for j in range(M):
            Jf[j, Ji[Jp[j]:Jp[j + 1]]] = Jx[Jp[j]:Jp[j + 1]]
        whole = np.concatenate([gf, [x @ x, jf], (Jf.T @ Jf).ravel()])
        assert np.allclose(part.numpy(), whole, rtol=1e-12, atol=1e-9 * np.abs(whole).max())
        dist.barrier(); dist.destroy_process_group()
        print("SUMS_OK", rank)
    """))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29518", str(script)],
                         capture_output=True, text=True, timeout=300)
    assert out.stdout.count("SUMS_OK") == 2, out.stdout[-2000:] + out.stderr[-2000:]
