"""World-size-2 test of the sample-sharded fit on CPU (gloo): each rank reduces its own row block to
partial statistics, ONE all-reduce combines them, every rank solves the same K x K system.
The per-rank statistics are built with numpy here (the kernels are covered by the -m gpu tests); what is
under test is the sharding arithmetic and the collective plumbing of pytuq_b200/parallel.py."""
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from oracle import pc_oracle as orc
    from pytuq_b200 import parallel
    from pytuq_b200.lreg import lsq, anl
    from pytuq_b200.lreg.stats import SuffStats
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)            # same data on every rank, each takes its shard
        n, d, p = 1001, 4, 3
        mi = orc.get_mi(p, d)
        x = rng.uniform(-1, 1, (n, d))
        A = orc.eval_bases(x, mi, "LU")
        c = rng.standard_normal(mi.shape[0])
        Y = np.stack([A @ c + 1e-2 * rng.standard_normal(n), A @ c[::-1]], axis=1)
        assert parallel.dist_info() == (rank, world)
        lo, hi = parallel.shard_bounds(n, rank, world)
        S_local = orc.suffstats(A[lo:hi], Y[lo:hi])
        S = parallel.allreduce_stats(S_local)
        S_full = orc.suffstats(A, Y)
        np.testing.assert_allclose(S, S_full, rtol=1e-13, atol=1e-12)
        st = SuffStats(S, mi.shape[0], 2)
        fit = lsq(); fit._fit_stats(st.select(0))
        ref = orc.lsq_fit(A, Y[:, 0])
        np.testing.assert_allclose(fit.cf, ref, rtol=1e-10, atol=1e-12)
        b = anl(); b._fit_stats(st.select(1))
        np.testing.assert_allclose(b.cf, c[::-1], rtol=1e-8, atol=1e-9)
        # every rank ends with the same coefficients (replicated solve)
        import torch
        t = torch.from_numpy(fit.cf.copy())
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        assert all(torch.equal(g, gathered[0]) for g in gathered)
        # QR statistic (8(f2)): R factors of the row shards combine to the R factor of the whole matrix
        from pytuq_b200.lreg import tsqr
        Z = torch.from_numpy(np.hstack([A, Y[:, :1]]))
        host_qr = lambda M_: torch.linalg.qr(M_, mode="r").R       # (the device kernel is covered by the -m gpu tests)
        R = tsqr.allgather_r(torch.linalg.qr(Z[lo:hi], mode="r").R, qr=host_qr)
        parts = [torch.linalg.qr(Z[i::3], mode="r").R for i in range(3)]
        Rt = tsqr.combine_r(parts, qr=host_qr)                      # odd count: balanced tree with a bye
        np.testing.assert_allclose((Rt.T @ Rt).numpy(), (Z.T @ Z).numpy(), rtol=1e-12, atol=1e-10)
        Rf = torch.linalg.qr(Z, mode="r").R
        np.testing.assert_allclose((R.T @ R).numpy(), (Rf.T @ Rf).numpy(), rtol=1e-12, atol=1e-10)
        cq, sv, rank_ = tsqr.lstsq_from_r(R, mi.shape[0])
        np.testing.assert_allclose(cq[:, 0], ref, rtol=1e-10, atol=1e-12)
        assert rank_ == mi.shape[0]
        out.put((rank, "ok"))
    except Exception as e:   # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_rows_exactly():
    sys.path.insert(0, ROOT)
    from pytuq_b200.parallel import shard_bounds
    for n in (0, 1, 7, 8, 9, 1000003):
        for w in (1, 2, 3, 8):
            edges = [shard_bounds(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_sharded_fit():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
