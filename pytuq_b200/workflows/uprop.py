"""Uncertainty propagation through a model into an output PC (drop-in for the reference's
``workflows/uprop.py``: ``uprop_proj`` 6-39, ``uprop_regr`` 41-69).

Both reduce to one fused statistics pass on the output basis:
* projection:  cf = A^T (w * y) / ||Psi||^2   is the A^T Y block with Y = w * y;
* regression:  cf = argmin ||A c - y||        from G = A^T A and A^T Y of all outputs at once.
"""
import numpy as np

from ..lreg.lreg import solve_normal
from ..lreg.stats import SuffStats


def _shared_basis(out_pc):
    first = out_pc.mindices[0]
    return all(m is first or np.array_equal(m, first) for m in out_pc.mindices)


def uprop_proj(in_pc, model, nqd, out_pc):
    """Galerkin projection on the tensor Gauss grid with nqd points per dimension."""
    qdpts, wghts = in_pc.quadGerm([nqd] * in_pc.sdim)
    yqdpts = model(in_pc.evalPC(qdpts))
    assert(yqdpts.shape[1] == out_pc.pdim)
    weighted = yqdpts * wghts[:, None]
    out_cfs_all = []
    if _shared_basis(out_pc):
        proj = SuffStats.from_pc(out_pc, qdpts, weighted, 0).B()
        normsq = out_pc.evalBasesNormsSq(0)
        out_cfs_all = [proj[:, odim] / normsq for odim in range(out_pc.pdim)]
    else:
        for odim in range(out_pc.pdim):
            proj = SuffStats.from_pc(out_pc, qdpts, weighted[:, odim], odim).b(0)
            out_cfs_all.append(proj / out_pc.evalBasesNormsSq(odim))
    out_pc.setCfs(out_cfs_all)


def uprop_regr(in_pc, model, nsam, out_pc):
    """Least-squares regression on nsam random germ samples."""
    samples_germ = in_pc.sampleGerm(nsam)
    ysamples = model(in_pc.evalPC(samples_germ))
    assert(ysamples.shape[1] == out_pc.pdim)
    if _shared_basis(out_pc):
        stats = SuffStats.from_pc(out_pc, samples_germ, ysamples, 0)
        sol = None
        if stats.dev_system() is not None:
            from ..lreg.lreg import solve_normal_device
            C = solve_normal_device(*stats.dev_system())
            sol = C.cpu().numpy() if C is not None else None
        if sol is None:
            sol = solve_normal(stats.G, stats.B())
        out_cfs_all = [np.ascontiguousarray(sol[:, odim]) for odim in range(out_pc.pdim)]
    else:
        out_cfs_all = []
        for odim in range(out_pc.pdim):
            stats = SuffStats.from_pc(out_pc, samples_germ, ysamples[:, odim], odim)
            out_cfs_all.append(solve_normal(stats.G, stats.b(0)))
    out_pc.setCfs(out_cfs_all)
