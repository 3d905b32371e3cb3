"""PC fitting workflow (drop-in for the reference's ``workflows/fits.py:16-81``).

The reference recomputes the same design matrix and the same Gram once per output
(fits.py:47-64).  Here one fused pass (K3) reduces (x, Y) to the statistics of all outputs
at once -- one shared G = A^T A and the K x M block B = A^T Y -- and each output's fitter
runs on its slice.  ``pc_ros`` (Rosenblatt-map workflow) is outside the PC hot path.
"""
import numpy as np

from ..rv.pcrv import PCRV
from ..utils.mindex import get_mi
from ..lreg.bcs import bcs
from ..lreg.anl import anl
from ..lreg.lreg import lsq
from ..lreg.stats import SuffStats


def _make_fitter(method, kwargs):
    if method == 'bcs':
        return bcs(eta=kwargs['eta']) if 'eta' in kwargs else bcs()
    if method == 'anl':
        return anl()
    if method == 'lsq':
        return lsq()
    raise ValueError(f"Fitting method '{method}' is not recognized.")


def pc_fit(x, y, order=3, pctype='LU', method='anl', **kwargs):
    """Fit a PC surrogate of total order `order` to (x (N,d), y (N,o)); returns
    (pcrv, [lreg per output]) exactly like the reference."""
    nsam, ndim = x.shape
    nsam_, nout = y.shape
    assert(nsam == nsam_)

    mindex = get_mi(order, ndim)
    pcrv = PCRV(nout, ndim, pctype, mi=mindex)

    # one pass over the samples for all outputs
    K = mindex.shape[0]
    batched, stats = None, None
    if method == 'lsq':
        # all outputs share the design matrix -> one batched solve on the device for the K x M block:
        # Cholesky of the shared Gram when the system is tall and well conditioned, else the QR statistic
        # (same switch as lsq.fit)
        from ..lreg.lreg import solve_normal_device
        from ..lreg import tsqr
        if nsam >= 4 * K:
            stats = SuffStats.from_pc(pcrv, x, y, 0)
            if stats.dev_system() is not None:
                G_t, B_t = stats.dev_system()
                C = solve_normal_device(G_t, B_t, 1.0e-3)
                if C is not None:
                    batched = C.cpu().numpy()
        if batched is None:
            batched = tsqr.lstsq_from_r(tsqr.shared_r(tsqr.r_factor_pc(pcrv, x, y, 0)), K)[0]
    else:
        stats = SuffStats.from_pc(pcrv, x, y, 0)

    mindices_list, cfs_list, lregs = [], [], []
    for iout in range(nout):
        print(f"Fitting output {iout+1} / {nout}")
        lreg = _make_fitter(method, kwargs)
        if batched is not None:
            lreg.cf = np.ascontiguousarray(batched[:, iout])
            lreg.cf_cov = np.zeros((mindex.shape[0], mindex.shape[0]))
            lreg.used = np.arange(mindex.shape[0])
            lreg.fitted = True
        else:
            lreg._fit_stats(stats.select(iout))
        mindices_list.append(mindex[lreg.used, :])
        cfs_list.append(lreg.cf)
        lregs.append(lreg)

    pcrv.setMiCfs(mindices_list, cfs_list)
    pcrv.setFunction()
    return pcrv, lregs
