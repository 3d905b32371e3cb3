"""CPU oracle for the PyTUQ polynomial-chaos hot path -- TEST INFRASTRUCTURE ONLY.

This module restates, in plain numpy, the arithmetic of the reference functions on
the path (file:line relative to the PyTUQ tree, ``src/pytuq/...``).  It exists to
check the CUDA path and to serve as the CPU baseline; the product package
``pytuq_b200`` never imports it.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may use it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real reference
from ``/root/reference/src`` (matplotlib stubbed) and stores its outputs on fixed
inputs in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every
function below against those vectors (bit-exact for basis values / evalPC /
multi-indices, 1e-12 for the dense solves) and against the known-answer vectors of
SURVEY.md Appendix A.

The loops deliberately keep the reference's structure (a Python loop over terms and
dimensions with length-N numpy multiplies), so that timing this module on the host
is an honest stand-in for timing the reference itself.
"""
import numpy as np
from scipy.linalg import lstsq as _scipy_lstsq

FAMILIES = ("LU", "nLU", "HG", "nHG")


# --------------------------------------------------------------------------- 1-D families
def rec_a(pctype, n):
    """a_n of the three-term recurrence (rv/pcrv.py:755,762,769,776)."""
    if pctype == "LU":
        return (2. * n + 1.) / (n + 1.)
    if pctype == "nLU":
        return np.sqrt((2. * n + 3.) / (2. * n + 1.)) * (2. * n + 1.) / (n + 1.)
    if pctype == "HG":
        return 1.
    if pctype == "nHG":
        return 1. / (n + 1.)
    raise ValueError(pctype)


def rec_b(pctype, n):
    """b_n of the three-term recurrence (rv/pcrv.py:756,763,770,777)."""
    if pctype == "LU":
        return -n / (n + 1.)
    if pctype == "nLU":
        return - np.sqrt((2. * n + 3.) / (2. * n - 1.)) * n / (n + 1.)
    if pctype == "HG":
        return -n
    if pctype == "nHG":
        return -1. / (n + 1.)
    raise ValueError(pctype)


def p1_of(pctype, x):
    """First-order polynomial (rv/pcrv.py:754,761,768,775)."""
    if pctype == "nLU":
        return x * np.sqrt(3.)
    if pctype in FAMILIES:
        return x
    raise ValueError(pctype)


def pc1d(pctype, x, order):
    """phi_0..phi_order at the points x (PC1d.__call__, rv/pcrv.py:784-804).

    Keeps the reference's evaluation order  a*x*p_n + b*p_{n-1}  =  ((a*x)*p_n)+(b*p_{n-1}).
    """
    vals = [np.ones_like(x)]
    if order >= 1:
        vals.append(p1_of(pctype, x))
    for n in range(1, order):
        vals.append(rec_a(pctype, n) * x * vals[-1] + rec_b(pctype, n) * vals[-2])
    return vals


def normsq_1d(pctype, order):
    """Squared norm of the 1-D basis function (PC1d.normsq, rv/pcrv.py:847-871)."""
    if pctype == "LU":
        return 1. / (2. * float(order) + 1.)
    if pctype == "HG":
        return 1.0 if order == 0 else float(np.prod(np.arange(1, order + 1)))
    if pctype in ("nLU", "nHG"):
        return 1.0
    raise ValueError(pctype)


def _types(pctype, sdim):
    return [pctype] * sdim if isinstance(pctype, str) else list(pctype)


# --------------------------------------------------------------------------- multi-indices
def get_npc(order, dim):
    """Number of total-order terms (utils/mindex.py:7-26)."""
    num = 1
    for i in range(order):
        num = num * (dim + i + 1)
    for i in range(order):
        num = num / (i + 1)
    return int(num)


def get_mi(order, dim):
    """Graded total-order multi-index table in the reference's row order (utils/mindex.py:29-62).

    Order q rows are generated from the order q-1 block: for each dimension j, every row
    among the last cnt[j] rows of the previous block gets +1 in dimension j, where cnt[j]
    is the number of order q-1 rows whose leading non-zero dimension is >= j.
    """
    npc = get_npc(order, dim)
    mi = np.zeros((npc, dim), dtype=int)
    if order == 0:
        return mi
    for j in range(dim):
        mi[1 + j, j] = 1
    filled = dim            # index of the last filled row
    cnt = np.ones(dim, dtype=int)
    for _ in range(2, order + 1):
        prev_last = filled
        # suffix sums, updated in place from the left exactly like the reference
        for j in range(dim):
            cnt[j] = cnt[j:].sum()
        for j in range(dim):
            for src in range(prev_last - cnt[j] + 1, prev_last + 1):
                filled += 1
                mi[filled] = mi[src]
                mi[filled, j] += 1
    return mi


# --------------------------------------------------------------------------- basis / evaluation
def eval_bases(x, mi, pctype, max_ord=None):
    """Design matrix A[n,k] = prod_i phi_{mi[k,i]}(x[n,i]) (PCRV.evalBases, rv/pcrv.py:413-442)."""
    nx, s = x.shape
    types = _types(pctype, s)
    if max_ord is None:
        max_ord = mi.max(axis=0)
    tabs = [pc1d(types[i], x[:, i], int(max_ord[i])) for i in range(s)]
    K = mi.shape[0]
    A = np.empty((nx, K))
    for k in range(K):
        prod = 1.0
        for i in range(s):
            prod = prod * tabs[i][mi[k, i]]
        A[:, k] = prod
    return A


def eval_pc(x, mis, cfs, pctype, max_ord=None):
    """y[n,j] = sum_k cf_j[k] prod_i phi(...) (PCRV.evalPC, rv/pcrv.py:469-498).

    Coefficient first, then the factors in dimension order, terms summed in k order.
    """
    nx, s = x.shape
    types = _types(pctype, s)
    if max_ord is None:
        max_ord = np.max([m.max(axis=0) for m in mis], axis=0)
    tabs = [pc1d(types[i], x[:, i], int(max_ord[i])) for i in range(s)]
    y = np.zeros((nx, len(mis)))
    for j, (mi, cf) in enumerate(zip(mis, cfs)):
        tot = np.zeros(nx)
        for k in range(mi.shape[0]):
            term = cf[k]
            for i in range(s):
                term = term * tabs[i][mi[k, i]]
            tot += term
        y[:, j] = tot
    return y


def norms_sq(mi, pctype):
    """||Psi_k||^2 (PCRV.evalBasesNormsSq, rv/pcrv.py:444-466)."""
    K, s = mi.shape
    types = _types(pctype, s)
    out = np.empty(K)
    for k in range(K):
        prod = 1.0
        for i in range(s):
            prod *= normsq_1d(types[i], mi[k, i])
        out[k] = prod
    return out


# --------------------------------------------------------------------------- moments / Sobol
def pc_mean(mi, cf):
    """rv/pcrv.py:202-215."""
    m = 0.0
    for k in np.where(mi.sum(axis=1) == 0)[0]:
        m += cf[k]
    return m


def pc_var(mi, cf, pctype):
    """rv/pcrv.py:217-231."""
    nsq = norms_sq(mi, pctype)
    v = 0.0
    for k in np.where(mi.sum(axis=1) > 0)[0]:
        v += cf[k] ** 2 * nsq[k]
    return v


def _safe_var(v):
    return v + 1.e-12 if v == 0.0 else v


def sobol_main(mi, cf, pctype):
    """rv/pcrv.py:233-256."""
    K, s = mi.shape
    var = _safe_var(pc_var(mi, cf, pctype))
    nsq = norms_sq(mi, pctype)
    out = np.zeros(s)
    for j in range(s):
        others = np.delete(mi, j, axis=1).sum(axis=1)
        for k in np.where(others == 0)[0]:
            if mi[k, j] > 0:
                out[j] += cf[k] ** 2 * nsq[k]
        out[j] /= var
    return out


def sobol_total(mi, cf, pctype):
    """rv/pcrv.py:258-280."""
    K, s = mi.shape
    var = _safe_var(pc_var(mi, cf, pctype))
    nsq = norms_sq(mi, pctype)
    out = np.zeros(s)
    for j in range(s):
        for k in range(K):
            if mi[k, j] > 0:
                out[j] += cf[k] ** 2 * nsq[k]
        out[j] /= var
    return out


def sobol_joint(mi, cf, pctype):
    """rv/pcrv.py:282-309."""
    K, s = mi.shape
    var = _safe_var(pc_var(mi, cf, pctype))
    nsq = norms_sq(mi, pctype)
    out = np.zeros((s, s))
    for j in range(s):
        for l in range(j + 1, s):
            for k in range(K):
                if mi[k, j] * mi[k, l] > 0:
                    out[j, l] += cf[k] ** 2 * nsq[k]
            out[j, l] /= var
            out[l, j] = out[j, l]
    return out


def sobol_group(mi, cf, pctype, group):
    """rv/pcrv.py:311-338."""
    var = _safe_var(pc_var(mi, cf, pctype))
    nsq = norms_sq(mi, pctype)
    outside = np.delete(mi, group, axis=1).sum(axis=1)
    g = 0.0
    for k in np.where(outside == 0)[0]:
        if mi[k, :].sum() > 0:
            g += cf[k] ** 2 * nsq[k]
    return g / var


# --------------------------------------------------------------------------- fitters (A-based)
def lsq_fit(A, y):
    """lsq.fita (lreg/lreg.py:328-339): scipy gelsd with cond 1e-13."""
    cf = _scipy_lstsq(A, y, 1.0e-13)[0]
    return cf


def anl_fit(A, y, datavar=None, prior_var=None, cov_nugget=0.0, method="full"):
    """anl.fita + _compute_sigmahatsq (lreg/anl.py:46-105).  Returns (cf, cf_cov, datavar)."""
    N, K = A.shape
    gram = np.matmul(A.T, A)
    if prior_var is not None:
        gram += (datavar / prior_var) * np.eye(K)
    gram += cov_nugget * np.eye(K)
    ginv = np.linalg.inv(gram)
    cf = np.dot(ginv, np.dot(A.T, y))
    if datavar is None:
        sig = (np.dot(y, y - np.dot(A, cf)) / 2.) / ((N - K) / 2. - 1.)
        if sig < 0.0:
            sig = 0.0
    else:
        sig = datavar
    if method == "full":
        cov = sig * ginv
    elif method == "vi":
        cov = np.diag(sig / np.diag(np.dot(A.T, A) + cov_nugget * np.diag(np.ones((K,)))))
    else:
        raise ValueError(method)
    return cf, cov, sig + 0.0


def bcs_fit(A, y, sigma2=None, eta=1.e-8, nugget=1.e-6):
    """Fast-RVM Bayesian compressive sensing on a materialised A (bcs_fit, lreg/bcs.py:58-265;
    the non-adaptive branch).  Returns (weights, errbars, used, sigma2 (1,), Sig)."""
    if sigma2 is None:
        sigma2 = max(np.std(y) ** 2 / 1.e+2, nugget)
    N, M = A.shape
    Aty = np.dot(A.T, y)
    A2 = np.sum(A ** 2, axis=0)
    ratio = (Aty ** 2 + nugget * np.ones_like(Aty)) / (A2 + nugget * np.ones_like(A2))

    index = [np.argmax(ratio)]
    alpha = A2[index] / (ratio[index] - sigma2)
    Asl = A[:, index]
    Sig = 1. / (alpha + np.dot(Asl.T, Asl) / sigma2 + nugget * np.ones((1, 1)))
    mu = np.zeros((1,))
    mu[0] = Sig[0, 0] * Aty[index[0]] / sigma2
    left = np.dot(A.T, Asl) / sigma2
    S = A2 / sigma2 - Sig[0, 0] * left[:, 0] ** 2
    Q = Aty / sigma2 - Sig[0, 0] * Aty[index[0]] / sigma2 * left[:, 0]

    itermax = 100
    mlhist = np.empty(itermax)
    for count in range(itermax):
        ss = S.copy()
        qq = Q.copy()
        ss[index] = alpha * S[index] / (alpha - S[index])
        qq[index] = alpha * Q[index] / (alpha - S[index])
        theta = qq ** 2 - ss

        ml = -np.inf * np.ones(M)
        pos = np.where(theta > 0)[0]
        # candidates already in the model: re-estimate
        ire, _, which = np.intersect1d(pos, index, return_indices=True)
        if len(ire) > 0:
            anew = ss[ire] ** 2 / theta[ire] + nugget
            delta = (alpha[which] - anew) / (anew * alpha[which])
            if (1 + S[ire] * delta <= 0.0).any():
                ml[ire] = -1.e+80
            else:
                ml[ire] = Q[ire] ** 2 * delta / (S[ire] * delta + 1) - np.log(1 + S[ire] * delta)
        # candidates not in the model: add
        iad = np.setdiff1d(pos, ire)
        if len(iad) > 0:
            if (S[iad] <= 0.0).any():
                ml[iad] = -1.e+80
            else:
                ml[iad] = (Q[iad] ** 2 - S[iad]) / S[iad] + np.log(S[iad] / (Q[iad] ** 2))
        # in the model with theta <= 0: delete
        nonpos = np.setdiff1d(np.arange(M), pos)
        ide, _, which = np.intersect1d(nonpos, index, return_indices=True)
        if len(ide) > 0:
            ml[ide] = Q[ide] ** 2 / (S[ide] - alpha[which]) - np.log(1. - S[ide] / alpha[which])

        idx = np.argmax(ml)
        mlhist[count] = ml[idx]
        if count > 1 and abs(mlhist[count] - mlhist[count - 1]) < abs(mlhist[count] - mlhist[0]) * eta:
            break

        which = np.where(index == idx)[0]
        if theta[idx] > 0:
            anew = ss[idx] ** 2 / theta[idx] + nugget
            if len(which) > 0:
                w = which[0]
                Sigii, mui, Sigi = Sig[w, w], mu[w], Sig[:, w]
                delta = anew - alpha[w]
                ki = delta / (1. + Sigii * delta)
                mu = mu - ki * mui * Sigi
                Sig = Sig - ki * np.dot(Sigi.reshape(-1, 1), Sigi.reshape(1, -1))
                comm = np.dot(A.T, np.dot(Asl, Sigi) / sigma2)
                S = S + ki * comm ** 2
                Q = Q + ki * mui * comm
                alpha[which] = anew
            else:
                Ai = A[:, idx]
                Sigii = 1. / (anew + S[idx])
                mui = Sigii * Q[idx]
                comm1 = np.dot(Sig, np.dot(Asl.T, Ai)) / sigma2
                ei = Ai - np.dot(Asl, comm1)
                off = -Sigii * comm1
                Sig = np.block([[Sig + Sigii * np.dot(comm1.reshape(-1, 1), comm1.reshape(1, -1)),
                                 off.reshape(-1, 1)],
                                [off.reshape(1, -1), Sigii]])
                mu = np.append(mu - mui * comm1, mui)
                comm2 = np.dot(A.T, ei) / sigma2
                S = S - Sigii * comm2 ** 2
                Q = Q - mui * comm2
                index = np.append(index, idx)
                alpha = np.append(alpha, anew)
                Asl = np.hstack((Asl, Ai.reshape(-1, 1)))
        else:
            if len(which) > 0 and len(index) > 1:
                w = which[0]
                Sigii, mui, Sigi = Sig[w, w], mu[w], Sig[:, w]
                Sig -= np.dot(Sigi.reshape(-1, 1), Sigi.reshape(1, -1)) / Sigii
                Sig = np.delete(np.delete(Sig, w, 0), w, 1)
                mu = np.delete(mu - (mui / Sigii) * Sigi, w)
                comm = np.dot(A.T, np.dot(Asl, Sigi)) / sigma2
                S = S + (comm ** 2 / Sigii)
                Q = Q + (mui / Sigii) * comm
                index = np.delete(index, w)
                alpha = np.delete(alpha, w)
                Asl = np.delete(Asl, w, 1)
            if len(which) > 0 and len(index) == 1:
                break

    used = index
    sigma2 = np.sum((y - np.dot(Asl, mu)) ** 2) / (N - len(index) + np.dot(alpha.reshape(1, -1), np.diag(Sig)))
    for i in range(Sig.shape[0]):
        if Sig[i, i] < 0.0:
            Sig[i, i] = 0.0
    errbars = np.sqrt(np.diag(Sig))
    return mu, errbars, np.asarray(used), sigma2, Sig


# --------------------------------------------------------------------------- sufficient statistics
def suffstats(A, Y):
    """S = Z^T Z with Z = [A | Y | 1]: the statistics K3/K4 produce (include/pcq.h)."""
    N = A.shape[0]
    Y = Y.reshape(N, -1)
    Z = np.hstack([A, Y, np.ones((N, 1))])
    return Z.T @ Z


# --------------------------------------------------------------------------- Philox germ (K5)
_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32_10(c, k):
    """Philox4x32-10 (Salmon et al. 2011) on integer arrays c[4], k[2] (uint64 holding 32-bit values)."""
    c0, c1, c2, c3 = [np.asarray(v, dtype=np.uint64) for v in c]
    k0, k1 = [np.asarray(v, dtype=np.uint64) for v in k]
    for _ in range(10):
        p0 = np.uint64(_M0) * c0
        p1 = np.uint64(_M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & np.uint64(_MASK)
        hi1, lo1 = p1 >> np.uint64(32), p1 & np.uint64(_MASK)
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0), lo1, (hi0 ^ c3 ^ k1), lo0
        k0 = (k0 + np.uint64(_W0)) & np.uint64(_MASK)
        k1 = (k1 + np.uint64(_W1)) & np.uint64(_MASK)
    return c0, c1, c2, c3


def germ_philox(seed, first_index, n, kinds):
    """The device germ stream of pcq_sample_germ_dev / pcq_propagate_dev (include/pcq.h)."""
    s = len(kinds)
    idx = np.arange(first_index, first_index + n, dtype=np.uint64)
    x = np.empty((n, s))
    for pair in range((s + 1) // 2):
        r0, r1, r2, r3 = philox4x32_10(
            (idx & np.uint64(_MASK), idx >> np.uint64(32), np.full(n, pair, np.uint64),
             np.full(n, 0x50435121, np.uint64)),
            (np.full(n, seed & _MASK, np.uint64), np.full(n, (seed >> 32) & _MASK, np.uint64)))
        a = (r1 << np.uint64(32)) | r0
        b = (r3 << np.uint64(32)) | r2
        u0 = (a >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
        u1 = (b >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
        rad = np.sqrt(-2.0 * np.log(1.0 - u0))
        i = 2 * pair
        x[:, i] = rad * np.cos(2.0 * np.pi * u1) if kinds[i] else 2.0 * u0 - 1.0
        if i + 1 < s:
            x[:, i + 1] = rad * np.sin(2.0 * np.pi * u1) if kinds[i + 1] else 2.0 * u1 - 1.0
    return x
