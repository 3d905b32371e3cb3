"""Bayesian compressive sensing (fast relevance-vector machine) in Gram form.

Drop-in for the reference's ``lreg/bcs.py`` (class ``bcs`` 11-52, ``bcs_fit`` 58-265).  The
reference touches the N x K matrix A in every one of its <= 100 iterations (O(N K) GEMVs);
every such product is a slice of the sufficient statistics K3 / K4 deliver:

    A^T y                       = b                    (bcs.py:83)
    sum(A**2, 0)                = diag G               (bcs.py:84)
    std(y)^2                    from sum y, y^T y, N   (bcs.py:79)
    Asl^T Asl                   = G[idx, idx]          (bcs.py:95)
    A^T Asl                     = G[:, idx]            (bcs.py:100)
    A^T (Asl v)                 = G[:, idx] v          (bcs.py:166,169,213)
    Asl^T A_i                   = G[idx, i]            (bcs.py:181)
    A^T (A_i - Asl c)           = G[:, i] - G[:, idx] c   (bcs.py:183,193)
    sum (y - Asl mu)^2          = y^T y - 2 mu.b_idx + mu^T G[idx,idx] mu   (bcs.py:229)

so after one streaming pass over the samples the iteration costs O(K k) and N can be 1e7.
Decision logic, constants (nugget 1e-6, itermax 100), return shapes and warnings are the
reference's.
"""
import numpy as np

from .lreg import lreg
from .stats import SuffStats


class bcs(lreg):
    r"""BCS regression; attributes ``eta, datavar_init`` as in the reference."""

    def __init__(self, eta=1.e-8, datavar_init=None):
        super().__init__()
        self.eta = eta
        self.used = None
        self.datavar_init = datavar_init

    def _fit_stats(self, stats):
        self.cf, errbars, self.used, self.datavar, basis, self.cf_cov = \
            bcs_fit_stats(stats, sigma2=self.datavar_init, eta=self.eta)
        self.fitted = True


def bcs_fit(A, y, sigma2=None, eta=1.e-8, adaptive=0, optimal=1, scale=0.1, nugget=1.e-6):
    r"""Reference signature (lreg/bcs.py:58): reduces (A, y) on the GPU (K4) and runs the
    Gram-form iteration.  Returns (weights, errbars, used, sigma2, basis, Sig)."""
    return bcs_fit_stats(SuffStats.from_matrix(A, y), sigma2=sigma2, eta=eta, adaptive=adaptive,
                         optimal=optimal, scale=scale, nugget=nugget)


def bcs_fit_stats(stats, sigma2=None, eta=1.e-8, adaptive=0, optimal=1, scale=0.1, nugget=1.e-6):
    r"""Fast-RVM iteration on sufficient statistics (one output)."""
    G = stats.G
    Aty = np.array(stats.b(0))
    N, M = stats.n, stats.K
    if sigma2 is None:
        sigma2 = max(stats.yvar(0) / 1.e+2, nugget)

    A2 = np.array(np.diag(G))
    ratio = (Aty ** 2 + nugget) / (A2 + nugget)
    index = np.array([np.argmax(ratio)])
    alpha = A2[index] / (ratio[index] - sigma2)
    Sig = 1. / (alpha + G[np.ix_(index, index)] / sigma2 + nugget)
    mu = np.array([Sig[0, 0] * Aty[index[0]] / sigma2])
    left = G[:, index[0]] / sigma2
    S = A2 / sigma2 - Sig[0, 0] * left ** 2
    Q = Aty / sigma2 - Sig[0, 0] * Aty[index[0]] / sigma2 * left

    itermax = 100
    mlhist = np.empty(itermax)
    for count in range(itermax):
        ss = S.copy()
        qq = Q.copy()
        ss[index] = alpha * S[index] / (alpha - S[index])
        qq[index] = alpha * Q[index] / (alpha - S[index])
        theta = qq ** 2 - ss

        ml = np.full(M, -np.inf)
        # membership masks instead of sorted-set operations (same candidate sets in the same ascending order, a
        # third of the iteration's time at K ~ 1300): pos_of[k] = position of term k in `index`, -1 outside the model
        positive = theta > 0
        pos_of = np.full(M, -1)
        pos_of[index] = np.arange(len(index))
        inside = pos_of >= 0
        # in the model and theta > 0: candidates for re-estimation
        ire = np.flatnonzero(positive & inside)
        which = pos_of[ire]
        if len(ire) > 0:
            anew = ss[ire] ** 2 / theta[ire] + nugget
            delta = (alpha[which] - anew) / (anew * alpha[which])
            if (1 + S[ire] * delta <= 0.0).any():
                ml[ire] = -1.e+80
            else:
                ml[ire] = Q[ire] ** 2 * delta / (S[ire] * delta + 1) - np.log(1 + S[ire] * delta)
        # outside the model and theta > 0: candidates for addition
        iad = np.flatnonzero(positive & ~inside)
        if len(iad) > 0:
            if (S[iad] <= 0.0).any():
                ml[iad] = -1.e+80
            else:
                ml[iad] = (Q[iad] ** 2 - S[iad]) / S[iad] + np.log(S[iad] / (Q[iad] ** 2))
        # in the model and theta <= 0: candidates for deletion
        ide = np.flatnonzero(~positive & inside)
        which = pos_of[ide]
        if len(ide) > 0:
            ml[ide] = Q[ide] ** 2 / (S[ide] - alpha[which]) - np.log(1. - S[ide] / alpha[which])

        idx = np.argmax(ml)
        mlhist[count] = ml[idx]
        if count > 1 and abs(mlhist[count] - mlhist[count - 1]) < abs(mlhist[count] - mlhist[0]) * eta:
            break

        which = np.where(index == idx)[0]
        if theta[idx] > 0:
            anew = ss[idx] ** 2 / theta[idx] + nugget
            if len(which) > 0:                      # re-estimate
                w = which[0]
                Sigii, mui, Sigi = Sig[w, w], mu[w], Sig[:, w].copy()
                delta = anew - alpha[w]
                ki = delta / (1. + Sigii * delta)
                mu = mu - ki * mui * Sigi
                Sig = Sig - ki * np.outer(Sigi, Sigi)
                comm = (G[:, index] @ Sigi) / sigma2
                S = S + ki * comm ** 2
                Q = Q + ki * mui * comm
                alpha[w] = anew
            else:                                   # add
                Sigii = 1. / (anew + S[idx])
                mui = Sigii * Q[idx]
                comm1 = (Sig @ G[index, idx]) / sigma2
                off = -Sigii * comm1
                Sig = np.block([[Sig + Sigii * np.outer(comm1, comm1), off.reshape(-1, 1)],
                                [off.reshape(1, -1), Sigii]])
                mu = np.append(mu - mui * comm1, mui)
                comm2 = (G[:, idx] - G[:, index] @ comm1) / sigma2
                S = S - Sigii * comm2 ** 2
                Q = Q - mui * comm2
                index = np.append(index, idx)
                alpha = np.append(alpha, anew)
        else:
            if len(which) > 0 and len(index) > 1:   # delete
                w = which[0]
                Sigii, mui = Sig[w, w], mu[w]
                Sigi_old = Sig[:, w].copy()
                Sig = Sig - np.outer(Sigi_old, Sigi_old) / Sigii
                # Reference quirk kept for parity (lreg/bcs.py:206-216): there `Sigi` is a VIEW of
                # column w and `Sig -= ...` updates it in place, so the mean / S / Q updates below
                # see the already-updated (numerically ~0) column, not the old one.
                Sigi = Sig[:, w].copy()
                Sig = np.delete(np.delete(Sig, w, 0), w, 1)
                mu = np.delete(mu - (mui / Sigii) * Sigi, w)
                comm = (G[:, index] @ Sigi) / sigma2
                S = S + (comm ** 2 / Sigii)
                Q = Q + (mui / Sigii) * comm
                index = np.delete(index, w)
                alpha = np.delete(alpha, w)
            if len(which) > 0 and len(index) == 1:
                break

    weights = mu
    used = index
    # re-estimated noise variance (bcs.py:229-231); keeps the reference's (1,) shape
    rss = stats.rss(mu, used, 0)
    sigma2 = rss / (N - len(index) + np.dot(alpha.reshape(1, -1), np.diag(Sig)))

    detsig = np.linalg.det(Sig)
    if detsig < -nugget:
        print(f"Warning: Sigma matrix has a determinant {detsig} that is not positive. Use with care.")
    if (np.diag(Sig) < 0.0).any():
        print("Warning: Sigma matrix has a negative diagonal element. Setting them to zero, but this may lead to inaccuracies.")
        Sig = Sig.copy()
        neg = np.diag(Sig) < 0.0
        Sig[neg, neg] = 0.0
    errbars = np.sqrt(np.diag(Sig))

    basis = None
    if adaptive:
        if optimal:
            D, V = np.linalg.eig(Sig)
            basis = V[:, np.argmax(D)]
        else:
            temp = G[np.ix_(used, used)] / sigma2
            Sig_inv = temp + scale * np.mean(np.diag(temp)) * np.eye(len(used))
            D, V = np.linalg.eig(Sig_inv)
            basis = V[:, np.argmin(D)]
    return weights, errbars, used, sigma2, basis, Sig
