"""Base class PCRV derives from (reference: rv/mrv.py:18-96; the distributions that live in
the reference's mrv.py are outside the PC path)."""
from itertools import product

import numpy as np


class MRV():
    def __init__(self, pdim):
        self.pdim = pdim

    def __repr__(self):
        return f"Multivariate Random Variable(dim={self.pdim})"

    def sample(self, nsam):
        raise NotImplementedError("Sampling is not implemented in the base class.")

    def pdf(self, x):
        raise NotImplementedError("PDF function is not implemented in the base class.")

    def logpdf(self, x):
        return np.log(self.pdf(x))

    def cdf(self, x):
        raise NotImplementedError("CDF function is not implemented in the base class.")

    def sample_indomain(self, nsam, domain, itry_max=1000):
        r"""``nsam`` samples of the variable restricted to the box ``domain`` (d, 2) by rejection: batches of ``nsam``
        draws from ``sample`` (for a PCRV: germ -> evalPC on the device), at most ``itry_max`` batches
        (reference: rv/mrv.py:49-74)."""
        domain = np.asarray(domain)
        kept = np.empty((0, self.pdim))
        tries = 0
        while kept.shape[0] < nsam:
            tries += 1
            batch = self.sample(nsam)
            inside = np.all(batch > domain[:, 0], axis=1) & np.all(batch < domain[:, 1], axis=1)
            kept = np.vstack((kept, batch[inside]))
            if tries > itry_max:
                print(f"WARNING: Rejection rate is too high, sampled only N={kept.shape[0]} points instead of requested {nsam}")
                return kept
        return kept[:nsam]

    def volume_indomain(self, domain):
        r"""Probability mass inside the box ``domain`` (d, 2) by inclusion-exclusion of ``cdf`` over its 2^d corners
        (reference: rv/mrv.py:76-96)."""
        domain = np.asarray(domain)
        volume = 0.0
        for pick in product(range(2), repeat=self.pdim):
            pick = np.array(pick)
            corner = np.where(pick == 1, domain[:, 1], domain[:, 0])
            volume += (-1) ** (self.pdim - int(pick.sum())) * self.cdf(corner)
        return volume
