"""Polynomial-chaos random variables on the B200 path.

Drop-in for the reference's ``rv/pcrv.py``: same classes (``PCRV``, ``PCRV_iid``,
``PCRV_mvn``, ``PC1d``), same signatures, numpy in / numpy out.  The arithmetic of

    PC1d.__call__      rv/pcrv.py:784-804   (three-term recurrence)
    PCRV.evalBases     rv/pcrv.py:413-442   (design matrix)                -> K1
    PCRV.evalPC        rv/pcrv.py:469-498   (surrogate evaluation, no A)   -> K2
    PCRV.sample        rv/pcrv.py:511-525

runs in libpcq's CUDA kernels through ctypes (include/pcq.h).  New methods that have no
reference counterpart expose the fused paths: ``suffStats`` (K3), ``propagate`` (K5).
Instances hold numpy data only (picklable, deep-copyable); native plans are cached by
content in ``pytuq_b200.device``.
"""
import functools
import itertools
import sys

import numpy as np
from scipy.special import erf
from scipy.stats import multivariate_normal

from .mrv import MRV
from ..func.func import Function
from .. import _native, device as _device

_SQRT3 = np.sqrt(3.)


def _is_device_tensor(a):
    return type(a).__module__.split(".")[0] == "torch" and bool(getattr(a, "is_cuda", False))


class PC1d():
    r"""One-dimensional orthogonal-polynomial family: 'LU' (Legendre), 'nLU' (normalised
    Legendre), 'HG' (probabilists' Hermite), 'nHG' (the reference's scaled Hermite He_n/n!).

    Attributes mirror the reference (rv/pcrv.py:732-782): ``a``, ``b`` (recurrence
    coefficient callables), ``p0``, ``p1``, ``sample``, ``domain``, ``pctype``.
    """
    # family -> (a_n, b_n, p1 scale, germ kind: 0 uniform(-1,1) / 1 normal, domain)
    _FAMILIES = {
        'LU': (lambda n: (2. * n + 1.) / (n + 1.),
               lambda n: -n / (n + 1.), 1.0, 0, (-1., 1.)),
        'nLU': (lambda n: np.sqrt((2. * n + 3.) / (2. * n + 1.)) * (2. * n + 1.) / (n + 1.),
                lambda n: - np.sqrt((2. * n + 3.) / (2. * n - 1.)) * n / (n + 1.), _SQRT3, 0, (-1., 1.)),
        'HG': (lambda n: 1., lambda n: -n, 1.0, 1, (-np.inf, np.inf)),
        'nHG': (lambda n: 1. / (n + 1.), lambda n: -1. / (n + 1.), 1.0, 1, (-np.inf, np.inf)),
    }

    def __init__(self, pctype='LU'):
        self.pctype = pctype
        if pctype not in self._FAMILIES:
            print(f'PC1d type {self.pctype} is not recognized. Exiting.')
            sys.exit()
        self.domain = np.array(self._FAMILIES[pctype][4])

    # the callables are properties so that instances stay picklable
    @property
    def a(self):
        return self._FAMILIES[self.pctype][0]

    @property
    def b(self):
        return self._FAMILIES[self.pctype][1]

    @property
    def p1_scale(self):
        return self._FAMILIES[self.pctype][2]

    @property
    def germ_kind(self):
        return self._FAMILIES[self.pctype][3]

    def p0(self, x):
        return np.ones_like(x)

    def p1(self, x):
        return x * _SQRT3 if self.pctype == 'nLU' else x

    def sample(self, m):
        return self.germSample(m)

    def rec_tables(self, pmax):
        """a_n, b_n for n = 0..pmax as float64, from the reference's own expressions
        (entry 0 is unused by the recurrence; b_0 of nLU is not finite and is stored as 0)."""
        ra = np.zeros(pmax + 1)
        rb = np.zeros(pmax + 1)
        for n in range(1, pmax + 1):
            ra[n] = float(self.a(n))
            rb[n] = float(self.b(n))
        return ra, rb

    def __call__(self, x, order):
        r"""phi_0..phi_order at the points x: list of order+1 arrays shaped like x.
        Evaluated on the device by K1 with the 1-D multi-index [[0],[1],...,[order]]."""
        x = np.asarray(x)
        pts = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 1)
        mi = np.arange(order + 1, dtype=np.int32).reshape(-1, 1)
        ra, rb = self.rec_tables(order)
        plan = _device.get_plan(mi, ra.reshape(-1, 1), rb.reshape(-1, 1), np.array([self.p1_scale]), order)
        out = np.empty((pts.shape[0], order + 1))
        _native.check(_native.lib().pcq_eval_bases(plan.handle, _native.ptr(pts), pts.shape[0], 1,
                                                   _native.ptr(out), order + 1), "pcq_eval_bases")
        return [np.ascontiguousarray(out[:, n]).reshape(x.shape) for n in range(order + 1)]

    def germCdf(self, x):
        """CDF of the germ (rv/pcrv.py:807-826)."""
        if self.germ_kind == 0:
            cdf = (x + 1.) / 2.
            cdf[cdf > 1] = 1.0
            cdf[cdf <= -1] = 0.0
            return cdf
        return (1.0 + erf(x / np.sqrt(2.0))) / 2.0

    def germSample(self, nsam):
        """nsam germ draws from numpy's global legacy generator, as the reference does
        (rv/pcrv.py:757,771,828-845) so that seeded streams agree."""
        if self.germ_kind == 0:
            return np.random.rand(nsam) * 2.0 - 1.0
        return np.random.randn(nsam)

    def normsq(self, ord):
        """Squared norm of the order-`ord` basis function (rv/pcrv.py:847-871)."""
        if self.pctype == 'LU':
            return 1. / (2. * float(ord) + 1.)
        if self.pctype == 'HG':
            return 1.0 if ord == 0 else float(np.prod(np.arange(1, ord + 1)))
        return 1.0

    def quad(self, k):
        """k-point Gauss rule of the family by Golub-Welsch (rv/pcrv.py:873-894)."""
        jac = np.zeros((k, k))
        for i in range(k - 1):
            jac[i + 1, i] = np.sqrt(-self.b(i + 1) / (self.a(i) * self.a(i + 1)))
        nodes, vecs = np.linalg.eigh(jac)
        return nodes, vecs[0, :] ** 2


class PCRV(MRV):
    r"""Multivariate PC random variable; attributes as in the reference (rv/pcrv.py:16-30):
    ``sdim, pdim, mindices, coefs, maxOrd, pind, rndind, detind, pctypes, PC1ds, function``."""

    def __init__(self, pdim, sdim, pctype, mi=None, cfs=None):
        assert(pdim > 0)
        super().__init__(pdim)
        self.function = None
        assert(sdim > 0)
        if mi is None:
            mi = np.zeros((1, sdim), dtype=int)
        self.setMiCfs(mi, cfs=cfs)
        if isinstance(pctype, str):
            self.pctypes = [pctype] * self.sdim
        elif isinstance(pctype, list):
            self.pctypes = pctype
        else:
            print(f"PC type {pctype} type is not recognized. Exiting.")
            sys.exit()
        assert(len(self.pctypes) == self.sdim)
        self.getParamIndices()
        self.getMaxOrders()
        self.PC1ds = [PC1d(pt) for pt in self.pctypes]

    def __repr__(self):
        return f"{self.pctypes} PC Random Variable(pdim={self.pdim}, sdim={self.sdim})"

    # ------------------------------------------------------------------ bookkeeping (host)
    def setMiCfs(self, mi, cfs=None):
        self.setMi(mi)
        self.setCfs(cfs=cfs)
        self.checkMiCfsizes()

    def setMi(self, mi):
        if isinstance(mi, np.ndarray):
            assert(mi.ndim == 2)
            self.mindices = [mi] * self.pdim
        elif isinstance(mi, list):
            assert(len(mi) == self.pdim)
            self.mindices = mi.copy()
        else:
            print(f"Multiindex {mi} type is not recognized. Exiting.")
            sys.exit()
        self.sdim = self.mindices[0].shape[1]
        colsums = [np.sum(m) for m in self.mindices]
        self.detind = [i for i, t in enumerate(colsums) if t == 0]
        self.rndind = [i for i, t in enumerate(colsums) if t != 0]

    def setCfs(self, cfs=None):
        if cfs is None:
            self.coefs = [np.zeros((m.shape[0],)) for m in self.mindices]
        elif isinstance(cfs, np.ndarray):
            if cfs.ndim == 1:
                self.coefs = [cfs] * self.pdim
            elif cfs.ndim == 2:
                assert(cfs.shape[0] == self.pdim)
                self.coefs = [cfs[j, :] for j in range(self.pdim)]
            else:
                print(f"Wrong shape {cfs.shape} of coefs array. Exiting.")
                sys.exit()
        elif isinstance(cfs, list):
            assert(len(cfs) == self.pdim)
            self.coefs = cfs.copy()
        else:
            print(f"Coefs {cfs} type is not recognized. Exiting.")
            sys.exit()

    def checkMiCfsizes(self):
        assert(len(self.mindices) == len(self.coefs))
        for m, c in zip(self.mindices, self.coefs):
            assert(isinstance(m, np.ndarray) and isinstance(c, np.ndarray))
            assert(m.ndim == 2 and c.ndim == 1)
            assert(m.shape[1] == self.sdim)
            assert(m.shape[0] == c.shape[0])

    def setRandomCfs(self):
        self.setCfs(cfs=[np.random.rand(m.shape[0],) for m in self.mindices])

    def getMaxOrders(self):
        self.maxOrd = np.max(np.array([np.max(m, axis=0) for m in self.mindices], dtype=int), axis=0)

    def getParamIndices(self):
        self.pind = [(i, k) for i in range(self.pdim) for k in range(self.coefs[i].shape[0])]

    def printInfo(self):
        for m, c in zip(self.mindices, self.coefs):
            print(f"{m} {c}")

    def cfsFlatten(self):
        return np.concatenate(self.coefs)

    def cfsUnflatten(self, cfs_flat):
        edges = np.cumsum([0] + [m.shape[0] for m in self.mindices])
        self.coefs = [cfs_flat[edges[i]:edges[i + 1]] for i in range(self.pdim)]

    # ------------------------------------------------------------------ native plumbing
    def _germ_kinds(self):
        return np.array([p.germ_kind for p in self.PC1ds], dtype=np.int32)

    def _check_orders(self, mindex):
        # the reference tabulates each dimension up to maxOrd (fixed at construction,
        # rv/pcrv.py:65) and indexes the tables with the live multi-index: a grown
        # multi-index raises IndexError there (rv/pcrv.py:433-438); keep that contract
        if (np.max(mindex, axis=0) > self.maxOrd).any():
            raise IndexError("list index out of range")
        if np.min(mindex) < 0:
            raise IndexError("negative order in multi-index")

    def _moment_tables(self, mindex):
        """(nord, a_n, b_n up to nord = twice the highest total order) for the moment form of the sufficient statistics
        (include/pcq.h: pcq_plan_enable_moments), or None where that form cannot apply."""
        nord = 2 * int(np.max(np.sum(mindex, axis=1)))
        if nord < 2 or nord > 32 or not (2 <= self.sdim <= 64):
            return None
        ra = np.zeros((nord + 1, self.sdim))
        rb = np.zeros((nord + 1, self.sdim))
        for i, p in enumerate(self.PC1ds):
            ra[:, i], rb[:, i] = p.rec_tables(nord)
        return nord, ra, rb

    def _plan(self, mindex, dev=None):
        pmax = int(np.max(mindex))
        ra = np.zeros((pmax + 1, self.sdim))
        rb = np.zeros((pmax + 1, self.sdim))
        p1 = np.empty(self.sdim)
        for i, p in enumerate(self.PC1ds):
            ra[:, i], rb[:, i] = p.rec_tables(pmax)
            p1[i] = p.p1_scale
        return _device.get_plan(mindex, ra, rb, p1, pmax, device=dev, ext=self._moment_tables(mindex))

    def _group(self, mindex, nuse=None):
        """pcq_group over device.devices() for this multi-index (several GPUs behind one host-array call), set to
        spread the next calls over its first `nuse` members."""
        devs = _device.devices()
        pmax = int(np.max(mindex))
        ra = np.zeros((pmax + 1, self.sdim))
        rb = np.zeros((pmax + 1, self.sdim))
        p1 = np.empty(self.sdim)
        for i, p in enumerate(self.PC1ds):
            ra[:, i], rb[:, i] = p.rec_tables(pmax)
            p1[i] = p.p1_scale
        grp = _device.get_group(mindex, ra, rb, p1, pmax, devs, ext=self._moment_tables(mindex))
        _native.check(_native.lib().pcq_group_set_active(grp.handle, len(devs) if nuse is None else int(nuse)),
                      "pcq_group_set_active")
        return grp

    @staticmethod
    def _as_input(x):
        if _is_device_tensor(x):       # resident samples on an entry point that returns host arrays
            x = x.cpu().numpy()
        return np.ascontiguousarray(x, dtype=np.float64)

    # ------------------------------------------------------------------ K1
    def evalBases(self, xi, jdim):
        r"""Design matrix (M, K_jdim) of the jdim-th multi-index at the germ points xi (M, s).
        Bit-identical to the reference's rv/pcrv.py:413-442."""
        nxi, xidim = xi.shape
        assert(xidim == self.sdim)
        assert(jdim >= 0 and jdim < self.pdim)
        mindex = self.mindices[jdim]
        npc, sdim_ = mindex.shape
        assert(sdim_ == self.sdim)
        self._check_orders(mindex)
        if _is_device_tensor(xi):
            return self._evalBases_resident(xi, mindex)
        x = self._as_input(xi)
        ybases = np.empty((nxi, npc))
        nuse = _device.fan_out(8.0 * nxi * npc, 128 << 20, nxi)       # PCIe bound: 128 MB of A per extra device
        if nuse > 1:
            _native.check(_native.lib().pcq_group_eval_bases(self._group(mindex, nuse).handle, _native.ptr(x), nxi,
                                                             self.sdim, _native.ptr(ybases), npc), "pcq_group_eval_bases")
            return ybases
        plan = self._plan(mindex)
        _native.check(_native.lib().pcq_eval_bases(plan.handle, _native.ptr(x), nxi, self.sdim,
                                                   _native.ptr(ybases), npc), "pcq_eval_bases")
        return ybases

    def evalBasesNormsSq(self, jdim):
        r"""||Psi_k||^2 for the jdim-th multi-index (rv/pcrv.py:444-466), same left-to-right
        product over dimensions, vectorised over k."""
        assert(jdim >= 0 and jdim < self.pdim)
        mindex = self.mindices[jdim]
        npc, sdim_ = mindex.shape
        assert(sdim_ == self.sdim)
        norms = np.ones((npc))
        top = int(np.max(mindex)) if npc else 0
        for i in range(self.sdim):
            table = np.array([self.PC1ds[i].normsq(n) for n in range(top + 1)])
            norms *= table[mindex[:, i]]
        return norms

    # ------------------------------------------------------------------ K2
    def evalPC(self, x, exact=True):
        r"""PC expansion at x (M, s) -> (M, d) (rv/pcrv.py:469-498); A is never formed.
        ``exact=True`` keeps the reference's rounding sequence (bit-identical output);
        ``exact=False`` shares one basis product per term across outputs and uses FMA."""
        assert(self.sdim == x.shape[1])
        nsam = x.shape[0]
        for mindex in self.mindices:
            self._check_orders(mindex)
        if _is_device_tensor(x):
            return self._evalPC_resident(x, exact)
        xin = self._as_input(x)
        y = np.zeros((nsam, self.pdim))
        lib = _native.lib()
        mode = 0 if exact else 1
        # outputs sharing one multi-index (the common case) go through one launch
        groups = {}
        for j, mindex in enumerate(self.mindices):
            groups.setdefault(id(mindex), []).append(j)
        for outs in groups.values():
            mindex = self.mindices[outs[0]]
            nuse = _device.fan_out(float(nsam) * mindex.shape[0] * len(outs), 4e9, nsam)
            if nuse > 1:
                group = self._group(mindex, nuse)
            else:
                plan = self._plan(mindex)
                _device.note_evalpc_work(plan, nsam, mindex.shape[0], len(outs), mode)
            runs = [list(g) for _, g in itertools.groupby(outs, key=lambda j, c=itertools.count(): j - next(c))]
            for run in runs:     # consecutive outputs -> adjacent columns of y
                cf = np.ascontiguousarray(np.stack([self.coefs[j] for j in run]), dtype=np.float64)
                yptr = y.ctypes.data + 8 * run[0]
                if nuse > 1:
                    _native.check(lib.pcq_group_eval_pc(group.handle, _native.ptr(xin), nsam, self.sdim,
                                                        _native.ptr(cf), len(run), yptr, self.pdim, mode),
                                  "pcq_group_eval_pc")
                else:
                    _native.check(lib.pcq_eval_pc(plan.handle, _native.ptr(xin), nsam, self.sdim,
                                                  _native.ptr(cf), len(run), yptr, self.pdim, mode),
                                  "pcq_eval_pc")
        return y

    # Samples that already live in HBM (torch CUDA tensors, e.g. from utils.io.loadtxt_device): same kernels through
    # the device-pointer entry points, results stay on that device as torch tensors.
    def _evalBases_resident(self, xi, mindex):
        import torch
        from ..parallel import _resident
        xt = _resident(xi)
        out = torch.empty((xt.shape[0], mindex.shape[0]), dtype=torch.float64, device=xt.device)
        plan = self._plan(mindex, dev=xt.device.index)
        with torch.cuda.device(xt.device):
            _native.check(_native.lib().pcq_eval_bases_dev(plan.handle, xt.data_ptr(), xt.shape[0], self.sdim,
                                                           out.data_ptr(), mindex.shape[0],
                                                           torch.cuda.current_stream().cuda_stream),
                          "pcq_eval_bases_dev")
        return out

    def _evalPC_resident(self, x, exact):
        import torch
        from ..parallel import _resident
        xt = _resident(x)
        nsam = xt.shape[0]
        y = torch.zeros((nsam, self.pdim), dtype=torch.float64, device=xt.device)
        mode = 0 if exact else 1
        groups = {}
        for j, mindex in enumerate(self.mindices):
            groups.setdefault(id(mindex), []).append(j)
        with torch.cuda.device(xt.device):
            stream = torch.cuda.current_stream().cuda_stream
            for outs in groups.values():
                mindex = self.mindices[outs[0]]
                plan = self._plan(mindex, dev=xt.device.index)
                _device.note_evalpc_work(plan, nsam, mindex.shape[0], len(outs), mode)
                runs = [list(g) for _, g in itertools.groupby(outs, key=lambda j, c=itertools.count(): j - next(c))]
                for run in runs:
                    cf = torch.from_numpy(np.ascontiguousarray(np.stack([self.coefs[j] for j in run]),
                                                               dtype=np.float64)).to(xt.device)
                    if nsam:
                        _native.check(_native.lib().pcq_eval_pc_dev(plan.handle, xt.data_ptr(), nsam, self.sdim,
                                                                    cf.data_ptr(), len(run), y.data_ptr() + 8 * run[0],
                                                                    self.pdim, mode, stream), "pcq_eval_pc_dev")
        return y

    def evalLinearForms(self, x, cfmat, jdim=0, exact=True, sumsq=False):
        r"""Y[n, m] = sum_k cfmat[m, k] Psi_k(x_n) on the jdim-th basis for M coefficient vectors at once
        (one K2 launch, A never formed).  With ``sumsq`` the M values are reduced on the device to
        sum_m Y[n, m]^2 -- the quadratic form Psi(x_n)^T (F F^T) Psi(x_n) for cfmat = F^T, i.e. the
        pushed-forward variance of a fit with coefficient covariance F F^T."""
        mindex = self.mindices[jdim]
        assert(self.sdim == x.shape[1] and cfmat.shape[1] == mindex.shape[0])
        self._check_orders(mindex)
        xin = self._as_input(x)
        cf = np.ascontiguousarray(cfmat, dtype=np.float64)
        nsam, M = xin.shape[0], cf.shape[0]
        out = np.zeros((nsam, 1 if sumsq else M))
        mode = (0 if exact and not sumsq else 1) | (2 if sumsq else 0)
        nuse = _device.fan_out(float(nsam) * mindex.shape[0] * M, 4e9, nsam)
        if nuse > 1:
            _native.check(_native.lib().pcq_group_eval_pc(self._group(mindex, nuse).handle, _native.ptr(xin), nsam,
                                                          self.sdim, _native.ptr(cf), M, _native.ptr(out), out.shape[1],
                                                          mode), "pcq_group_eval_pc")
        else:
            plan = self._plan(mindex)           # (kept referenced across the native call)
            _native.check(_native.lib().pcq_eval_pc(plan.handle, _native.ptr(xin), nsam, self.sdim,
                                                    _native.ptr(cf), M, _native.ptr(out), out.shape[1], mode),
                          "pcq_eval_pc")
        return out[:, 0] if sumsq else out

    def evalQuadForm(self, x, Cmat, jdim=0):
        r"""v[n] = Psi(x_n)^T C Psi(x_n) on the jdim-th basis for a symmetric (K, K) matrix C (its upper triangle
        is what is read): K6, a triangular GEMM on the FP64 tensor pipe over packed design-matrix chunks --
        N K^2 flops, the pushed-forward variance of a fit with full coefficient covariance C.  Rounding can
        leave tiny negative values for a semi-definite C; callers that take square roots clip at zero."""
        mindex = self.mindices[jdim]
        K = mindex.shape[0]
        assert(self.sdim == x.shape[1] and Cmat.shape == (K, K))
        self._check_orders(mindex)
        xin = self._as_input(x)
        nsam = xin.shape[0]
        out = np.zeros(nsam)
        if nsam:
            cm = np.ascontiguousarray(Cmat, dtype=np.float64)
            nuse = _device.fan_out(float(nsam) * K * K, 5e11, nsam)
            if nuse > 1:
                _native.check(_native.lib().pcq_group_quadform(self._group(mindex, nuse).handle, _native.ptr(xin), nsam,
                                                               self.sdim, _native.ptr(cm), K, _native.ptr(out), 1),
                              "pcq_group_quadform")
            else:
                plan = self._plan(mindex)
                _native.check(_native.lib().pcq_quadform(plan.handle, _native.ptr(xin), nsam, self.sdim,
                                                         _native.ptr(cm), K, _native.ptr(out), 1), "pcq_quadform")
        return out

    def specialize(self, exact=True, fast=True):
        r"""Opt-in: compile (NVRTC) evalPC kernels specialised for the current multi-indices -- term list
        unrolled, 1-D tables in registers.  Bit-identical in exact mode and 3-4x faster, at ~3 ms of compile
        time per term: worth it for propagation jobs and repeated evaluation (MCMC), not for one small call.
        ``evalPC`` (up to 8 outputs per multi-index) and ``propagate`` use them automatically afterwards.
        Returns False (and keeps the generic kernels) when NVRTC is not available."""
        modes = (1 if exact else 0) | (2 if fast else 0)
        seen = set()
        ok = True
        for mindex in self.mindices:
            if id(mindex) in seen:
                continue
            seen.add(id(mindex))
            self._check_orders(mindex)
            plan = self._plan(mindex)
            st = _native.lib().pcq_plan_specialize(plan.handle, modes)
            if st == 0 and len(_device.devices()) > 1:     # and on every GPU this process spreads host-array calls over
                st = _native.lib().pcq_group_specialize(self._group(mindex).handle, modes)
            if st == -5:       # PCQ_ERR_UNSUPPORTED
                ok = False
            else:
                _native.check(st, "pcq_plan_specialize")
        return ok

    def setFunction(self):
        self.function = Function(name='PC Function')
        domain = np.array([p.domain for p in self.PC1ds])
        self.function.setDimDom(domain=np.clip(domain, -5.0, 5.0))
        self.function.setCall(self.evalPC)
        self.function.outdim = self.pdim

    # ------------------------------------------------------------------ sampling
    def sampleGerm(self, nsam=1, seed=None):
        r"""Germ samples (nsam, s) from numpy's global legacy generator, column by column
        (rv/pcrv.py:360-377): the stream is inherently sequential, so it stays on the host;
        the counter-based device stream is ``propagate``'s."""
        if seed is not None:
            np.random.seed(seed)
        germSam = np.empty((nsam, self.sdim))
        for i in range(self.sdim):
            germSam[:, i] = self.PC1ds[i].sample(nsam)
        return germSam

    def sample(self, nsam, seed=None):
        r"""Germ -> evalPC (rv/pcrv.py:511-525; the seed is applied twice there, too)."""
        if seed is not None:
            np.random.seed(seed)
        return self.evalPC(self.sampleGerm(nsam, seed=seed))

    def quadGerm(self, pts=None):
        r"""Full tensor-product Gauss quadrature of the germ (rv/pcrv.py:379-410)."""
        if pts is None:
            pts = [2] * self.sdim
        assert(len(pts) == self.sdim)
        rules = [self.PC1ds[i].quad(pts[i]) for i in range(self.sdim)]
        weights = functools.reduce(np.outer, [r[1] for r in rules]).ravel()
        points = np.array(list(itertools.product(*[r[0] for r in rules])))
        return points, weights

    # ------------------------------------------------------------------ K3 / K5 (new entry points)
    def suffStats(self, x, y=None, jdim=0):
        r"""Sufficient statistics of a regression on the jdim-th basis without forming A:
        S = sum_n z_n z_n^T, z_n = [Psi(x_n) | y_n | 1]  (see include/pcq.h, K3).
        Returns the (K+M+1, K+M+1) float64 matrix."""
        from ..lreg.stats import SuffStats
        return SuffStats.from_pc(self, x, y, jdim)

    def propagate(self, nsam, seed=0, exact=False, return_samples=False, first_index=0, specialize="auto"):
        r"""Monte Carlo forward propagation with the germ drawn on the device (Philox4x32-10
        keyed by (seed, sample index, dimension)) and evalPC fused behind it; nothing of
        size nsam leaves the GPU unless return_samples.  Returns a dict with per-output
        ``mean``, ``var`` (population), ``min``, ``max``, ``sum``, ``sumsq``, ``n``."""
        from ..parallel import propagate as _prop
        return _prop(self, nsam, seed=seed, exact=exact, return_samples=return_samples,
                     first_index=first_index, specialize=specialize)

    # ------------------------------------------------------------------ moments / Sobol (host, K x s)
    def _weights(self, i):
        # c_k^2 ||Psi_k||^2, the variance contribution of every term
        return self.coefs[i] ** 2 * self.evalBasesNormsSq(i)

    def computeMean(self):
        mean = np.zeros((self.pdim))
        for i in range(self.pdim):
            const = np.sum(self.mindices[i], axis=1) == 0
            mean[i] = np.sum(self.coefs[i][const])
        return mean

    def computeVar(self):
        var = np.zeros((self.pdim))
        for i in range(self.pdim):
            active = np.sum(self.mindices[i], axis=1) > 0
            var[i] = np.sum(self._weights(i)[active])
        return var

    def _safe_var(self):
        var = self.computeVar()
        var[var == 0.0] += 1.e-12
        return var

    def computeSens(self):
        r"""Main Sobol indices (d, s): terms that involve dimension j only (rv/pcrv.py:233-256)."""
        var = self._safe_var()
        mainsens = np.zeros((self.pdim, self.sdim))
        for i in range(self.pdim):
            mi = self.mindices[i]
            w = self._weights(i)
            nnz = np.count_nonzero(mi, axis=1)
            single = nnz == 1
            mainsens[i] = ((mi[single] > 0) * w[single, None]).sum(axis=0) / var[i]
        return mainsens

    def computeTotSens(self):
        r"""Total Sobol indices (d, s): every term that involves dimension j (rv/pcrv.py:258-280)."""
        var = self._safe_var()
        totsens = np.zeros((self.pdim, self.sdim))
        for i in range(self.pdim):
            totsens[i] = ((self.mindices[i] > 0).T.astype(np.float64) @ self._weights(i)) / var[i]
        return totsens

    def computeJointSens(self):
        r"""Joint (total-sense) Sobol indices (d, s, s), zero diagonal (rv/pcrv.py:282-309)."""
        var = self._safe_var()
        jointsens = np.zeros((self.pdim, self.sdim, self.sdim))
        for i in range(self.pdim):
            act = (self.mindices[i] > 0).astype(np.float64)
            joint = (act.T * self._weights(i)) @ act
            np.fill_diagonal(joint, 0.0)
            jointsens[i] = joint / var[i]
        return jointsens

    def computeGroupSens(self, paramIndices):
        r"""Sensitivity of a group of dimensions: terms living inside the group (rv/pcrv.py:311-338)."""
        var = self._safe_var()
        groupsens = np.zeros((self.pdim, ))
        for i in range(self.pdim):
            mi = self.mindices[i]
            outside = np.sum(np.delete(mi, paramIndices, axis=1), axis=1) == 0
            inside = outside & (np.sum(mi, axis=1) > 0)
            groupsens[i] = np.sum(self._weights(i)[inside]) / var[i]
        return groupsens

    # ------------------------------------------------------------------ structural operations (host)
    def slice_1d(self, jdim):
        assert(jdim >= 0 and jdim < self.sdim)
        mis, cfs = [], []
        for mi, cf in zip(self.mindices, self.coefs):
            keep = np.flatnonzero(np.sum(np.delete(mi, jdim, axis=1), axis=1) == 0)
            mis.append(mi[keep, jdim][:, np.newaxis])
            cfs.append(cf[keep])
        return PCRV(self.pdim, 1, self.pctypes[jdim], mi=mis, cfs=cfs)

    def compressPC(self):
        drop = np.where(self.maxOrd == 0)[0]
        mis = [np.delete(mi, drop, axis=1) for mi in self.mindices]
        types = [t for i, t in enumerate(self.pctypes) if i not in drop]
        return PCRV(self.pdim, self.sdim - len(drop), types, mi=mis, cfs=self.coefs)

    def compressMI(self):
        for i in range(self.pdim):
            uniq, inv = np.unique(self.mindices[i], return_inverse=True, axis=0)
            merged = np.zeros((uniq.shape[0],))
            np.add.at(merged, np.asarray(inv).ravel(), self.coefs[i])
            self.mindices[i] = uniq.copy()
            self.coefs[i] = merged

    def slicePC(self, fixind=None, nominal=None):
        if nominal is None:
            nominal = np.zeros((self.sdim))
        if fixind is None:
            fixind = []
        assert(len(fixind) < self.sdim)
        keep = list(set(range(self.sdim)) - set(fixind))
        tabs = []
        for i in fixind:
            assert(i < self.sdim)
            tabs.append(self.PC1ds[i](np.asarray(nominal[i], dtype=np.float64).reshape(1, 1), self.maxOrd[i]))
        cfs_new = []
        for mi, cf in zip(self.mindices, self.coefs):
            scaled = np.array(cf, dtype=np.float64, copy=True)
            for t, i in zip(tabs, fixind):
                vals = np.array([float(v.ravel()[0]) for v in t])
                scaled = scaled * vals[mi[:, i]]
            cfs_new.append(scaled)
        out = PCRV(self.pdim, self.sdim - len(fixind), [self.pctypes[i] for i in keep],
                   mi=[mi[:, keep] for mi in self.mindices], cfs=cfs_new)
        out.compressMI()
        return out


class PCRV_iid(PCRV):
    """One germ per physical dimension (rv/pcrv.py:632-655)."""

    def __init__(self, pdim, pctype, orders=None, cfs=None):
        if orders is None:
            orders = np.ones((pdim,), dtype=int)
        mindices = []
        for ii in range(pdim):
            mi = np.zeros((orders[ii] + 1, pdim), dtype=int)
            mi[:, ii] = np.arange(orders[ii] + 1)
            mindices.append(mi)
        super().__init__(pdim, pdim, pctype, mi=mindices, cfs=cfs)


class PCRV_mvn(PCRV):
    """Multivariate normal as a first-order Hermite PC (rv/pcrv.py:661-726)."""

    def __init__(self, pdim, rndind=None, mean=None, cov=None):
        if rndind is None:
            rndind = range(pdim)
        if len(rndind) > 0:
            assert(len(rndind) == len(set(rndind)))
            assert(max(rndind) < pdim and min(rndind) >= 0)
        self.mean = np.zeros(pdim,) if mean is None else mean
        self.cov = np.eye(len(rndind)) if cov is None else cov
        assert(self.mean.shape[0] == pdim)
        assert(self.cov.shape == (len(rndind), len(rndind)))
        lower = _safe_cholesky(self.cov)
        mindices, cfs = [], []
        r = 0
        for ii in range(pdim):
            if ii in rndind:
                mi = np.zeros((r + 2, pdim), dtype=int)
                mi[1:r + 2, :r + 1] = np.eye(r + 1)
                cf = np.concatenate(([self.mean[ii]], lower[r, :r + 1]))
                r += 1
            else:
                mi = np.zeros((1, pdim), dtype=int)
                cf = self.mean[ii] * np.ones((1,))
            mindices.append(mi)
            cfs.append(cf)
        super().__init__(pdim, pdim, "HG", mi=mindices, cfs=cfs)

    def pdf(self, x):
        return multivariate_normal.pdf(x, mean=self.mean, cov=self.cov)


def _safe_cholesky(cov):
    """Lower factor of a covariance, tolerant of zero eigenvalues (utils/xutils.py:264-292)."""
    assert(cov.shape[0] == cov.shape[1])
    assert(np.linalg.norm(cov - cov.T) < 1.e-14)
    if cov.shape[0] == 0:
        return np.zeros((0, 0))
    lam = np.min(np.linalg.eigvals(cov))
    if lam < 0:
        print("The matrix is not a covariance matrix (negative eigenvalues). Exiting.")
        sys.exit()
    if lam < 1e-14:
        print("Small/near-zero eigenvalue: replacing Cholesky with SVD+QR.")
        _, s, vd = np.linalg.svd(cov, hermitian=True)
        lower = np.linalg.qr(np.dot(np.diag(np.sqrt(s)), vd))[1].T
        lower = np.dot(lower, np.diag(np.sign(np.diag(lower))))
    else:
        lower = np.linalg.cholesky(cov)
    assert(np.linalg.norm(cov - np.dot(lower, lower.T)) < 1.e-12)
    return lower
