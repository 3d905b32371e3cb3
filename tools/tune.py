"""Times the hot kernels (CUDA events, device-resident inputs) for the BASELINE shapes and the
compile-time variants selectable through PCQ_* environment switches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pytuq_b200 import _native
from pytuq_b200.rv.pcrv import PCRV
from pytuq_b200.utils.mindex import get_mi

lib = _native.lib()
st = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit(fn, reps=5, warm=2):
    ts = []
    for it in range(warm + reps):
        e0.record(); fn(); e1.record(); e1.synchronize()
        if it >= warm:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def setenv(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)


which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["k1", "k2", "k3"]
shapes = [("C2", "HG", 20, 3, 1_250_000 if "k3chunk" in which else 1_000_000), ("C1", "LU", 10, 4, 1_000_000), ("C3", "LU", 50, 2, 500_000),
          ("C4", "HG", 20, 4, 200_000), ("C5", "LU", 30, 3, 200_000)]
for name, fam, d, p, n in shapes:
    mi = get_mi(p, d)
    K = mi.shape[0]
    pc = PCRV(1, d, fam, mi=mi)
    plan = pc._plan(mi)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = (torch.randn((n, d), dtype=torch.float64, device="cuda", generator=g) if fam == "HG"
         else torch.rand((n, d), dtype=torch.float64, device="cuda", generator=g) * 2 - 1)
    c = torch.randn(K, dtype=torch.float64, device="cuda", generator=g)
    y = torch.empty((n, 1), dtype=torch.float64, device="cuda")
    if "k1" in which:
        n1 = min(n, int(6e9 // (8 * K)))
        A = torch.empty((n1, K), dtype=torch.float64, device="cuda")
        for v in (None, 1):
            setenv(PCQ_K1_V1=v)
            ms = timeit(lambda: _native.check(lib.pcq_eval_bases_dev(plan.handle, x.data_ptr(), n1, d, A.data_ptr(), K, st)))
            print(f"{name} K1 {'v1' if v else 'v2'}: {ms:8.3f} ms  {8.0 * n1 * (d + K) / ms / 1e6:8.1f} GB/s  {n1 * K / ms / 1e6:8.1f} Gevals/s", flush=True)
        setenv(PCQ_K1_V1=None)
        del A
    if "k2" in which:
        for mode in (0, 1):
            for cfg in (0, 1, 2, 3):
                setenv(PCQ_K2_CFG=cfg)
                try:
                    ms = timeit(lambda: _native.check(lib.pcq_eval_pc_dev(plan.handle, x.data_ptr(), n, d, c.data_ptr(), 1, y.data_ptr(), 1, mode, st)))
                except Exception as ex:
                    print(f"{name} K2 mode{mode} cfg{cfg}: failed {ex}"); continue
                fl = n * (4.0 * d * (p - 1) + 3.0 * K)
                print(f"{name} K2 mode{mode} cfg{cfg}: {ms:8.3f} ms  {n / ms / 1e3:9.1f} Msamples/s  {fl / ms / 1e9:6.2f} TFLOP/s(alg)", flush=True)
            setenv(PCQ_K2_CFG=None)
            ms = timeit(lambda: _native.check(lib.pcq_eval_pc_dev(plan.handle, x.data_ptr(), n, d, c.data_ptr(), 1, y.data_ptr(), 1, mode, st)))
            print(f"{name} K2 mode{mode} auto: {ms:8.3f} ms", flush=True)
    if "k2m" in which:
        for M in (16, 64, 256):
            cm = torch.randn((M, K), dtype=torch.float64, device="cuda", generator=g)
            nn = min(n, 200_000)
            ym = torch.empty((nn, M), dtype=torch.float64, device="cuda")
            for cfg in (None, 0, 1, 2, 3):
                setenv(PCQ_K2_CFG=cfg)
                try:
                    ms = timeit(lambda: _native.check(lib.pcq_eval_pc_dev(plan.handle, x.data_ptr(), nn, d, cm.data_ptr(), M, ym.data_ptr(), M, 1, st)), reps=3, warm=1)
                except Exception as ex:
                    print(f"{name} K2 M={M} cfg{cfg}: failed"); continue
                print(f"{name} K2 fast M={M} cfg={cfg}: n={nn} {ms:8.3f} ms  {2.0 * nn * K * M / ms / 1e9:6.2f} TFLOP/s (2NKM)", flush=True)
            setenv(PCQ_K2_CFG=None)
            del cm, ym
    if "k3chunk" in which and name == "C2":
        S = torch.empty((K + 2, K + 2), dtype=torch.float64, device="cuda")
        for mb in (0, 24, 48, 96):
            setenv(PCQ_GRAM_CHUNK_MB=mb)
            ms = timeit(lambda: _native.check(lib.pcq_suffstats_dev(plan.handle, x.data_ptr(), n, d, y.data_ptr(), 1, 1, S.data_ptr(), 0, st)), reps=3, warm=1)
            fl = float(n) * K * (K + 1) + 2.0 * n * K + 2.0 * n
            print(f"{name} K3 chunk={mb}MB: n={n} {ms:8.3f} ms  {fl / ms / 1e9:6.2f} TFLOP/s(alg)", flush=True)
        setenv(PCQ_GRAM_CHUNK_MB=None)
        del S
    if "k3only" in which:
        S = torch.empty((K + 2, K + 2), dtype=torch.float64, device="cuda")
        n3 = min(n, 400_000 if K < 3000 else 60_000)
        ms = timeit(lambda: _native.check(lib.pcq_suffstats_dev(plan.handle, x.data_ptr(), n3, d, y.data_ptr(), 1, 1, S.data_ptr(), 0, st)), reps=3, warm=1)
        fl = float(n3) * K * (K + 1) + 2.0 * n3 * K + 2.0 * n3
        print(f"{name} K3 env: n={n3} {ms:8.3f} ms  {fl / ms / 1e9:6.2f} TFLOP/s(alg)", flush=True)
        del S
    if "k3" in which:
        S = torch.empty((K + 2, K + 2), dtype=torch.float64, device="cuda")
        n3 = min(n, 400_000 if K < 3000 else 60_000)
        for tag, env in (("pack+syrk", dict(PCQ_GRAM_SYRK=None)), ("fused generator", dict(PCQ_GRAM_SYRK=0))):
            setenv(**env)
            ms = timeit(lambda: _native.check(lib.pcq_suffstats_dev(plan.handle, x.data_ptr(), n3, d, y.data_ptr(), 1, 1, S.data_ptr(), 0, st)), reps=3, warm=1)
            fl = float(n3) * K * (K + 1) + 2.0 * n3 * K + 2.0 * n3
            print(f"{name} K3 {tag}: n={n3} {ms:8.3f} ms  {fl / ms / 1e9:6.2f} TFLOP/s(alg)", flush=True)
        setenv(PCQ_GRAM_SYRK=None)
        del S
    if "k4" in which:
        n4 = min(n, int(4e9 // (8 * K)))
        A = torch.empty((n4, K), dtype=torch.float64, device="cuda")
        _native.check(lib.pcq_eval_bases_dev(plan.handle, x.data_ptr(), n4, d, A.data_ptr(), K, st))
        S = torch.empty((K + 2, K + 2), dtype=torch.float64, device="cuda")
        for tag, env in (("pack+syrk", dict(PCQ_GRAM_SYRK=None)), ("v1 generic", dict(PCQ_GRAM_SYRK=0))):
            setenv(**env)
            ms = timeit(lambda: _native.check(lib.pcq_gram_dev(0, A.data_ptr(), n4, K, K, y.data_ptr(), 1, 1, S.data_ptr(), 0, st)), reps=3, warm=1)
            fl = float(n4) * K * (K + 1) + 2.0 * n4 * K + 2.0 * n4
            print(f"{name} K4 {tag}: n={n4} {ms:8.3f} ms  {fl / ms / 1e9:6.2f} TFLOP/s(alg)", flush=True)
        setenv(PCQ_GRAM_SYRK=None)
        del S, A
    if "k6" in which:
        n6 = min(n, 400_000 if K < 3000 else 60_000)
        Cm = torch.randn((K, K), dtype=torch.float64, device="cuda", generator=g)
        Cm = Cm @ Cm.T / K
        v = torch.empty(n6, dtype=torch.float64, device="cuda")
        ms = timeit(lambda: _native.check(lib.pcq_quadform_dev(plan.handle, x.data_ptr(), n6, d, Cm.data_ptr(), K, v.data_ptr(), 1, st)), reps=3, warm=1)
        print(f"{name} K6 quadform: n={n6} {ms:8.3f} ms  {float(n6) * K * K / ms / 1e9:6.2f} TFLOP/s (N K^2)", flush=True)
        F = torch.linalg.cholesky(Cm + 1e-9 * torch.eye(K, dtype=torch.float64, device="cuda")).T.contiguous()
        n6b = min(n6, 100_000)
        ms = timeit(lambda: _native.check(lib.pcq_eval_pc_dev(plan.handle, x.data_ptr(), n6b, d, F.data_ptr(), K, v.data_ptr(), 1, 3, st)), reps=2, warm=1)
        print(f"{name} K2 quadratic-form mode (factor, {K} outputs): n={n6b} {ms:8.3f} ms  -> {ms * n6 / n6b:8.3f} ms at n={n6}", flush=True)
        del Cm, F, v
    if "k2lf" in which:
        for M in (64, 256):
            cm = torch.randn((M, K), dtype=torch.float64, device="cuda", generator=g)
            nn = min(n, 200_000)
            ym = torch.empty((nn, M), dtype=torch.float64, device="cuda")
            ms = timeit(lambda: _native.check(lib.pcq_eval_pc_dev(plan.handle, x.data_ptr(), nn, d, cm.data_ptr(), M, ym.data_ptr(), M, 1, st)), reps=3, warm=1)
            print(f"{name} evalPC fast M={M} (tensor pipe): n={nn} {ms:8.3f} ms  {2.0 * nn * K * M / ms / 1e9:6.2f} TFLOP/s (2NKM)", flush=True)
            del cm, ym
