"""Breakdown of the text -> HBM route (pcq_text_open / pcq_text_parse_dev): upload + scan vs conversion, against the
host reader + upload.  python tools/bench_text.py [rows] [cols]   (needs a GPU)"""
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from pytuq_b200 import _native  # noqa: E402
from pytuq_b200.utils import io  # noqa: E402


def main():
    rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
    cols = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    x = np.random.default_rng(2).standard_normal((rows, cols))
    lib = _native.lib()
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        f = os.path.join(d, "x.txt")
        io.savetxt(f, x)
        size = os.path.getsize(f)
        out = torch.empty((rows, cols), dtype=torch.float64, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        t_open, t_parse = [], []
        for _ in range(4):
            h, r, c = C.c_void_p(), C.c_int64(), C.c_int64()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _native.check(lib.pcq_text_open(0, os.fsencode(f), C.byref(h), C.byref(r), C.byref(c)))
            t1 = time.perf_counter()
            _native.check(lib.pcq_text_parse_dev(h, out.data_ptr(), cols, stream))
            t2 = time.perf_counter()
            lib.pcq_text_close(h)
            t_open.append(t1 - t0)
            t_parse.append(t2 - t1)
        same = bool(np.array_equal(out.cpu().numpy(), x))
        t0 = time.perf_counter()
        host = io.loadtxt(f, ndmin=2)
        t1 = time.perf_counter()
        up = torch.from_numpy(host).cuda()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        del up
        # files -> coefficients: the C2 fit (HG, order 3) with the samples read from their files
        fit_line = None
        if cols == 20:
            from pytuq_b200.rv.pcrv import PCRV
            from pytuq_b200.utils.mindex import get_mi
            from pytuq_b200.lreg import lsq, PCBasisEvaluator
            mi = get_mi(3, cols)
            pc = PCRV(1, cols, "HG", mi=mi)
            c_true = np.random.default_rng(7).standard_normal(mi.shape[0]) / (1.0 + mi.sum(axis=1)) ** 2
            pc.setCfs([c_true])
            fy = os.path.join(d, "y.txt")
            io.savetxt(fy, pc.evalPC(x, exact=False)[:, 0] + 1e-2 * np.random.default_rng(8).standard_normal(rows))

            def fit_from(load):
                fitter = lsq()
                fitter.setBasisEvaluator(PCBasisEvaluator(0), (pc,))
                t = time.perf_counter()
                fitter.fit(load(f), load(fy))
                torch.cuda.synchronize()
                return time.perf_counter() - t, fitter.cf
            fit_from(io.loadtxt_device)
            t_dev, cf_dev = fit_from(io.loadtxt_device)
            t_host, cf_host = fit_from(lambda name: io.loadtxt(name, ndmin=2))
            fit_line = {"terms": int(mi.shape[0]), "device_route_s": t_dev, "host_reader_route_s": t_host,
                        "max_abs_coef_diff": float(np.max(np.abs(cf_dev - cf_host))),
                        "max_abs_coef_error_vs_truth": float(np.max(np.abs(cf_dev - c_true)))}
    print(json.dumps({
        "what": "text -> HBM, %d x %d float64 (%d bytes of text)" % (rows, cols, size), "values_identical": same,
        "open_upload_scan_s": min(t_open[1:]), "parse_s": min(t_parse[1:]), "first_call_s": t_open[0] + t_parse[0],
        "total_s": min(a + b for a, b in zip(t_open[1:], t_parse[1:])),
        "text_GB_per_s": size / min(a + b for a, b in zip(t_open[1:], t_parse[1:])) / 1e9,
        "Mvalues_per_s": rows * cols / min(a + b for a, b in zip(t_open[1:], t_parse[1:])) / 1e6,
        "host_route": {"loadtxt_s": t1 - t0, "upload_s": t2 - t1, "cores": os.cpu_count()},
        "files_to_coefficients": fit_line,
    }))


if __name__ == "__main__":
    main()
