"""Times the sample-file I/O (pytuq_b200.utils.io, include/pcq_io.h) against numpy.savetxt / numpy.loadtxt on
the same matrix and checks the files / values are identical.  Host-only.  python tools/bench_io.py [rows] [cols]"""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pytuq_b200.utils import io  # noqa: E402


def best(fn, reps):
    out = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        out.append(time.perf_counter() - t)
    return min(out)


def main():
    rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
    cols = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    x = np.random.default_rng(2).standard_normal((rows, cols))
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=base) as d:
        f_np, f_us = os.path.join(d, "numpy.txt"), os.path.join(d, "native.txt")
        t_np_w = best(lambda: np.savetxt(f_np, x), 1)
        t_us_w = best(lambda: io.savetxt(f_us, x), 3)
        t_us_w1 = best(lambda: io.savetxt(f_us, x, threads=1), 1)
        same = open(f_np, "rb").read() == open(f_us, "rb").read()
        got = {}
        t_np_r = best(lambda: got.__setitem__("np", np.loadtxt(f_np)), 1)
        t_us_r = best(lambda: got.__setitem__("us", io.loadtxt(f_np)), 3)
        t_us_r1 = best(lambda: got.__setitem__("us1", io.loadtxt(f_np, threads=1)), 1)
        size = os.path.getsize(f_np)
        gpu = None
        try:
            import torch
            have_gpu = torch.cuda.is_available()
        except ImportError:
            have_gpu = False
        if have_gpu:
            # text -> HBM on the device (pcq_text.cu) against host parse + upload of the binary matrix
            def dev_route():
                got["dev"] = io.loadtxt_device(f_np)
                torch.cuda.synchronize()

            def host_route():
                got["up"] = torch.from_numpy(io.loadtxt(f_np, ndmin=2)).cuda()
                torch.cuda.synchronize()
            dev_route()
            t_dev = best(dev_route, 3)
            t_up = best(host_route, 3)
            gpu = {"loadtxt_device_s": t_dev, "host_loadtxt_plus_upload_s": t_up, "speedup_vs_host_route": t_up / t_dev,
                   "speedup_vs_numpy": None, "Mvalues_per_s": rows * cols / t_dev / 1e6,
                   "text_GB_per_s": size / t_dev / 1e9,
                   "values_identical": bool(torch.equal(got["dev"], got["up"]))}
    n = rows * cols
    print(json.dumps({
        "what": "np.savetxt / np.loadtxt (default options) vs pytuq_b200.utils.io on a %d x %d float64 matrix" % (rows, cols),
        "cores": os.cpu_count(), "file_bytes": size,
        "files_identical": bool(same),
        "values_identical": bool(np.array_equal(got["np"], got["us"]) and np.array_equal(got["us"], x)
                                 and np.array_equal(got["us1"], x)),
        "savetxt": {"numpy_s": t_np_w, "native_s": t_us_w, "native_1thread_s": t_us_w1,
                    "speedup": t_np_w / t_us_w, "native_Mvalues_per_s": n / t_us_w / 1e6},
        "to_hbm": gpu and dict(gpu, speedup_vs_numpy=t_np_r / gpu["loadtxt_device_s"]),
        "loadtxt": {"numpy_s": t_np_r, "native_s": t_us_r, "native_1thread_s": t_us_r1,
                    "speedup": t_np_r / t_us_r, "native_Mvalues_per_s": n / t_us_r / 1e6},
    }))


if __name__ == "__main__":
    main()
