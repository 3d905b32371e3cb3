"""Times the cubins made by tools/k2spec/variants.py on the GPU box (driver API through ctypes, torch for memory) and
checks them against the generic exact kernel."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from pytuq_b200 import _native
from pytuq_b200.rv.pcrv import PCRV
from pytuq_b200.utils.mindex import get_mi

EXP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "exp")
cu = C.CDLL("libcuda.so.1")
man = json.load(open(os.path.join(EXP, "manifest.json")))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
only = sys.argv[2:] 
torch.cuda.init()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
res = {}
for var in man:
    if only and var["tag"] not in only:
        continue
    fam, d, p = var["fam"], var["d"], var["p"]
    mi = get_mi(p, d)
    K = mi.shape[0]
    x = torch.randn((n, d), dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    cf = torch.randn(K, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(2)) / \
        (1.0 + torch.from_numpy(mi.sum(axis=1)).to(dev)) ** 2
    y = torch.zeros((n, 1), dtype=torch.float64, device=dev)
    order = torch.tensor(var["order"], dtype=torch.int64, device=dev)
    cfp = cf[order].contiguous()
    funcs = []
    for k in var["kernels"]:
        mod = C.c_void_p()
        data = open(os.path.join(EXP, k["cubin"]), "rb").read()
        rc = cu.cuModuleLoadData(C.byref(mod), data)
        assert rc == 0, ("cuModuleLoadData", rc, k["cubin"])
        fn = C.c_void_p()
        assert cu.cuModuleGetFunction(C.byref(fn), mod, k["name"].encode()) == 0
        dptr, sz = C.c_uint64(), C.c_size_t()
        assert cu.cuModuleGetGlobal_v2(C.byref(dptr), C.byref(sz), mod, (k["name"] + "_cf").encode()) == 0
        sub = cfp[k["j0"]:k["j0"] + k["n"]].contiguous()
        assert cu.cuMemcpyDtoD_v2(dptr, C.c_uint64(sub.data_ptr()), C.c_size_t(8 * k["n"])) == 0
        funcs.append((fn, sub))
    per = var["threads"] * var["spt"]
    nblocks = (n + per - 1) // per
    grid = nblocks if var.get('noloop') else min(nblocks, 148 * 8)
    st = torch.cuda.current_stream().cuda_stream

    def launch():
        for fn, sub in funcs:
            args = [C.c_uint64(x.data_ptr()), C.c_longlong(n), C.c_longlong(d), C.c_uint64(y.data_ptr()), C.c_longlong(1),
                    C.c_uint64(sub.data_ptr())]
            argv = (C.c_void_p * len(args))(*[C.cast(C.pointer(a), C.c_void_p) for a in args])
            rc = cu.cuLaunchKernel(fn, grid, 1, 1, var["threads"], 1, 1, 0, C.c_void_p(st), argv, None)
            assert rc == 0, ("cuLaunchKernel", rc)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for it in range(6):
        e0.record(); launch(); e1.record(); e1.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    # check against the generic exact kernel on a slice
    pc = PCRV(1, d, fam, mi=mi, cfs=cf.cpu().numpy())
    m = min(n, 20000)
    ref = pc.evalPC(x[:m].cpu().numpy())[:, 0]
    err = float(np.max(np.abs(y[:m, 0].cpu().numpy() - ref)) / np.max(np.abs(ref)))
    res[var["tag"]] = dict(ms=ms, samples_per_s=n / ms * 1e3, rel_err=err)
    print(f"{var['tag']:22s} {ms:8.3f} ms  {n / ms / 1e6:8.2f} Gsamples/s  err {err:.1e}", flush=True)
json.dump(res, open(os.path.join(os.path.dirname(EXP), "..", "gpurun_out", "k2_variants.json"), "w"))
