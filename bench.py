#!/usr/bin/env python
"""Headline benchmark of the PC hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (config.workload): BASELINE.json's metric configuration -- Hermite PCE, d=20, order 3
(K=1771 terms), basis evaluation + least-squares fit on N = 10M synthetic Gaussian samples.
STRONG scaling: the 10M samples are the whole job at every GPU count (N=1: all 10M on one GPU;
N GPUs: 10M/N contiguous rows per rank).  `--samples-per-gpu R` switches to weak scaling (R rows per rank).

A step = one pass of the hot path over the rank's sample block:
    fused basis-evaluation + Gram/A^T y statistics (K3, DMMA; in moment form -- csrc/pcq_moment.cu -- where that is
    faster, as at this shape)  ->  one all-reduce of the statistics (N>1)  ->  K x K normal-equation solve (cuSOLVER Cholesky on the device, a plain
    library call: the factorisation is not sample-parallel)  ->  coefficients.
`value` is basis evaluations (samples x terms) per second for that whole step with x, y
already resident in HBM; `e2e` is the same metric through the public API
(`lsq.fit(x, y)` with numpy host arrays: pinned double-buffered H2D of x, y and D2H of the
coefficients inside the timed region).  `roofline` is for the dominant kernel (K3) against the FP64 tensor (DMMA)
peak measured live on the same GPU (MEASURED_PEAKS.json has no FP64 number): `achieved` counts the multiply-adds the
DMMA pipe EXECUTES in the moment form (the pipe's utilisation); the same step counted with the dense N K^2 flops of
SURVEY 8(d) is `dense_equivalent` (above the roof: the moment form does ~5x fewer multiply-adds), and the dense pack +
SYRK form itself is timed beside it in `kernels.K3_pack_syrk_form`.  The K1 (materialised A) and K2 (evalPC) kernels are
reported in `kernels` with their own roofs.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv[1:]:
    # the CPU reference arm uses every core this process may run on: torchrun exports OMP_NUM_THREADS=1, and OpenBLAS
    # sizes its thread pool from the environment when numpy/scipy are first imported -- so fix it before that import
    _cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "GOTO_NUM_THREADS"):
        os.environ[_v] = str(_cores)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D, P, FAMILY = 20, 3, "HG"
N_TOTAL = 10_000_000
WORKLOAD = "C2/target: Hermite PCE d=20 p=3 K=1771, basis eval + lstsq fit on N=10M Gaussian samples (strong scaling over 1/2/4/8 GPUs)"


def shard_rows(n_total, rank, world):
    base, extra = divmod(int(n_total), int(world))
    return base + (1 if rank < extra else 0)


def make_config(args, world, K):
    """The `config` object of the JSON line; identical for the B200 arm and the reference arm."""
    if args.samples_per_gpu:
        n_total, scaling = args.samples_per_gpu * world, "weak"
        workload = WORKLOAD.replace("N=10M", f"N={args.samples_per_gpu}/GPU").replace("strong", "weak")
    else:
        n_total, scaling = args.samples, "strong"
        workload = WORKLOAD if n_total == N_TOTAL else WORKLOAD.replace("N=10M", f"N={n_total}")
    per_gpu = -(-n_total // world)
    cfg = {"workload": workload, "samples": int(n_total), "terms": int(K), "dim": D, "order": P, "family": FAMILY,
           "parallelism": f"sample-sharded x{world}, one all-reduce of the packed statistics",
           "l2": "inputs (x,y = %.0f MB per GPU) larger than the 126 MB L2; no explicit flush" % (per_gpu * (D + 1) * 8 / 1e6)}
    return cfg, n_total, scaling


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--samples", type=int, default=N_TOTAL, help="total samples of the job (strong scaling)")
    ap.add_argument("--samples-per-gpu", type=int, default=0, help="rows per rank: weak scaling instead")
    ap.add_argument("--cpu-samples", type=int, default=20000,
                    help="rows of the workload one step of the CPU reference arm is timed on")
    ap.add_argument("--cpu-probe-samples", type=int, default=100000,
                    help="rows of the single larger pass that checks the CPU path scales linearly (0 = skip)")
    ap.add_argument("--no-extras", action="store_true", help="skip K1/K2/peak side measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); smax.append(float(r[1])); power.append(float(r[2]))
            except ValueError:
                continue
            for name, flag in zip(names, r[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU reference arm
class CpuReference:
    """The reference's own CPU implementation of the path, timed on the host cores.

    kind "reference": the UNMODIFIED PyTUQ copied to oracle/_ref (oracle/Makefile; matplotlib stubbed), driven through
    its public API exactly as surrogates/pce.py:378-383 does -- lsq().setBasisEvaluator(closure over
    PCRV.evalBases) -> evalBases (rv/pcrv.py:413-442) + lsq.fita (scipy gelsd, lreg/lreg.py:328-339).
    kind "port": the numpy oracle restating the same two functions, only when oracle/_ref is absent.
    BLAS threads are set explicitly (threadpoolctl) to the cores this process may use, so that torchrun's
    OMP_NUM_THREADS=1 cannot skew the lstsq timing; evalBases is single-threaded Python + numpy in the reference."""

    def __init__(self):
        from oracle import ref_loader
        self.cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        try:
            from threadpoolctl import threadpool_limits
            self._limits = threadpool_limits(limits=self.cores)
        except Exception:
            self._limits = None
        if ref_loader.available():
            R = ref_loader.import_reference()
            self.kind = "reference"
            self.mi = R["get_mi"](P, D)
            self._pcrv = R["PCRV"](1, D, FAMILY, mi=self.mi)
            self._lsq = R["lsq"]
        else:
            from oracle import pc_oracle as orc
            self.kind = "port"
            self.mi = orc.get_mi(P, D)
            self._orc = orc
        self.K = self.mi.shape[0]

    def blas_threads(self):
        try:
            from threadpoolctl import threadpool_info
            return max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
        except Exception:
            return self.cores

    def bases(self, x):
        if self.kind == "reference":
            return self._pcrv.evalBases(x, 0)
        return self._orc.eval_bases(x, self.mi, FAMILY)

    def one_pass(self, x, y):
        """evalBases + lsq.fita on (x, y); returns (cf, seconds of evalBases, seconds of the lstsq)."""
        t0 = time.perf_counter()
        A = self.bases(x)
        t1 = time.perf_counter()
        if self.kind == "reference":
            fit = self._lsq()
            fit.fita(A, y)
            cf = fit.cf
        else:
            cf = self._orc.lsq_fit(A, y)
        t2 = time.perf_counter()
        return cf, t1 - t0, t2 - t1

    def describe(self, n, tb, tf):
        what = ("unmodified PyTUQ (oracle/_ref) PCRV.evalBases + lsq.fita" if self.kind == "reference"
                else "oracle port (numpy restatement of PCRV.evalBases + lsq.fita)")
        return (f"{what} on {n} rows of the workload: basis {tb:.2f} s (single-threaded Python/numpy loop), "
                f"lstsq {tf:.2f} s ({self.blas_threads()} BLAS threads of {self.cores} usable cores)")


def synth_host(n, seed, mi):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, D))
    c_true = rng.standard_normal(mi.shape[0]) / (1.0 + mi.sum(axis=1)) ** 2
    return rng, x, c_true


def cpu_probe(ref, x, y, sizes):
    """Per-unit CPU throughput at several row counts (the reference's cost is linear in N: K*s length-N
    multiplies + an O(N K^2) SVD least squares) -> {rows: {...}}."""
    out = {}
    for n in sizes:
        if n <= 0 or n > x.shape[0]:
            continue
        _, tb, tf = ref.one_pass(x[:n], y[:n])
        out[str(n)] = {"basis_s": tb, "lstsq_s": tf, "evals_per_s": n * ref.K / (tb + tf)}
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation (oracle/_ref; the oracle port only when that copy is
    absent) on a bounded sample of the SAME workload per step, all usable host cores, the requested steps/warm-up.
    `value` is per-unit throughput (basis evaluations per second); 10M rows would need a 141.7 GB design matrix and
    hours of SVD, so the 10M figure is the linear extrapolation stated in `extrapolated_seconds_at_full_size`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    ref = CpuReference()
    mi, K = ref.mi, ref.K
    config, n_total, scaling = make_config(args, max(world, args.gpus), K)
    n = args.cpu_samples
    nmax = max(n, args.cpu_probe_samples)
    _, x, c_true = synth_host(nmax, 2, mi)
    rng = np.random.default_rng(99)
    y = np.empty(nmax)
    for r0 in range(0, nmax, 20000):                 # untimed: only to synthesise the targets
        y[r0:r0 + 20000] = ref.bases(x[r0:r0 + 20000]) @ c_true
    y += 1e-2 * rng.standard_normal(nmax)
    times = []
    t_begin = time.perf_counter()
    for it in range(args.warmup + args.steps):
        cf, tb, tf = ref.one_pass(x[:n], y[:n])
        if it >= args.warmup:
            times.append((tb, tf))
    tb = float(np.mean([t[0] for t in times])); tf = float(np.mean([t[1] for t in times]))
    value = n * K / (tb + tf)
    probe = cpu_probe(ref, x, y, sorted({n // 2, args.cpu_probe_samples} - {0, n}))
    probe[str(n)] = {"basis_s": tb, "lstsq_s": tf, "evals_per_s": value}
    err = float(np.max(np.abs(cf - c_true)))
    unit = "basis evals/s (samples*terms)"
    line = {
        "impl": "reference", "metric": "pc_basis_evals_per_s", "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (tb + tf),
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "fit_seconds": tf, "basis_seconds": tb,
        "max_abs_coef_error_vs_truth": err,
        "cpu_baseline": {"value": value, "unit": unit, "cores": ref.cores, "kind": ref.kind,
                         "sample": ref.describe(n, tb, tf) + f"; {args.steps} timed steps after {args.warmup} warm-up",
                         "blas_threads": ref.blas_threads(),
                         "linearity": probe,
                         "extrapolated_seconds_at_full_size": n_total * K / value},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "wall_s": time.perf_counter() - t_begin,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- B200 arm
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    from pytuq_b200 import _native, device as pdev
    from pytuq_b200.rv.pcrv import PCRV
    from pytuq_b200.utils.mindex import get_mi
    from pytuq_b200.lreg import lsq, PCBasisEvaluator
    from pytuq_b200.lreg.lreg import solve_normal, solve_normal_device

    # stdout must carry exactly one JSON line, but libraries write there too (NCCL's version banner on
    # communicator creation): route fd 1 to stderr while working and restore it for the final print
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    pdev.set_device(local)
    pdev.set_devices([local])          # this arm counts GPUs by ranks: the public API must stay on the rank's own device
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _native.lib()
    dev = torch.device("cuda", local)

    mi = get_mi(P, D)
    K = mi.shape[0]
    config, n_total, scaling = make_config(args, world, K)
    n = args.samples_per_gpu if args.samples_per_gpu else shard_rows(n_total, rank, world)
    pc = PCRV(1, D, FAMILY, mi=mi)
    plan = pc._plan(mi, dev=local)

    # synthetic inputs, generated on the device for the HBM-resident arm
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    x = torch.randn((n, D), dtype=torch.float64, device=dev, generator=g)
    gc = torch.Generator(device=dev).manual_seed(7)
    c_true = torch.randn(K, dtype=torch.float64, device=dev, generator=gc) / \
        (1.0 + torch.from_numpy(mi.sum(axis=1)).to(dev)) ** 2
    y = torch.empty((n, 1), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    _native.check(lib.pcq_eval_pc_dev(plan.handle, x.data_ptr(), n, D, c_true.data_ptr(), 1, y.data_ptr(), 1, 1, stream))
    y += 1e-2 * torch.randn((n, 1), dtype=torch.float64, device=dev, generator=g)
    L = K + 2
    S = torch.empty((L, L), dtype=torch.float64, device=dev)
    S_host = torch.empty((L, L), dtype=torch.float64).pin_memory()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step():
        """K3 -> (all-reduce) -> K x K Cholesky solve on the device -> D2H of the K coefficients.
        Returns (cf, ms of K3, ms of all-reduce, s of solve incl. the final sync)."""
        ev[0].record()
        _native.check(lib.pcq_suffstats_dev(plan.handle, x.data_ptr(), n, D, y.data_ptr(), 1, 1, S.data_ptr(), 0, stream))
        ev[1].record()
        if world > 1:
            dist.all_reduce(S, op=dist.ReduceOp.SUM)
        ev[2].record()
        C = solve_normal_device(S[:K, :K], S[:K, K])          # cuSOLVER potrf/potrs on the device
        if C is None:                                        # not positive definite: host eigen path
            S_host.copy_(S)
            Sh = S_host.numpy()
            cf = solve_normal(Sh[:K, :K], Sh[:K, K])
        else:
            cf = C.cpu().numpy()                             # D2H of K doubles; syncs the stream
        ev[3].record()
        ev[3].synchronize()
        return cf, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), 1e-3 * ev[2].elapsed_time(ev[3])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        cf, *_ = step()
    sampler = ClockSampler(local)
    launches0 = lib.pcq_launch_count()
    barrier()
    if rank == 0:
        sampler.start()
    t_start = time.perf_counter()
    k3_ms, ar_ms, solve_s = [], [], []
    for _ in range(args.steps):
        cf, a, b, c = step()
        k3_ms.append(a); ar_ms.append(b); solve_s.append(c)
    barrier()
    elapsed = time.perf_counter() - t_start
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.pcq_launch_count() - launches0
    tmax = torch.tensor([elapsed, float(np.mean(k3_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    elapsed, k3_mean_ms = float(tmax[0]), float(tmax[1])
    ms_per_step = 1e3 * elapsed / args.steps
    total_evals = float(n_total) * K
    value = total_evals / (elapsed / args.steps)
    err = float(np.max(np.abs(cf - c_true.cpu().numpy())))

    # ---------------- end to end through the public API (host numpy in, coefficients out)
    xh = x.cpu().numpy()
    yh = y.cpu().numpy()[:, 0].copy()
    fitter = lsq()
    fitter.setBasisEvaluator(PCBasisEvaluator(0), (pc,))

    def e2e_step():
        if world > 1:
            from pytuq_b200.parallel import suffstats_pc
            st = suffstats_pc(pc, xh, yh, 0, distributed=True)
            fitter._fit_stats(st)
        else:
            fitter.fit(xh, yh)
        return fitter.cf

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        cf_e2e = e2e_step()
    barrier()
    e2e_t = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = total_evals / float(e2e_t[0])
    e2e_err = float(np.max(np.abs(cf_e2e - cf)))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel + side kernels (rank 0)
    flops = float(n) * K * (K + 1) + 2.0 * n * K * 1 + 2.0 * n * 1          # dense form, SURVEY 8(d)
    minfo = (ctypes.c_double * 8)()
    _native.check(lib.pcq_moment_info(plan.handle, 1, n, minfo))
    moment_form = bool(minfo[5])
    flops_exec = 2.0 * n * minfo[2] if moment_form else flops              # multiply-adds the DMMA pipe executes
    achieved = flops_exec / (k3_mean_ms * 1e-3) / 1e12
    peaks = {}
    kernels = {}
    if not args.no_extras:
        for kind, name in ((1, "fp64_dmma_tflops"), (0, "fp64_dfma_tflops"), (2, "hbm_copy_gbs")):
            best = 0.0
            for _ in range(2):
                v = ctypes.c_double()
                _native.check(lib.pcq_measure_peak(local, kind, ctypes.byref(v)))
                best = max(best, v.value)
            peaks[name] = best
        # K3 in its dense form (pack + SYRK, PCQ_GRAM_MOMENT=0) on the same resident samples: the N K^2 contraction
        # the moment form replaces, against the same DMMA roof
        if moment_form:
            os.environ["PCQ_GRAM_MOMENT"] = "0"
            try:
                ts = []
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                S2 = torch.empty_like(S)
                for it in range(3):
                    e0.record()
                    _native.check(lib.pcq_suffstats_dev(plan.handle, x.data_ptr(), n, D, y.data_ptr(), 1, 1, S2.data_ptr(), 0, stream))
                    e1.record(); e1.synchronize()
                    if it >= 1:
                        ts.append(e0.elapsed_time(e1))
                dgn = torch.sqrt(torch.abs(torch.diagonal(S2)))
                kernels["K3_pack_syrk_form"] = {"rows": n, "ms": float(np.mean(ts)), "TFLOPs_algorithmic": flops / (np.mean(ts) * 1e-3) / 1e12,
                                                "bound": "tensor", "max_rel_diff_of_the_statistics_vs_moment_form":
                                                float(torch.max(torch.abs(S2 - S) / torch.outer(dgn, dgn)))}
                del S2
            finally:
                os.environ.pop("PCQ_GRAM_MOMENT", None)
        # K1: materialised design matrix, device resident (N*K*8 bytes must fit: use 2^19 rows = 7.4 GB)
        n1 = 1 << 19
        A = torch.empty((n1, K), dtype=torch.float64, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for it in range(6):
            e0.record()
            _native.check(lib.pcq_eval_bases_dev(plan.handle, x.data_ptr(), n1, D, A.data_ptr(), K, stream))
            e1.record(); e1.synchronize()
            if it >= 2:
                ts.append(e0.elapsed_time(e1))
        k1_ms = float(np.mean(ts))
        k1_bytes = 8.0 * n1 * (D + K)
        kernels["K1_evalBases"] = {"rows": n1, "ms": k1_ms, "evals_per_s": n1 * K / (k1_ms * 1e-3),
                                   "GBps": k1_bytes / (k1_ms * 1e-3) / 1e9, "bound": "hbm"}
        del A
        # K2: surrogate evaluation without A (exact and fast modes)
        yy = torch.empty((n, 1), dtype=torch.float64, device=dev)
        for mode, tag in ((0, "K2_evalPC_exact"), (1, "K2_evalPC_fast")):
            ts = []
            for it in range(5):
                e0.record()
                _native.check(lib.pcq_eval_pc_dev(plan.handle, x.data_ptr(), n, D, c_true.data_ptr(), 1,
                                                  yy.data_ptr(), 1, mode, stream))
                e1.record(); e1.synchronize()
                if it >= 2:
                    ts.append(e0.elapsed_time(e1))
            ms = float(np.mean(ts))
            fl = n * (4.0 * D * (P - 1) + 3.0 * K)
            kernels[tag] = {"rows": n, "ms": ms, "samples_per_s": n / (ms * 1e-3),
                            "TFLOPs_algorithmic": fl / (ms * 1e-3) / 1e12, "bound": "fp64"}
        # K2 specialised per multi-index (opt-in, NVRTC): same call, after pcq_plan_specialize
        t0 = time.perf_counter()
        rc = lib.pcq_plan_specialize(plan.handle, 3)
        t_comp = time.perf_counter() - t0
        if rc == 0:
            for mode, tag in ((0, "K2s_evalPC_exact_specialised"), (1, "K2s_evalPC_fast_specialised")):
                ts = []
                for it in range(5):
                    e0.record()
                    _native.check(lib.pcq_eval_pc_dev(plan.handle, x.data_ptr(), n, D, c_true.data_ptr(), 1,
                                                      yy.data_ptr(), 1, mode, stream))
                    e1.record(); e1.synchronize()
                    if it >= 2:
                        ts.append(e0.elapsed_time(e1))
                ms = float(np.mean(ts))
                fl = n * (4.0 * D * (P - 1) + 3.0 * K)
                kernels[tag] = {"rows": n, "ms": ms, "samples_per_s": n / (ms * 1e-3),
                                "TFLOPs_algorithmic": fl / (ms * 1e-3) / 1e12, "bound": "fp64",
                                "nvrtc_compile_s": t_comp}
                if mode == 1:
                    # Horner form over the multi-index trie: ONE fused multiply-add per term, i.e. 2 (K - 1) executed
                    # flops where the survey's F_min counts 3 K -- the algorithmic fraction can exceed 1, the executed
                    # one is the pipe utilisation (a lower bound: table recurrences and node adds not counted)
                    kernels[tag]["form"] = "Horner over the multi-index trie, 1 DFMA per term (csrc/pcq_spec.cu)"
                    kernels[tag]["TFLOPs_executed_min"] = n * 2.0 * (K - 1) / (ms * 1e-3) / 1e12
        # K6: predictive variance psi^T C psi as a triangular GEMM on the tensor pipe (N K^2 flops)
        n6 = min(n, 400_000)
        Cm = torch.randn((K, K), dtype=torch.float64, device=dev)
        Cm = Cm @ Cm.T / K
        v6 = torch.empty(n6, dtype=torch.float64, device=dev)
        ts = []
        for it in range(4):
            e0.record()
            _native.check(lib.pcq_quadform_dev(plan.handle, x.data_ptr(), n6, D, Cm.data_ptr(), K, v6.data_ptr(), 1, stream))
            e1.record(); e1.synchronize()
            if it >= 1:
                ts.append(e0.elapsed_time(e1))
        ms = float(np.mean(ts))
        kernels["K6_quadform_variance"] = {"rows": n6, "ms": ms, "samples_per_s": n6 / (ms * 1e-3),
                                           "TFLOPs_algorithmic": float(n6) * K * K / (ms * 1e-3) / 1e12, "bound": "tensor"}
        del Cm, v6
        # samples read from their text file straight into HBM (csrc/pcq_text.cu): a side measurement of the
        # on-disk format next to the path; it can never break the line
        if world == 1:
            try:
                import tempfile
                from pytuq_b200.utils import io as pio
                with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmpd:
                    ftxt = os.path.join(tmpd, "x.txt")
                    nt = min(n, 1_250_000)
                    pio.savetxt(ftxt, xh[:nt])
                    nbytes = os.path.getsize(ftxt)
                    pio.loadtxt_device(ftxt, device=local)
                    tt = []
                    for _ in range(2):
                        torch.cuda.synchronize()
                        t0 = time.perf_counter()
                        xd = pio.loadtxt_device(ftxt, device=local)
                        torch.cuda.synchronize()
                        tt.append(time.perf_counter() - t0)
                    same = bool(torch.equal(xd, x[:nt]))
                    del xd
                kernels["text_to_hbm"] = {"rows": nt, "cols": D, "text_bytes": nbytes, "ms": 1e3 * min(tt),
                                          "values_per_s": nt * D / min(tt), "text_GBps": nbytes / min(tt) / 1e9,
                                          "bit_identical_to_the_resident_samples": same, "bound": "host-to-device copy of the text",
                                          "api": "pytuq_b200.utils.io.loadtxt_device (pcq_text_open + pcq_text_parse_dev)"}
            except Exception as exc:                     # noqa: BLE001
                kernels["text_to_hbm"] = {"error": repr(exc)}
    peak = peaks.get("fp64_dmma_tflops")
    # DRAM bytes per step of the dominant kernel: NOT a live counter -- the committed ncu capture of this very
    # command (bytes per sample x this rank's samples; the SYRK works chunk by chunk, so its traffic is linear in N)
    traffic, traffic_source = None, None
    for fn in (("r02_q_moment_traffic.json", "r02_p_moment_traffic.json", "r02_n_moment_traffic.json", "r02_moment_traffic.json") if moment_form else ()) + ("r02_k3_traffic.json", "r01_j_k3_traffic.json"):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", fn)))
            per_sample = tj.get("dram_bytes_per_sample_main_kernel",
                                tj["dram_bytes_per_step_main_kernel"] / tj.get("samples_per_step", 1_250_000))
            traffic = per_sample * n
            traffic_source = (f"committed ncu capture profiles/{fn} (dram__bytes_read.sum + dram__bytes_write.sum of the "
                              f"{'pcq_mom_ws_kernel' if 'moment' in fn else 'pcq_syrk_kernel'} launches of one step: {per_sample:.0f} B per sample) x {n} samples; "
                              "not measured in this run")
            break
        except Exception:
            continue
    mp = {}
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    if "K1_evalBases" in kernels:
        hb = mp.get("hbm_gbs", 6650.0)
        kernels["K1_evalBases"]["frac_of_hbm_peak"] = kernels["K1_evalBases"]["GBps"] / hb
        kernels["K1_evalBases"]["hbm_peak_GBps"] = hb
        kernels["K1_evalBases"]["hbm_peak_source"] = "MEASURED_PEAKS.json" if "hbm_gbs" in mp else "fallback"
    if peaks.get("fp64_dfma_tflops"):
        for tag in ("K2_evalPC_exact", "K2_evalPC_fast", "K2s_evalPC_exact_specialised", "K2s_evalPC_fast_specialised"):
            if tag in kernels:
                kernels[tag]["frac_of_fp64_peak"] = kernels[tag]["TFLOPs_algorithmic"] / peaks["fp64_dfma_tflops"]
                if "TFLOPs_executed_min" in kernels[tag]:
                    kernels[tag]["executed_frac_of_fp64_peak"] = kernels[tag]["TFLOPs_executed_min"] / peaks["fp64_dfma_tflops"]

    # ---------------- CPU baseline on a bounded sample (rank 0, N=1 only): the reference itself (oracle/_ref)
    cpu = None
    if world == 1 and not args.no_cpu:
        ref = CpuReference()
        nc = min(args.cpu_samples, n)
        npr = min(args.cpu_probe_samples, n)
        probe = cpu_probe(ref, xh, yh, sorted({nc, npr} - {0}))
        big = probe[str(max(nc, npr))]
        tb, tf = big["basis_s"], big["lstsq_s"]
        cpu = {"value": big["evals_per_s"], "unit": "basis evals/s (samples*terms)", "cores": ref.cores, "kind": ref.kind,
               "sample": ref.describe(max(nc, npr), tb, tf), "blas_threads": ref.blas_threads(),
               "basis_evals_per_s": max(nc, npr) * K / tb, "fit_seconds": tf, "linearity": probe,
               "extrapolated_seconds_at_full_size": n_total * K / big["evals_per_s"],
               "max_abs_coef_diff_vs_gpu_on_the_same_rows": None}
        try:      # same rows through the drop-in: the coefficient parity of the two arms, for the record
            chk = lsq()
            chk.setBasisEvaluator(PCBasisEvaluator(0), (pc,))
            chk.fit(xh[:nc], yh[:nc])
            cf_ref, _, _ = ref.one_pass(xh[:nc], yh[:nc])
            cpu["max_abs_coef_diff_vs_gpu_on_the_same_rows"] = float(np.max(np.abs(chk.cf - cf_ref)))
        except Exception as exc:                         # noqa: BLE001
            cpu["max_abs_coef_diff_vs_gpu_on_the_same_rows"] = repr(exc)

    peak_source = ("FP64 DMMA m8n8k4 register-chain micro-benchmark run live on this GPU "
                   "(pcq_measure_peak); MEASURED_PEAKS.json has no FP64 figure -- the same measurement looped "
                   "for 4 s with a clock record is committed as profiles/r02_fp64_peaks.json (DMMA 36.95 burst / "
                   "36.85 sustained, DFMA 36.49 / 36.42 TFLOP/s at 1965 MHz, no throttle reason)")
    if moment_form:
        roofline = {"bound": "tensor",
                    "kernel": "pcq_mom_ws_kernel<FM,FN> (K3 in moment form, csrc/pcq_moment.cu, warp-specialised: 4 producer warps generate the operands in shared memory, 8 consumer warps issue the DMMAs; the %d moments the Gram entries are linear "
                              "combinations of, as staircase GEMMs on DMMA -- %.0f executed multiply-adds per sample instead of "
                              "K(K+1)/2 + ... = %.0f; %d units of eight warp tasks; per 6 GB table chunk one launch per task "
                              "shape on forked streams, fed by pcq_mom_table_kernel; then fix-up and the K x K reconstruction; "
                              "launch list profiles/r02_q_moment_launches.csv)"
                              % (int(minfo[0]), minfo[2], flops / (2.0 * n), int(minfo[1])),
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if peak else None,
                    "achieved_note": "multiply-adds EXECUTED on the DMMA pipe (planner count, fragment padding included) x 2 / "
                                     "duration of the whole pcq_suffstats_dev call (table + moment kernels + fix-up + "
                                     "reconstruction launches), CUDA events on the launching stream: the pipe utilisation",
                    "executed_flops_per_launch": flops_exec,
                    "algorithmic_minimum_flops": 2.0 * n * minfo[0],
                    "algorithmic_minimum_frac": (2.0 * n * minfo[0] / (k3_mean_ms * 1e-3) / 1e12 / peak) if peak else None,
                    "dense_equivalent": {"flops_per_launch_survey_8d": flops, "TFLOPs": flops / (k3_mean_ms * 1e-3) / 1e12,
                                         "frac": (flops / (k3_mean_ms * 1e-3) / 1e12 / peak) if peak else None,
                                         "note": "the step counted with the dense N K (K+1) flops of SURVEY 8(d): above the DMMA roof "
                                                 "because the moment form needs ~5x fewer multiply-adds for the same statistics; the "
                                                 "dense form itself (pack + SYRK) is kernels.K3_pack_syrk_form"},
                    "traffic": traffic, "traffic_source": traffic_source,
                    "traffic_note": "x, y are read once by the table kernel, which writes %d B per sample of 1-D values; every "
                                    "unit (%d at this shape) streams that table once through L2; algorithmic input bytes = %d"
                                    % (int(minfo[6]), int(minfo[1]), 8 * n * (D + 1)),
                    "peak_source": peak_source}
    else:
        roofline = {"bound": "tensor", "kernel": "pcq_syrk_kernel (K3: DMMA SYRK over the packed design matrix, TMA-fed; "
                                                 "profiles/r02_k3_traffic.json, profiles/r02_launches_bench.csv)",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": (achieved / peak) if peak else None, "traffic": traffic,
                    "traffic_source": traffic_source,
                    "achieved_note": "algorithmic flops of one step / duration of the whole pcq_suffstats_dev call "
                                     "(pack + SYRK + fix-up + mirror launches), CUDA events on the launching stream",
                    "algorithmic_flops_per_launch": flops,
                    "peak_source": peak_source}

    line = {
        "metric": "pc_basis_evals_per_s", "value": value, "unit": "basis evals/s (samples*terms)",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "samples_per_gpu": n,
        "fit_seconds": ms_per_step * 1e-3,
        "breakdown_ms": {"k3_suffstats": k3_mean_ms, "allreduce": float(np.mean(ar_ms)),
                         "solve_kxk": 1e3 * float(np.mean(solve_s))},
        "max_abs_coef_error_vs_truth": err,
        "roofline": roofline,
        "peaks_measured_live": peaks,
        "kernels": kernels,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "basis evals/s (samples*terms)",
                "h2d_bytes_per_step": int(n * (D + 1) * 8), "d2h_bytes_per_step": int(K * 8),
                "api": "pytuq_b200.lreg.lsq.fit(x, y) with numpy host arrays",
                "max_abs_coef_diff_vs_resident": e2e_err},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
