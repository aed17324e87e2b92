#!/usr/bin/env python
"""bench.py -- KPCA fit benchmark (BASELINE.json metric: "KPCA fit s & Gram TFLOP/s, RBF
n=64k d=256 k=16, at 1/2/4/8 B200 vs host CPU").

A "step" is one complete fit of the named configuration: Gram + fused kernel map,
double centring, top-k eigensolve (and the collectives between ranks at N > 1).

    python bench.py --gpus N --steps K --warmup W          # our arm (one rank per GPU via torchrun)
    python bench.py --impl reference --steps K --warmup W   # CPU arm: the oracle restatement of
                                                            # the reference's dense dsyevd path
    python bench.py --config C3|C4|C5 --gpus N              # the other BASELINE.json configurations

value  = seconds per fit with X already resident in HBM and V, lambda left in HBM
e2e    = the same through the host-buffer C ABI call (pinned host X in, V / lambda out)
The whole job is ONE problem sharded by row blocks, so scaling is "strong".

What is measured and what is modelled is kept apart in the JSON line:
  cpu_baseline / --impl reference   the dense CPU path TIMED at config.n_measured (default 4096) and
                                    EXTRAPOLATED to the full n (the dense path needs 137 GB and hours
                                    at n = 65,536): labelled as such in `config.workload` and `sample`
  measured_pair                     GPU and CPU BOTH measured on the same input at --pair-n (8192):
                                    a ratio in which neither side is a model
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="HEADLINE")
    ap.add_argument("--n", type=int, default=0, help="override n (debug)")
    ap.add_argument("--cpu-sample-n", type=int, default=4096)
    ap.add_argument("--pair-n", type=int, default=8192,
                    help="size of the measured GPU-vs-CPU pair (0 = skip; 16384 = BASELINE configs[1] size)")
    ap.add_argument("--transform-m", type=int, default=262144, help="queries of the transform block (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] if power else sm
        return {"sm_mhz": float(np.median(busy)) if busy else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "power_w_max": max(power) if power else None, "samples": len(sm)}


def load_cfg(args):
    from fxb200.synth import CONFIGS
    cfg = dict(CONFIGS[args.config])
    cfg["name"] = args.config
    if args.n:
        cfg["n"] = args.n
    return cfg


def okernel(cfg):
    import oracle as O
    return {"gaussian": O.GaussianKernel, "polynomial": O.PolynomialKernel,
            "sigmoid": O.SigmoidKernel}[cfg["kernel"]](*cfg["params"])


def blas_threads():
    from threadpoolctl import threadpool_info, threadpool_limits
    # torchrun exports OMP_NUM_THREADS=1: give the CPU arm every host core explicitly
    threadpool_limits(limits=os.cpu_count(), user_api="blas")
    return max([p.get("num_threads", 0) for p in threadpool_info() if p.get("user_api") == "blas"] + [0]) \
        or os.cpu_count()


def cpu_dense_fit(cfg, n_s):
    """One fit of the oracle's dense reference path (Gram + map, centring, dsyevd 'V','U', reverse,
    scale -- kpca.jl:49-57) at n_s observations of the configuration's generator: (gram, centre,
    eig seconds, leading eigenvalues)."""
    import oracle as O
    from fxb200.synth import mix
    Xs, _ = mix(n_s, cfg["d"], cfg["c"], cfg["sep"], cfg["seed"])
    kern = okernel(cfg)
    t0 = time.perf_counter()
    K = O.kernelmatrix(kern, Xs)
    t1 = time.perf_counter()
    Kc, _, _ = O.centerkernelmatrix(K)
    del K
    t2 = time.perf_counter()
    V, L = O.components_eig(Kc)
    L = L / (n_s - 1)
    t3 = time.perf_counter()
    return (t1 - t0, t2 - t1, t3 - t2), L[:cfg["k"]].copy(), V[:, :cfg["k"]].copy(), Kc


def cpu_fit_sample(cfg, n_s, steps=1):
    """Times the dense CPU path on a bounded sample (n_s rows) and extrapolates to the full n:
    Gram ~ n^2 d, centring ~ n^2, dsyevd ~ n^3.  The extrapolation is a MODEL and is labelled so."""
    threads = blas_threads()
    parts = [cpu_dense_fit(cfg, n_s)[0] for _ in range(steps)]
    g, c, e = [float(np.median([p[i] for p in parts])) for i in range(3)]
    r = cfg["n"] / float(n_s)
    extrap = g * r * r + c * r * r + e * r ** 3
    return dict(measured_s=g + c + e, gram_s=g, center_s=c, eig_s=e, extrapolated_s=extrap,
                cores=int(threads), n_sample=n_s,
                sample="oracle dense fit (NumPy GEMM+map, centring, LAPACK dsyevd) MEASURED at n=%d, d=%d "
                       "(%.2f s); value EXTRAPOLATED to n=%d as gram*r^2 + center*r^2 + dsyevd*r^3, r=%g "
                       "(the dense path needs ~137 GB and hours at the full size). Julia/BLAS thread "
                       "settings: N/A (no Julia in the image); OpenBLAS threads=%d"
                       % (n_s, cfg["d"], g + c + e, cfg["n"], r, threads))


def workload_config(cfg, gpus, reference_n=None):
    w = {"workload": "KPCA fit %s n=%d d=%d k=%d (Gram+map, double centring, top-k eigensolve)"
                     % (cfg["kernel"], cfg["n"], cfg["d"], cfg["k"]),
         "kernel": cfg["kernel"], "kernel_params": list(cfg["params"]), "n": cfg["n"], "d": cfg["d"],
         "k": cfg["k"], "generator": "mix(n,d,c=%d,sep=%g,seed=%d)" % (cfg["c"], cfg["sep"], cfg["seed"]),
         "sharding": "row-blocks x%d" % gpus,
         "l2": "inputs_larger_than_L2 (K is %.1f GB per fit)" % (cfg["n"] ** 2 * 8 / 1e9)}
    if reference_n is not None:
        w["n_measured"] = int(reference_n)
        w["workload"] += (" -- CPU dense dsyevd path MEASURED at n=%d and EXTRAPOLATED from n=%d to n=%d "
                          "(not a measurement at the full size)" % (reference_n, reference_n, cfg["n"]))
    return w


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = load_cfg(args)
    vals, meas = [], []
    info = None
    for i in range(args.warmup + args.steps):
        info = cpu_fit_sample(cfg, args.cpu_sample_n, 1)
        if i >= args.warmup:
            vals.append(info["extrapolated_s"])
            meas.append(info["measured_s"])
    v = float(np.median(vals)) if vals else info["extrapolated_s"]
    line = {
        "impl": "reference", "metric": "kpca_fit_seconds", "value": v, "unit": "s",
        "value_is": "extrapolated from n_measured=%d (see config.workload / cpu_baseline.sample)" % args.cpu_sample_n,
        "measured_value": {"n": args.cpu_sample_n, "seconds": float(np.median(meas)) if meas else info["measured_s"]},
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": v * 1e3,
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(cfg, args.gpus, reference_n=args.cpu_sample_n),
        "cpu_baseline": {"value": v, "unit": "s", "cores": info["cores"], "kind": "port",
                         "sample": info["sample"], "measured_sample_s": info["measured_s"],
                         "n_measured": args.cpu_sample_n, "extrapolated": True},
        "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


class Harness:
    """torch plumbing shared by the configurations: device, process group, library context, timing."""

    def __init__(self):
        import torch
        import fxb200
        self.torch, self.fx = torch, fxb200
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU fallback "
                             "(use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            self.dist = dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            ids = [fxb200.Context.unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            self.ctx = fxb200.Context.distributed(self.local, self.rank, self.world, ids[0])
        else:
            self.ctx = fxb200.Context(self.local)
        self.stream = torch.cuda.current_stream()
        self.ctx.set_stream(self.stream.cuda_stream)
        self.lib = self.ctx._lib

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.dist is None:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps):
        """barrier + sync, `steps` calls bracketed by CUDA events on the launching stream, sync, MAX
        over ranks.  Returns total ms."""
        torch = self.torch
        self.barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(steps):
            fn()
        e1.record(self.stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        self.barrier()
        return self.max_over_ranks(ms)

    def close(self):
        self.ctx.close()
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def live_peaks(h):
    """fp64 DMMA issue peak (TFLOP/s) and stream copy (GB/s) measured on this GPU in this run.
    The register-only DMMA chain runs 10 x 0.2 s back to back: the best repetition is the BURST peak,
    the median of the second half the SUSTAINED one (what the GPU holds under its power cap; the
    Gram kernel is timed inside a long step, so its roofline uses the sustained figure --
    B200_PROFILING.md)."""
    ms = np.asarray(h.fx.bench_kernel(4, reps=10, ctx=h.ctx))
    flops = h.ctx.sm_count * 8 * 8 * 20000 * 8 * 512.0
    peak_burst = flops / float(ms.min()) / 1e9
    peak_sustained = flops / float(np.median(ms[5:])) / 1e9
    ms = h.fx.bench_kernel(6, n=1 << 28, reps=5, ctx=h.ctx)
    peak_copy = 2 * (1 << 28) * 8 / float(np.min(ms)) / 1e6
    return peak_sustained, peak_burst, peak_copy


def hbm_peak():
    try:
        pf = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        if "hbm_gbs" in pf:
            return pf["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def fixture_parity(cfg, lam, trace):
    """Headline only: eigenvalues against the committed CPU ARPACK result on the same input
    (tests/golden/large_headline_eigsh.npz, minted by tests/golden/make_large_golden.py)."""
    name = {"HEADLINE": "large_headline_eigsh.npz", "C2": "large_c2_dsyevd.npz"}.get(cfg["name"])
    if name is None or cfg["n"] != __import__("fxb200").synth.CONFIGS[cfg["name"]]["n"]:
        return None
    try:
        g = np.load(os.path.join(ROOT, "tests", "golden", name))
    except Exception:
        return None
    k = len(lam)
    return {"fixture": "tests/golden/" + name, "how": str(g["how"]),
            "lambda_max_rel": float(np.abs(np.asarray(lam) / g["L"][:k] - 1).max()),
            "trace_rel": float(abs(trace / float(g["trace"]) - 1))}


def run_kpca(args):
    h = Harness()
    torch, fxb200, lib, ctx = h.torch, h.fx, h.lib, h.ctx
    from fxb200.synth import mix
    world, rank = h.world, h.rank

    cfg = load_cfg(args)
    n, d, k = cfg["n"], cfg["d"], cfg["k"]
    X, _ = mix(n, d, cfg["c"], cfg["sep"], cfg["seed"])          # column-major float64
    kern = fxb200.from_spec(cfg["kernel"], cfg["params"])
    kid, kp = kern.kernel_id, (C.c_double * 3)(*kern.params())

    # host pinned copy (e2e arm), pageable copy (what a Julia Matrix{Float64} is) and device-resident
    # copy (value arm); torch tensors hold the column-major data as a (d, n) row-major array
    Xpage = torch.from_numpy(np.ascontiguousarray(X.T))
    Xh = Xpage.pin_memory()
    Xd = Xh.cuda(non_blocking=False)
    Vd = torch.empty((k, n), dtype=torch.float64, device="cuda")
    Ld = torch.empty(k, dtype=torch.float64, device="cuda")
    Vh = torch.empty((k, n), dtype=torch.float64).pin_memory()
    Lh = torch.empty(k, dtype=torch.float64).pin_memory()
    Vpage = torch.empty((k, n), dtype=torch.float64)
    Lpage = torch.empty(k, dtype=torch.float64)
    st = fxb200.Stats()
    model_h = C.c_void_p()

    def fit_dev(model=None):
        rc = lib.fx_kpca_fit_dev_f64(ctx._h, C.c_void_p(Xd.data_ptr()), n, d, n, kid, kp, k, 1, 0.0, 0.0,
                                     None, C.c_void_p(Vd.data_ptr()), n, C.c_void_p(Ld.data_ptr()),
                                     C.byref(model) if model is not None else None, C.byref(st))
        fxb200._lib.check(rc, "fx_kpca_fit_dev_f64")

    def fit_host(Xt=Xh, Vt=Vh, Lt=Lh):
        rc = lib.fx_kpca_fit_f64(ctx._h, C.c_void_p(Xt.data_ptr()), n, d, n, kid, kp, k, 1, 0.0, 0.0, None,
                                 C.c_void_p(Vt.data_ptr()), n, C.c_void_p(Lt.data_ptr()), None, C.byref(st))
        fxb200._lib.check(rc, "fx_kpca_fit_f64")

    keys = ("gram_ms", "center_ms", "symm_ms", "ortho_ms", "rr_ms", "comm_ms", "eig_ms", "h2d_ms", "d2h_ms",
            "total_ms")

    def timed(fn, steps):
        acc = {"launches": 0, "symm_launches": 0}

        def one():
            fn()
            acc["launches"] += st.kernel_launches
            acc["symm_launches"] += st.symm_launches
            acc["steps"] = st.steps
            for key in keys:
                acc[key] = acc.get(key, 0.0) + getattr(st, key)
        ms = h.timed(one, steps)
        return ms, acc

    sampler = ClockSampler(h.local)
    if rank == 0:
        sampler.start()                       # nvidia-smi needs ~1 s to start: launch before warm-up
    # the very first call pays the ~40 GB workspace cudaMalloc, module load and peer mapping
    h.barrier()
    t0 = time.perf_counter()
    fit_dev()
    torch.cuda.synchronize()
    cold_ms = h.max_over_ranks((time.perf_counter() - t0) * 1e3)
    for _ in range(max(0, args.warmup - 1)):
        fit_dev()
    ms_dev, acc = timed(fit_dev, args.steps)
    maxres, conv, block, trace = st.max_residual, st.converged, st.block, st.trace
    peer_access = ctx.peer_access
    lam_dev = Ld.cpu().numpy().copy()

    for _ in range(max(1, min(args.warmup, 2))):
        fit_host()
    ms_e2e, acc_h = timed(fit_host, args.steps)
    h2d, d2h = st.h2d_bytes, st.d2h_bytes
    lam_host = Lh.numpy().copy()
    fit_host(Xpage, Vpage, Lpage)
    ms_page, _ = timed(lambda: fit_host(Xpage, Vpage, Lpage), max(1, min(args.steps, 3)))
    ms_page /= max(1, min(args.steps, 3))
    clocks = sampler.stop() if rank == 0 else None

    # ---- out-of-sample transform of the fitted model, queries sharded over the ranks -------------
    transform = None
    if args.transform_m > 0:
        m_tot = args.transform_m
        q0, mloc = fxb200.sharding.query_partition(m_tot, world)[rank]
        fit_dev(model_h)
        gen = torch.Generator(device="cuda"); gen.manual_seed(1234 + rank)
        reps_needed = (mloc + n - 1) // n
        Zd = (Xd.repeat(1, reps_needed)[:, :mloc] +
              0.05 * torch.randn((d, mloc), dtype=torch.float64, device="cuda", generator=gen)).contiguous()
        Td = torch.empty((k, mloc), dtype=torch.float64, device="cuda")
        tst = fxb200.Stats()

        def tr():
            rc = lib.fx_kpca_transform_dev_f64(ctx._h, model_h, C.c_void_p(Zd.data_ptr()), mloc, mloc,
                                               C.c_void_p(Td.data_ptr()), mloc, C.byref(tst))
            fxb200._lib.check(rc, "fx_kpca_transform_dev_f64")
        tr()
        ms_tr = h.timed(tr, 3) / 3
        kern_ms = h.max_over_ranks(tst.transform_ms)
        flops = 2.0 * m_tot * n * (d + k)
        # spot check, bias-free: T(z_a) - T(z_b) = (k(z_a, X) - k(z_b, X)) V against the CPU oracle kernel
        spot = None
        if rank == 0:
            import oracle as O
            qs = np.arange(0, min(mloc, 4096), 64)[:32]
            Zs = Zd[:, torch.as_tensor(qs, device="cuda")].cpu().numpy().T
            Ts = Td[:, torch.as_tensor(qs, device="cuda")].cpu().numpy().T
            Vh_ = Vd.cpu().numpy().T
            Kq = O.kernelmatrix(okernel(cfg), Zs, np.ascontiguousarray(X))
            ref = Kq @ Vh_
            dref, dgpu = ref - ref[:1], Ts - Ts[:1]
            spot = float(np.abs(dgpu - dref).max() / np.abs(ref).max())
        transform = {"queries": m_tot, "per_rank": mloc, "ms": ms_tr, "kernel_ms": kern_ms,
                     "tflops": flops / ms_tr / 1e9, "kernel_tflops": flops / kern_ms / 1e9,
                     "algorithmic_flops": "2*m*n*(d+k) = %.4g (+ m*n = %.3g kernel-map evaluations)" % (flops, float(m_tot) * n),
                     "api": "fx_kpca_transform_dev_f64 (device-resident queries, query-sharded over ranks)",
                     "oracle_spot_check_rel": spot}
        lib.fx_model_free(model_h)

    per_rank = None
    if h.dist is not None:
        t = torch.tensor([acc["gram_ms"] / args.steps, acc["comm_ms"] / args.steps], dtype=torch.float64, device="cuda")
        out = [torch.zeros_like(t) for _ in range(world)]
        h.dist.all_gather(out, t)
        per_rank = {"gram_ms": [round(float(o[0]), 3) for o in out], "comm_ms": [round(float(o[1]), 3) for o in out]}
    peak_dmma = peak_burst = peak_copy = None
    if rank == 0:
        peak_dmma, peak_burst, peak_copy = live_peaks(h)
    h.barrier()

    # ---- measured pair: GPU and CPU dense path on the SAME input at a size the CPU finishes -------
    pair = None
    if rank == 0 and world == 1 and args.pair_n > 0 and not args.no_cpu_baseline:
        pair = measured_pair(h, cfg, args.pair_n)

    if rank == 0:
        hbm, hbm_src = hbm_peak()
        K = args.steps
        ld = fxb200.sharding.row_partition(n, world)[0]
        gram_ms = acc["gram_ms"] / K
        center_ms = acc["center_ms"] / K
        symm_ms = acc["symm_ms"] / max(1, acc["symm_launches"])
        gram_flops = 2.0 * ld * n * d                        # algorithmic, per rank, per launch
        # The Gram kernel uses the symmetry of K: tiles left of the diagonal are produced by mirror
        # stores (across ranks over NVLink peer mappings when peer_access).  The roofline uses the
        # EXECUTED flops (so frac <= 1); gram_tflops keeps the algorithmic 2 n^2 d ("no symmetry credit").
        t_exec, t_all = fxb200.sharding.gram_tiles_executed(n, world, 0, peer_access)
        exec_frac = t_exec / t_all
        gram_tf_alg = gram_flops / gram_ms / 1e9
        gram_tf = exec_frac * gram_tf_alg
        traffic = None                                       # DRAM bytes per Gram launch from the committed
        try:                                                 # ncu --set full capture of the same shape
            tj = json.load(open(os.path.join(ROOT, "profiles", "gram_traffic.json")))
            key = "n%d_d%d_rows%d" % (n, d, ld)
            if key in tj:
                traffic = tj[key]
        except Exception:
            pass
        line = {
            "metric": "kpca_fit_seconds", "value": ms_dev / K / 1e3, "unit": "s", "n_gpus": world,
            "steps": K, "warmup": args.warmup, "ms_per_step": ms_dev / K, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(cfg, world),
            "e2e": {"value": ms_e2e / K / 1e3, "unit": "s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "api": "fx_kpca_fit_f64 (host pinned X in, V/lambda out)"},
            "e2e_pageable": {"value": ms_page / 1e3, "unit": "s",
                             "note": "same call with ordinary (unpinned) host arrays, as a Julia Matrix{Float64} is"},
            "cold_first_call_ms": cold_ms,
            "cold_first_call_note": "first fit of the process: workspace cudaMalloc (grow-only arena, reused by "
                                    "later calls), lazy module load, peer mapping; excluded from value / e2e",
            "gpu_launches": int(acc["launches"] + acc_h["launches"]),
            "gram_tflops": gram_tf_alg * world,
            "gram_executed_fraction": exec_frac,
            "peer_access": bool(peer_access),
            "roofline": {"bound": "tensor", "kernel": "gram_dmma_kernel", "achieved": gram_tf,
                         "peak": peak_dmma, "unit": "TFLOP/s", "frac": gram_tf / peak_dmma,
                         "peak_burst": peak_burst, "frac_of_burst_peak": gram_tf / peak_burst,
                         "traffic": traffic,
                         "peak_source": "fp64 DMMA.8x8x4 issue-rate microbenchmark measured live in this run "
                                        "(fx_bench_kernel 4, register-only chains): sustained = median of the "
                                        "second half of 10 x 0.2 s back to back, burst = best repetition; "
                                        "MEASURED_PEAKS.json has no fp64 entry",
                         "algorithmic": "2*rows_per_rank*n*d flops per launch = %.4g; executed = %.4g of it "
                                        "(symmetric tiles mirrored, across ranks over NVLink peer stores when "
                                        "peer_access)" % (gram_flops, exec_frac)},
            "roofline_secondary": [
                {"kernel": "symm_dmma_kernel<CENTER> (centring fused with the first K*V pass)", "bound": "hbm",
                 "achieved": 16.0 * ld * n / center_ms / 1e6,
                 "peak": hbm, "unit": "GB/s", "frac": 16.0 * ld * n / center_ms / 1e6 / hbm,
                 "frac_of_nominal_8TBs": 16.0 * ld * n / center_ms / 1e6 / 8000.0,
                 "peak_source": hbm_src, "copy_GBs_live": peak_copy,
                 "note": "one read + one write of K; center_ms also covers the row-sum partial reduce and the "
                         "means kernel; the K*V product of this pass is free (HBM-bound)"},
                {"kernel": "symm_dmma_kernel", "bound": "hbm", "achieved": 8.0 * ld * n / symm_ms / 1e6,
                 "peak": hbm, "unit": "GB/s", "frac": 8.0 * ld * n / symm_ms / 1e6 / hbm,
                 "frac_of_nominal_8TBs": 8.0 * ld * n / symm_ms / 1e6 / 8000.0,
                 "tflops": 2.0 * ld * n * block / symm_ms / 1e9, "block": int(block),
                 "launches_per_fit": acc["symm_launches"] / K}],
            "phases_ms_per_fit": {k2: acc[k2] / K for k2 in ("gram_ms", "center_ms", "symm_ms", "ortho_ms",
                                                             "rr_ms", "comm_ms", "eig_ms")},
            "e2e_phases_ms_per_fit": {k2: acc_h[k2] / K for k2 in ("h2d_ms", "d2h_ms", "gram_ms", "eig_ms")},
            "per_rank_ms_per_fit": per_rank,
            "solver": {"kv_passes": int(acc["steps"]), "block": int(block), "converged": int(conv),
                       "max_rel_residual": float(maxres)},
            "parity": {"host_vs_device_lambda_max_rel_diff": float(np.abs(lam_host / lam_dev - 1).max()),
                       "vs_cpu_fixture": fixture_parity(cfg, lam_dev, trace),
                       "lambda": lam_dev.tolist(), "trace": float(trace)},
            "clocks": clocks,
        }
        if line["parity"]["vs_cpu_fixture"]:
            line["parity"]["vs_cpu_eigsh_max_rel"] = line["parity"]["vs_cpu_fixture"]["lambda_max_rel"]
        if transform:
            transform["frac_of_dmma_peak"] = transform["kernel_tflops"] / world / peak_dmma
            line["transform"] = transform
        if pair:
            line["measured_pair"] = pair
        if not args.no_cpu_baseline and world == 1:       # CPU baseline: rank 0 at N = 1 only
            cpu = cpu_fit_sample(cfg, args.cpu_sample_n, 1)
            line["cpu_baseline"] = {"value": cpu["extrapolated_s"], "unit": "s", "cores": cpu["cores"],
                                    "kind": "port", "sample": cpu["sample"],
                                    "measured_sample_s": cpu["measured_s"], "n_measured": cpu["n_sample"],
                                    "extrapolated": True}
        print(json.dumps(line))
    h.close()


def measured_pair(h, cfg, n_p):
    """GPU fit (host buffers in and out, fx_kpca_fit_f64) and the CPU dense path (oracle: dsyevd) both
    TIMED on the same n_p x d input of the configuration's generator, plus a CPU iterative top-k
    (ARPACK eigsh on the CPU-built centred matrix) as the fairer CPU baseline.  No extrapolation."""
    fxb200 = h.fx
    from fxb200.synth import mix
    from scipy.sparse.linalg import eigsh
    threads = blas_threads()
    k = cfg["k"]
    Xp, _ = mix(n_p, cfg["d"], cfg["c"], cfg["sep"], cfg["seed"])
    kern = fxb200.from_spec(cfg["kernel"], cfg["params"])
    fxb200.kpca(Xp, kern, k=k, ctx=h.ctx)                       # warm-up (workspace exists already)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        V, L = fxb200.kpca(Xp, kern, k=k, ctx=h.ctx)
        ts.append(time.perf_counter() - t0)
    gpu_s = float(np.median(ts))
    (g, c, e), Lo, Vo, Kc = cpu_dense_fit(cfg, n_p)
    t0 = time.perf_counter()
    w, _ = eigsh(Kc, k=k, which="LA", tol=1e-12)
    eigsh_s = time.perf_counter() - t0
    Qa, _ = np.linalg.qr(V)
    ang = float(np.linalg.norm(Vo - Qa @ (Qa.T @ Vo), 2))
    cpu_s = g + c + e
    return {"n": n_p, "d": cfg["d"], "k": k, "gpu_e2e_s": gpu_s, "cpu_dense_s": cpu_s,
            "cpu_dense_parts_s": {"gram_map": g, "centre": c, "dsyevd": e},
            "cpu_topk_eigsh_s": g + c + eigsh_s, "cpu_cores": int(threads),
            "speedup_vs_cpu_dense": cpu_s / gpu_s, "speedup_vs_cpu_topk_eigsh": (g + c + eigsh_s) / gpu_s,
            "lambda_max_rel_vs_dsyevd": float(np.abs(L / Lo - 1).max()), "subspace_angle_vs_dsyevd": ang,
            "note": "both sides measured (wall clock, host buffers in and out) on the same input; "
                    "GPU = fxb200.kpca through fx_kpca_fit_f64, CPU = oracle dense path (LAPACK dsyevd, all n pairs)"}


def run_klda(args):
    """BASELINE.json configs[3]: KLDA RBF, 10 classes, n = 32,768, d = 512 (fx_klda_fit_f64, host API)."""
    h = Harness()
    torch, fxb200, ctx = h.torch, h.fx, h.ctx
    from fxb200.synth import mix
    cfg = load_cfg(args)
    n, d, c = cfg["n"], cfg["d"], cfg["c"]
    X, y0 = mix(n, d, c, cfg["sep"], cfg["seed"])
    y = y0 + 1
    freq = np.bincount(y0, minlength=c) / n
    kern = fxb200.from_spec(cfg["kernel"], cfg["params"])
    res = {}

    def fit():
        res["A"], res["L"], res["st"] = fxb200.klda(X, kern, y, c, freq, eps_w=-1e-6, ctx=ctx, return_stats=True)
    sampler = ClockSampler(h.local)
    if h.rank == 0:
        sampler.start()
    for _ in range(max(1, min(args.warmup, 2))):
        fit()
    dev_ms = []

    def one():
        fit()
        dev_ms.append(res["st"]["total_ms"])
    steps = max(1, min(args.steps, 3))
    h.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize()
    wall_ms = h.max_over_ranks((time.perf_counter() - t0) * 1e3 / steps)
    dev = h.max_over_ranks(float(np.mean(dev_ms)))
    clocks = sampler.stop() if h.rank == 0 else None
    st = res["st"]
    if h.rank == 0:
        ld = fxb200.sharding.row_partition(n, h.world)[0]
        hbm, hbm_src = hbm_peak()
        passes = st["symm_launches"]
        line = {"metric": "klda_fit_seconds", "value": dev / 1e3, "unit": "s", "n_gpus": h.world, "steps": steps,
                "warmup": args.warmup, "ms_per_step": dev, "higher_is_better": False, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "KLDA fit gaussian n=%d d=%d classes=%d eps_w=1e-6*tr(Sigma_w)/n "
                                       "(Gram+map, class scatter operators, generalized eigenproblem)" % (n, d, c),
                           "n": n, "d": d, "classes": c, "sharding": "row-blocks x%d" % h.world,
                           "generator": "mix(n,d,c=%d,sep=%g,seed=%d)" % (c, cfg["sep"], cfg["seed"]),
                           "l2": "inputs_larger_than_L2 (K is %.1f GB)" % (n * n * 8 / 1e9)},
                "e2e": {"value": wall_ms / 1e3, "unit": "s", "h2d_bytes_per_step": int(st["h2d_bytes"]),
                        "d2h_bytes_per_step": int(st["d2h_bytes"]), "api": "fx_klda_fit_f64 (host buffers)"},
                "gpu_launches": int(st["kernel_launches"]) * steps,
                "roofline": {"bound": "hbm", "kernel": "symm_dmma_kernel (K*V passes of the Sigma_w solves)",
                             "achieved": 8.0 * ld * n * passes / st["symm_ms"] / 1e6, "peak": hbm, "unit": "GB/s",
                             "frac": 8.0 * ld * n * passes / st["symm_ms"] / 1e6 / hbm, "traffic": None,
                             "peak_source": hbm_src, "passes": int(passes)},
                "phases_ms_per_fit": {k2: st[k2] for k2 in ("gram_ms", "symm_ms", "ortho_ms", "rr_ms", "comm_ms", "eig_ms")},
                "solver": {"iterations": int(st["cg_iters"]), "kv_passes": int(passes), "converged": int(st["converged"]),
                           "max_rel_residual": float(st["max_residual"]), "dense_path": int(st["dense_path"])},
                "parity": {"lambda": res["L"].tolist(), "trace_sigma_w": float(st["trace"])},
                "clocks": clocks}
        print(json.dumps(line))
    h.close()


def run_c5(args):
    """BASELINE.json configs[4]: RBF fit at n = 131,072, d = 64, k = 32 (K = 137 GB, row-sharded) and the
    out-of-sample transform of 1M queries, query-sharded (kpca.jl:64-68)."""
    if args.transform_m == 262144:
        args.transform_m = 1000000
    run_kpca(args)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "C4":
        run_klda(args)
    elif args.config == "C5":
        run_c5(args)
    else:
        run_kpca(args)


if __name__ == "__main__":
    main()
