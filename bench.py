#!/usr/bin/env python
"""bench.py -- KPCA fit benchmark (BASELINE.json metric: "KPCA fit s & Gram TFLOP/s, RBF
n=64k d=256 k=16, at 1/2/4/8 B200 vs host CPU").

A "step" is one complete fit of the named configuration: Gram + fused kernel map,
double centring, top-k eigensolve (and the collectives between ranks at N > 1).

    python bench.py --gpus N --steps K --warmup W          # our arm (one rank per GPU via torchrun)
    python bench.py --impl reference --steps K --warmup W   # CPU arm: the oracle restatement of
                                                            # the reference's dense dsyevd path

value  = seconds per fit with X already resident in HBM and V, lambda left in HBM
e2e    = the same through the host-buffer C ABI call (pinned host X in, V / lambda out)
The whole job is ONE problem sharded by row blocks, so scaling is "strong".
"""
import argparse
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
    if args.n:
        cfg["n"] = args.n
    return cfg


def cpu_fit_sample(cfg, n_s, steps=1):
    """Times the oracle's dense reference path (Gram + map, centring, dsyevd 'V','U', reverse,
    scale -- kpca.jl:49-57) on a bounded sample and extrapolates to the full n:
    Gram ~ n^2 d, centring ~ n^2, dsyevd ~ n^3."""
    import oracle as O
    from fxb200.synth import mix
    from threadpoolctl import threadpool_info, threadpool_limits
    # torchrun exports OMP_NUM_THREADS=1: give the CPU arm every host core explicitly
    threadpool_limits(limits=os.cpu_count(), user_api="blas")
    Xs, _ = mix(n_s, cfg["d"], cfg["c"], cfg["sep"], cfg["seed"])
    okern = {"gaussian": O.GaussianKernel, "polynomial": O.PolynomialKernel,
             "sigmoid": O.SigmoidKernel}[cfg["kernel"]](*cfg["params"])
    parts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        K = O.kernelmatrix(okern, Xs)
        t1 = time.perf_counter()
        Kc, _, _ = O.centerkernelmatrix(K)
        t2 = time.perf_counter()
        V, L = O.components_eig(Kc)
        L = L / (n_s - 1)
        t3 = time.perf_counter()
        parts.append((t1 - t0, t2 - t1, t3 - t2))
    g, c, e = [float(np.median([p[i] for p in parts])) for i in range(3)]
    r = cfg["n"] / float(n_s)
    extrap = g * r * r + c * r * r + e * r ** 3
    threads = max([p.get("num_threads", 0) for p in threadpool_info() if p.get("user_api") == "blas"] + [0]) or os.cpu_count()
    return dict(measured_s=g + c + e, gram_s=g, center_s=c, eig_s=e, extrapolated_s=extrap,
                cores=int(threads), n_sample=n_s,
                sample="oracle dense fit (NumPy GEMM+map, centring, LAPACK dsyevd) at n=%d, d=%d; "
                       "extrapolated to n=%d as gram*r^2 + center*r^2 + dsyevd*r^3, r=%g. Julia/BLAS "
                       "thread settings: N/A (no Julia in the image); OpenBLAS threads=%d"
                       % (n_s, cfg["d"], cfg["n"], r, threads))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = load_cfg(args)
    vals = []
    info = None
    for i in range(args.warmup + args.steps):
        info = cpu_fit_sample(cfg, args.cpu_sample_n, 1)
        if i >= args.warmup:
            vals.append(info["extrapolated_s"])
    v = float(np.median(vals)) if vals else info["extrapolated_s"]
    line = {
        "impl": "reference", "metric": "kpca_fit_seconds", "value": v, "unit": "s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": v * 1e3,
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(cfg, args.gpus),
        "cpu_baseline": {"value": v, "unit": "s", "cores": info["cores"], "kind": "port",
                         "sample": info["sample"], "measured_sample_s": info["measured_s"]},
        "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(cfg, gpus):
    return {"workload": "KPCA fit %s n=%d d=%d k=%d (Gram+map, double centring, top-k eigensolve)"
                        % (cfg["kernel"], cfg["n"], cfg["d"], cfg["k"]),
            "kernel": cfg["kernel"], "kernel_params": list(cfg["params"]), "n": cfg["n"], "d": cfg["d"],
            "k": cfg["k"], "generator": "mix(n,d,c=%d,sep=%g,seed=%d)" % (cfg["c"], cfg["sep"], cfg["seed"]),
            "sharding": "row-blocks x%d" % gpus,
            "l2": "inputs_larger_than_L2 (K is %.1f GB per fit)" % (cfg["n"] ** 2 * 8 / 1e9)}


def run_ours(args):
    import torch
    import fxb200
    from fxb200.synth import mix
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ids = [fxb200.Context.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx = fxb200.Context.distributed(local, rank, world, ids[0])
    else:
        ctx = fxb200.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    lib = ctx._lib

    cfg = load_cfg(args)
    n, d, k = cfg["n"], cfg["d"], cfg["k"]
    X, _ = mix(n, d, cfg["c"], cfg["sep"], cfg["seed"])          # column-major float64
    kern = fxb200.from_spec(cfg["kernel"], cfg["params"])
    kid, kp = kern.kernel_id, (C.c_double * 3)(*kern.params())

    # host pinned copy (e2e arm) and device-resident copy (value arm); torch tensors hold the
    # column-major data as a (d, n) row-major array
    Xh = torch.from_numpy(np.ascontiguousarray(X.T)).pin_memory()
    Xd = Xh.cuda(non_blocking=False)
    Vd = torch.empty((k, n), dtype=torch.float64, device="cuda")
    Ld = torch.empty(k, dtype=torch.float64, device="cuda")
    Vh = torch.empty((k, n), dtype=torch.float64).pin_memory()
    Lh = torch.empty(k, dtype=torch.float64).pin_memory()
    st = fxb200.Stats()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def fit_dev():
        rc = lib.fx_kpca_fit_dev_f64(ctx._h, C.c_void_p(Xd.data_ptr()), n, d, n, kid, kp, k, 1, 0.0, 0.0,
                                     None, C.c_void_p(Vd.data_ptr()), n, C.c_void_p(Ld.data_ptr()), None,
                                     C.byref(st))
        fxb200._lib.check(rc, "fx_kpca_fit_dev_f64")

    def fit_host():
        rc = lib.fx_kpca_fit_f64(ctx._h, C.c_void_p(Xh.data_ptr()), n, d, n, kid, kp, k, 1, 0.0, 0.0, None,
                                 C.c_void_p(Vh.data_ptr()), n, C.c_void_p(Lh.data_ptr()), None, C.byref(st))
        fxb200._lib.check(rc, "fx_kpca_fit_f64")

    def timed(fn, steps):
        acc = {}
        launches = 0
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
            launches += st.kernel_launches
            for key in ("gram_ms", "center_ms", "symm_ms", "ortho_ms", "rr_ms", "comm_ms", "eig_ms",
                        "h2d_ms", "d2h_ms", "total_ms"):
                acc[key] = acc.get(key, 0.0) + getattr(st, key)
            acc["symm_launches"] = acc.get("symm_launches", 0) + st.symm_launches
            acc["steps"] = st.steps
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        barrier()
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, acc, launches

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                       # nvidia-smi needs ~1 s to start: launch before warm-up
    for _ in range(args.warmup):
        fit_dev()
    ms_dev, acc, launches = timed(fit_dev, args.steps)
    maxres, conv, block, trace = st.max_residual, st.converged, st.block, st.trace
    peer_access = ctx.peer_access
    lam_dev = Ld.cpu().numpy().copy()

    for _ in range(max(1, min(args.warmup, 2))):
        fit_host()
    ms_e2e, acc_h, launches_h = timed(fit_host, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    h2d, d2h = st.h2d_bytes, st.d2h_bytes
    lam_host = Lh.numpy().copy()

    # live roofline denominators on this GPU: fp64 DMMA issue peak and stream copy
    # per-rank Gram / collective time (ranks wait for the slowest one inside comm_ms)
    per_rank = None
    if dist is not None:
        t = torch.tensor([acc["gram_ms"] / args.steps, acc["comm_ms"] / args.steps], dtype=torch.float64, device="cuda")
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        per_rank = {"gram_ms": [round(float(o[0]), 3) for o in out], "comm_ms": [round(float(o[1]), 3) for o in out]}
    peak_dmma = peak_copy = None
    cpu = None
    if rank == 0:
        ms = fxb200.bench_kernel(4, reps=5, ctx=ctx)
        peak_dmma = ctx.sm_count * 8 * 8 * 20000 * 8 * 512.0 / float(np.min(ms)) / 1e9
        ms = fxb200.bench_kernel(6, n=1 << 28, reps=5, ctx=ctx)
        peak_copy = 2 * (1 << 28) * 8 / float(np.min(ms)) / 1e6
    barrier()

    if rank == 0:
        peaks_file = {}
        try:
            peaks_file = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks_file.get("hbm_gbs", 6650.0)
        hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks_file else "fallback 6.65 TB/s"
        K = args.steps
        ld = fxb200.sharding.row_partition(n, world)[0]
        gram_ms = acc["gram_ms"] / K
        center_ms = acc["center_ms"] / K
        symm_ms = acc["symm_ms"] / max(1, acc["symm_launches"])
        gram_flops = 2.0 * ld * n * d                        # algorithmic, per rank, per launch
        # The Gram kernel uses the symmetry of K inside each rank's square diagonal block: only the
        # 128x64 tiles on / right of the diagonal are computed there, the rest is written by mirror stores.  The roofline uses the
        # EXECUTED flops (so frac <= 1); gram_tflops keeps the algorithmic 2 n^2 d ("no symmetry credit").
        # With several ranks and NVLink peer mappings the off-diagonal blocks are shared out as well.
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
            "gpu_launches": int(launches + launches_h),
            "gram_tflops": gram_tf_alg * world,
            "gram_executed_fraction": exec_frac,
            "peer_access": bool(peer_access),
            "roofline": {"bound": "tensor", "kernel": "gram_dmma_kernel", "achieved": gram_tf,
                         "peak": peak_dmma, "unit": "TFLOP/s", "frac": gram_tf / peak_dmma,
                         "traffic": traffic,
                         "peak_source": "fp64 DMMA.8x8x4 issue-rate microbenchmark measured live in this run "
                                        "(fx_bench_kernel 4); MEASURED_PEAKS.json has no fp64 entry",
                         "algorithmic": "2*rows_per_rank*n*d flops per launch = %.4g; executed = %.4g of it "
                                        "(symmetric tiles mirrored, across ranks over NVLink peer stores when "
                                        "peer_access)" % (gram_flops, exec_frac)},
            "roofline_secondary": [
                {"kernel": "symm_dmma_kernel<CENTER> (centring fused with the first K*V pass)", "bound": "hbm",
                 "achieved": 16.0 * ld * n / center_ms / 1e6,
                 "peak": hbm_peak, "unit": "GB/s", "frac": 16.0 * ld * n / center_ms / 1e6 / hbm_peak,
                 "frac_of_nominal_8TBs": 16.0 * ld * n / center_ms / 1e6 / 8000.0,
                 "peak_source": hbm_src, "copy_GBs_live": peak_copy,
                 "note": "one read + one write of K; center_ms also covers the row-sum partial reduce and the "
                         "means kernel; the K*V product of this pass is free (HBM-bound)"},
                {"kernel": "symm_dmma_kernel", "bound": "hbm", "achieved": 8.0 * ld * n / symm_ms / 1e6,
                 "peak": hbm_peak, "unit": "GB/s", "frac": 8.0 * ld * n / symm_ms / 1e6 / hbm_peak,
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
                       "lambda": lam_dev.tolist(), "trace": float(trace)},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:       # CPU baseline: rank 0 at N = 1 only
            cpu = cpu_fit_sample(cfg, args.cpu_sample_n, 1)
            line["cpu_baseline"] = {"value": cpu["extrapolated_s"], "unit": "s", "cores": cpu["cores"],
                                    "kind": "port", "sample": cpu["sample"],
                                    "measured_sample_s": cpu["measured_s"]}
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
