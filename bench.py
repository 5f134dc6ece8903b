#!/usr/bin/env python
"""bench.py -- REML log-likelihood evaluations per second on the 5k-interval single-block workload
(BASELINE.json configs[1]) and the other hot-path workloads, on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W [--workload all|s5|cl|krige] [--impl ours|reference]

A step (default workload `s5`) is one forward-difference gradient of the REML objective as the optimiser asks for
it (gradnllIAK3D, fitIAK3D.R:1865-1886): npar + 1 likelihood evaluations -- covariance build + Cholesky + solves
each -- sent to the GPU as one batch.  `value` times the library calls with the dataset resident in HBM; `e2e`
goes through the reference-facing call with host buffers (W re-uploaded from host memory each step, results read
back, REML tail on the host).  Single-block REML does not shard ("replicas only"), so --gpus N runs N
independent replicas and reports weak scaling.  The two workloads that DO shard -- the 200k-interval composite likelihood
(configs[3]) and block kriging of the 1M-cell grid (configs[4]) -- run after it in the same call (default `--workload all`)
and are reported as the sub-records `cl200` and `p1m` of the same JSON line: value, ms_per_step, roofline, e2e, per-rank
unit counts and the time of the all-reduce, at whatever N the call was launched with (strong scaling: total work fixed).
`--workload cl|krige|s5` runs one of them alone.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PARBNDS = dict(axBnds=(0.05, 60.0), nuxBnds=(0.05, 20.0), adBnds=(0.02, 1.2), tau2Bnds=(0.01, 20.0),
               sxd1Bnds=(1e-8, 1.0 - 1e-8), maxd=2.0)
METRIC = "REML loglik evals/sec at n obs"


# ---------------------------------------------------------------------------------------------
# clocks during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    _REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                try:
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self._REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2.0)
        if not self.samples:
            try:    # one-shot fallback through the CLI
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                a, b = [float(x) for x in out.strip().split(",")[:2]]
                self.samples, self.max_mhz = [a], b
            except Exception:
                pass
        med = float(np.median(self.samples)) if self.samples else None
        return dict(sm_mhz=med, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(self.samples))


# ---------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------
def s5_inputs(n):
    from iak3d_b200 import synth
    ds = synth.make_dataset(n, synth.SEEDS["S5"])
    P, modelx, types, cmeOpt = synth.default_pars("matern0.5")
    # unconstrained vector of the matern / prodSum / cd1 / cme model: ax, nux, ad, cx0, cx1, cd1, cxd1, cme (npar = 8)
    from iak3d_b200 import api
    pars = np.array([float(api.logitab(P["ax"], *PARBNDS["axBnds"])), float(api.logitab(P["nux"], *PARBNDS["nuxBnds"])),
                     float(api.logitab(P["ad"], *PARBNDS["adBnds"])), np.log(P["cx0"]), np.log(P["cx1"]), np.log(P["cd1"]),
                     float(api.logitab(P["cxd1"], *PARBNDS["sxd1Bnds"])), np.log(P["cme"])])
    return ds, pars, modelx, types, cmeOpt


def s5_config(n, p, nxU, ndIU, modelx, types, cmeOpt, nb, world):
    """`config` of the S5 line -- one function for both arms, so that the reference arm describes the same workload"""
    return dict(workload=f"synthetic {n}-interval single-block REML: iaCovMatern build + Cholesky loglik (BASELINE configs[1])",
                n=n, p=p, nxU=int(nxU), ndIU=int(ndIU), modelx=modelx, nux=0.5, nud=0.5, sdfdTypes=list(types), cmeOpt=cmeOpt,
                evals_per_step=nb, step="one forward-difference gradient (npar+1 evaluations: gradnllIAK3D, fitIAK3D.R:1865-1886)",
                parallelism="replicas only" if world > 1 else "1 GPU",
                l2="working set (factor buffers %.0f MB) exceeds the 126 MB L2; no flush needed" % (nb * 8.0 * n * n / 1e6))


def fd_candidates(pars, step_no, delta=1e-6):
    """pars and its npar forward-difference neighbours; the base point moves a little every step, as it does along
    an optimiser's path (no two steps evaluate the same matrices)."""
    base = pars + 1e-3 * np.sin(0.7 * step_no + np.arange(pars.size))
    out = [base.copy()]
    for i in range(pars.size):
        v = base.copy(); v[i] += delta
        out.append(v)
    return out


def run_s5(args, rank, local_rank, world, comm):
    import ctypes as C
    import iak3d_b200 as iak
    from iak3d_b200 import api, synth
    from iak3d_b200._lib import Pars, make_model, dptr, iptr

    n = args.n
    ds, pars, modelx, types, cmeOpt = s5_inputs(n)
    ctx = iak.Context(local_rank)
    W = np.asfortranarray(np.column_stack([ds["XData"], ds["zData"]]))
    p = ds["XData"].shape[1]; q = p + 1
    sm = iak.SetupMats(ds["xData"], ds["dIData"], W=W, ctx=ctx)
    nb = pars.size + 1
    lib = ctx.lib
    dsarr = (C.c_void_p * nb)(*[sm.h] * nb)
    lnd = np.empty(nb); G = np.empty((nb, q * q)); st = np.empty(nb, np.int32)

    model = make_model(modelx, 0.5, *types, cmeOpt, True, PARBNDS)

    def structs(step_no):
        # the unconstrained vectors of one gradient batch -> parsBTfmd structs, by the library's readParsIAK3D (iak3d_read_pars)
        arr = (Pars * nb)()
        for i, v in enumerate(fd_candidates(pars, step_no)):
            vv = np.ascontiguousarray(v, dtype=np.float64)
            ctx.check(lib.iak3d_read_pars(C.byref(model), dptr(vv), vv.size, C.byref(arr[i])))
        return arr

    def step_resident(step_no):
        ctx.check(lib.iak3d_nll_terms_batch_enqueue(ctx.h, nb, dsarr, structs(step_no)))
        ctx.check(lib.iak3d_nll_terms_batch_fetch(ctx.h, dptr(lnd), dptr(G), iptr(st)))

    kw_e2e = dict(modelx=modelx, nud=0.5, sdfdType_cd1=types[0], sdfdType_cxd0=types[1], sdfdType_cxd1=types[2], cmeOpt=cmeOpt, prodSum=True,
                  setupMats=sm, parBnds=PARBNDS, useReml=True)

    def step_e2e(step_no):
        """the reference-facing call: host W in (uploaded every step), the forward-difference gradient and the value out --
        gradnllIAK3D as optim() calls it (npar + 1 = 9 evaluations in one batched launch; iak3d_gradnll_fd)"""
        sm.upload_W(W)
        g, v0 = api.gradnllIAK3D(fd_candidates(pars, step_no)[0], ds["zData"], ds["XData"], analytic=False, rtn_nll=True, **kw_e2e)
        return [v0] + list(g)

    peaks = dict(dfma=ctx.fp64_peak(0), dmma=ctx.fp64_peak(1), dfma_and_dmma_together=ctx.fp64_peak(2))
    # a write-only stream (cudaMemsetAsync, 16 x 256 MiB): the ceiling of a kernel that only writes, as the covariance build does
    ctx.flush_l2(); ctx.sync(); ctx.timer_start()
    for _ in range(16):
        ctx.flush_l2()
    write_only_gbs = 16 * (256 << 20) / (ctx.timer_stop() * 1e-3) / 1e9
    for w in range(args.warmup):
        step_resident(-1 - w)
    assert np.all(st == 0), f"warm-up evaluations failed: {st}"
    ctx.sync()

    # ---- device-resident timing (value) ----
    clk = ClockSampler(local_rank)
    comm.barrier(); ctx.sync()
    l0 = ctx.launch_count()
    clk.start()
    stage = np.zeros(4)
    ctx.timer_start()
    for k in range(args.steps):
        step_resident(k)
        ms4, flops, cov_bytes = ctx.last_stage_ms()
        stage += np.array(ms4)
    ms_total = ctx.timer_stop()
    comm.barrier(); ctx.sync()
    clocks = clk.stop()
    launches = ctx.launch_count() - l0
    ms_total = float(comm.allreduce_max(np.array([ms_total]))[0])
    stage /= args.steps

    # ---- single-evaluation latency (the optimiser's fn call) ----
    one = (C.c_void_p * 1)(sm.h)
    for k in range(3):                      # untimed: the task list of a one-evaluation batch is built on first use
        ctx.check(lib.iak3d_nll_terms_batch_enqueue(ctx.h, 1, one, structs(-1 - k)))
        ctx.check(lib.iak3d_nll_terms_batch_fetch(ctx.h, dptr(lnd), dptr(G), iptr(st)))
    ctx.sync(); ctx.timer_start()
    reps = max(3, min(20, args.steps))
    for k in range(reps):
        ctx.check(lib.iak3d_nll_terms_batch_enqueue(ctx.h, 1, one, structs(k)))
        ctx.check(lib.iak3d_nll_terms_batch_fetch(ctx.h, dptr(lnd), dptr(G), iptr(st)))
    single_ms = ctx.timer_stop() / reps

    # ---- the Edgeroi-sized call (BASELINE configs[0]: n ~ 1200): what optim() pays per function evaluation ----
    dsE = synth.make_dataset(1450, synth.SEEDS["E"])
    keepE = dsE["prof"] < dsE["prof"].max() - 59                       # 60 hold-out profiles, getEdgeroiData.R:87
    smE = api.setupIAK3D(dsE["xData"][keepE], dsE["dIData"][keepE], ctx=ctx)
    PE, mxE, tyE, cmE = synth.default_pars("edgeroi")
    kwE = dict(modelx=mxE, nud=0.5, sdfdType_cd1=tyE[0], sdfdType_cxd0=tyE[1], sdfdType_cxd1=tyE[2], cmeOpt=cmE, setupMats=smE, parBnds=PARBNDS,
               useReml=True)
    zE, XE = dsE["zData"][keepE], dsE["XData"][keepE]
    for k in range(3):
        api.nllIAK3D(None, zE, XE, parsBTfmd=PE, **kwE)
    stageE = np.zeros(4)
    ctx.sync(); t0 = time.perf_counter()
    for k in range(50):
        PEk = dict(PE); PEk["ax"] = PE["ax"] * (1.0 + 1e-3 * k)
        vE = api.nllIAK3D(None, zE, XE, parsBTfmd=PEk, **kwE)
        stageE += np.array(ctx.last_stage_ms()[0])
    edgeroi_ms = 1e3 * (time.perf_counter() - t0) / 50
    stageE /= 50
    assert np.isfinite(vE) and vE < 1e90
    parsE = np.array([float(api.logitab(PE["ax"], *PARBNDS["axBnds"])), float(api.logitab(PE["ad"], *PARBNDS["adBnds"])), np.log(PE["cx0"]),
                      np.log(PE["cx1"]), float(api.logitab(PE["cxd1"], *PARBNDS["sxd1Bnds"])), np.log(PE["cme"])])
    kwV = {k: v for k, v in kwE.items()}
    for k in range(3):
        api.nllIAK3D(parsE, zE, XE, prodSum=True, **kwV); api.gradnllIAK3D(parsE, zE, XE, prodSum=True, **kwV)
    ctx.sync(); t0 = time.perf_counter()
    for k in range(50):
        api.nllIAK3D(parsE * (1.0 + 1e-4 * k), zE, XE, prodSum=True, **kwV)
    vec_ms = 1e3 * (time.perf_counter() - t0) / 50
    ctx.sync(); t0 = time.perf_counter()
    for k in range(20):
        gE = api.gradnllIAK3D(parsE * (1.0 + 1e-4 * k), zE, XE, prodSum=True, **kwV)
    gradE_ms = 1e3 * (time.perf_counter() - t0) / 20
    assert np.all(np.isfinite(gE))
    edgeroi = dict(n=int(keepE.sum()), nll_call_wall_ms=edgeroi_ms, nll_vector_call_wall_ms=vec_ms, gradient_call_wall_ms=gradE_ms,
                   gradient_evals=int(parsE.size + 1), device_ms=float(stageE.sum()),
                   stage_ms=dict(tables=stageE[0], cov_build=stageE[1], factorise=stageE[2], finish=stageE[3]),
                   what="the optimiser's callbacks through the host mirror at Edgeroi size: nll_call = nllIAK3D(parsBTfmd given; "
                        "iak3d_nll_reml), nll_vector_call = nllIAK3D(pars) (iak3d_nll_vector: transform + device terms + REML tail in one "
                        "library call), gradient_call = gradnllIAK3D (iak3d_gradnll_fd: 7 evaluations in one launch); 19 block columns, each "
                        "a warp-specialised 64 x 64 diagonal factorisation on the critical path")
    smE.close()

    # ---- the same gradient by the analytic pass (iak3d_nll_grad; informational, not part of `value`) ----
    kwg = dict(modelx=modelx, nud=0.5, sdfdType_cd1=types[0], sdfdType_cxd0=types[1], sdfdType_cxd1=types[2], cmeOpt=cmeOpt, prodSum=True,
               setupMats=sm, parBnds=PARBNDS, useReml=True)
    grad_ms = {}
    for name, an in (("forward_difference_batch", False), ("analytic_pass", True)):
        for w in range(4):
            api.gradnllIAK3D(pars, ds["zData"], ds["XData"], analytic=an, **kwg)
        ctx.sync(); t0 = time.perf_counter()
        for k in range(reps):
            gk = api.gradnllIAK3D(pars + 1e-4 * np.sin(k), ds["zData"], ds["XData"], analytic=an, **kwg)
        ctx.sync(); grad_ms[name] = 1e3 * (time.perf_counter() - t0) / reps
        grad_ms[name + "_norm"] = float(np.linalg.norm(gk))
    sm.upload_W(W)

    # ---- end to end through the plugin call (e2e) ----
    for w in range(min(args.warmup, 3)):
        step_e2e(-1 - w)
    comm.barrier(); ctx.sync()
    t0 = time.perf_counter()
    for k in range(args.steps):
        vals = step_e2e(k)
    ctx.sync()
    e2e_s = time.perf_counter() - t0
    e2e_s = float(comm.allreduce_max(np.array([e2e_s]))[0])
    assert all(np.isfinite(v) and v < 1e90 for v in vals)

    evals = nb * args.steps * world
    value = evals / (ms_total * 1e-3)
    tf = (nb * n ** 3 / 3.0) / (stage[2] * 1e-3) / 1e12
    cov_gbs = (nb * 8.0 * n * (n + 1) / 2.0) / (stage[1] * 1e-3) / 1e9
    hbm_peak, hbm_src = measured_peak("hbm_gbs", 6650.0)
    out = dict(
        metric=METRIC, value=value, unit="evals/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=ms_total / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
        config=s5_config(n, p, sm.nxU, sm.ndIU, modelx, types, cmeOpt, nb, world),
        single_eval_ms=single_ms, single_eval_per_s=1e3 / single_ms, edgeroi_sized=edgeroi, gradient_wall_ms=grad_ms, fp64_peaks_tflops=peaks,
        stage_ms=dict(tables=stage[0], cov_build=stage[1], factorise=stage[2], finish=stage[3]),
        roofline=dict(kernel="tile_pair_kernel<0> (tile Cholesky + bordered solves, FP64 DMMA; two half-size CTAs per SM, chol_pair.cuh)",
                      bound="tensor", achieved=tf,
                      peak=peaks["dmma"], unit="TFLOP/s", frac=tf / peaks["dmma"], traffic=ncu_traffic("tile_pair_kernel", n, nb),
                      flops_per_launch=nb * n ** 3 / 3.0,
                      peak_source="FP64 mma.sync m8n8k4 microbenchmark run inside this bench (MEASURED_PEAKS.json holds no FP64 "
                                  "figure); DFMA microbenchmark = %.1f TFLOP/s" % peaks["dfma"]),
        roofline_cov_build=dict(kernel="cov_factor_kernel", bound="hbm", achieved=cov_gbs, peak=hbm_peak, unit="GB/s",
                                frac=cov_gbs / hbm_peak, bytes_per_launch=8.0 * n * (n + 1) / 2.0, peak_source=hbm_src,
                                traffic=ncu_traffic("cov_factor_kernel", n, nb), write_only_peak=write_only_gbs,
                                frac_of_write_only=cov_gbs / write_only_gbs,
                                note="stage time: one batched launch + the table expansion.  peak = the measured COPY bandwidth (read + write "
                                     "bytes); the kernel only writes, write_only_peak = cudaMemsetAsync of 4 GiB in this run; history of the "
                                     "kernel in profiles/cov_build_history_r02.txt"),
        e2e=dict(value=evals / e2e_s, unit="evals/s", h2d_bytes_per_step=int(W.nbytes + nb * C.sizeof(Pars)),
                 d2h_bytes_per_step=int(nb * (2 + q * q) * 8)),
        gpu_launches=int(launches), clocks=clocks)
    return out, ctx


def ncu_traffic(kernel, n, evals):
    """DRAM bytes per launch of `kernel` from the committed ncu capture, when one exists for exactly this shape."""
    try:
        path = next(pth for pth in (os.path.join(ROOT, "profiles", f"traffic_r0{r}.json") for r in (2, 1)) if os.path.exists(pth))
        with open(path) as f:
            e = json.load(f)[kernel]
        return float(e["bytes"]) if (e["n"] == n and e["evals_per_launch"] == evals) else None
    except Exception:
        return None


def measured_peak(key, fallback):
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)[key]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return fallback, "fallback (B200_PROFILING.md)"


def host_threads():
    """Give the host BLAS every core of the box and say how many that is.  torchrun exports OMP_NUM_THREADS=1 to its
    workers, which OpenBLAS reads when it is loaded -- threadpoolctl overrides it at run time."""
    cores = os.cpu_count() or 1
    info = dict(cores=cores, blas="unknown", blas_threads=None)
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=cores)
        libs = [d for d in threadpool_info() if d.get("user_api") == "blas"]
        if libs:
            info.update(blas=f"{libs[0].get('internal_api')} {libs[0].get('version')}", blas_threads=int(libs[0].get("num_threads", 0)))
    except Exception as e:                      # noqa: BLE001
        info["note"] = f"threadpoolctl unavailable ({e})"
    return info


def oracle_s5(n):
    """The reference algorithm's CPU implementation of one S5 evaluation (oracle port; R is not installed on the GPU box:
    profiles/r_probe_gpubox_r02.txt) and the shapes of its setup."""
    from oracle import iak3d_oracle as orc
    ds, pars, modelx, types, cmeOpt = s5_inputs(n)
    sm = orc.setupIAK3D(ds["xData"], ds["dIData"])

    def f(v):
        return orc.nllIAK3D(v, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, sm, PARBNDS, useReml=True)
    return f, pars, ds, modelx, types, cmeOpt, int(np.max(sm["xid"])) + 1, int(np.max(sm["did"])) + 1


def cpu_baseline_s5(n, budget_s=20.0):
    """The oracle (NumPy/SciPy port of the reference algorithm; R is not installed) on the host cores."""
    th = host_threads()
    f, pars, *_ = oracle_s5(n)
    f(pars)
    t0 = time.perf_counter(); k = 0
    while True:
        f(fd_candidates(pars, k)[k % (pars.size + 1)]); k += 1
        if time.perf_counter() - t0 > budget_s or k >= 20:
            break
    dt = time.perf_counter() - t0
    return dict(value=k / dt, unit="evals/s", cores=th["blas_threads"] or th["cores"], kind="port", blas=th["blas"],
                sample=f"{k} single evaluations of the same n={n} workload by oracle/iak3d_oracle.nllIAK3D (NumPy/SciPy, {th['blas']} with "
                       f"{th['blas_threads']} threads on {th['cores']} host cores); R is not installed on the box, so this is a port, not the R "
                       "reference")


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the SAME step on the host cores -- gradnllIAK3D's serial loop
    over the npar + 1 candidates (fitIAK3D.R:1865-1886), each a full nllIAK3D (covariance build + Cholesky + REML tail) by
    the oracle port.  Same config, warm-up and step definition as the repo arm; rank 0 alone runs it."""
    if rank != 0:
        return None
    th = host_threads()
    n = args.n
    f, pars, ds, modelx, types, cmeOpt, nxU, ndIU = oracle_s5(n)
    nb = pars.size + 1
    t0 = time.perf_counter(); f(pars); t1 = time.perf_counter() - t0
    budget = 240.0                                  # bounded: the whole run ends within a few minutes
    total = args.steps + args.warmup
    steps, warm = args.steps, args.warmup
    if total * nb * t1 > budget:                    # keep the step definition, shorten the run, and say so
        warm = max(0, min(args.warmup, 1))
        steps = int(max(1, min(args.steps, (budget / (nb * max(t1, 1e-3))) - warm)))
    for w in range(warm):
        for v in fd_candidates(pars, -1 - w):
            f(v)
    t0 = time.perf_counter()
    for k in range(steps):
        vals = [f(v) for v in fd_candidates(pars, k)]
    dt = time.perf_counter() - t0
    assert all(np.isfinite(v) and v < 1e90 for v in vals)
    v = steps * nb / dt
    cb = dict(value=v, unit="evals/s", cores=th["blas_threads"] or th["cores"], kind="port", blas=th["blas"],
              sample=f"{steps} steps of {nb} evaluations (one forward-difference gradient each) of the n={n} workload, oracle port "
                     f"(NumPy/SciPy, {th['blas']} with {th['blas_threads']} threads on {th['cores']} host cores); the reference is R and R "
                     "is not installed on the box (profiles/r_probe_gpubox_r02.txt)")
    return dict(impl="reference", metric=METRIC, value=v, unit="evals/s", n_gpus=world, steps=steps, steps_requested=args.steps,
                warmup=warm, ms_per_step=1e3 * dt / steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                data="synthetic", config=s5_config(n, ds["XData"].shape[1], nxU, ndIU, modelx, types, cmeOpt, nb, world),
                cpu_baseline=cb, e2e=dict(value=v, unit="evals/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)


def sub_args(args, steps, warmup, **kw):
    a = argparse.Namespace(**vars(args))
    a.steps, a.warmup = steps, warmup
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "s5", "cl", "krige"])
    ap.add_argument("--nobs", dest="n", type=int, default=5000, help="observations (torchrun's parser trips over a bare --n)")
    ap.add_argument("--cl-nobs", dest="cl_n", type=int, default=200000, help="observations of the composite-likelihood workload")
    ap.add_argument("--cells", type=int, default=1000000, help="grid cells of the kriging workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    from iak3d_b200 import parallel
    rank, local_rank, world = parallel.dist_env()
    if args.impl == "reference":
        out = run_reference(args, rank, world)
        if out is not None:
            print(json.dumps(out), flush=True)
        return 0
    parallel.init()
    comm = parallel.BlockComm()
    import bench_extra
    cpu = (rank == 0 and world == 1 and not args.no_cpu_baseline)
    if args.workload in ("all", "s5"):
        out, ctx = run_s5(args, rank, local_rank, world, comm)
        ctx.close()
        if cpu:
            out["cpu_baseline"] = cpu_baseline_s5(args.n)
        if args.workload == "all":
            # the workloads that shard: a few steps each (one composite-likelihood evaluation is 286 factorisations of ~4000,
            # one kriging step 6e6 predictions), sized so that the whole call stays within minutes at any N
            sub, ctx = bench_extra.run(sub_args(args, max(2, min(args.steps, 5)), 3, workload="cl", n=args.cl_n), rank, local_rank, world, comm)
            ctx.close()
            if cpu:
                sub["cpu_baseline"] = bench_extra.cpu_baseline_cl(args.cl_n)
            out["cl200"] = sub
            sub, ctx = bench_extra.run(sub_args(args, max(1, min(args.steps, 2)), 3, workload="krige"), rank, local_rank, world, comm)
            ctx.close()
            if cpu:
                sub["cpu_baseline"] = bench_extra.cpu_baseline_krige(args.n)
            out["p1m"] = sub
    else:
        if args.workload == "cl" and args.n == 5000:
            args.n = args.cl_n
        out, ctx = bench_extra.run(args, rank, local_rank, world, comm)
        ctx.close()
        if cpu:
            out["cpu_baseline"] = bench_extra.cpu_baseline_krige(args.n) if args.workload == "krige" else bench_extra.cpu_baseline_cl(args.n)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
