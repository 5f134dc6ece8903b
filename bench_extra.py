"""The two workloads of the path that shard across GPUs (BASELINE.json configs[3], configs[4]); driven by
`bench.py --workload cl|krige`.

  cl     synthetic composite likelihood (compositeLikIAK3D.R:6-307, compLikOptn = 2: adjacent block pairs, ML): the
         block pairs are dealt to the ranks (LPT on n^3), each rank evaluates its pairs in ONE batch launch, and the
         sums are combined by one all-reduce of (p+1)^2 + 3 doubles (NCCL through torch.distributed).  Strong scaling.
  krige  block kriging of a cells x 6 GlobalSoilMap-interval grid (predictIAK3D.R:5-287, egEdgeroi.R:396) against a
         kept fit of the 5k-interval dataset; every rank holds the factor and predicts a contiguous tile of the grid;
         no collective in the data path.  Strong scaling.
"""
from __future__ import annotations

import time

import numpy as np

from bench import PARBNDS, ClockSampler


def _cl(args, rank, local_rank, world, comm):
    import iak3d_b200 as iak
    from iak3d_b200 import api, synth
    n = args.n
    nblocks = max(4, int(round(n / 2000.0)))
    ctx = iak.Context(local_rank)
    ds = synth.make_dataset(n, synth.SEEDS["CL200"])
    blk = synth.make_blocks(ds["xData"], ds["prof"], nblocks, synth.SEEDS["CL200"])
    blk["compLikOptn"] = 2
    P0, modelx, types, cmeOpt = synth.default_pars("edgeroi")
    cl = iak.setupIAK3D_CL(ds["xData"], ds["dIData"], ds["zData"], ds["XData"], blk, ctx=ctx, rank=rank, world=world)
    sizes = [u["n"] for u in cl["units"]]
    p = ds["XData"].shape[1]
    tm = {}

    def step(k):
        P = dict(P0); P["ax"] = P0["ax"] * (1.0 + 1e-3 * np.sin(0.7 * k)); P["cme"] = P0["cme"] * (1.0 + 1e-3 * np.cos(0.3 * k))
        return iak.nllIAK3D_CL(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, False, cl, parsBTfmd=P, comm=comm,
                               timings=tm)

    peak = ctx.fp64_peak(1)
    for w in range(args.warmup):
        v = step(-1 - w)
    assert np.isfinite(v) and v < 1e90
    tm.clear()
    clk = ClockSampler(local_rank)
    comm.barrier(); ctx.sync()
    l0 = ctx.launch_count(); clk.start()
    stage = np.zeros(4)
    t0 = time.perf_counter()
    ctx.timer_start()
    for k in range(args.steps):
        step(k)
        if cl["units"]:
            stage += np.array(ctx.last_stage_ms()[0])
    ms = ctx.timer_stop()
    wall = time.perf_counter() - t0
    comm.barrier(); ctx.sync()
    clocks = clk.stop()
    launches = ctx.launch_count() - l0
    ms_rank = comm.gather_scalar(ms / args.steps)
    ms = float(comm.allreduce_max(np.array([ms]))[0])
    wall = float(comm.allreduce_max(np.array([wall]))[0])
    flops_rank = sum(float(s) ** 3 / 3.0 for s in sizes)
    flops = float(comm.allreduce_sum(np.array([flops_rank]))[0])
    stage /= args.steps
    fac_rank = comm.gather_scalar(stage[2]); pre_rank = comm.gather_scalar(stage[0] + stage[1]); units_rank = comm.gather_scalar(len(sizes))
    rows_rank = comm.gather_scalar(sum(sizes))
    tf_rank = flops_rank / max(stage[2], 1e-9) / 1e9
    # the optimiser's gradient call: npar + 1 composite-likelihood evaluations as batched launches + ONE all-reduce
    pars = np.array([float(api.logitab(P0["ax"], *PARBNDS["axBnds"])), float(api.logitab(P0["ad"], *PARBNDS["adBnds"])), np.log(P0["cx0"]),
                     np.log(P0["cx1"]), float(api.logitab(P0["cxd1"], *PARBNDS["sxd1Bnds"])), np.log(P0["cme"])])
    gargs = (ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, False, cl)
    tg = {}
    iak.gradnllIAK3D_CL(pars, *gargs, comm=comm)
    comm.barrier(); ctx.sync(); t0 = time.perf_counter()
    g = iak.gradnllIAK3D_CL(pars * (1 + 1e-6), *gargs, comm=comm, timings=tg)
    grad_s = float(comm.allreduce_max(np.array([time.perf_counter() - t0]))[0])
    assert np.all(np.isfinite(g))
    calls = max(tm.get("allreduce_calls", 0), 1)
    out = dict(metric="composite-likelihood evals/sec at n obs", value=args.steps / (ms * 1e-3), unit="evals/s", n_gpus=world,
               steps=args.steps, warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True, scaling="strong",
               vs_baseline=None, dtype="f64", data="synthetic",
               config=dict(workload=f"synthetic {n}-interval composite likelihood, {nblocks} Voronoi blocks, compLikOptn=2 "
                                    "(adjacent pairs, ML) (BASELINE configs[3])", n=n, p=p, blocks=nblocks,
                           block_pairs=int(cl["n_units_total"]), mean_pair_size=float(np.mean(sizes)) if sizes else 0,
                           parallelism=f"block pairs over {world} GPU(s) (LPT on n^3), sums formed on the device, one all-reduce of "
                                       f"{(p + 1) ** 2 + 2} doubles per evaluation",
                           l2="per-rank factor buffers exceed the 126 MB L2; no flush needed"),
               units_per_s=int(cl["n_units_total"]) * args.steps / (ms * 1e-3),
               per_rank=dict(pairs=[int(x) for x in units_rank], rows=[int(x) for x in rows_rank], ms_per_step=[float(x) for x in ms_rank],
                             factorise_ms=[float(x) for x in fac_rank], tables_build_ms=[float(x) for x in pre_rank]),
               allreduce=dict(us_per_eval=tm.get("allreduce_us", 0.0) / calls, calls_per_eval=1 if world > 1 else 0,
                              doubles=(p + 1) ** 2 + 2, kind=tm.get("allreduce_kind", "none"),
                              where="on the device buffer the library summed into, one D2H after it; kind 'peer' = the library's own "
                                    "kernel over CUDA-IPC mailboxes (NVLink loads, rank-ordered sum), 'nccl' = ncclAllReduce"),
               gradient=dict(ms=1e3 * grad_s, evals=int(pars.size + 1), evals_per_s=(pars.size + 1) / grad_s,
                             allreduce_calls=int(tg.get("allreduce_calls", 0)), allreduce_us=float(tg.get("allreduce_us", 0.0)),
                             what="gradnllIAK3D_CL: npar + 1 evaluations as (units x parameter sets) batched launches + ONE all-reduce"),
               stage_ms=dict(tables=stage[0], cov_build=stage[1], factorise=stage[2], finish=stage[3]),
               roofline=dict(kernel="tile_pair_kernel<0> (rank 0's share)", bound="tensor", achieved=tf_rank, peak=peak, unit="TFLOP/s",
                             frac=tf_rank / peak, traffic=None, flops_per_launch=flops_rank,
                             peak_source="FP64 DMMA microbenchmark in this run"),
               whole_job_tflops=flops * args.steps / (ms * 1e-3) / 1e12,
               e2e=dict(value=args.steps / wall, unit="evals/s", h2d_bytes_per_step=int(len(sizes) * (192 + 12)),
                        d2h_bytes_per_step=int(((p + 1) ** 2 + 2) * 8),
                        note="wall clock of the reference-facing call (nllIAK3D_CL: host parameters in, host nll out, all-reduce "
                             "included); block data are resident like setupMats"),
               gpu_launches=int(launches), clocks=clocks)
    return out, ctx


def _krige(args, rank, local_rank, world, comm):
    import iak3d_b200 as iak
    from iak3d_b200 import parallel, synth
    n = args.n
    cells = getattr(args, "cells", 1000000)
    ctx = iak.Context(local_rank)
    ds = synth.make_dataset(n, synth.SEEDS["S5"])
    P, modelx, types, cmeOpt = synth.default_pars("edgeroi")
    sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)
    # the trend model: makeXvX with modelX = c('const', 4 covariates, 'd', 'd2', 'd3') (makeXvX.R:294-353) -- p = 8 like the other
    # workloads.  The design of the DATA rows is made by the same device code as the design of the map (iak.makeXvX).
    names = ["const", "c1", "c2", "c3", "c4", "d", "d2", "d3"]
    rngc = np.random.default_rng(synth.SEEDS["S5"] + 7)
    covp = rngc.standard_normal((int(ds["prof"].max()) + 1, 4))[ds["prof"]]
    XData = iak.makeXvX({f"c{i + 1}": covp[:, i] for i in range(4)}, ds["dIData"], names, nDiscPts=1000, XLims=None, ctx=ctx)["X"]
    fit = iak.nllIAK3D(None, ds["zData"], XData, modelx=modelx, nud=0.5, sdfdType_cd1=types[0], sdfdType_cxd0=types[1],
                       sdfdType_cxd1=types[2], cmeOpt=cmeOpt, setupMats=sm, parBnds=PARBNDS, rtnAll=True, attachBigMats=True, parsBTfmd=P)
    side = int(np.ceil(np.sqrt(cells)))
    lo, hi = parallel.grid_tile(cells, rank, world)
    xMap = synth.make_grid(side)[lo:hi]
    m = hi - lo
    nd = len(synth.GSM_INTERVALS)
    rngm = np.random.default_rng(synth.SEEDS["P1M"] + 7)
    kk = rngm.uniform(0.02, 0.2, (4, 2)); ph = rngm.uniform(0, 2 * np.pi, 4)
    covMap = np.asfortranarray(np.cos(xMap @ kk.T + ph[None, :]))             # covariates of the map cells (m x 4)
    p8 = len(names)
    XL = np.vstack([np.full(p8, -np.inf), np.full(p8, np.inf)])               # predictIAK3D.R:82-84 (constrainX4Pred = FALSE)
    dm = iak.DesignModel(names, ["c1", "c2", "c3", "c4"], None, (), 0, 10, XL, ctx=ctx)
    # the row-by-row staging the reference's loops imply (host-made X): kept as the comparison leg
    xy = np.tile(xMap, (nd, 1))
    dI = np.repeat(synth.GSM_INTERVALS, m, axis=0)
    X = np.vstack([dm.eval(covMap, np.repeat(synth.GSM_INTERVALS[i:i + 1], m, axis=0)) for i in range(nd)])
    xy, dI, X = np.asfortranarray(xy), np.asfortranarray(dI), np.asfortranarray(X)
    xMapF = np.asfortranarray(xMap)
    k = fit["fit"]
    peak = ctx.fp64_peak(1)
    k.stage(xy, dI, X)
    for w in range(max(1, min(args.warmup, 2))):
        k.enqueue(); z, v = k.fetch()
    assert np.all(np.isfinite(z)) and np.all(np.isfinite(v))
    clk = ClockSampler(local_rank)
    comm.barrier(); ctx.sync()
    l0 = ctx.launch_count(); clk.start()
    ctx.timer_start()
    for s in range(args.steps):
        k.enqueue()
    ms = ctx.timer_stop()
    z, v = k.fetch()
    comm.barrier(); ctx.sync()
    clocks = clk.stop()
    launches = ctx.launch_count() - l0
    ms = float(comm.allreduce_max(np.array([ms]))[0])
    # end to end through the plugin call: host arrays in, z and v out.  (1) the grid call: per-cell coordinates and covariates
    # in, the design rows made on the device (iak3d_predict_stage_grid == predictIAK3D's loops with makeXvX inside);
    # (2) the row-by-row call with a host-made design matrix (what round 1 measured)
    comm.barrier(); t0 = time.perf_counter()
    for s in range(args.steps):
        k.stage_grid(dm, xMapF, covMap, synth.GSM_INTERVALS); k.enqueue(); z3, v3 = k.fetch()
    e2e_grid_s = float(comm.allreduce_max(np.array([time.perf_counter() - t0]))[0])
    comm.barrier(); t0 = time.perf_counter()
    for s in range(args.steps):
        z2, v2 = k.predict(xy, dI, X)
    e2e_s = float(comm.allreduce_max(np.array([time.perf_counter() - t0]))[0])
    assert np.allclose(z3, z2, rtol=1e-9, atol=1e-9) and np.allclose(v3, v2, rtol=1e-9, atol=1e-9)
    cells_rank = comm.gather_scalar(m)
    preds = cells * nd * args.steps
    tf = (m * nd * float(n) ** 2) * args.steps / (ms * 1e-3) / 1e12
    out = dict(metric="kriging preds/sec", value=preds / (ms * 1e-3), unit="preds/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
               ms_per_step=ms / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
               config=dict(workload=f"block kriging of a {cells}-cell grid x {nd} GlobalSoilMap intervals with variances against the "
                                    f"{n}-interval fit (BASELINE configs[4])", n=n, p=ds["XData"].shape[1], cells=cells, intervals=nd,
                           parallelism=f"grid tiles over {world} GPU(s), no data-path collective",
                           l2="each 16384-support panel (0.7 GB) exceeds the 126 MB L2; no flush needed"),
               roofline=dict(kernel="tile_pair_kernel<1> (kriging panels; rank 0's tile)", bound="tensor", achieved=tf, peak=peak, unit="TFLOP/s",
                             frac=tf / peak, traffic=None, flops_per_launch=16384.0 * n * n,
                             note="n^2 flops per prediction (X = Ckh L^-T); the reference's Ckh %*% iC costs 2 n^2",
                             peak_source="FP64 DMMA microbenchmark in this run"),
               per_rank=dict(cells=[int(x) for x in cells_rank]), allreduce=dict(us_per_eval=0.0, calls_per_eval=0, where="no collective in the data path"),
               e2e=dict(value=preds / e2e_grid_s, unit="preds/s",
                        h2d_bytes_per_step=int(xMapF.nbytes + covMap.nbytes + synth.GSM_INTERVALS.nbytes), d2h_bytes_per_step=int(2 * m * nd * 8),
                        note="KeptFit.stage_grid + enqueue + fetch: cell coordinates and covariates in, design matrix made on the "
                             "device (makeXvX, mlr model const + 4 covariates + d + d2 + d3), zMap and vMap out"),
               e2e_host_design=dict(value=preds / e2e_s, unit="preds/s", h2d_bytes_per_step=int(xy.nbytes + dI.nbytes + X.nbytes),
                                    d2h_bytes_per_step=int(2 * m * nd * 8), note="KeptFit.predict with a host-made m x p design matrix"),
               gpu_launches=int(launches), clocks=clocks)
    return out, ctx


def cpu_baseline_krige(n, sample=2000):
    """The oracle's predictIAK3D (setCIAK3D2 + Ckh %*% [iCX, iCz, iC], predictIAK3D.R:179-255) on a bounded sample of the
    grid, host cores; the fit it predicts from (one nllIAK3D with rtnAll) is outside the timed region like on the GPU."""
    import os
    from oracle import iak3d_oracle as orc
    from iak3d_b200 import synth
    ds = synth.make_dataset(n, synth.SEEDS["S5"])
    P, modelx, types, cmeOpt = synth.default_pars("edgeroi")
    sm = orc.setupIAK3D(ds["xData"], ds["dIData"])
    fit = orc.nllIAK3D(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, sm, PARBNDS, useReml=True, rtnAll=True, parsBTfmd=P)
    fit["xData"], fit["dIData"] = ds["xData"], ds["dIData"]
    side = int(np.ceil(np.sqrt(1000000)))
    xMap = synth.make_grid(side)[:sample]
    dI = synth.GSM_INTERVALS[1:2]
    Xfn = lambda i, idx: synth.grid_design(xMap[idx], dI[i], synth.SEEDS["P1M"])
    orc.predictIAK3D(xMap[:200], dI, lambda i, idx: Xfn(i, idx), fit, modelx, types, cmeOpt)
    t0 = time.perf_counter()
    orc.predictIAK3D(xMap, dI, Xfn, fit, modelx, types, cmeOpt)
    dt = time.perf_counter() - t0
    return dict(value=sample / dt, unit="preds/s", cores=os.cpu_count(), kind="port",
                sample=f"{sample} grid cells x 1 depth interval against the same n={n} fit by oracle/iak3d_oracle.predictIAK3D (NumPy/SciPy, "
                       "OpenBLAS threads = host cores); R is not installed, so this is a port, not the R reference")


def cpu_baseline_cl(n, npairs=3):
    """The oracle's per-pair work of nllIAK3D_CL (setCIAK3D + Cholesky terms of one adjacent pair, compositeLikIAK3D.R:54-64)
    on a bounded sample of the pairs, scaled to the whole evaluation by the pairs' n^3."""
    import os
    from oracle import iak3d_oracle as orc
    from iak3d_b200 import synth
    nblocks = max(4, int(round(n / 2000.0)))
    ds = synth.make_dataset(n, synth.SEEDS["CL200"])
    blk = synth.make_blocks(ds["xData"], ds["prof"], nblocks, synth.SEEDS["CL200"])
    P, modelx, types, cmeOpt = synth.default_pars("edgeroi")
    pairs = np.asarray(blk["subsetPairsAdj"]).reshape(-1, 2)
    sizes = np.array([len(blk["listBlocks"][a]) + len(blk["listBlocks"][b]) for a, b in pairs], float)
    pick = np.argsort(np.abs(sizes - np.median(sizes)))[:npairs]
    dt = 0.0
    for k in pick:
        rows = np.concatenate([blk["listBlocks"][pairs[k][0]], blk["listBlocks"][pairs[k][1]]])
        sm = orc.setupIAK3D(ds["xData"][rows], ds["dIData"][rows])
        t0 = time.perf_counter()
        orc.nllIAK3D(None, ds["zData"][rows], ds["XData"][rows], modelx, 0.5, *types, cmeOpt, True, sm, PARBNDS, useReml=False,
                     forCompLik=True, parsBTfmd=P)
        dt += time.perf_counter() - t0
    est = dt * float(np.sum(sizes ** 3)) / float(np.sum(sizes[pick] ** 3))
    return dict(value=1.0 / est, unit="evals/s", cores=os.cpu_count(), kind="port",
                sample=f"{npairs} median-sized adjacent pairs of the {len(pairs)} (oracle/iak3d_oracle.nllIAK3D, forCompLik) timed on the host "
                       f"cores and scaled by sum n_kl^3 to one whole evaluation ({est:.1f} s); setup of the pair datasets excluded as on the GPU")


def run(args, rank, local_rank, world, comm):
    if args.workload == "cl":
        return _cl(args, rank, local_rank, world, comm)
    return _krige(args, rank, local_rank, world, comm)
