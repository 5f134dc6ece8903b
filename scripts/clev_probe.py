"""Eidsvik composite-likelihood block prediction (predMatsIAK3D_CLEV_ByBlock) at block sizes of BASELINE configs[3]
(~2000 data per block): time, memory and sanity of a prediction grid after a compLikOptn-2 fit."""
import json, sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import iak3d_b200 as iak
from iak3d_b200 import synth
from helpers import PARBNDS
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
side = int(sys.argv[2]) if len(sys.argv) > 2 else 100
optn = int(sys.argv[3]) if len(sys.argv) > 3 else 2
nblocks = max(4, int(round(n / 2000.0)))
ctx = iak.Context(0)
ds = synth.make_dataset(n, synth.SEEDS["CL200"])
blk = synth.make_blocks(ds["xData"], ds["prof"], nblocks, synth.SEEDS["CL200"])
blk["compLikOptn"] = optn
P, modelx, types, cmeOpt = synth.default_pars("edgeroi")
t0 = time.time()
cl = iak.setupIAK3D_CL(ds["xData"], ds["dIData"], ds["zData"], ds["XData"], blk, ctx=ctx)
fit = iak.nllIAK3D_CL(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, False, cl, parsBTfmd=P, rtnAll=True)
t_fit = time.time() - t0
xMap = synth.make_grid(side)
dIMap = synth.GSM_INTERVALS[1:2]
Xfn = lambda i, idx: synth.grid_design(xMap[idx], dIMap[i], synth.SEEDS["P1M"])
bm = iak.getBlocksFromLocations(xMap, blk)
cnt = np.bincount(bm, minlength=nblocks)
nadj = np.zeros(nblocks, int)
for a, b in np.asarray(blk["subsetPairsAdj"]).reshape(-1, 2):
    nadj[a] += 1; nadj[b] += 1
out = dict(n=n, nblocks=nblocks, optn=optn, map_points=len(xMap), fit_s=t_fit, nll=fit["nll"], map_per_block_max=int(cnt.max()),
           adj_max=int(nadj.max()), block_max=int(max(len(b) for b in blk["listBlocks"])))
for rep in range(2):
    ctx.sync(); l0 = ctx.launch_count(); t0 = time.time()
    pg = iak.predictIAK3D_CL(xMap, dIMap, Xfn, fit)
    ctx.sync(); dt = time.time() - t0
    out[f"predict_s_{rep}"] = dt; out["launches"] = int(ctx.launch_count() - l0)
z, v = pg["zMap"][0], pg["vMap"][0]
out.update(preds_per_s=len(xMap) / dt, z_range=[float(z.min()), float(z.max())], zData_range=[float(ds["zData"].min()), float(ds["zData"].max())],
           v_range=[float(v.min()), float(v.max())], finite=bool(np.all(np.isfinite(z)) and np.all(np.isfinite(v))))
print("RESULT " + json.dumps(out))
if len(sys.argv) > 4 and sys.argv[4] == "profile":
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    iak.predictIAK3D_CL(xMap, dIMap, Xfn, fit)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)
