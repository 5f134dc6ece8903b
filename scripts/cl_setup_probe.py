"""Time of setupIAK3D_CL (per-pair device datasets, iaCovMatern.R:551-575) and of one evaluation at BASELINE configs[3]."""
import sys, time, json, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import iak3d_b200 as iak
from iak3d_b200 import synth
from helpers import PARBNDS
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
ctx = iak.Context(0)
t0 = time.time(); ds = synth.make_dataset(n, synth.SEEDS["CL200"]); t_ds = time.time() - t0
nblocks = max(4, int(round(n / 2000.0)))
t0 = time.time(); blk = synth.make_blocks(ds["xData"], ds["prof"], nblocks, synth.SEEDS["CL200"]); t_blk = time.time() - t0
blk["compLikOptn"] = 2
P, modelx, types, cmeOpt = synth.default_pars("edgeroi")
t0 = time.time(); cl = iak.setupIAK3D_CL(ds["xData"], ds["dIData"], ds["zData"], ds["XData"], blk, ctx=ctx); ctx.sync(); t_setup = time.time() - t0
ts = []
for r in range(3):
    t0 = time.time(); v = iak.nllIAK3D_CL(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, False, cl, parsBTfmd=P); ts.append(time.time() - t0)
print("RESULT " + json.dumps(dict(n=n, units=len(cl["units"]), synth_dataset_s=t_ds, synth_blocks_s=t_blk, setupIAK3D_CL_s=t_setup, eval_s=ts, nll=v)))
