"""Where the fixed per-evaluation cost of nllIAK3D_CL goes (one rank's share of CL200 at world = 8, on one GPU): wall time of
the call against the device stage times, and the host phases inside cl_terms_sets."""
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import iak3d_b200 as iak
from iak3d_b200 import synth, api
from helpers import PARBNDS
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ctx = iak.Context(0)
ds = synth.make_dataset(n, synth.SEEDS["CL200"])
blk = synth.make_blocks(ds["xData"], ds["prof"], max(4, int(round(n / 2000.0))), synth.SEEDS["CL200"])
blk["compLikOptn"] = 2
P0, modelx, types, cmeOpt = synth.default_pars("edgeroi")
cl = iak.setupIAK3D_CL(ds["xData"], ds["dIData"], ds["zData"], ds["XData"], blk, ctx=ctx, rank=0, world=world)
print("units on this rank:", len(cl["units"]))
def step(k, tm=None):
    P = dict(P0); P["ax"] = P0["ax"] * (1.0 + 1e-3 * np.sin(0.7 * k))
    return iak.nllIAK3D_CL(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, False, cl, parsBTfmd=P, timings=tm)
for k in range(3): step(-k - 1)
walls, devs = [], []
for k in range(10):
    ctx.sync(); t0 = time.perf_counter()
    step(k)
    walls.append(time.perf_counter() - t0)
    devs.append(sum(ctx.last_stage_ms()[0]))
print("wall ms per call: median %.3f min %.3f ; device stages ms: median %.3f ; difference %.3f ms" %
      (1e3 * np.median(walls), 1e3 * min(walls), np.median(devs), 1e3 * np.median(walls) - np.median(devs)))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for k in range(10): step(100 + k)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
