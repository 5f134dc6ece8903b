"""Per-task timeline of one kriging panel (tile tasks in mode 1, IAK3D_PROF=1): K-loop rate, epilogue phases and the
gap between consecutive tasks of one CTA."""
import os, sys, ctypes as C, numpy as np
os.environ["IAK3D_PROF"] = "1"
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import iak3d_b200 as iak
from iak3d_b200 import synth
from helpers import PARBNDS
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 37888
ctx = iak.Context(0)
ds = synth.make_dataset(n, synth.SEEDS["S5"])
P, modelx, types, cmeOpt = synth.default_pars("edgeroi")
sm = iak.setupIAK3D(ds["xData"], ds["dIData"], ctx=ctx)
fit = iak.nllIAK3D(None, ds["zData"], ds["XData"], modelx=modelx, nud=0.5, sdfdType_cd1=types[0], sdfdType_cxd0=types[1],
                   sdfdType_cxd1=types[2], cmeOpt=cmeOpt, setupMats=sm, parBnds=PARBNDS, rtnAll=True, attachBigMats=True, parsBTfmd=P)
side = int(np.ceil(np.sqrt(m)))
xMap = synth.make_grid(side)[:m]
dI = np.repeat(synth.GSM_INTERVALS[1:2], m, axis=0)
X = synth.grid_design(xMap, synth.GSM_INTERVALS[1], 3)
k = fit["fit"]
k.stage(xMap, dI, X)
for r in range(3):
    k.enqueue(); z, v = k.fetch()
nt = C.c_int64()
ctx.check(ctx.lib.iak3d_last_task_profile(ctx.h, None, 0, C.byref(nt)))
buf = np.zeros((nt.value, 11), np.int64)
ctx.check(ctx.lib.iak3d_last_task_profile(ctx.h, buf.ctypes.data_as(C.POINTER(C.c_int64)), nt.value, C.byref(nt)))
t0 = buf[:, 3].min()
T = (buf[:, 3:9] - t0) / 1e3
cta = buf[:, 10] >> 48
kw = ((buf[:, 10] >> 1) & ((1 << 47) - 1)) / 1965.0
nch = buf[:, 2] * 2
span = T[:, 5].max()
print("tasks", nt.value, "span us %.1f" % span, " ideal MMA us (all chunks + 2 for the solve)/148: %.1f" % ((nch + 2).sum() * 2.084 / 148))
names = ["kloop", "E1", "inv fetch", "E3E4E5", "publish"]
d = np.diff(T, axis=1)
print("mean us per phase:", {a: round(float(b), 2) for a, b in zip(names, d.mean(0))})
mm = nch >= 20
rate = d[mm, 0] / nch[mm]
print("K-loop us/chunk: p10 %.3f median %.3f mean %.3f p90 %.3f; operand wait %.1f%% of the K-loop" %
      (np.percentile(rate, 10), np.median(rate), rate.mean(), np.percentile(rate, 90), 100 * kw[mm].sum() / d[mm, 0].sum()))
e3 = (buf[:, 9] - t0) / 1e3          # mode 1: slot 6 holds the end of E3 (the solve against the inverse diagonal block)
print("E3 (solve) %.2f us, E4+E5 (copy-out, border reduction) %.2f us" % ((e3 - T[:, 3]).mean(), (T[:, 4] - e3).mean()))
gaps = []
for c in np.unique(cta):
    i = np.where(cta == c)[0]
    i = i[np.argsort(T[i, 0])]
    gaps.extend(T[i[1:], 0] - T[i[:-1], 5])
gaps = np.array(gaps)
print("gap between a CTA's tasks (publish -> next start): mean %.2f median %.2f p90 %.2f us" % (gaps.mean(), np.median(gaps), np.percentile(gaps, 90)))
busy = (T[:, 5] - T[:, 0]).sum()
print("busy/(span*148) %.3f; kloop share of span %.3f; epilogue share %.3f; gap share %.3f" %
      (busy / (span * 148), d[:, 0].sum() / (span * 148), d[:, 1:].sum() / (span * 148), gaps.sum() / (span * 148)))
first = np.array([T[cta == c, 0].min() for c in np.unique(cta)]); last = np.array([T[cta == c, 5].max() for c in np.unique(cta)])
print("CTA start spread %.1f us, finish spread: min %.1f max %.1f" % (first.max() - first.min(), last.min(), last.max()))
