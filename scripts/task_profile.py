"""Per-task timeline of one factorisation (IAK3D_PROF=1): where the critical path goes."""
import os, sys, ctypes as C, numpy as np
os.environ["IAK3D_PROF"] = "1"
sys.path.insert(0, '.')
import iak3d_b200 as iak
from iak3d_b200 import synth
from iak3d_b200._lib import make_pars, Pars, dptr, iptr
ctx = iak.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ds = synth.make_dataset(n, synth.SEEDS["S5"])
W = np.column_stack([ds["XData"], ds["zData"]])
sm = iak.SetupMats(ds["xData"], ds["dIData"], W=W, ctx=ctx)
q = W.shape[1]
P, modelx, types, cmeOpt = synth.default_pars("matern0.5")
arr = (Pars * nb)(*[make_pars(P, modelx, *types, cmeOpt)] * nb)
dsarr = (C.c_void_p * nb)(*[sm.h] * nb)
lnd = np.empty(nb); G = np.empty((nb, q * q)); st = np.empty(nb, np.int32)
for r in range(3):
    ctx.check(ctx.lib.iak3d_nll_terms_batch(ctx.h, nb, dsarr, arr, dptr(lnd), dptr(G), iptr(st)))
nt = C.c_int64()
ctx.check(ctx.lib.iak3d_last_task_profile(ctx.h, None, 0, C.byref(nt)))
buf = np.zeros((nt.value, 11), np.int64)
ctx.check(ctx.lib.iak3d_last_task_profile(ctx.h, buf.ctypes.data_as(C.POINTER(C.c_int64)), nt.value, C.byref(nt)))
t0 = buf[:, 3].min()
T = (buf[:, 3:9] - t0) / 1e3; dep = (buf[:, 9] - t0) / 1e3
isd = (buf[:, 10] & 1) == 1
print("tasks", nt.value, "span us", T[:, 5].max(), "stage ms", ctx.last_stage_ms()[0])
names = ["kloop", "E1", "E2(potf2/inv)", "E3E4", "publish"]
for lab, m in (("diag", isd), ("offdiag", ~isd)):
    d = np.diff(T[m], axis=1)
    print(lab, "count", m.sum(), " mean us per phase:", {k: round(float(v), 2) for k, v in zip(names, d.mean(0))},
          " dep->kloop_end us:", round(float((T[m, 1] - dep[m]).mean()), 2))
# critical path per column for problem 0: diag published(j) -> diag published(j+1)
p0 = buf[:, 0] == 0
dj = {int(r[2]): i for i, r in enumerate(buf) if r[0] == 0 and (r[10] & 1)}
js = sorted(dj)
gaps = [T[dj[j + 1], 5] - T[dj[j], 5] for j in js[:-1]]
print("diag-to-diag gap us: mean %.2f  (even->odd %.2f, odd->even %.2f)" % (np.mean(gaps), np.mean(gaps[0::2]), np.mean(gaps[1::2])))
j = js[len(js) // 2]
i = dj[j]
print("example diag task j=%d:" % j, np.round(T[i], 2), "dep", round(dep[i], 2))
sub = [k for k, r in enumerate(buf) if r[0] == 0 and r[2] == j and not (r[10] & 1)]
for k in sub[:3]:
    print("   offdiag r=%d:" % buf[k, 1], np.round(T[k], 2), "dep", round(dep[k], 2))
# K-loop rate: us per 32-deep chunk (ideal: 4096 clk = 2.08 us at 1.965 GHz)
nch = buf[:, 2] * 2
m = nch >= 20
rate = (T[m, 1] - T[m, 0]) / nch[m]
print("K-loop us/chunk: min %.3f  p10 %.3f  median %.3f  mean %.3f  p90 %.3f" % (rate.min(), np.percentile(rate, 10), np.median(rate), rate.mean(), np.percentile(rate, 90)))
busy = (T[:, 5] - T[:, 0]).sum(); span = T[:, 5].max()
print("sum of task durations / (span x 148 SMs) = %.3f ; ideal MMA time for all chunks / (span x 148) = %.3f" % (busy / (span * 148), nch.sum() * 2.084 / (span * 148)))
kw = ((buf[:, 10] >> 1) & ((1 << 47) - 1)) / 1965.0   # us at 1965 MHz
print("K-loop operand-wait of warp 0: mean %.1f us per task = %.1f%% of the K-loop; per chunk %.3f us" % (kw[m].mean(), 100 * kw[m].sum() / (T[m, 1] - T[m, 0]).sum(), (kw[m] / nch[m]).mean()))
