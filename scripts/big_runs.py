"""Full-size runs of BASELINE.json configs[2] (S40) on one B200: REML terms of the 40k-interval single block, checked
against LAPACK on the host (size-independent identities + lndet), then the Newton-Raphson terms (gradHessIAK3D)."""
import json, sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import scipy.linalg as sla
import iak3d_b200 as iak
from iak3d_b200 import synth
from helpers import PARBNDS
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
do_nr = (len(sys.argv) > 2 and sys.argv[2] == "nr")
check = (len(sys.argv) > 3 and sys.argv[3] == "check")
out = dict(n=n)
ctx = iak.Context(0)
t0 = time.time()
ds = synth.make_dataset(n, synth.SEEDS["S40"])
P, modelx, types, cmeOpt = synth.default_pars("matern0.5")
W = np.column_stack([ds["XData"], ds["zData"]])
sm = iak.SetupMats(ds["xData"], ds["dIData"], W=W, ctx=ctx)
out["setup_s"] = time.time() - t0; out["nxU"] = sm.nxU; out["ndIU"] = sm.ndIU
kw = dict(modelx=modelx, nud=0.5, sdfdType_cd1=0, sdfdType_cxd0=0, sdfdType_cxd1=0, cmeOpt=cmeOpt, setupMats=sm, parBnds=PARBNDS, parsBTfmd=P)
for rep in range(3):
    t0 = time.time()
    nll = iak.nllIAK3D(None, ds["zData"], ds["XData"], **kw)
    dt = time.time() - t0
    ms, fl, by = ctx.last_stage_ms()
    print(f"rep {rep}: nll={nll:.6f} wall {dt:.3f}s stages ms {np.round(ms, 2)} chol TF {fl / ms[2] / 1e9:.2f} cov GB/s {by / ms[1] / 1e6:.0f}", flush=True)
out.update(nll=nll, eval_wall_s=dt, stage_ms=list(ms), chol_tflops=fl / ms[2] / 1e9, cov_gbs=by / ms[1] / 1e6)
if check:
    t0 = time.time()
    terms = iak.nllIAK3D(None, ds["zData"], ds["XData"], forCompLik=True, attachBigMats=True, **kw)
    A = terms["A"]
    c = sla.cho_factor(A, lower=True, overwrite_a=False, check_finite=False)
    lnd = 2.0 * np.sum(np.log(np.diag(c[0])))
    want = W.T @ sla.cho_solve(c, W, check_finite=False)
    out["lapack_s"] = time.time() - t0
    out["lndet_rel_err"] = abs(terms["lndetA"] - lnd) / abs(lnd)
    out["WiAW_scaled_err"] = float(np.max(np.abs(terms["WiAW"] - want)) / np.max(np.abs(want)))
    out["A_iAW_minus_W_scaled_err"] = float(np.max(np.abs(A @ terms["iAW"] - W)) / np.max(np.abs(W)))
    print("check:", {k: out[k] for k in ("lapack_s", "lndet_rel_err", "WiAW_scaled_err", "A_iAW_minus_W_scaled_err")}, flush=True)
    del A, c, terms
if do_nr:
    th = np.log([P[k] for k in ("cx0", "cx1", "cd1", "cxd0", "cxd1", "cme")])
    t0 = time.time()
    g = iak.gradHessIAK3D(sm, P, modelx, types, th)
    dt = time.time() - t0
    flops = 3 * n ** 3 / 3.0 + 6 * 2.0 * n ** 3        # chol + L^-T + U U' (triangular-aware: n^3/3 each) + six T dC_i
    out.update(nr_wall_s=dt, nr_nll=g["nll"], nr_grad=list(map(float, g["grad"])), nr_tflops=flops / dt / 1e12)
    print(f"gradHess: {dt:.2f}s nll={g['nll']:.6f} ~{flops / dt / 1e12:.1f} TFLOP/s overall; grad={np.round(g['grad'], 4)}", flush=True)
    print("fim diag", np.round(np.diag(g["fim"]), 4))
print("RESULT " + json.dumps(out))
