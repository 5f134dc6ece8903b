"""Writes tests/golden/oracle_cases.json: end-to-end values of the hot path on small seeded datasets, computed by the CPU
oracle (oracle/iak3d_oracle.py) at the time of writing.  They are NOT reference outputs (R cannot run in the build image;
DESIGN.md section 2) -- they freeze the oracle so that neither it nor the CUDA path can drift unnoticed: the CPU suite
checks the oracle against this file, the GPU suite checks the library against it.

    python scripts/make_golden.py        # rewrites the file; review the diff before committing
"""
import json, os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import iak3d_oracle as orc
from iak3d_b200 import synth
from helpers import PARBNDS, case_pars

CASES = [  # (name, n, seed, kind, nonstat, nud, useReml)
    ("edgeroi-reml", 180, 901, "edgeroi", False, 0.5, True),
    ("matern-free-nu-nonstat-ml", 150, 902, "matern0.8", True, 1.5, False),
    ("matern-half-nonstat-reml", 130, 903, "matern0.5", True, 2.5, True),
]


def compute(case):
    name, n, seed, kind, nonstat, nud, useReml = case
    ds = synth.make_dataset(n, seed)
    P, modelx, types, cmeOpt = case_pars(kind, nonstat, nud)
    sm = orc.setupIAK3D(ds["xData"], ds["dIData"])
    t = orc.nllIAK3D(None, ds["zData"], ds["XData"], modelx, nud, *types, cmeOpt, True, sm, PARBNDS, useReml=useReml, forCompLik=True, parsBTfmd=P)
    fit = orc.nllIAK3D(None, ds["zData"], ds["XData"], modelx, nud, *types, cmeOpt, True, sm, PARBNDS, useReml=useReml, rtnAll=True, parsBTfmd=P)
    fit["xData"], fit["dIData"] = ds["xData"], ds["dIData"]
    xMap = synth.make_grid(4)[:12]
    dIMap = synth.GSM_INTERVALS[[0, 3]]
    Xfn = lambda i, idx: synth.grid_design(xMap[idx], dIMap[i], 3)
    z, v, lo, hi = orc.predictIAK3D(xMap, dIMap, Xfn, fit, modelx, types, cmeOpt)
    C, s2 = orc.setCIAK3D(P, modelx, *types, cmeOpt, sm)
    pick = [(0, 0), (1, 0), (n // 2, n // 3), (n - 1, n - 2), (n - 1, n - 1)]
    return dict(name=name, n=n, seed=seed, kind=kind, nonstat=nonstat, nud=nud, useReml=useReml,
                lndetA=float(t["lndetA"]), WiAW=np.asarray(t["WiAW"]).tolist(), nll=float(fit["nll"]), cxdhat=float(fit["cxdhat"]),
                betahat=np.asarray(fit["betahat"]).reshape(-1).tolist(), C_entries=[[i, j, float(C[i, j])] for i, j in pick],
                sigma2Vec_head=np.asarray(s2)[:4].tolist(), zMap=np.asarray(z).tolist(), vMap=np.asarray(v).tolist())


if __name__ == "__main__":
    out = dict(source="oracle/iak3d_oracle.py via scripts/make_golden.py (oracle-frozen values, not R outputs)", cases=[compute(c) for c in CASES])
    with open(os.path.join(ROOT, "tests", "golden", "oracle_cases.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["cases"]), "cases")
