import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import iak3d_b200 as iak
from iak3d_b200 import synth
from oracle import iak3d_oracle as orc
from helpers import case_pars
ctx = iak.Context(0)
n, ncomp = 400, 6
ds = synth.make_dataset(n, 4)
P, modelx, types, cmeOpt = case_pars("matern0.5", False, 0.5)
W = np.column_stack([ds["XData"], ds["zData"]])
sm = iak.SetupMats(ds["xData"], ds["dIData"], W=W, ctx=ctx)
keys = ["cx0", "cx1", "cd1", "cxd0", "cxd1", "cme"][:ncomp]
th = np.log([max(P[k], 1e-3) for k in keys]) + 0.1 * np.arange(ncomp)
got = iak.gradHessIAK3D(sm, P, modelx, types, th)
osm = orc.setupIAK3D(ds["xData"], ds["dIData"])
dA = orc.component_matrices(P, modelx, types, osm)[:ncomp]
want = orc.gradHessIAK3D(dA, th, W)
np.set_printoptions(linewidth=200, precision=5)
print("fim got\n", got["fim"], "\nfim want\n", want["fim"])
print("hess got\n", got["hess"], "\nhess want\n", want["hess"])
print("q got", got["hess"] + got["fim"], "\nq want", want["hess"] + want["fim"])
