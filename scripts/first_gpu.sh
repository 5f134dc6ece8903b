set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
timeout 300 python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import iak3d_b200 as iak
from iak3d_b200 import synth
from oracle import iak3d_oracle as orc
from helpers import PARBNDS, case_pars
ctx = iak.Context(0)
print(ctx.device_info())
print('peak DFMA', ctx.fp64_peak(0), 'DMMA', ctx.fp64_peak(1))
for n in (50, 200, 1000):
    ds = synth.make_dataset(n, 41)
    P, modelx, types, cmeOpt = case_pars('matern0.5', False, 0.5)
    sm = iak.setupIAK3D(ds['xData'], ds['dIData'], ctx=ctx)
    t=time.time()
    got = iak.nllIAK3D(None, ds['zData'], ds['XData'], modelx=modelx, nud=0.5, cmeOpt=cmeOpt, setupMats=sm, parBnds=PARBNDS, forCompLik=True, parsBTfmd=P)
    print(n, 'gpu time', time.time()-t)
    want = orc.nllIAK3D(None, ds['zData'], ds['XData'], modelx, 0.5, *types, cmeOpt, True, orc.setupIAK3D(ds['xData'], ds['dIData']), PARBNDS, forCompLik=True, parsBTfmd=P)
    print(n, got['lndetA'], want['lndetA'], np.max(np.abs(got['WiAW']-want['WiAW']))/np.max(np.abs(want['WiAW'])))
PY
echo EXIT $?
