"""GPU: the composite likelihood's batched gradient and its multi-rank path.
 * gradnllIAK3D_CL sends (units x parameter sets) through batched launches with the per-set sums formed on the device and one
   exchange (fitIAK3D.R:1819-1911 forks npar + 1 evaluations of nllIAK3D_CL): it must equal the npar + 1 separate evaluations
   bit for bit, and the oracle's forward differences within the noise of a 1e-6 difference quotient.
 * two ranks (two processes, gloo, both on cuda:0 -- the box has one GPU) shard the units, sum on the device and exchange once:
   the result equals the single-rank evaluation and the oracle (compositeLikIAK3D.R:54-64)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import iak3d_b200 as iak
from iak3d_b200 import api, synth
from oracle import iak3d_oracle as orc
from helpers import PARBNDS, case_pars

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _edgeroi_vector(P):
    # ax, ad, cx0, cx1, cxd1, cme of the spherical / no-cd1 / cmeOpt = 1 model (setInitsIAK3D.R:584-604)
    return np.array([float(api.logitab(P["ax"], *PARBNDS["axBnds"])), float(api.logitab(P["ad"], *PARBNDS["adBnds"])), np.log(P["cx0"]),
                     np.log(P["cx1"]), float(api.logitab(P["cxd1"], *PARBNDS["sxd1Bnds"])), np.log(P["cme"])])


@pytest.mark.parametrize("optn,useReml", [(2, False), (1, True), (4, True)])
def test_cl_gradient_batch_equals_separate_evaluations(ctx, optn, useReml):
    ds = synth.make_dataset(1500, 61)
    P, modelx, types, cmeOpt = case_pars("edgeroi", False, 0.5)
    blk = synth.make_blocks(ds["xData"], ds["prof"], 7, 61)
    blk["compLikOptn"] = optn
    cl = iak.setupIAK3D_CL(ds["xData"], ds["dIData"], ds["zData"], ds["XData"], blk, ctx=ctx)
    pars = _edgeroi_vector(P)
    args = (ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, useReml, cl)
    g, v0 = iak.gradnllIAK3D_CL(pars, *args, rtn_nll=True)
    vals = [iak.nllIAK3D_CL(pars, *args)]
    for k in range(pars.size):
        v = pars.copy(); v[k] += 1e-6
        vals.append(iak.nllIAK3D_CL(v, *args))
    assert v0 == vals[0]
    assert np.array_equal(g, (np.array(vals[1:]) - vals[0]) / 1e-6)
    want0 = orc.nllIAK3D_CL(pars, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, useReml, blk, ds["xData"], ds["dIData"])
    assert abs(v0 - want0) <= 1e-8 * abs(want0)
    gw = []
    for k in range(pars.size):
        v = pars.copy(); v[k] += 1e-6
        gw.append((orc.nllIAK3D_CL(v, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, useReml, blk, ds["xData"],
                                   ds["dIData"]) - want0) / 1e-6)
    assert np.max(np.abs(g - np.array(gw))) <= 2e-3 * max(np.max(np.abs(gw)), 1.0)     # 1e-6 differences of values good to ~1e-10 relative


def test_cl_failed_unit_gives_badnll(ctx):
    ds = synth.make_dataset(900, 61)
    P, modelx, types, cmeOpt = case_pars("edgeroi", False, 0.5)
    P = dict(P); P["cx1"] = float("inf")
    blk = synth.make_blocks(ds["xData"], ds["prof"], 5, 61)
    blk["compLikOptn"] = 2
    cl = iak.setupIAK3D_CL(ds["xData"], ds["dIData"], ds["zData"], ds["XData"], blk, ctx=ctx)
    assert iak.nllIAK3D_CL(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, False, cl, parsBTfmd=P) == api.BADNLL


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, optn, useReml, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK="0", WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import iak3d_b200 as iak
    from iak3d_b200 import parallel, synth
    from helpers import PARBNDS, case_pars
    parallel.init(backend="gloo")
    comm = parallel.BlockComm()
    ctx = iak.Context(0)
    ds = synth.make_dataset(1500, 61)
    P, modelx, types, cmeOpt = case_pars("edgeroi", False, 0.5)
    blk = synth.make_blocks(ds["xData"], ds["prof"], 7, 61)
    blk["compLikOptn"] = optn
    cl = iak.setupIAK3D_CL(ds["xData"], ds["dIData"], ds["zData"], ds["XData"], blk, ctx=ctx, rank=rank, world=world)
    v = iak.nllIAK3D_CL(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, useReml, cl, parsBTfmd=P, comm=comm)
    q.put((rank, float(v), len(cl["units"]), int(cl["n_units_total"])))
    comm.barrier()
    ctx.close()
    import torch.distributed as dist
    dist.destroy_process_group()


@pytest.mark.parametrize("optn,useReml", [(2, False), (1, True)])
def test_cl_two_ranks_share_one_gpu(ctx, optn, useReml):
    ds = synth.make_dataset(1500, 61)
    P, modelx, types, cmeOpt = case_pars("edgeroi", False, 0.5)
    blk = synth.make_blocks(ds["xData"], ds["prof"], 7, 61)
    blk["compLikOptn"] = optn
    want = orc.nllIAK3D_CL(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, useReml, blk, ds["xData"],
                           ds["dIData"], parsBTfmd=P)
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_worker, args=(r, 2, port, optn, useReml, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1]                                   # both ranks hold the same value
    assert abs(res[0][1] - want) <= 1e-8 * abs(want)
    assert res[0][2] + res[1][2] == res[0][3] and min(res[0][2], res[1][2]) >= 1


@pytest.mark.parametrize("optn,useReml", [(2, False), (1, True)])
def test_cl_several_contexts_in_one_process(ctx, optn, useReml):
    """What an R session can do: several device contexts driven from one process (a host thread per device inside the library,
    iak3d_nll_terms_batch_wsum_multi).  The box has one GPU: both contexts sit on cuda:0."""
    ds = synth.make_dataset(1500, 61)
    P, modelx, types, cmeOpt = case_pars("edgeroi", False, 0.5)
    blk = synth.make_blocks(ds["xData"], ds["prof"], 7, 61)
    blk["compLikOptn"] = optn
    import torch
    ctx2 = iak.Context(1 if torch.cuda.device_count() > 1 else 0)       # a second GPU when the box has one
    try:
        cl1 = iak.setupIAK3D_CL(ds["xData"], ds["dIData"], ds["zData"], ds["XData"], blk, ctx=ctx)
        cl2 = iak.setupIAK3D_CL(ds["xData"], ds["dIData"], ds["zData"], ds["XData"], blk, ctx=[ctx, ctx2])
        assert len({id(u["setup"].ctx) for u in cl2["units"]}) == 2
        args = (None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, useReml)
        v1 = iak.nllIAK3D_CL(*args, cl1, parsBTfmd=P)
        v2 = iak.nllIAK3D_CL(*args, cl2, parsBTfmd=P)
        want = orc.nllIAK3D_CL(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, useReml, blk, ds["xData"],
                               ds["dIData"], parsBTfmd=P)
        assert abs(v2 - want) <= 1e-8 * abs(want) and abs(v2 - v1) <= 1e-11 * abs(v1)
        pars = _edgeroi_vector(P)
        g1 = iak.gradnllIAK3D_CL(pars, *args[1:], cl1)
        g2 = iak.gradnllIAK3D_CL(pars, *args[1:], cl2)
        assert np.max(np.abs(g1 - g2)) <= 1e-3 * max(1.0, np.max(np.abs(g1)))
    finally:
        for u in cl2["units"]:
            u["setup"].close()
        ctx2.close()
