"""The N > 1 host logic on CPU: two processes, gloo.  Each rank plans the composite-likelihood units, takes its
share, and combines the per-unit terms with the package's own exchange step (cl_reduce_and_tail + BlockComm); the
per-unit terms themselves come from the oracle here (on a GPU box they come from libiak3d -- tests/test_gpu_predict.py)
so that this test checks exactly the sharding, the weights and the all-reduce."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, optn, useReml, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import iak3d_b200 as iak
    from iak3d_b200 import parallel, synth
    from oracle import iak3d_oracle as orc
    from helpers import PARBNDS, case_pars
    parallel.init(backend="gloo")
    comm = parallel.BlockComm()
    ds = synth.make_dataset(900, 61)
    P, modelx, types, cmeOpt = case_pars("edgeroi", False, 0.5)
    blk = synth.make_blocks(ds["xData"], ds["prof"], 6, 61)
    blk["compLikOptn"] = optn
    n, p = ds["XData"].shape
    units = iak.cl_plan_units(blk, world)
    S = np.zeros((p + 1, p + 1)); lnd = 0.0; n_SUM = 0.0
    mine = 0
    for u in units:
        if u["owner"] != rank:
            continue
        mine += 1
        rows = u["rows"]
        t = orc.nllIAK3D(None, ds["zData"][rows], ds["XData"][rows], modelx, 0.5, *types, cmeOpt, True,
                         orc.setupIAK3D(ds["xData"][rows], ds["dIData"][rows]), PARBNDS, forCompLik=True, parsBTfmd=P)
        S += u["weight"] * t["WiAW"]; lnd += u["weight"] * t["lndetA"]
        if u["kind"] == "pair":
            n_SUM += len(rows)
    got = iak.cl_reduce_and_tail(S, lnd, n_SUM, 0.0, n, p, optn, len(blk["listBlocks"]), useReml, comm, rtnTerms=True)
    want = orc.nllIAK3D_CL(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, useReml, blk, ds["xData"],
                           ds["dIData"], parsBTfmd=P, rtnTerms=True)
    ok = (abs(got["nll"] - want["nll"]) <= 1e-10 * abs(want["nll"])
          and np.allclose(got["XziAXz_SUM"], want["XziAXz_SUM"], rtol=1e-10, atol=0)
          and abs(got["n_SUM"] - want["n_SUM"]) < 0.5)
    # the time reduction bench.py uses
    mx = comm.allreduce_max(np.array([float(rank + 1)]))[0]
    q.put((rank, bool(ok), mine, len(units), float(mx), got["nll"]))
    comm.barrier()
    import torch.distributed as dist
    dist.destroy_process_group()


@pytest.mark.parametrize("optn,useReml", [(2, False), (1, True), (4, True)])
def test_composite_likelihood_sharded_over_two_ranks(optn, useReml):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, optn, useReml, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(r[1] for r in res), res
    assert res[0][2] + res[1][2] == res[0][3] and min(res[0][2], res[1][2]) >= 1      # every unit owned exactly once
    assert res[0][4] == 2.0 and res[1][4] == 2.0                                       # max over ranks
    assert res[0][5] == res[1][5]                                                      # both ranks hold the same nll


def test_unit_plan_balances_and_covers():
    sys.path.insert(0, ROOT)
    import iak3d_b200 as iak
    from iak3d_b200 import synth
    ds = synth.make_dataset(3000, 7)
    blk = synth.make_blocks(ds["xData"], ds["prof"], 12, 7)
    for optn in (1, 2, 3, 4):
        blk["compLikOptn"] = optn
        for world in (1, 2, 4, 8):
            units = iak.cl_plan_units(blk, world)
            load = np.zeros(world)
            for u in units:
                load[u["owner"]] += len(u["rows"]) ** 3
            assert set(u["owner"] for u in units) <= set(range(world))
            if len(units) >= 4 * world:
                assert load.max() <= 1.6 * load.mean()
        npairs = len(blk["subsetPairsAdj"])
        if optn == 2:
            assert len(units) == npairs and all(u["kind"] == "pair" for u in units)
        if optn == 4:
            assert len(units) == 12 and all(u["weight"] == 1.0 for u in units)


def _peer_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from iak3d_b200 import parallel
    parallel.init(backend="gloo")
    comm = parallel.BlockComm()
    first = comm.attach_peer(None)                      # no device context on this machine: every rank must agree on "no"
    second = comm.attach_peer(None)                     # and the answer is remembered (the call is collective only once)
    s = comm.allreduce_sum(np.array([1.0 + rank, 10.0]))
    q.put((rank, first, second, comm.device_buffer(8) is None, list(s)))
    comm.barrier()
    import torch.distributed as dist
    dist.destroy_process_group()


def test_peer_exchange_declines_without_device_contexts():
    """The peer-memory exchange is negotiated collectively; ranks without a CUDA context (this CPU box, or a rank that owns
    no unit) make every rank fall back to the process group's all-reduce."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert r[1] is False and r[2] is False and r[3] is True and r[4] == [3.0, 20.0]
