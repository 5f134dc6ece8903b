"""GPU: the exchange step over peer memory (csrc/peer.cu).  The box has one GPU, so the ranks are processes that share cuda:0
and map each other's mailboxes through CUDA IPC -- the same code path as one process per GPU over NVLink.
 * many all-reduces in a row (both mailbox banks, ranks arriving at different times): every rank must hold the SAME bits, equal
   to the sum taken in rank order;
 * the composite likelihood through it equals the host exchange (gloo) bit for bit and the oracle within 1e-8."""
import ctypes as C
import os
import socket
import sys
import time

import numpy as np
import pytest
import torch.multiprocessing as mp

from iak3d_b200 import synth
from oracle import iak3d_oracle as orc
from helpers import PARBNDS, case_pars

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(rank, world, port):
    os.environ.update(RANK=str(rank), LOCAL_RANK="0", WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import iak3d_b200 as iak
    from iak3d_b200 import parallel
    parallel.init(backend="gloo")
    return iak, parallel.BlockComm(), iak.Context(0)


def _stress_worker(rank, world, port, rounds, q):
    iak, comm, ctx = _setup(rank, world, port)
    ok = comm.attach_peer(ctx, max_doubles=512)
    import torch
    same = True
    if ok:
        rng = np.random.default_rng(7)
        for it in range(rounds):
            k = int(rng.integers(1, 513))
            vals = rng.standard_normal((world, k)) * 10.0 ** rng.integers(-8, 8, size=(world, k))
            want = np.zeros(k)
            for r in range(world):                                  # rank order, like the kernel
                want = want + vals[r]
            buf = comm.device_buffer(k)
            buf[:k] = torch.from_numpy(vals[rank]).cuda()
            if (it + rank) % 7 == 0:
                time.sleep(0.002 * (rank + 1))                      # ranks arrive apart
            got = comm.allreduce_device(k)
            same = same and np.array_equal(got, want) and np.array_equal(buf[:k].cpu().numpy(), want)
    q.put((rank, bool(ok), bool(same)))
    comm.barrier()
    comm.close(); ctx.close()
    import torch.distributed as dist
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_allreduce_is_exact_and_identical_on_every_rank(world):
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_stress_worker, args=(r, world, port, 300, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "the ranks could not map each other's mailboxes"
    assert all(r[2] for r in res)


def _cl_worker(rank, world, port, kind, q):
    os.environ["IAK3D_ALLREDUCE"] = kind                            # "peer" | "nccl" (= no peers: host sums over gloo here)
    iak, comm, ctx = _setup(rank, world, port)
    from iak3d_b200 import synth
    from helpers import PARBNDS, case_pars
    ds = synth.make_dataset(1500, 61)
    P, modelx, types, cmeOpt = case_pars("edgeroi", False, 0.5)
    blk = synth.make_blocks(ds["xData"], ds["prof"], 7, 61)
    blk["compLikOptn"] = 2
    cl = iak.setupIAK3D_CL(ds["xData"], ds["dIData"], ds["zData"], ds["XData"], blk, ctx=ctx, rank=rank, world=world)
    tm = {}
    vals = [float(iak.nllIAK3D_CL(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, False, cl, parsBTfmd=P, comm=comm,
                                  timings=tm)) for _ in range(3)]
    q.put((rank, vals, tm.get("allreduce_kind"), tm.get("allreduce_us", 0.0) / max(tm.get("allreduce_calls", 1), 1)))
    comm.barrier()
    comm.close(); ctx.close()
    import torch.distributed as dist
    dist.destroy_process_group()


def test_cl_through_peer_memory_equals_host_exchange():
    ds = synth.make_dataset(1500, 61)
    P, modelx, types, cmeOpt = case_pars("edgeroi", False, 0.5)
    blk = synth.make_blocks(ds["xData"], ds["prof"], 7, 61)
    blk["compLikOptn"] = 2
    want = orc.nllIAK3D_CL(None, ds["zData"], ds["XData"], modelx, 0.5, *types, cmeOpt, True, PARBNDS, False, blk, ds["xData"], ds["dIData"],
                           parsBTfmd=P)
    out = {}
    for kind in ("peer", "nccl"):
        mpc = mp.get_context("spawn")
        q = mpc.Queue()
        port = _free_port()
        procs = [mpc.Process(target=_cl_worker, args=(r, 2, port, kind, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = sorted(q.get(timeout=600) for _ in range(2))
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
        assert res[0][1] == res[1][1] and len(set(res[0][1])) == 1       # same on both ranks, same on every call
        out[kind] = res
    assert out["peer"][0][2] == "peer" and out["nccl"][0][2] is None      # the peer run really went through the mailboxes
    assert out["peer"][0][1][0] == out["nccl"][0][1][0]                   # rank-ordered sums: the same bits as the host exchange
    assert abs(out["peer"][0][1][0] - want) <= 1e-8 * abs(want)
    print("peer all-reduce us per evaluation:", out["peer"][0][3], out["peer"][1][3])
