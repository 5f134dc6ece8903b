"""Newton-Raphson terms (gradHessIAK3D, nrUpdatesIAK3D.R:118-216) of one n-interval block over N GPUs:
torchrun --nproc-per-node N scripts/nr_shard.py [n].  Every rank holds the dataset; the six T dC_i products and the 21 pair
traces are cut by tile-pair ownership (iak3d_gradhess_shard), one all-reduce of 42 doubles joins them.  Prints one JSON line
(rank 0): wall time of the call = max over ranks, the single-GPU time of the same call when N == 1."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, '.')
import iak3d_b200 as iak
from iak3d_b200 import parallel, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
rank, local_rank, world = parallel.init()
comm = parallel.BlockComm()
ctx = iak.Context(local_rank)
ds = synth.make_dataset(n, synth.SEEDS["S40"])
P, modelx, types, cmeOpt = synth.default_pars("matern0.5")
W = np.column_stack([ds["XData"], ds["zData"]])
sm = iak.SetupMats(ds["xData"], ds["dIData"], W=W, ctx=ctx)
th = np.log([P[k] for k in ("cx0", "cx1", "cd1", "cxd0", "cxd1", "cme")])
times = []
for rep in range(2):
    comm.barrier(); ctx.sync()
    t0 = time.perf_counter()
    g = iak.gradHessIAK3D(sm, P, modelx, types, th, comm=comm if world > 1 else None)
    ctx.sync()
    dt = float(comm.allreduce_max(np.array([time.perf_counter() - t0]))[0])
    times.append(dt)
flops = 3 * n ** 3 / 3.0 + 6 * 2.0 * n ** 3
if rank == 0:
    print(json.dumps(dict(metric="gradHessIAK3D calls/sec at n obs", n=n, n_gpus=world, wall_s=min(times), walls=times,
                          value=1.0 / min(times), unit="calls/s", algorithmic_tflops=flops / min(times) / 1e12, nll=g["nll"],
                          grad=[float(x) for x in g["grad"]], fim_diag=[float(x) for x in np.diag(g["fim"])],
                          note="chol + inverse (n^3 flops) on every rank, 12 n^3 flops of products and traces split by tile-pair "
                               "ownership, one all-reduce of 42 doubles")), flush=True)
comm.barrier()
ctx.close()
if world > 1:
    import torch.distributed as dist
    dist.destroy_process_group()
