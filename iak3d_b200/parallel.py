"""Multi-GPU plumbing: one process per GPU (torchrun), `torch.distributed` for the single exchange step of the
path -- the sum of the composite-likelihood terms (compositeLikIAK3D.R:63-64: Reduce('+') over the mcmapply
workers) -- and contiguous tiling of prediction grids (no collective in the data path).

Backend: NCCL over NVLink on GPUs; gloo for the CPU tests of the host logic.
"""
from __future__ import annotations

import os

import numpy as np


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment; returns (rank, local_rank, world)."""
    rank, local_rank, world = dist_env()
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if backend == "nccl":
                torch.cuda.set_device(local_rank)
                dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
            else:
                dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


class BlockComm:
    """Sum / max / gather of small float64 host vectors across ranks (identity when world == 1)."""

    def __init__(self, device=None):
        self.rank, self.local_rank, self.world = dist_env()
        self._dev = device
        if self.world > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                raise RuntimeError("BlockComm: call parallel.init() first")

    def _tensor(self, buf):
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(np.ascontiguousarray(buf, dtype=np.float64).copy())
        if dist.get_backend() == "nccl":
            t = t.cuda(self.local_rank if self._dev is None else self._dev)
        return t

    def allreduce_sum(self, buf):
        if self.world == 1:
            return np.asarray(buf, dtype=np.float64)
        import torch.distributed as dist
        t = self._tensor(buf)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def device_buffer(self, k):
        """Persistent float64 CUDA tensor of at least k elements on this rank's GPU (NCCL only; None otherwise): the
        library writes the composite-likelihood sums straight into it (iak3d_nll_terms_batch_wsum) and NCCL reduces it in
        place -- no host staging before the exchange."""
        import torch
        import torch.distributed as dist
        if self.world == 1 or dist.get_backend() != "nccl":
            return None
        if getattr(self, "_dbuf", None) is None or self._dbuf.numel() < k:
            self._dbuf = torch.zeros(max(int(k), 1024), dtype=torch.float64, device=torch.device("cuda", self.local_rank if self._dev is None else self._dev))
            self._hbuf = torch.zeros(self._dbuf.numel(), dtype=torch.float64).pin_memory()
            self._ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        return self._dbuf

    def allreduce_device(self, k, timings=None):
        """In-place NCCL all-reduce (sum) of the first k elements of the device buffer, then ONE device-to-host copy into
        pinned memory.  `timings` (a dict) accumulates the device time of the collective in `allreduce_us`."""
        import torch
        import torch.distributed as dist
        t = self._dbuf[:k]
        if timings is not None:
            self._ev[0].record()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        if timings is not None:
            self._ev[1].record()
        self._hbuf[:k].copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        if timings is not None:
            timings["allreduce_us"] = timings.get("allreduce_us", 0.0) + 1e3 * self._ev[0].elapsed_time(self._ev[1])
            timings["allreduce_calls"] = timings.get("allreduce_calls", 0) + 1
        return self._hbuf[:k].numpy().copy()

    def allreduce_max(self, buf):
        if self.world == 1:
            return np.asarray(buf, dtype=np.float64)
        import torch.distributed as dist
        t = self._tensor(buf)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.cpu().numpy()

    def gather_scalar(self, x):
        """x of every rank, in rank order (one-hot all-reduce: bench.py's per-rank records)"""
        v = np.zeros(self.world)
        v[self.rank] = float(x)
        return self.allreduce_sum(v)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()


def grid_tile(n_cells, rank, world, align=128):
    """Contiguous slice [lo, hi) of a prediction grid for this rank; slices are multiples of `align` cells except
    the last (row blocks of the kriging panel are 128 supports)."""
    per = -(-n_cells // world)
    per = -(-per // align) * align
    lo = min(rank * per, n_cells)
    hi = min(lo + per, n_cells)
    return lo, hi
