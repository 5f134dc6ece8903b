"""Multi-GPU plumbing: one process per GPU (torchrun), `torch.distributed` for the single exchange step of the
path -- the sum of the composite-likelihood terms (compositeLikIAK3D.R:63-64: Reduce('+') over the mcmapply
workers) -- and contiguous tiling of prediction grids (no collective in the data path).

Backend: NCCL over NVLink on GPUs for the plumbing (rendezvous, barriers, bench records); gloo for the CPU tests of the host
logic.  The exchange step itself runs through the library's own peer-memory kernel (csrc/peer.cu: mailboxes mapped over CUDA
IPC, one single-CTA kernel per rank) once `BlockComm.attach_peer` has connected the ranks; NCCL's all-reduce stays as the
fallback (and as the comparison: IAK3D_ALLREDUCE=nccl).
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

    def attach_peer(self, ctx, max_doubles=8192):
        """Connect this rank's library context with its peers' for the peer-memory all-reduce (collective: every rank calls it
        once, with its context or None).  Returns True when ALL ranks are connected; otherwise every rank keeps NCCL."""
        if getattr(self, "_peer_tried", False):
            return getattr(self, "_peer", None) is not None
        self._peer_tried = True
        self._peer = None
        import torch
        import torch.distributed as dist
        if self.world == 1 or os.environ.get("IAK3D_ALLREDUCE", "peer") == "nccl":
            return False
        import ctypes as C
        # the handles travel over whatever the process group is (NCCL wants device tensors, gloo host tensors)
        dev = torch.device("cuda", self.local_rank if self._dev is None else self._dev) if dist.get_backend() == "nccl" else torch.device("cpu")
        mine = torch.zeros(65, dtype=torch.uint8)
        peer = C.c_void_p()
        if ctx is not None:
            hbuf = (C.c_ubyte * 64)()
            if ctx.lib.iak3d_peer_create(ctx.h, self.rank, self.world, int(max_doubles), C.byref(peer), hbuf) == 0:
                mine[:64] = torch.frombuffer(bytearray(hbuf), dtype=torch.uint8)
                mine[64] = 1
        allh = [torch.zeros(65, dtype=torch.uint8, device=dev) for _ in range(self.world)]
        dist.all_gather(allh, mine.to(dev))
        allh = torch.stack(allh).cpu()
        ok = bool(allh[:, 64].all())
        if ok:
            raw = (C.c_ubyte * (64 * self.world)).from_buffer_copy(allh[:, :64].contiguous().numpy().tobytes())
            ok = ctx.lib.iak3d_peer_connect(peer, raw) == 0
        agree = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        if int(agree.item()) == 1:
            self._peer, self._peer_ctx, self._peer_cap = peer, ctx, int(max_doubles)
            return True
        if peer.value:
            ctx.lib.iak3d_peer_destroy(peer)
        return False

    def close(self):
        if getattr(self, "_peer", None) is not None:
            self._peer_ctx.lib.iak3d_peer_destroy(self._peer)
            self._peer = None

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
        """Persistent float64 CUDA tensor of at least k elements on this rank's GPU (None when the ranks have neither NCCL nor
        connected peer mailboxes): the library writes the composite-likelihood sums straight into it
        (iak3d_nll_terms_batch_wsum) and the exchange reduces it in place -- no host staging before it."""
        import torch
        import torch.distributed as dist
        if self.world == 1 or (dist.get_backend() != "nccl" and getattr(self, "_peer", None) is None):
            return None
        if getattr(self, "_dbuf", None) is None or self._dbuf.numel() < k:
            self._dbuf = torch.zeros(max(int(k), 8192), dtype=torch.float64, device=torch.device("cuda", self.local_rank if self._dev is None else self._dev))
            self._hbuf = torch.zeros(self._dbuf.numel(), dtype=torch.float64).pin_memory()
            self._ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        return self._dbuf

    def allreduce_device(self, k, timings=None):
        """In-place all-reduce (sum) of the first k elements of the device buffer, then ONE device-to-host copy into pinned
        memory: the library's peer-memory kernel when the ranks are connected (attach_peer), NCCL otherwise.  `timings` (a
        dict) accumulates the device time of the exchange in `allreduce_us`."""
        import torch
        import torch.distributed as dist
        t = self._dbuf[:k]
        if getattr(self, "_peer", None) is None and dist.get_backend() != "nccl":
            raise RuntimeError("allreduce_device needs NCCL or connected peers")
        if getattr(self, "_peer", None) is not None and k <= self._peer_cap:
            import ctypes as C
            torch.cuda.current_stream().synchronize()         # torch may have filled the buffer (a rank without units)
            out = np.empty(k)
            us = C.c_double()
            ctx = self._peer_ctx
            ctx.check(ctx.lib.iak3d_peer_allreduce(self._peer, C.c_void_p(t.data_ptr()), k, out.ctypes.data_as(C.POINTER(C.c_double)), C.byref(us)))
            if timings is not None:
                timings["allreduce_us"] = timings.get("allreduce_us", 0.0) + us.value
                timings["allreduce_calls"] = timings.get("allreduce_calls", 0) + 1
                timings["allreduce_kind"] = "peer"
            return out
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
            timings["allreduce_kind"] = "nccl"
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
