"""
Multi-GPU plumbing: one process per GPU, chromosomes sharded across ranks, no data-path exchange.

The only collective on the path is the final reduction of per-chromosome counters (hit counts / covered bases):
``Comm.allreduce_sum_i64`` -- NCCL over NVLink through libbxb200 (``bxg_comm_*``) on the GPU box, or gloo on CPU
for the world_size-2 host-logic tests.  Rendezvous (shipping the 128-byte ncclUniqueId) uses torch.distributed's
store when launched by torchrun; torch is plumbing only and is imported only when WORLD_SIZE > 1.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np


def lpt_assign(weights, nranks):
    """Longest-processing-time greedy: heaviest unit first onto the least loaded rank.
    -> list (per rank) of unit indices, in ascending unit order."""
    w = np.asarray(weights, np.float64)
    load = np.zeros(nranks)
    shards = [[] for _ in range(nranks)]
    for u in np.argsort(-w, kind="stable"):
        r = int(np.argmin(load))
        shards[r].append(int(u))
        load[r] += w[u]
    return [sorted(s) for s in shards]


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


class Comm:
    """backend 'nccl' (libbxb200 + NCCL, GPU) | 'gloo' (torch.distributed on CPU, tests) | 'single'."""

    def __init__(self, backend=None):
        self.rank, self.world, self.local_rank = env_rank()
        if self.world == 1:
            self.backend = "single"
            return
        self.backend = backend or "nccl"
        import torch.distributed as dist
        self._dist = dist
        if not dist.is_initialized():
            dist.init_process_group(backend="gloo")     # CPU store/bootstrap only; the data path never uses it
        if self.backend == "nccl":
            from . import _lib
            L = _lib.lib()
            buf = C.create_string_buffer(128)
            if self.rank == 0:
                _lib.check(L.bxg_comm_unique_id(buf))
            ids = [bytes(buf.raw)]
            dist.broadcast_object_list(ids, src=0)
            _lib.check(L.bxg_comm_init(ids[0], self.world, self.rank))
            self._L, self._check = L, _lib.check

    def allreduce_sum_i64(self, a):
        a = np.ascontiguousarray(a, np.int64)
        if self.backend == "single":
            return a
        if self.backend == "nccl":
            self._check(self._L.bxg_comm_allreduce_i64(a.ctypes.data_as(C.c_void_p), a.size))
            return a
        import torch
        t = torch.from_numpy(a)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM)
        return a

    def allreduce_max_f64(self, a):
        a = np.ascontiguousarray(a, np.float64)
        if self.backend == "single":
            return a
        if self.backend == "nccl":
            self._check(self._L.bxg_comm_allreduce_max_f64(a.ctypes.data_as(C.c_void_p), a.size))
            return a
        import torch
        t = torch.from_numpy(a)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX)
        return a

    def barrier(self):
        if self.backend == "single":
            return
        if self.backend == "nccl":
            self._check(self._L.bxg_comm_barrier())
        else:
            self._dist.barrier()

    def close(self):
        if self.backend == "nccl":
            self._L.bxg_comm_destroy()
        if self.backend != "single" and self._dist.is_initialized():
            self._dist.destroy_process_group()
