"""
Multi-GPU plumbing: one process per GPU, chromosomes sharded across ranks, no data-path exchange.

The only collective on the path is the final reduction of per-chromosome counters (hit counts / covered bases):
``Comm.allreduce_sum_i64`` -- NCCL over NVLink through libbxb200 (``bxg_comm_*``) on the GPU box, or gloo on CPU
for the world_size-2 host-logic tests.  The NCCL backend needs no PyTorch: the 128-byte ncclUniqueId travels from
rank 0 to the other ranks of the node over an abstract unix-domain socket (``exchange_id``); torch is imported only
by the 'gloo' backend, i.e. by the CPU tests.
"""
from __future__ import annotations

import ctypes as C
import os
import socket
import time

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


def rendezvous_name():
    """Name shared by the ranks of ONE launch on one node: the launcher's pid (torchrun is the parent of every local
    rank) plus MASTER_PORT; BXB200_RENDEZVOUS overrides it for other launchers."""
    return os.environ.get("BXB200_RENDEZVOUS") or "bxb200-%d-%s" % (os.getppid(), os.environ.get("MASTER_PORT", "0"))


def exchange_id(payload, rank, world, name=None, timeout=120.0):
    """Broadcast `payload` (bytes, significant on rank 0) to all ranks of this node: rank 0 listens on an abstract
    unix-domain socket (no file, vanishes with the process -- nothing stale to trip over) and serves world-1 peers."""
    if world <= 1:
        return payload
    addr = "\0" + (name or rendezvous_name())
    if rank == 0:
        with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as srv:
            srv.bind(addr)
            srv.listen(world)
            srv.settimeout(timeout)
            for _ in range(world - 1):
                conn, _ = srv.accept()
                with conn:
                    conn.sendall(len(payload).to_bytes(4, "little") + payload)
        return payload
    deadline = time.monotonic() + timeout
    while True:
        try:
            with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as c:
                c.connect(addr)
                c.settimeout(timeout)
                buf = b""
                while len(buf) < 4 or len(buf) < 4 + int.from_bytes(buf[:4], "little"):
                    chunk = c.recv(4096)
                    if not chunk:
                        raise ConnectionError("rendezvous peer closed early")
                    buf += chunk
                return buf[4:4 + int.from_bytes(buf[:4], "little")]
        except (ConnectionRefusedError, FileNotFoundError):
            if time.monotonic() > deadline:
                raise TimeoutError("rank 0 never opened the rendezvous socket %r" % addr[1:])
            time.sleep(0.05)


class Comm:
    """backend 'nccl' (libbxb200 + NCCL, GPU; no torch) | 'gloo' (torch.distributed on CPU, tests) | 'single'."""

    def __init__(self, backend=None):
        self.rank, self.world, self.local_rank = env_rank()
        self._dist = None
        if self.world == 1:
            self.backend = "single"
            return
        self.backend = backend or "nccl"
        if self.backend == "nccl":
            from . import _lib
            L = _lib.lib()
            buf = C.create_string_buffer(128)
            if self.rank == 0:
                _lib.check(L.bxg_comm_unique_id(buf))
            uid = exchange_id(bytes(buf.raw), self.rank, self.world)
            _lib.check(L.bxg_comm_init(uid, self.world, self.rank))
            self._L, self._check = L, _lib.check
        else:
            import torch.distributed as dist
            self._dist = dist
            if not dist.is_initialized():
                dist.init_process_group(backend="gloo")

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
        if self._dist is not None and self._dist.is_initialized():
            self._dist.destroy_process_group()
