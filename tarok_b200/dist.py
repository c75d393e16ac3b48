"""Multi-GPU: independent game shards, one collective.

Games never interact, so N GPUs = N disjoint contiguous ranges of GLOBAL game ids, one process per
GPU.  Every random draw is keyed by the global game id (Philox counter), and seat rotation depends only
on it (Tarok.py:34), so the set of games played -- and therefore every statistic -- is identical for
any GPU count.  The only exchange on the path is the all-reduce of the 32-entry statistics vector
(per-player returns = ``Tarok.rezultati``, Tarok.py:59-61; contract histogram; counters) after scoring:
``torch.distributed`` all_reduce, i.e. NCCL over NVLink on a B200 box (gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard(total_games: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range of global game ids owned by ``rank``: (first id, count).  Ranges tile [0, total)."""
    base, rem = divmod(int(total_games), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def bind_to_gpu_numa(device_index: int) -> bool:
    """Pin this process to the CPUs NVML reports as local to the GPU, so that pinned host buffers (and the
    upload/download DMA of the host-buffer entry point) stay on the GPU's NUMA node.  Best effort."""
    try:
        import os
        import pynvml as N
        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(device_index)
        words = N.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return True
    except Exception:
        pass
    return False


def allreduce_stats(stats: torch.Tensor) -> torch.Tensor:
    """Sum the statistics vector over all ranks, in place; no-op without an initialised process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


class NcclComm:
    """A raw NCCL communicator (ncclComm_t) for ``tarok_allreduce_stats``.  With an initialised ``torch.distributed`` group
    rank 0's ncclUniqueId is broadcast over it and every rank calls ncclCommInitRank; without one (a single process) it is a
    one-rank communicator, so the C-ABI collective runs the same code path at every GPU count.  ctypes only."""

    def __init__(self, device: int):
        import ctypes as C

        class UniqueId(C.Structure):
            _fields_ = [("internal", C.c_byte * 128)]

        self._C, self._nccl = C, C.CDLL("libnccl.so.2")
        grouped = dist.is_available() and dist.is_initialized()
        rank, world = (dist.get_rank(), dist.get_world_size()) if grouped else (0, 1)
        uid = UniqueId()
        if rank == 0:
            self._ok(self._nccl.ncclGetUniqueId(C.byref(uid)))
        if grouped and world > 1:
            buf = torch.tensor(list(bytes(uid.internal)), dtype=torch.uint8, device=torch.device("cuda", device))
            dist.broadcast(buf, src=0)
            C.memmove(C.byref(uid), bytes(buf.cpu().tolist()), 128)
        self.comm = C.c_void_p()
        torch.cuda.set_device(device)
        self._nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
        self._ok(self._nccl.ncclCommInitRank(C.byref(self.comm), world, uid, rank))
        self.world = world

    def _ok(self, rc):
        if rc != 0:
            raise RuntimeError("NCCL error %d" % rc)

    def allreduce_stats(self, env, out: torch.Tensor = None) -> torch.Tensor:
        """Sum of ``env``'s statistics vector over all ranks (int64 [32] on the device), via the C ABI; enqueued on the
        current stream, no synchronisation."""
        if out is None:
            out = torch.empty(32, dtype=torch.int64, device=env.torch_device)
        env._check(env._lib.tarok_allreduce_stats(env._h, self.comm, self._C.c_void_p(out.data_ptr()), env._stream()))
        return out

    def close(self):
        if self.comm:
            self._nccl.ncclCommDestroy(self.comm)
            self.comm = self._C.c_void_p()


class ShardedTarok:
    """This rank's shard of a ``total_games`` batch on its own GPU (one process per GPU)."""

    def __init__(self, total_games: int, seed: int, rank: int = 0, world: int = 1, device: int = 0):
        from .env import TarokEnv

        self.total, self.rank, self.world = int(total_games), rank, world
        self.first, self.count = shard(total_games, rank, world)
        self.env = TarokEnv(self.count, seed=seed, device=device)
        self._glob = torch.zeros(32, dtype=torch.int64, device=self.env.torch_device)

    def rollout(self, mode: int, batch_index: int = 0, fused: bool = False) -> torch.Tensor:
        """Plays batch ``batch_index`` (global ids batch_index*total + [first, first+count)) and returns the
        statistics vector summed over all ranks (device tensor, int64 [32])."""
        self.env.reset_stats()
        self.env.rollout(mode, first_game_id=batch_index * self.total + self.first, fused=fused)
        self._glob.copy_(self.env.stats_dev)
        return allreduce_stats(self._glob)

    def close(self):
        self.env.close()
