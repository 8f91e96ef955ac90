"""Data-parallel plumbing: one process per GPU, torch.distributed over NCCL/NVLink.

The path shards on batch: every op is per-sample (InstanceNorm has no cross-sample statistic), so
  * inference = independent replicas, no collective;
  * training  = rank r takes samples [r*B/W, (r+1)*B/W) of lr / hr / label-noise; losses are means, so the
    global gradient is (1/W) * sum_r grad_r: ONE all-reduce(sum) per network per step on the flat fp32
    gradient buffer (D 18.7 MB after trainer.py:180, G 3.7 MB after :195), the 1/W folded into AdamW.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = "nccl") -> Tuple[int, int, int]:
    """torchrun contract: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT."""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, local, world


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    if n % world:
        raise ValueError(f"global batch {n} is not divisible by world size {world} (equal shards keep the mean exact)")
    per = n // world
    return rank * per, (rank + 1) * per


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    a, b = shard_range(t.shape[0], rank, world)
    return t[a:b]


def shard_noise(noise: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    return {k: shard_batch(v, rank, world) for k, v in noise.items()}


def allreduce_flat(flat_grad: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the flat gradient buffer over ranks in place (NCCL over NVLink on GPUs, gloo in CPU tests)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def broadcast_flat(flat: torch.Tensor, group=None, src: int = 0) -> torch.Tensor:
    """Rank `src`'s flat parameter / Adam-moment buffer to every rank (replica initialisation and resume)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src=src, group=group)
    return flat


def rank_and_world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


class FlatComm:
    """The gradient exchange of the GAN step: in-place sum of a flat fp32 buffer over the ranks.

    On CUDA it is libfsr_b200's own NCCL communicator (fsr_nccl_* in include/fsr_b200.h: ncclAllReduce on the CURRENT
    stream, so the call can sit inside the step's CUDA graph and on the side stream of the overlap window).  The
    128-byte NCCL id travels from rank 0 over the existing torch.distributed group (any backend).  FSR_NCCL_CAPI=0, a
    CPU device or a missing libnccl fall back to torch.distributed.all_reduce on the same group."""

    def __init__(self, device: torch.device, group=None):
        self.group = group
        self.rank, self.world = rank_and_world(group)
        self._h: Optional[ctypes.c_void_p] = None
        self._lib = None
        if self.world > 1 and device.type == "cuda" and os.environ.get("FSR_NCCL_CAPI", "1") != "0":
            from . import _lib as L
            lib = L.load()
            if lib.fsr_nccl_available():
                buf = ctypes.create_string_buffer(128)
                if self.rank == 0:
                    L.check(lib.fsr_nccl_unique_id(buf), "fsr_nccl_unique_id")
                obj = [buf.raw if self.rank == 0 else None]
                src = dist.get_global_rank(group, 0) if group is not None else 0
                dist.broadcast_object_list(obj, src=src, group=group)
                h = ctypes.c_void_p()
                with torch.cuda.device(device):
                    L.check(lib.fsr_nccl_init(obj[0], self.rank, self.world, ctypes.byref(h)), "fsr_nccl_init")
                self._h, self._lib, self._L = h, lib, L

    @property
    def native(self) -> bool:
        return self._h is not None

    def allreduce(self, flat: torch.Tensor) -> torch.Tensor:
        if self._h is not None:
            assert flat.dtype == torch.float32 and flat.is_contiguous()
            self._L.check(self._lib.fsr_nccl_allreduce(self._h, flat.data_ptr(), flat.numel(), self._L.stream_ptr(flat.device)),
                          "fsr_nccl_allreduce")
            return flat
        return allreduce_flat(flat, self.group)

    def broadcast(self, flat: torch.Tensor, src: int = 0) -> torch.Tensor:
        if self._h is not None:
            self._L.check(self._lib.fsr_nccl_broadcast(self._h, flat.data_ptr(), flat.numel(), src, self._L.stream_ptr(flat.device)),
                          "fsr_nccl_broadcast")
            return flat
        return broadcast_flat(flat, self.group, src)

    def close(self):
        if self._h is not None:
            self._lib.fsr_nccl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
