"""Rendezvous plumbing for one-process-per-GPU launches under torchrun (bench.py, multi-GPU tests).

torch.distributed (gloo, CPU) is used only to hand the 128-byte NCCL unique id from rank 0 to the other ranks, for
barriers and for the max-over-ranks of timings.  The data path's single collective (the gradient all-reduce) is NCCL
inside libdtxtune, never torch."""
from __future__ import annotations

import os
from typing import Callable, Optional


class Rendezvous:
    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29531")
            if not dist.is_initialized():
                dist.init_process_group("gloo", rank=self.rank, world_size=self.world)
            self.dist = dist

    def broadcast_bytes(self, make: Callable[[], bytes]) -> Optional[bytes]:
        """rank 0 calls make(); every rank gets the bytes. world == 1 -> None (no id needed)."""
        if self.dist is None:
            return None
        box = [make() if self.rank == 0 else None]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]

    def barrier(self) -> None:
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, x: float) -> float:
        if self.dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])

    def mean_over_ranks(self, x: float) -> float:
        if self.dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t[0]) / self.world

    def sum_over_ranks(self, x: float) -> float:
        if self.dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t[0])

    def gather_objects(self, obj) -> list:
        """Every rank's `obj`, in rank order, on every rank."""
        if self.dist is None:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def close(self) -> None:
        if self.dist is not None and self.dist.is_initialized():
            self.dist.destroy_process_group()
