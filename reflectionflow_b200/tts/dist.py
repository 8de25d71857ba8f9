"""One process per GPU.  The outer loop needs exactly one exchange step per round:
an all-gather of fixed-size score records, plus an all-gather of the candidates' packed latents
(the parents of the next round).  NCCL on GPUs; gloo on CPU for the host-logic tests."""
from __future__ import annotations

import os
from typing import Any, List, Sequence, Tuple

import torch

from .search import gather_records


class DistCtx:
    def __init__(self, rank: int = 0, world: int = 1, device="cpu"):
        self.rank, self.world, self.device = rank, world, torch.device(device)

    @classmethod
    def from_env(cls, device=None) -> "DistCtx":
        """torchrun-style env (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
        import torch.distributed as dist
        world = int(os.environ.get("WORLD_SIZE", "1"))
        rank = int(os.environ.get("RANK", "0"))
        local = int(os.environ.get("LOCAL_RANK", "0"))
        if device is None:
            device = f"cuda:{local}" if torch.cuda.is_available() else "cpu"
        dev = torch.device(device)
        if dev.type == "cuda":
            torch.cuda.set_device(dev)
        if world > 1 and not dist.is_initialized():
            if dev.type == "cuda":
                dist.init_process_group("nccl", device_id=dev)
            else:
                dist.init_process_group("gloo")
        return cls(rank, world, dev)

    def my_candidates(self, n: int) -> List[int]:
        return [i for i in range(n) if i % self.world == self.rank]

    def per_rank(self, n: int) -> int:
        return (n + self.world - 1) // self.world

    def gather_records(self, local: Sequence[Tuple[int, int, int, float]], n: int):
        """-> all (cand_id, seed, label, score) records, identical and cand_id-sorted on every rank."""
        return gather_records(list(local), self.per_rank(n), self.rank, self.world, self.device)

    def gather_latents(self, local: Sequence[Tuple[int, torch.Tensor]], n: int, shape, dtype):
        """local: [(cand_id, tensor[shape])] -> list of n tensors (on self.device), every rank."""
        import torch.distributed as dist
        per = self.per_rank(n)
        slab = torch.zeros((per,) + tuple(shape), dtype=dtype, device=self.device)
        ids = torch.full((per,), -1, dtype=torch.int64, device=self.device)
        for j, (cid, t) in enumerate(local):
            slab[j] = t.to(self.device, dtype).reshape(shape)
            ids[j] = cid
        if self.world > 1:
            all_slab = torch.empty((self.world * per,) + tuple(shape), dtype=dtype, device=self.device)
            all_ids = torch.empty(self.world * per, dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(all_slab, slab)
            dist.all_gather_into_tensor(all_ids, ids)
        else:
            all_slab, all_ids = slab, ids
        out = [None] * n
        for j, cid in enumerate(all_ids.tolist()):
            if cid >= 0:
                out[cid] = all_slab[j]
        assert all(o is not None for o in out), "a candidate's latents were not contributed"
        return out

    def broadcast_object(self, obj: Any, src: int = 0) -> Any:
        if self.world == 1:
            return obj
        import torch.distributed as dist
        box = [obj if self.rank == src else None]
        dist.broadcast_object_list(box, src=src, device=self.device if self.device.type == "cuda" else None)
        return box[0]

    def gather_objects(self, local: Any) -> List[Any]:
        """every rank's (small, picklable) object, in rank order, on every rank"""
        if self.world == 1:
            return [local]
        import torch.distributed as dist
        out = [None] * self.world
        dist.all_gather_object(out, local)
        return out

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
