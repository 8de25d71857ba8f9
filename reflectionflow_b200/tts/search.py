"""Selection, bookkeeping and sharding logic of the outer loop — pure functions so that they can be
pinned by table tests (reference: tts/tts_reflectionflow.py:144-182,359-448) — plus the one
collective the B200 design adds: an all-gather of fixed-size score records per round
(SURVEY.md §8e)."""
from __future__ import annotations

import struct
from typing import Dict, List, Sequence, Tuple

import torch

MAX_SEED = 2 ** 31 - 1  # np.iinfo(np.int32).max (tts_reflectionflow.py:44)


# ------------------------------------------------------------------ sharding
def shard_candidates(num_candidates: int, rank: int, world: int) -> List[int]:
    """candidate i of a round -> rank i mod world (every rank derives the same table)."""
    return [i for i in range(num_candidates) if i % world == rank]


# ------------------------------------------------------------------ stub verifier
def stub_verifier_score(latents: torch.Tensor) -> torch.Tensor:
    """Deterministic scalar 'verifier' for offline runs: a fixed linear functional of the final
    latent evaluated in fp64 on the device (ties are measure-zero).  Returns [B] fp64."""
    x = latents.to(torch.float64).reshape(latents.shape[0], -1)
    n = x.shape[1]
    w = torch.cos(torch.arange(n, device=x.device, dtype=torch.float64) * 0.6180339887498949)
    return (x * w).sum(dim=1) / n


# ------------------------------------------------------------------ score exchange
RECORD_BYTES = 32  # cand_id:int32, label:int32, seed:int64, score:float64, pad


def pack_record(cand_id: int, seed: int, label: int, score: float) -> bytes:
    return struct.pack("<iiqd8x", int(cand_id), int(label), int(seed), float(score))


def unpack_record(b: bytes) -> Tuple[int, int, int, float]:
    cand_id, label, seed, score = struct.unpack("<iiqd8x", b)
    return cand_id, seed, label, score


def gather_records(local: Sequence[Tuple[int, int, int, float]], per_rank: int, rank: int,
                   world: int, device) -> List[Tuple[int, int, int, float]]:
    """One all-gather of `per_rank` fixed-size records from every rank (NCCL on GPUs, gloo on
    CPU).  Slots a rank does not fill carry cand_id = -1 and are dropped.  Every rank returns the
    same list, sorted by cand_id."""
    import torch.distributed as dist
    buf = bytearray()
    for i in range(per_rank):
        buf += pack_record(*local[i]) if i < len(local) else pack_record(-1, 0, 0, 0.0)
    t = torch.frombuffer(bytes(buf), dtype=torch.uint8).clone().to(device)
    if world > 1:
        out = torch.empty(world * t.numel(), dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(out, t)
    else:
        out = t
    raw = out.cpu().numpy().tobytes()
    recs = [unpack_record(raw[o:o + RECORD_BYTES]) for o in range(0, len(raw), RECORD_BYTES)]
    return sorted([r for r in recs if r[0] >= 0], key=lambda r: r[0])


def gather_scores(score: torch.Tensor, rank: int, world: int, device):
    """bench helper: one record per rank."""
    return gather_records([(rank, 0, 1, float(score.reshape(-1)[0].item()))], 1, rank, world, device)
