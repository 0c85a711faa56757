"""Utterance sharding for multi-GPU generation: one process per GPU, each holding a full weight replica and a slice of
the request batch. Utterances are independent (no cross-utterance state), so there is NO collective on the hot path;
torch.distributed (NCCL on GPUs, gloo in the CPU tests) only gathers per-rank counts and timings at the end."""
from __future__ import annotations

from typing import Iterable, List, Sequence


def estimate_frames(n_words: int) -> int:
    """Upper bound on generated frames of a sentence: the reference's cap int((words + 2) * 12.5) (src/pocket_tts.cpp:429-430)."""
    return int((n_words + 2.0) * 12.5)


def shard_utterances(costs: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-processing-time greedy assignment of utterances (by estimated frames) to ranks.
    Deterministic: ties broken by utterance index, then rank index. Returns per-rank lists of utterance indices."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(costs)), key=lambda i: (-int(costs[i]), i))
    loads = [0] * world_size
    counts = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], counts[k], k))
        out[r].append(i)
        loads[r] += int(costs[i]); counts[r] += 1
    for lst in out:
        lst.sort()
    return out


def imbalance(costs: Sequence[int], shards: Iterable[Sequence[int]]) -> float:
    loads = [sum(int(costs[i]) for i in s) for s in shards]
    mean = sum(loads) / max(len(loads), 1)
    return (max(loads) / mean - 1.0) if mean > 0 else 0.0


def gather_stats(values: Sequence[float], device=None):
    """All-gather a small vector of per-rank numbers (frames, utterances, elapsed ms, ...). Returns a [world, len] tensor.
    Works without an initialised process group (single process)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t[None].cpu()
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out).cpu()
