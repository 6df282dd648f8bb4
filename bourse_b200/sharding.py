"""Multi-GPU plumbing: envs are independent, so a run shards them over ranks with no per-step collective.

One process per GPU (torch.distributed: NCCL on the GPU box, gloo in CPU tests).  Each rank owns the
contiguous block of global env ids returned by `shard_range`; the agent RNG is keyed by GLOBAL env id
(`env_id_base` of the handle), so results do not depend on the number of ranks.  The only collective is
the end-of-run all-gather of a small statistics vector (`gather_stats`).
"""
from __future__ import annotations

import typing

STAT_KEYS = ("instructions", "orders_created", "trades", "traded_volume", "env_steps", "transitions", "error_envs")


def shard_range(n_envs_total: int, world_size: int, rank: int, multiple: int = 1) -> typing.Tuple[int, int]:
    """(first global env id, number of envs) of `rank`: contiguous blocks, remainder spread over the first ranks.

    `multiple`: the blocks are made of whole groups of that many consecutive envs — the books of one multi-asset market
    (`bb_config.assets`) must stay on one GPU, and `env_id_base` must be a multiple of `assets` for the market-keyed RNG."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    if multiple < 1 or n_envs_total % multiple:
        raise ValueError("n_envs_total must be a multiple of `multiple`")
    q, r = divmod(n_envs_total // multiple, world_size)
    count = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return base * multiple, count * multiple


def gather_stats(stats: dict, elapsed_ms: float, l1_checksum: int, device=None) -> dict:
    """All-gather every rank's statistics; returns the whole-job aggregate (identical on every rank).

    Work counters are summed, the elapsed time is the MAX over ranks (the job finishes with its slowest
    shard), checksums are returned per rank in rank order.
    """
    import torch
    import torch.distributed as dist

    vec = torch.tensor([float(stats[k]) for k in STAT_KEYS] + [float(elapsed_ms), float(l1_checksum >> 32),
                                                               float(l1_checksum & 0xFFFFFFFF)],
                       dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        out = [torch.empty_like(vec) for _ in range(dist.get_world_size())]
        dist.all_gather(out, vec)
        allv = torch.stack(out).cpu()
    else:
        allv = vec.cpu()[None]
    agg = {k: int(allv[:, i].sum().item()) for i, k in enumerate(STAT_KEYS)}
    n = len(STAT_KEYS)
    agg["elapsed_ms_max"] = float(allv[:, n].max().item())
    agg["elapsed_ms_per_rank"] = [float(x) for x in allv[:, n]]
    agg["l1_checksums"] = [(int(hi) << 32) | int(lo) for hi, lo in zip(allv[:, n + 1], allv[:, n + 2])]
    agg["world_size"] = int(allv.shape[0])
    return agg
