"""Multi-GPU plumbing: envs are independent, so a run shards them over ranks with no per-step collective.

One process per GPU.  Each rank owns the contiguous block of global env ids returned by `shard_range`; the agent RNG is
keyed by GLOBAL env id (`env_id_base` of the handle), so results do not depend on the number of ranks.  The only
collective is the end-of-run all-gather of every shard's statistics block: `gather_env_stats`, which goes through the C
ABI (`bb_comm_*` / `bb_gather_stats`: ncclAllGather inside the library, no torch).
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


def aggregate(rows: typing.Sequence[dict]) -> dict:
    """Whole-job aggregate of per-rank records {STAT_KEYS..., elapsed_ms, l1_checksum}: work counters are summed, the
    elapsed time is the MAX over ranks (the job finishes with its slowest shard), checksums stay per rank in rank order."""
    agg = {k: int(sum(int(r[k]) for r in rows)) for k in STAT_KEYS}
    agg["elapsed_ms_max"] = float(max(r["elapsed_ms"] for r in rows))
    agg["elapsed_ms_per_rank"] = [float(r["elapsed_ms"]) for r in rows]
    agg["l1_checksums"] = [int(r["l1_checksum"]) for r in rows]
    agg["world_size"] = len(rows)
    return agg


class Comm:
    """The job's communicator behind the C ABI (`bb_comm_*`, csrc/comm.cu): NCCL over NVLink / NVSwitch, loaded by the
    library itself — no torch.  One process per GPU: `Comm.from_env(device)` reads RANK / WORLD_SIZE as torchrun (or any
    launcher) sets them; rank 0 creates the NCCL unique id and hands it to the other ranks through a file under the
    system temp directory (single node).  world size 1 needs no NCCL and creates no communicator."""

    def __init__(self, handle, rank: int, world: int):
        self._h, self.rank, self.world = handle, rank, world

    @classmethod
    def from_env(cls, device: int, tag: str = "") -> "Comm":
        import ctypes as C
        import os
        import tempfile
        import time

        from . import abi

        rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        if world == 1:
            return cls(None, 0, 1)
        lib = abi.load()
        job = os.environ.get("TORCHELASTIC_RUN_ID", "") + "_" + os.environ.get("MASTER_PORT", "0") + "_" + str(os.getppid()) + tag
        path = os.path.join(tempfile.gettempdir(), f"bourse_b200_nccl_id_{job}")
        ident = (C.c_ubyte * 128)()
        if rank == 0:
            if lib.bb_comm_unique_id(ident) != abi.BB_OK:
                raise RuntimeError(lib.bb_comm_last_error().decode())
            with open(path + ".tmp", "wb") as f:
                f.write(bytes(ident))
            os.replace(path + ".tmp", path)
        else:
            t0 = time.time()
            while not os.path.exists(path):
                if time.time() - t0 > 120:
                    raise RuntimeError("timed out waiting for rank 0's NCCL id")
                time.sleep(0.01)
            with open(path, "rb") as f:
                ident = (C.c_ubyte * 128).from_buffer_copy(f.read(128))
        h = C.c_void_p()
        rc = lib.bb_comm_init_rank(ident, world, rank, device, C.byref(h))
        if rc != abi.BB_OK:
            raise RuntimeError(f"bb_comm_init_rank failed ({rc}): {lib.bb_comm_last_error().decode()}")
        comm = cls(h, rank, world)
        comm._path = path if rank == 0 else None
        return comm

    def close(self):
        if self._h:
            from . import abi
            abi.load().bb_comm_destroy(self._h)
            self._h = None
            if getattr(self, "_path", None):
                try:
                    import os
                    os.remove(self._path)
                except OSError:
                    pass


def gather_env_stats(env, elapsed_ms: float, comm: typing.Optional[Comm]) -> dict:
    """`bb_stats` of this rank's `BatchedEnv`, all-gathered over the job's communicator (bb_gather_stats: ncclAllGather inside
    the library) and aggregated — identical on every rank.  The run's ONLY collective."""
    import ctypes as C

    from . import abi

    if comm is None or comm.world == 1:
        st = env.stats()
        return aggregate([dict(st, elapsed_ms=elapsed_ms)])
    lib = abi.load()
    handles = (C.c_void_p * 1)(env._h)
    ms_in, ms_out = (C.c_double * 1)(elapsed_ms), (C.c_double * comm.world)()
    out = (abi.Stats * comm.world)()
    rc = lib.bb_gather_stats(comm._h, handles, 1, ms_in, out, ms_out)
    if rc != abi.BB_OK:
        raise RuntimeError(f"bb_gather_stats failed ({rc}): {lib.bb_comm_last_error().decode()}")
    rows = [dict({n: int(getattr(out[k], n)) for n, _ in abi.Stats._fields_}, elapsed_ms=float(ms_out[k])) for k in range(comm.world)]
    return aggregate(rows)


def gather_stats(stats: dict, elapsed_ms: float, l1_checksum: int, device=None) -> dict:
    """The same aggregate over `torch.distributed` — used by the CPU suite only (two gloo ranks with the oracle standing in
    for the GPU: tests/test_sharding_gloo.py); the product path is `gather_env_stats`."""
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
    n = len(STAT_KEYS)
    rows = [dict({k: int(allv[r, i].item()) for i, k in enumerate(STAT_KEYS)}, elapsed_ms=float(allv[r, n].item()),
                 l1_checksum=(int(allv[r, n + 1].item()) << 32) | int(allv[r, n + 2].item())) for r in range(allv.shape[0])]
    return aggregate(rows)
