"""CPU suite: the N>1 path with two gloo ranks.  The CUDA compute is replaced by the oracle (keyed
Philox contract) so the host-side sharding, global-env-id keying and the statistics all-gather are
exercised end to end without a GPU."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from bourse_b200.sharding import shard_range

    for total, world in [(4096, 1), (4096, 8), (65536, 8), (10, 4), (3, 8), (0, 2)]:
        seen = []
        for r in range(world):
            base, cnt = shard_range(total, world, r)
            seen += list(range(base, base + cnt))
        assert seen == list(range(total))
        sizes = [shard_range(total, world, r)[1] for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, os.environ["BB_ROOT"])
    from bourse_b200 import workloads
    from bourse_b200.sharding import shard_range, gather_stats
    from oracle import oracle as orc

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_total, n_steps, seed = 12, 20, 101
    assets = int(os.environ.get("BB_ASSETS", "1"))
    base, cnt = shard_range(n_total, world, rank, multiple=assets)
    instr = trades = 0
    fnv = 0xcbf29ce484222325
    l1_rows = []
    if assets == 1:
        groups = workloads.c3_groups()
        for e in range(base, base + cnt):
            env = orc.StepEnvNumpy(0, 0, 1, 1_000_000)
            env.set_groups(groups)
            env.run_agents(n_steps, seed, env_id=e, keyed=True)     # keyed by GLOBAL env id
            instr += env.n_instructions(); trades += len(env.get_trades())
            l1_rows.append(env.level_1_data())
    else:   # multi-asset markets: whole markets per rank, keyed by GLOBAL market id = global book id // assets
        groups, g_assets = workloads.market_example_groups()
        assert base % assets == 0 and cnt % assets == 0
        for m in range(base // assets, (base + cnt) // assets):
            env = orc.MarketEnv(0, 0, [1] * assets, 1_000_000)
            env.set_groups(groups, g_assets)
            env.run_agents(n_steps, seed, market_id=m, keyed=True)
            instr += env.n_instructions(); trades += sum(len(env.get_trades(a)) for a in range(assets))
            l1_rows += [env.history(a)[-1, :9] for a in range(assets)]
    for b in np.concatenate(l1_rows).astype(np.uint32).tobytes():
        fnv = ((fnv ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    stats = dict(instructions=instr, orders_created=0, trades=trades, traded_volume=0, env_steps=cnt * n_steps,
                 transitions=0, error_envs=0)
    agg = gather_stats(stats, elapsed_ms=10.0 * (rank + 1), l1_checksum=fnv)
    gathered = [None] * world
    dist.all_gather_object(gathered, np.concatenate(l1_rows).tolist())
    if rank == 0:
        agg["l1_all"] = sum(gathered, [])
        print("RESULT " + json.dumps(agg))
    dist.barrier()
    dist.destroy_process_group()
""")


def run_world(world_size, tmp_path, assets=1):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, BB_ROOT=ROOT, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1", BB_ASSETS=str(assets))
    port = 29500 + (os.getpid() % 2000) + world_size + 10 * assets
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world_size}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1]
    import json
    return json.loads(line[len("RESULT "):])


def test_two_ranks_equal_one_rank(tmp_path):
    one = run_world(1, tmp_path)
    two = run_world(2, tmp_path)
    # sharding does not change any result: same per-env market data, same aggregate counters
    assert two["l1_all"] == one["l1_all"]
    assert two["instructions"] == one["instructions"] and two["trades"] == one["trades"]
    assert two["env_steps"] == one["env_steps"] == 12 * 20
    assert two["world_size"] == 2 and len(two["l1_checksums"]) == 2
    assert two["elapsed_ms_max"] == 20.0 and two["elapsed_ms_per_rank"] == [10.0, 20.0]   # max over ranks


def test_two_ranks_equal_one_rank_markets(tmp_path):
    """Two-asset markets (6 markets = 12 books): ranks own whole markets, results keyed by global market id."""
    one = run_world(1, tmp_path, assets=2)
    two = run_world(2, tmp_path, assets=2)
    assert two["l1_all"] == one["l1_all"] and len(one["l1_all"]) == 12 * 9
    assert two["instructions"] == one["instructions"] > 0 and two["trades"] == one["trades"] > 0
    assert two["env_steps"] == one["env_steps"] == 12 * 20
