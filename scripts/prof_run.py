#!/usr/bin/env python
"""Small fixed-size runs of one workload / engine for ncu captures (one launch of the kernel of interest, no timing).

    python scripts/prof_run.py c5 [n_books] [deep|paged]        # launch 1 = pre-load (skip it with ncu -s 1), launch 2 = the timed phase
    python scripts/prof_run.py c3|c4|market [n_envs] [n_steps] [dense|paged]
    python scripts/prof_run.py c2 [n_books] [deep|paged]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bourse_b200 import abi, core, workloads  # noqa: E402


def main():
    wl = sys.argv[1]
    if wl == "c5":
        import torch
        n_books = int(sys.argv[2]) if len(sys.argv) > 2 else 16
        eng = sys.argv[3] if len(sys.argv) > 3 else "deep"
        # (a lighter variant for ncu captures with many books: PROF_REST=200000 PROF_STEPS=30 keeps the save / restore small)
        n_rest, n_steps, per_step = int(os.environ.get("PROF_REST", 1_000_000)), int(os.environ.get("PROF_STEPS", 100)), 10_000
        s = workloads.c5_stream(n_rest, n_steps, per_step, seed=100)
        dev = torch.device("cuda", 0)
        d1 = torch.from_numpy(s[:n_rest].view(np.uint8)).to(dev).repeat(n_books)
        d2 = torch.from_numpy(s[n_rest:].view(np.uint8)).to(dev).repeat(n_books)
        o1 = torch.arange(0, n_books + 1, dtype=torch.int64, device=dev) * n_rest
        o2 = torch.arange(0, n_books + 1, dtype=torch.int64, device=dev) * (n_steps * per_step)
        kw = dict(price_window=(7936, 12160), deep_chunks=98304) if eng == "deep" else dict(pages_smem=192, pages_total=192)
        env = core.BatchedEnv(n_books, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L2, max_orders=n_rest + n_steps * per_step * 8 // 10 + 1024,
                              max_trades=n_steps * per_step + 1024, max_steps=n_steps, max_queue=32, **kw)
        torch.cuda.synchronize()
        env.replay_device(d1.data_ptr(), o1.data_ptr())
        env.replay_device(d2.data_ptr(), o2.data_ptr())
        env.synchronize()
        print(env.stats())
    elif wl == "c2":
        import torch
        n_books = int(sys.argv[2]) if len(sys.argv) > 2 else 1
        eng = sys.argv[3] if len(sys.argv) > 3 else "deep"
        n_ev = 1_000_000
        s = workloads.replay_stream(n_ev, 0, tick_size=1, trading_windows=False)
        dev = torch.device("cuda", 0)
        d = torch.from_numpy(s.view(np.uint8)).to(dev).repeat(n_books)
        off = torch.arange(0, n_books + 1, dtype=torch.int64, device=dev) * n_ev
        kw = dict(price_window=(896, 1152), deep_chunks=65536) if eng == "deep" else dict(pages_smem=64, pages_total=64)
        env = core.BatchedEnv(n_books, 0, 0, 1, 100_000, obs_words=abi.OBS_L2, max_orders=1 << 20, max_trades=1 << 20,
                              max_steps=n_ev // 64 + 8, max_queue=32, **kw)
        env.replay_device(d.data_ptr(), off.data_ptr())
        env.synchronize()
        print(env.stats())
    else:
        n_envs = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
        n_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 60
        eng = sys.argv[4] if len(sys.argv) > 4 else "dense"
        if wl == "market":
            groups, assets = workloads.market_example_groups()
            kw = dict(price_window=(20, 180), live_cap=128) if eng == "dense" else {}
            env = core.BatchedEnv(n_envs, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L2, max_orders=8192, max_trades=8192, max_steps=n_steps,
                                  max_queue=80, assets=2, **kw)
            env.set_agents(groups, assets=assets)
        else:
            groups = workloads.c3_groups() if wl == "c3" else workloads.c4_groups()
            kw = {}
            if eng == "dense":
                kw = dict(price_window=(20, 180), live_cap=128) if wl == "c3" else {}
            env = core.BatchedEnv(n_envs, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L1 if wl == "c3" else abi.OBS_L2, max_orders=8192,
                                  max_trades=8192, max_steps=n_steps, max_queue=128 if wl == "c3" else 256, **kw)
            env.set_agents(groups)
        env.run_agents(n_steps, 101)
        print(env.stats(), np.unique(env.env_errors()))


if __name__ == "__main__":
    main()
