"""Generates the committed golden fixtures.  Run in the BUILD container (needs /root/reference):

    python tests/golden/make_golden.py

1. c1_random_trades.npz — BASELINE config C1.  The reference's OWN Python layer
   (/root/reference/src/bourse: step_sim.run + agents.RandomAgent, imported unmodified) drives a
   `bourse.core` whose classes are the C++ oracle's (the Rust extension cannot be built here).  This pins
   (a) our Python mirror of the runner/agents against the reference's Python code and (b) the CUDA core
   against the oracle on the example the reference ships (examples/random_trades.py:4-20).
2. replay_*.npz — config C2 streams: oracle results (trade log, order table, emitted L2 records) for fixed
   generator seeds, so the GPU suite can check the CUDA path against committed vectors.
3. agents_c3.npz / agents_c4.npz — keyed (Philox) agent-driven runs of the oracle: level-2 history per env.
"""
import hashlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from bourse_b200 import workloads  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def trades_array(trades):
    return np.array(trades, dtype=np.uint64).reshape(-1, 6)


def orders_array(orders):
    return np.array([[int(x) for x in o] for o in orders], dtype=np.uint64).reshape(-1, 9)


def make_c1():
    ref_src = "/root/reference/src"
    if not os.path.isdir(ref_src):
        raise SystemExit("reference not mounted; fixtures can only be regenerated in the build container")
    core = types.ModuleType("bourse.core")
    core.StepEnv, core.StepEnvNumpy, core.OrderBook = orc.StepEnv, orc.StepEnvNumpy, orc.OrderBook
    import importlib.abc
    import importlib.machinery

    class CoreFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):  # serves `bourse.core` from the oracle
        def find_spec(self, name, path, target=None):
            return importlib.machinery.ModuleSpec(name, self) if name == "bourse.core" else None

        def create_module(self, spec):
            return core

        def exec_module(self, module):
            pass

    sys.meta_path.insert(0, CoreFinder())
    sys.path.insert(0, ref_src)
    import bourse  # the reference's pure-Python package, unmodified
    from bourse.step_sim.agents import RandomAgent

    out = {}
    for tick in (2, 1):  # the example file uses 2, BASELINE.json's config string says 1
        agents = [RandomAgent(i, 0.5, (10, 100), (20, 50), tick) for i in range(100)]
        env = bourse.core.StepEnv(101, 0, tick, 100_000)
        data = bourse.step_sim.run(env, agents, 200, 101, show_progress=False)
        for k, v in data.items():
            out[f"t{tick}/{k}"] = np.asarray(v, dtype=np.uint32)
        out[f"t{tick}/trades"] = trades_array(env.get_trades())
        out[f"t{tick}/orders"] = orders_array(env.get_orders())
    np.savez_compressed(os.path.join(HERE, "c1_random_trades.npz"), **out)
    print("c1:", {k: v.shape for k, v in list(out.items())[:3]}, "trades", out["t2/trades"].shape)


def make_replay():
    for name, kw in {"strict_t1": dict(seed=0, tick_size=1, time_mode="strict"),
                     "strict_t2": dict(seed=1, tick_size=2, time_mode="strict"),
                     "flat_t1": dict(seed=2, tick_size=1, time_mode="flat", min_vol=0),
                     "jitter_t1": dict(seed=3, tick_size=1, time_mode="jitter")}.items():
        n = 20000
        s = workloads.replay_stream(n, **kw)
        ob = orc.OrderBook(0, kw["tick_size"])
        obs = ob.replay(s, obs_cap=n)
        trades, orders = trades_array(ob.get_trades()), orders_array(ob.get_orders())
        np.savez_compressed(os.path.join(HERE, f"replay_{name}.npz"), stream_sha=np.frombuffer(
            hashlib.sha256(s.tobytes()).digest(), dtype=np.uint8), trades=trades, orders=orders, obs=obs,
            l1=np.array(ob._l1(), dtype=np.uint32), gen=np.array([n, kw["seed"], kw["tick_size"]]))
        print(name, "trades", trades.shape, "orders", orders.shape, "obs", obs.shape)


def make_agents():
    for name, groups, n_envs, n_steps in (("c3", workloads.c3_groups(), 6, 48), ("c4", workloads.c4_groups(), 4, 48)):
        hist, ntr = [], []
        for e in range(n_envs):
            env = orc.StepEnvNumpy(0, 0, 1, 1_000_000)
            env.set_groups(groups)
            env.run_agents(n_steps, 101, env_id=e, keyed=True)
            hist.append(env._history())
            ntr.append(len(env.get_trades()))
        np.savez_compressed(os.path.join(HERE, f"agents_{name}.npz"), hist=np.stack(hist), n_trades=np.array(ntr))
        print(name, np.stack(hist).shape, ntr)


if __name__ == "__main__":
    make_c1()
    make_replay()
    make_agents()
