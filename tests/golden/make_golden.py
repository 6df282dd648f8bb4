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
4. market_example.npz — the multi-asset example population (crates/step_sim/examples/multi_asset/main.rs) plus a
   Momentum / Noise twin mix, market-keyed runs of the oracle's MarketSim: level-2 history and trade count per (market, asset).
5. vector_env.npz — a fixed block of action rows (orders, cancels, modifies, no-ops, off-tick prices) through one oracle
   StepEnv per env: the ids and level-2 record after every step (the device-resident loop must reproduce them).
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


def market_populations():
    ex_g, ex_a = workloads.market_example_groups()
    mixed_g = [orc.random_group(40, (40, 60), (10, 20), 2, 0.7), orc.random_group(30, (40, 60), (10, 20), 2, 0.7),
               orc.random_group(25, (30, 70), (5, 30), 2, 0.5), orc.momentum_group(100, 12, 1, 0.1, 10, 1.0, 5.0, 0.5, 1.0, 0.0, 1.0),
               orc.noise_group(200, 9, 1, 0.3, 0.1, 0.2, 7, 0.5, 0.8)]
    return {"example": (ex_g, ex_a, 2), "mixed": (mixed_g, [0, 1, 2, 1, 0], 3)}


def make_market():
    out = {}
    for name, (groups, assets, n_assets) in market_populations().items():
        n_markets, n_steps = 3, 40
        hist, ntr = [], []
        for m in range(n_markets):
            env = orc.MarketEnv(0, 0, [1] * n_assets, 1_000_000)
            env.set_groups(groups, assets)
            env.run_agents(n_steps, 101, market_id=m, keyed=True)
            hist.append(np.stack([env.history(a) for a in range(n_assets)]))
            ntr.append([len(env.get_trades(a)) for a in range(n_assets)])
        out[f"{name}/hist"], out[f"{name}/n_trades"] = np.stack(hist), np.array(ntr)
        print("market", name, out[f"{name}/hist"].shape, ntr)
    np.savez_compressed(os.path.join(HERE, "market_example.npz"), **out)


def vector_env_actions(n_envs=6, rows=5, n_steps=12, tick=2, seed=17):
    """Deterministic action block shared by the generator and the tests: [steps, envs, rows] packed rows."""
    from bourse_b200 import abi, gym
    rng = np.random.default_rng(seed)
    shape = (n_steps, n_envs, rows)
    u = rng.random(shape)
    op = np.where(u < 0.6, abi.OP_NEW, np.where(u < 0.75, abi.OP_CANCEL, np.where(u < 0.9, abi.OP_MODIFY, abi.OP_NOOP))).astype(np.uint32)
    op[0] = np.where(op[0] == abi.OP_NOOP, abi.OP_NOOP, abi.OP_NEW)          # nothing to cancel yet
    op[:, :, 0] = abi.OP_NEW
    issued = np.maximum(1, np.arange(n_steps))[:, None, None]                   # >= 1 order per earlier step
    price = tick * rng.integers(45, 56, shape) + (rng.random(shape) < 0.04)     # a few prices off the tick grid
    return gym.pack_actions(op, bid=rng.random(shape) < 0.5, vol=rng.integers(1, 40, shape), trader=rng.integers(0, 99, shape),
                            price=price, order_id=(rng.random(shape) * issued).astype(np.uint64), market=rng.random(shape) < 0.06,
                            has_price=rng.random(shape) < 0.6, has_vol=rng.random(shape) < 0.6), tick


def make_vector_env():
    from bourse_b200 import abi
    acts, tick = vector_env_actions()
    n_steps, n_envs, rows = acts.shape
    envs = [orc.StepEnv(9 + e, 0, tick, 1000) for e in range(n_envs)]
    ids = np.full((n_steps, n_envs, rows), abi.NO_ID, np.uint64)
    obs = np.zeros((n_steps, n_envs, 45), np.uint32)
    bad = np.zeros(n_envs, bool)
    for s in range(n_steps):
        for e, env in enumerate(envs):
            for r in range(rows):
                a = acts[s, e, r]
                op, f = int(a["op_flags"]) & 0xFF, int(a["op_flags"])
                if op == abi.OP_NEW:
                    try:
                        ids[s, e, r] = env.place_order(bool(f & abi.F_BID), int(a["vol"]), int(a["trader"]),
                                                       None if f & abi.F_MARKET else int(a["price"]))
                    except ValueError:
                        bad[e] = True
                elif op == abi.OP_CANCEL:
                    env.cancel_order(int(a["order_id"]))
                elif op == abi.OP_MODIFY:
                    env.modify_order(int(a["order_id"]), int(a["price"]) if f & abi.F_HAS_PRICE else None,
                                     int(a["vol"]) if f & abi.F_HAS_VOL else None)
            env.step()
            obs[s, e] = env.level_2_data_array()
    np.savez_compressed(os.path.join(HERE, "vector_env.npz"), actions_sha=np.frombuffer(hashlib.sha256(acts.tobytes()).digest(), dtype=np.uint8),
                        ids=ids, obs=obs, price_error=bad, n_trades=np.array([len(e.get_trades()) for e in envs]))
    print("vector_env", ids.shape, obs.shape, "price errors in envs", np.flatnonzero(bad), "trades", [len(e.get_trades()) for e in envs])


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, fn in (("c1", make_c1), ("replay", make_replay), ("agents", make_agents), ("market", make_market), ("vector_env", make_vector_env)):
        if not only or name in only:
            fn()
