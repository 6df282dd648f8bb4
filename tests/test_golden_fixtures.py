"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py).

CPU part: the oracle and our Python agent mirror reproduce the fixtures (the C1 fixture was produced by
the REFERENCE's own Python runner/agents, imported unmodified, on top of the oracle core).
GPU part: the CUDA path reproduces the same fixtures bit for bit."""
import hashlib
import os

import numpy as np
import pytest

from bourse_b200 import abi, workloads

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REPLAYS = {"strict_t1": dict(seed=0, tick_size=1, time_mode="strict"), "strict_t2": dict(seed=1, tick_size=2, time_mode="strict"),
           "flat_t1": dict(seed=2, tick_size=1, time_mode="flat", min_vol=0), "jitter_t1": dict(seed=3, tick_size=1, time_mode="jitter")}


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def t_arr(trades):
    return np.array(trades, dtype=np.uint64).reshape(-1, 6)


def o_arr(orders):
    return np.array([[int(x) for x in o] for o in orders], dtype=np.uint64).reshape(-1, 9)


def run_c1(env_cls, tick):
    from bourse_b200.step_sim.agents import RandomAgent

    agents = [RandomAgent(i, 0.5, (10, 100), (20, 50), tick) for i in range(100)]
    env = env_cls(101, 0, tick, 100_000)
    rng = np.random.default_rng(101)
    for _ in range(200):
        for a in agents:
            a.update(rng, env)
        env.step()
    return env


def check_c1(env, tick):
    g = load("c1_random_trades")
    data = env.get_market_data()
    assert len(data) == 45
    for k, v in data.items():
        assert np.array_equal(v, g[f"t{tick}/{k}"]), k
    assert np.array_equal(t_arr(env.get_trades()), g[f"t{tick}/trades"])
    assert np.array_equal(o_arr(env.get_orders()), g[f"t{tick}/orders"])


def replay_stream(name):
    g = load("replay_" + name)
    s = workloads.replay_stream(int(g["gen"][0]), **REPLAYS[name])
    assert np.array_equal(np.frombuffer(hashlib.sha256(s.tobytes()).digest(), dtype=np.uint8), g["stream_sha"]), \
        "stream generator changed: regenerate the fixtures"
    return g, s


# ------------------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("tick", [2, 1])
def test_c1_python_mirror_matches_reference_python_layer(oracle, tick):
    check_c1(run_c1(oracle.StepEnv, tick), tick)


@pytest.mark.parametrize("name", sorted(REPLAYS))
def test_oracle_reproduces_replay_fixture(oracle, name):
    g, s = replay_stream(name)
    ob = oracle.OrderBook(0, REPLAYS[name]["tick_size"])
    obs = ob.replay(s, obs_cap=len(s))
    assert np.array_equal(t_arr(ob.get_trades()), g["trades"]) and np.array_equal(o_arr(ob.get_orders()), g["orders"])
    assert np.array_equal(obs, g["obs"]) and list(g["l1"]) == ob._l1()


def test_oracle_reproduces_agent_fixtures(oracle):
    for name, groups in (("c3", workloads.c3_groups()), ("c4", workloads.c4_groups())):
        g = load("agents_" + name)
        for e in range(g["hist"].shape[0]):
            env = oracle.StepEnvNumpy(0, 0, 1, 1_000_000)
            env.set_groups(groups)
            env.run_agents(g["hist"].shape[1], 101, env_id=e, keyed=True)
            assert np.array_equal(env._history(), g["hist"][e]) and len(env.get_trades()) == g["n_trades"][e]


def _golden_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name", ["example", "mixed"])
def test_oracle_reproduces_market_fixture(oracle, name):
    g = load("market_example")
    groups, assets, n_assets = _golden_module().market_populations()[name]
    hist = g[f"{name}/hist"]
    for m in range(hist.shape[0]):
        env = oracle.MarketEnv(0, 0, [1] * n_assets, 1_000_000)
        env.set_groups(groups, assets)
        env.run_agents(hist.shape[2], 101, market_id=m, keyed=True)
        for a in range(n_assets):
            assert np.array_equal(env.history(a), hist[m, a]) and len(env.get_trades(a)) == g[f"{name}/n_trades"][m, a]


def vector_env_fixture():
    g = load("vector_env")
    acts, tick = _golden_module().vector_env_actions()
    assert np.array_equal(np.frombuffer(hashlib.sha256(acts.tobytes()).digest(), dtype=np.uint8), g["actions_sha"]), \
        "action generator changed: regenerate the fixtures"
    return g, acts, tick


def test_vector_env_fixture_is_consistent():
    g, acts, tick = vector_env_fixture()
    n_steps, n_envs, rows = acts.shape
    assert g["ids"].shape == (n_steps, n_envs, rows) and g["obs"].shape == (n_steps, n_envs, 45)
    new = (acts["op_flags"] & 0xFF) == abi.OP_NEW
    assert (g["ids"][~new] == abi.NO_ID).all()                      # only NEW rows create ids
    off_tick = new & ((acts["op_flags"] & abi.F_MARKET) == 0) & (acts["price"] % tick != 0)
    assert (g["ids"][off_tick] == abi.NO_ID).all() and np.array_equal(off_tick.any(axis=(0, 2)), g["price_error"])
    ok = new & ~off_tick
    for e in range(n_envs):                                         # ids are dense per env, in row order
        assert np.array_equal(g["ids"][:, e][ok[:, e]], np.arange(ok[:, e].sum(), dtype=np.uint64))


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", [("example", dict(price_window=(0, 256), live_cap=128)), ("example", dict()), ("mixed", dict())],
                         ids=["example_dense", "example_fast", "mixed_fast"])
def test_cuda_reproduces_market_fixture(name, kw):
    from bourse_b200 import core
    g = load("market_example")
    groups, assets, n_assets = _golden_module().market_populations()[name]
    hist = g[f"{name}/hist"]
    n_markets, n_steps = hist.shape[0], hist.shape[2]
    env = core.BatchedEnv(n_markets * n_assets, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L2, max_orders=8192, max_trades=8192,
                          max_steps=n_steps, max_queue=128, assets=n_assets, **kw)
    env.set_agents(groups, assets=assets)
    env.run_agents(n_steps, 101)
    assert not env.env_errors().any()
    assert np.array_equal(env.history_all(n_steps).reshape(hist.shape), hist)
    assert [[env.n_trades(m * n_assets + a) for a in range(n_assets)] for m in range(n_markets)] == g[f"{name}/n_trades"].tolist()


@pytest.mark.gpu
def test_cuda_reproduces_vector_env_fixture():
    from bourse_b200 import gym
    g, acts, tick = vector_env_fixture()
    n_steps, n_envs, rows = acts.shape
    v = gym.VectorEnv(n_envs, rows, 9, 0, tick, 1000, max_orders=256, max_trades=512, max_steps=32)
    v.reset()
    for s in range(n_steps):
        obs, ids = v.step(acts[s])
        assert np.array_equal(ids.numpy(), g["ids"][s]) and np.array_equal(obs.numpy(), g["obs"][s]), s
    assert np.array_equal((v.env.env_errors() & 0x200) != 0, g["price_error"])
    assert [v.env.n_trades(e) for e in range(n_envs)] == list(g["n_trades"])
    v.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tick", [2, 1])
def test_c1_cuda_matches_fixture(core, tick):
    import bourse_b200
    from bourse_b200.step_sim.agents import RandomAgent

    env = run_c1(core.StepEnv, tick)
    check_c1(env, tick)
    # and through the runner entry point the example uses (examples/random_trades.py:7-20)
    agents = [RandomAgent(i, 0.5, (10, 100), (20, 50), tick) for i in range(100)]
    data = bourse_b200.step_sim.run(core.StepEnv(101, 0, tick, 100_000), agents, 200, 101)
    g = load("c1_random_trades")
    assert all(np.array_equal(v, g[f"t{tick}/{k}"]) for k, v in data.items())


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(REPLAYS))
def test_cuda_reproduces_replay_fixture(core, name):
    g, s = replay_stream(name)
    ob = core.OrderBook(0, REPLAYS[name]["tick_size"], max_orders=len(s), max_trades=4 * len(s), max_steps=len(s) // 64 + 8)
    obs = ob.replay(s)
    assert np.array_equal(t_arr(ob.get_trades()), g["trades"]) and np.array_equal(o_arr(ob.get_orders()), g["orders"])
    assert np.array_equal(obs, g["obs"]) and list(g["l1"]) == ob._l1()


@pytest.mark.gpu
def test_cuda_reproduces_agent_fixtures(core):
    g = load("agents_c3")
    n_envs, n_steps = g["hist"].shape[:2]
    env = core.BatchedEnv(n_envs, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L2, max_orders=4096, max_trades=8192,
                          max_steps=n_steps, max_queue=128)
    env.set_agents(workloads.c3_groups())
    env.run_agents(n_steps, 101)
    assert np.array_equal(env.history_all(n_steps), g["hist"])
    assert [env.n_trades(e) for e in range(n_envs)] == list(g["n_trades"])
