"""The reference's order table, trade log and per-step records grow without bound (orderbook.rs:113-115, data.rs:9-57); the
single-env drop-in classes follow through bb_reserve.  Started with deliberately tiny capacities, they must reproduce the
oracle over runs that outgrow every table several times."""
import numpy as np
import pytest

from bourse_b200 import abi, workloads

pytestmark = pytest.mark.gpu


def test_order_book_grows_like_a_vec(core, oracle):
    rng = np.random.default_rng(4)
    g = core.OrderBook(0, 1, max_orders=128, max_trades=64, max_steps=8)
    o = oracle.OrderBook(0, 1)
    for i in range(3000):
        for ob in (g, o):
            ob.set_time(i + 1)
        bid, vol, price = bool(rng.random() < 0.5), int(rng.integers(1, 30)), int(rng.integers(95, 106))
        u = rng.random()
        if u < 0.75 or i < 10:
            assert g.place_order(bid, vol, 7, price=price) == o.place_order(bid, vol, 7, price=price)
        elif u < 0.9:
            k = int(rng.integers(0, len(o.get_orders())))
            g.cancel_order(k); o.cancel_order(k)
        else:
            k = int(rng.integers(0, len(o.get_orders())))
            g.modify_order(k, new_price=price, new_vol=vol); o.modify_order(k, new_price=price, new_vol=vol)
    assert g._env.max_orders > 2048 and g._env.max_trades > 64      # grew several times
    assert g.get_orders() == o.get_orders() and g.get_trades() == o.get_trades()
    assert len(o.get_trades()) > 500 and g.bid_ask() == o.bid_ask()


def test_order_book_replay_reserves_ahead(core, oracle):
    n = 20000
    s = workloads.replay_stream(n, 3, tick_size=1)
    g = core.OrderBook(0, 1, max_orders=64, max_trades=64, max_steps=4)
    o = oracle.OrderBook(0, 1)
    obs_cpu = o.replay(s, obs_cap=n)
    half = n // 2
    obs_gpu = np.concatenate([g.replay(s[:half]), g.replay(s[half:])])
    assert np.array_equal(obs_gpu, obs_cpu)
    assert g.get_orders() == o.get_orders() and g.get_trades() == o.get_trades()


def test_step_env_grows_like_a_vec(core, oracle):
    from bourse_b200.step_sim.agents import RandomAgent
    g = core.StepEnv(11, 0, 2, 1000, max_orders=256, max_trades=128, max_steps=16, max_queue=64)
    o = oracle.StepEnv(11, 0, 2, 1000)
    agents_g = [RandomAgent(i, 0.5, (10, 100), (20, 50), 2) for i in range(30)]
    agents_o = [RandomAgent(i, 0.5, (10, 100), (20, 50), 2) for i in range(30)]
    rg, ro = np.random.default_rng(5), np.random.default_rng(5)
    for _ in range(400):
        for a in agents_g:
            a.update(rg, g)
        for a in agents_o:
            a.update(ro, o)
        g.step(); o.step()
    assert g._env.max_steps >= 400 and g._env.max_orders > 256
    dg, do = g.get_market_data(), o.get_market_data()
    assert all(np.array_equal(dg[k], do[k]) for k in do) and len(do["bid_price"]) == 400
    assert g.get_orders() == o.get_orders() and g.get_trades() == o.get_trades()


def test_reserve_keeps_contents_of_every_env(core, oracle):
    groups = workloads.c3_groups()
    e = core.BatchedEnv(6, 0, 0, 1, 1_000_000, obs_words=abi.OBS_L2, max_orders=2048, max_trades=4096, max_steps=24, max_queue=128)
    e.set_agents(groups)
    e.run_agents(24, 9)
    e.reserve(max_orders=8192, max_trades=16384, max_steps=64)
    e.run_agents(40, 9)
    assert not e.env_errors().any()
    for env in range(6):
        ce = oracle.StepEnvNumpy(0, 0, 1, 1_000_000)
        ce.set_groups(groups)
        ce.run_agents(64, 9, env_id=env, keyed=True)
        assert np.array_equal(e.history(env), ce._history()) and e.get_trades(env) == ce.get_trades() and e.get_orders(env) == ce.get_orders()
