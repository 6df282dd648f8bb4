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


def test_step_queue_grows_like_a_vec(core, oracle):
    """The reference's per-step transaction queue is an unbounded Vec (env.rs:93-96, 121).  A StepEnv created with room for 32
    transactions per step takes steps of up to 700 (bb_reserve_queue ahead of the step) and shuffles them exactly like the oracle;
    StepEnvNumpy's array submissions follow the same route."""
    from bourse_b200.step_sim.agents import RandomAgent
    g = core.StepEnv(23, 0, 1, 100_000, max_queue=32)
    o = oracle.StepEnv(23, 0, 1, 100_000)
    rg, ro = np.random.default_rng(8), np.random.default_rng(8)
    for n_agents in (20, 90, 700, 40, 300):
        ag = [RandomAgent(i, 0.9, (10, 100), (40, 70), 1) for i in range(n_agents)]
        ao = [RandomAgent(i, 0.9, (10, 100), (40, 70), 1) for i in range(n_agents)]
        for _ in range(3):
            for a in ag:
                a.update(rg, g)
            for a in ao:
                a.update(ro, o)
            g.step(); o.step()
    assert g._env.max_queue >= 600
    dg, do = g.get_market_data(), o.get_market_data()
    assert all(np.array_equal(dg[k], do[k]) for k in do) and len(do["bid_price"]) == 15
    assert g.get_orders() == o.get_orders() and g.get_trades() == o.get_trades() and len(o.get_trades()) > 300

    gn, on = core.StepEnvNumpy(5, 0, 1, 100_000, max_queue=16), oracle.StepEnvNumpy(5, 0, 1, 100_000)
    rng = np.random.default_rng(2)
    for n in (10, 500, 64):
        orders = ((rng.random(n) < 0.5), rng.integers(1, 40, n).astype(np.uint32), np.arange(n, dtype=np.uint32),
                  rng.integers(90, 111, n).astype(np.uint32))
        assert np.array_equal(gn.submit_limit_orders(orders), on.submit_limit_orders(orders))
        gn.step(); on.step()
    assert gn.get_orders() == on.get_orders() and gn.get_trades() == on.get_trades()


def test_an_overlong_step_queue_is_refused_before_anything_is_applied(core, oracle):
    """Batched handles are sized by the caller: bb_step answers BB_ECAP (MemoryError) while an env's queue exceeds max_queue, the
    queue and the books stay as they were, and after bb_reserve_queue the same step runs and equals the oracle."""
    e = core.BatchedEnv(3, 7, 0, 1, 1000, obs_words=abi.OBS_L2, max_queue=8)
    n = 20
    rng = np.random.default_rng(1)
    side, vol, price = (rng.random(n) < 0.5), rng.integers(1, 30, n).astype(np.uint32), rng.integers(95, 106, n).astype(np.uint32)
    e.submit(np.full(n, abi.ACT_NEW, np.uint32), side, vol, np.arange(n, dtype=np.uint32), price, env=np.full(n, 1, np.uint32))
    with pytest.raises(MemoryError, match="max_queue"):
        e.step(1)
    assert not e.env_errors().any() and e.n_trades(1) == 0 and e.time(1) == 0
    with pytest.raises(MemoryError):
        e.reserve_queue(60000)          # (does not fit next to the book image with four books per CTA ... nor with one)
    e.reserve_queue(64)
    e.step(1)
    o = oracle.StepEnvNumpy(7 + 1, 0, 1, 1000)   # env 1 of a handle seeded 7 draws from seed 7 + env id
    o.submit_limit_orders((side, vol, np.arange(n, dtype=np.uint32), price))
    o.step()
    assert not e.env_errors().any()
    assert e.get_orders(1) == o.get_orders() and e.get_trades(1) == o.get_trades() and np.array_equal(e.level_2_data()[1], o.level_2_data())
