"""CPU suite for the oracle's market agent twins (oracle/env.hpp MarketSim): the invariants the reference's agent tests
assert (crates/step_sim/src/agents/random_agent.rs:255-296, noise_agent.rs:377-425, momentum_agent.rs:419-444), restated
for RandomMarketAgents / NoiseMarketAgent / MomentumMarketAgent, plus structural checks of the market-wide queue."""
import numpy as np
import pytest

from oracle import oracle as orc


def _example_groups():
    # crates/step_sim/examples/multi_asset/main.rs:15-20
    g = [orc.random_group(50, (40, 60), (10, 20), 2, 0.8), orc.random_group(50, (10, 90), (50, 70), 2, 0.2)]
    return g + g, [0, 0, 1, 1]


@pytest.mark.parametrize("keyed", [True, False])
def test_activity_rate(keyed):
    # random_agent.rs:255-267: rate 0 => nothing queued, rate 1 => every agent queues one instruction
    m = orc.MarketEnv(0, 0, [1, 1], 1000)
    m.set_groups([orc.random_group(5, (10, 20), (20, 30), 1, 0.0), orc.random_group(7, (10, 20), (20, 30), 1, 1.0)], [0, 1])
    m.run_agents(1, 101, keyed=keyed)
    assert m.n_instructions() == 7
    assert len(m.get_orders(0)) == 0 and len(m.get_orders(1)) == 7


@pytest.mark.parametrize("keyed", [True, False])
def test_order_place_then_cancel(keyed):
    # random_agent.rs:270-296 on asset 1 of a two-asset market: bids and asks far apart never trade, so after two steps
    # every order placed in the first step is cancelled by its owner in the second
    m = orc.MarketEnv(0, 0, [1, 1], 1000)
    m.set_groups([orc.random_group(4, (10, 11), (20, 30), 1, 1.0)], [1])
    m.run_agents(1, 7, keyed=keyed)
    o = m.get_orders(1)
    assert len(o) == 4 and all(x[1] in (1, 2) for x in o)      # Active, or Filled where a bid met an ask at the same price
    n_active = sum(x[1] == 1 for x in o)
    m.run_agents(1, 7, keyed=keyed)
    o2 = m.get_orders(1)
    assert sum(x[1] == 3 for x in o2[:4]) == n_active            # the Active ones are now Cancelled
    assert len(o2) == 4 + (4 - n_active)                         # agents whose order had been filled placed a new one
    assert len(m.get_orders(0)) == 0


def test_noise_place_and_cancel():
    # noise_agent.rs:377-425: p_limit = 1 => one limit order per agent per step; p_cancel = 1 => all cancelled next step
    m = orc.MarketEnv(0, 0, [2, 2], 1000)
    m.set_groups([orc.noise_group(10, 4, 2, 1.0, 0.0, 1.0, 100, 0.0, 1.0)], [0])
    m.run_agents(1, 3)
    o = m.get_orders(0)
    assert len(o) == 4 and [x[7] for x in o] == [10, 11, 12, 13] and all(x[6] % 2 == 0 for x in o)
    m.run_agents(1, 3)
    o = m.get_orders(0)
    assert len(o) == 8 and all(x[1] != 1 for x in o[:4])  # first batch gone (cancelled, or filled on the way)


def test_momentum_no_orders_without_history():
    # momentum_agent.rs:419-444: the first update has no last price => no orders
    m = orc.MarketEnv(0, 0, [1, 1], 1000)
    m.set_groups([orc.momentum_group(0, 5, 1, 0.1, 10, 1.0, 5.0, 0.5, 1.0, 0.0, 1.0)], [1])
    m.run_agents(1, 11)
    assert m.n_instructions() == 0


def test_event_times_are_market_wide():
    """Event i of the market's shuffled queue runs at start + i on its asset's book (market_env.rs:116-121): within a
    step the arrival times of the two assets' orders interleave without repeats."""
    groups, assets = _example_groups()
    m = orc.MarketEnv(0, 0, [1, 1], 1_000_000)
    m.set_groups(groups, assets)
    m.run_agents(5, 101)
    arr = [np.array([o[2] for o in m.get_orders(a) if o[1] != 0]) for a in (0, 1)]
    both = np.concatenate(arr)
    assert len(np.unique(both)) == len(both)
    step0 = [a[a < 1_000_000] for a in arr]
    n0 = len(step0[0]) + len(step0[1])
    assert n0 > 0 and max(s.max() for s in step0) < m.n_instructions()
    assert len(step0[0]) and len(step0[1])


def test_keyed_run_is_split_invariant_and_market_keyed():
    groups, assets = _example_groups()
    a = orc.MarketEnv(0, 0, [1, 1], 1_000_000); a.set_groups(groups, assets); a.run_agents(40, 101, market_id=5)
    b = orc.MarketEnv(0, 0, [1, 1], 1_000_000); b.set_groups(groups, assets)
    b.run_agents(15, 101, market_id=5); b.run_agents(25, 101, market_id=5)
    c = orc.MarketEnv(0, 0, [1, 1], 1_000_000); c.set_groups(groups, assets); c.run_agents(40, 101, market_id=6)
    for asset in (0, 1):
        assert np.array_equal(a.history(asset), b.history(asset)) and a.get_trades(asset) == b.get_trades(asset)
        assert not np.array_equal(a.history(asset), c.history(asset))


def test_bench_entry_counts():
    groups, assets = _example_groups()
    r = orc.bench_market_agents(8, 2, 50, 101, groups, assets, 2, keyed=True)
    tot = 0
    for mk in range(8):
        m = orc.MarketEnv(0, 0, [1, 1], 1_000_000); m.set_groups(groups, assets); m.run_agents(50, 101, market_id=mk)
        tot += m.n_instructions()
    assert r["instructions"] == tot and r["env_steps"] == 8 * 2 * 50
