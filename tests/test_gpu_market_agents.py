"""GPU suite for the in-kernel market agent twins (k_sim<.., MKT>, bb_set_agents_market): RandomMarketAgents,
MomentumMarketAgent and NoiseMarketAgent driving multi-asset markets inside one persistent kernel, bit-exact against
the oracle's MarketSim (same Philox contract keyed by market) on both engines, every (market, asset) compared on its
level-2 history, trade log and order table."""
import numpy as np
import pytest

from bourse_b200 import market

pytestmark = pytest.mark.gpu

MOM = market.MomentumParams(tick_size=1, p_cancel=0.1, trade_vol=10, decay=1.0, demand=5.0, scale=0.5, order_ratio=1.0,
                            price_dist_mu=0.0, price_dist_sigma=1.0)
NOISE = market.NoiseAgentParams(tick_size=1, p_limit=0.3, p_market=0.1, p_cancel=0.2, trade_vol=7, price_dist_mu=0.5,
                                price_dist_sigma=0.8)


def example_agents():
    """crates/step_sim/examples/multi_asset/main.rs:15-20"""
    return [market.RandomMarketAgents(0, 50, (40, 60), (10, 20), 2, 0.8), market.RandomMarketAgents(0, 50, (10, 90), (50, 70), 2, 0.2),
            market.RandomMarketAgents(1, 50, (40, 60), (10, 20), 2, 0.8), market.RandomMarketAgents(1, 50, (10, 90), (50, 70), 2, 0.2)]


def mixed_agents(n_assets, noise=True):
    """noise=False: the population for a dense price window — no noise traders (they quote around the mid price of a
    possibly empty book, 2^31) and momentum traders that only send market orders (order_ratio 0)."""
    out = []
    for a in range(n_assets):
        out.append(market.RandomMarketAgents(a, 40 - 5 * a, (40, 60), (10, 20), 2, 0.7))
    out.append(market.MomentumMarketAgent(100, 12, n_assets - 1, MOM if noise else MOM._replace(order_ratio=0.0)))
    if noise:
        out.append(market.NoiseMarketAgent(0, 200, 9, NOISE))
    out.append(market.RandomMarketAgents(n_assets - 1, 17, (10, 90), (50, 70), 2, 0.3))
    return out


def _oracle_market(oracle, agents, n_assets, step_size, runs, seed, market_id):
    m = oracle.MarketEnv(0, 0, [1] * n_assets, step_size)
    m.set_groups([a.group for a in agents], [a.asset for a in agents])
    for n in runs:
        m.run_agents(n, seed, market_id=market_id)
    return m


def _compare(g, oracle, agents, n_assets, n_markets, step_size, runs, seed, id_base=0):
    assert not g.env_errors().any(), g.env_errors()
    n_tr = 0
    for mk in range(n_markets):
        o = _oracle_market(oracle, agents, n_assets, step_size, runs, seed, id_base + mk)
        for a in range(n_assets):
            assert np.array_equal(g.get_level_2_data_history(a, mk), o.history(a)), (mk, a)
            assert g.get_trades(a, mk) == o.get_trades(a), (mk, a)
            assert g.get_orders(a, mk) == o.get_orders(a), (mk, a)
            n_tr += len(o.get_trades(a))
    return n_tr


ENGINES = [dict(), dict(pages_smem=16, pages_total=64), dict(price_window=(0, 256), live_cap=128)]
ENGINE_IDS = ["fast", "paged", "dense"]


@pytest.mark.parametrize("kw", ENGINES, ids=ENGINE_IDS)
def test_multi_asset_example_bit_exact(oracle, kw):
    """The reference's multi_asset example population, 7 lockstep markets (so the last CTA is half empty)."""
    n_markets, n_steps, seed = 7, 60, 101
    g = market.MarketEnv(0, 0, [1, 1], 1_000_000, n_markets=n_markets, max_orders=8192, max_trades=8192, max_queue=128, **kw)
    agents = example_agents()
    market.market_sim_runner(g, agents, seed, n_steps)
    n_tr = _compare(g, oracle, agents, 2, n_markets, 1_000_000, [n_steps], seed)
    assert n_tr > 10_000
    s = g.stats()
    assert s["env_steps"] == n_markets * 2 * n_steps and s["error_envs"] == 0


@pytest.mark.parametrize("n_assets", [2, 3, 4])
@pytest.mark.parametrize("kw", [ENGINES[0], dict(price_window=(0, 1024), live_cap=254)], ids=["fast", "dense_l"])
def test_mixed_agent_twins_bit_exact(oracle, n_assets, kw):
    """Random + Momentum + Noise twins on 2, 3 (three-warp CTAs) and 4 assets, launches split 11 + 19."""
    n_markets, seed = 5, 2024
    g = market.MarketEnv(0, 0, [1] * n_assets, 100_000, n_markets=n_markets, max_orders=8192, max_trades=8192, max_queue=128, **kw)
    agents = mixed_agents(n_assets, noise="price_window" not in kw)
    g.set_agents(agents)
    g.run_agents(11, seed)
    g.run_agents(19, seed)
    n_tr = _compare(g, oracle, agents, n_assets, n_markets, 100_000, [30], seed)
    assert n_tr > 1000


def test_market_rng_is_keyed_by_global_market_id(oracle):
    """A shard starting at env_id_base = 6 books (market 3) reproduces markets 3.. of the unsharded run."""
    from bourse_b200 import abi, core
    agents = example_agents()
    e = core.BatchedEnv(8, 0, 0, 1, 1_000_000, assets=2, env_id_base=6, obs_words=abi.OBS_L2, max_orders=4096, max_trades=4096,
                        max_steps=32, max_queue=128)
    e.set_agents([a.group for a in agents], assets=[a.asset for a in agents])
    e.run_agents(25, 9)
    assert not e.env_errors().any()
    for mk in range(4):
        o = _oracle_market(oracle, agents, 2, 1_000_000, [25], 9, 3 + mk)
        for a in range(2):
            assert np.array_equal(e.history(2 * mk + a), o.history(a))
            assert e.get_trades(2 * mk + a) == o.get_trades(a)


def test_host_instructions_between_agent_launches(oracle):
    """Host-queued MarketEnv instructions and in-kernel agent steps interleave on the same books."""
    agents = example_agents()
    g = market.MarketEnv(0, 0, [1, 1], 1_000_000, max_orders=4096, max_trades=4096, max_queue=128)
    g.set_agents(agents)
    g.run_agents(10, 5)
    before = [len(g.get_orders(a)) for a in (0, 1)]
    oid = g.place_order(1, True, 5, 999, 2)
    g.step()
    assert g.order_status(oid) == 1 and oid[1] == before[1]
    g.run_agents(10, 5)
    assert not g.env_errors().any()
    assert g.get_orders(1)[oid[1]][7] == 999 and len(g.get_level_2_data_history(0)) == 21


def test_market_agent_argument_checks():
    from bourse_b200 import core
    agents = example_agents()
    e = core.BatchedEnv(2, 0, 0, 1, 1000, max_orders=256, max_trades=256, max_steps=8, max_queue=16)
    with pytest.raises(ValueError, match="assets > 1"):
        e.set_agents([a.group for a in agents], assets=[a.asset for a in agents])
    e = core.BatchedEnv(4, 0, 0, 1, 1000, assets=2, max_orders=256, max_trades=256, max_steps=8, max_queue=16)
    with pytest.raises(ValueError, match="out of range"):
        e.set_agents([agents[0].group], assets=[2])
    e = core.BatchedEnv(10, 0, 0, 1, 1000, assets=5, max_orders=256, max_trades=256, max_steps=8, max_queue=16)
    with pytest.raises(ValueError, match="at most 4 assets"):
        e.set_agents([agents[0].group], assets=[0])
