"""GPU suite for the multi-asset types (bourse_b200.market, bb_config.assets): the reference's known-answer tests for
Market / MarketEnv, and random multi-asset, multi-market runs bit-exact against the oracle (shared shuffled queue,
global event times, per-asset records)."""
import types

import numpy as np
import pytest

from bourse_b200 import market

from . import scenarios_market as sm

pytestmark = pytest.mark.gpu


def _ns(**kw):
    def env(*a, **k): return market.MarketEnv(*a, **{**kw, **k})
    def mk(*a, **k): return market.Market(*a, max_orders=4096, max_trades=4096, max_steps=16, **k)
    return types.SimpleNamespace(Market=mk, MarketEnv=env)


@pytest.mark.parametrize("scenario", sm.ALL_MARKET + sm.ALL_MARKET_ENV, ids=lambda f: f.__name__)
def test_reference_known_answers_market(scenario):
    scenario(_ns(max_orders=4096, max_trades=4096, max_queue=64))


@pytest.mark.parametrize("scenario", sm.ALL_MARKET_ENV, ids=lambda f: f.__name__)
def test_reference_known_answers_market_env_dense(scenario):
    scenario(_ns(max_orders=4096, max_trades=4096, max_queue=64, price_window=(0, 256), live_cap=128))


def _drive(rng, envs, n_assets, n_steps, n_markets=1):
    """Feed identical random instructions to every env in `envs` (objects with the MarketEnv instruction API taking a
    `market=` keyword or wrapped to accept it)."""
    ids = [[[] for _ in range(n_assets)] for _ in range(n_markets)]
    for step in range(n_steps):
        for mk in range(n_markets):
            for _ in range(int(rng.integers(0, 25))):
                a = int(rng.integers(n_assets))
                u = rng.random()
                if u < 0.6 or not ids[mk][a]:
                    bid, vol = bool(rng.random() < 0.5), int(rng.integers(1, 40))
                    price = None if rng.random() < 0.08 else int(rng.integers(80, 121))
                    trader = int(rng.integers(100))
                    got = {e.place_order(a, bid, vol, trader, price, market=mk) for e in envs}
                    assert len(got) == 1
                    ids[mk][a].append(got.pop()[1])
                elif u < 0.8:
                    i = ids[mk][a][int(rng.integers(len(ids[mk][a])))]
                    for e in envs:
                        e.cancel_order((a, i), market=mk)
                else:
                    i = ids[mk][a][int(rng.integers(len(ids[mk][a])))]
                    p = None if rng.random() < 0.3 else int(rng.integers(80, 121))
                    v = None if (p is not None and rng.random() < 0.3) else int(rng.integers(1, 40))
                    for e in envs:
                        e.modify_order((a, i), p, v, market=mk)
        for e in envs:
            e.step()


class _OracleMarkets:
    """`n_markets` oracle MarketEnvs behind the batched interface (market m shuffles with seed + m)."""

    def __init__(self, oracle, seed, n_assets, n_markets, step_size):
        self.m = [oracle.MarketEnv(seed + k, 0, [1] * n_assets, step_size) for k in range(n_markets)]

    def place_order(self, a, bid, vol, trader, price, market=0): return self.m[market].place_order(a, bid, vol, trader, price)
    def cancel_order(self, oid, market=0): self.m[market].cancel_order(oid)
    def modify_order(self, oid, p, v, market=0): self.m[market].modify_order(oid, p, v)

    def step(self):
        for x in self.m:
            x.step()


@pytest.mark.parametrize("kw", [dict(), dict(price_window=(64, 192), live_cap=254)], ids=["paged", "dense"])
def test_random_markets_bit_exact(oracle, kw):
    n_assets, n_markets, n_steps, seed = 3, 4, 40, 77
    g = market.MarketEnv(seed, 0, [1] * n_assets, 10_000, n_markets=n_markets, max_orders=4096, max_trades=8192, max_queue=128, **kw)
    o = _OracleMarkets(oracle, seed, n_assets, n_markets, 10_000)
    _drive(np.random.default_rng(3), [g, o], n_assets, n_steps, n_markets)
    assert not g.env_errors().any()
    n_tr = 0
    for mk in range(n_markets):
        assert g.time(mk) == o.m[mk].time() == n_steps * 10_000
        for a in range(n_assets):
            assert np.array_equal(g.get_level_2_data_history(a, mk), o.m[mk].history(a)), (mk, a)
            assert g.get_orders(a, mk) == o.m[mk].get_orders(a), (mk, a)
            assert g.get_trades(a, mk) == o.m[mk].get_trades(a), (mk, a)
            n_tr += len(g.get_trades(a, mk))
        assert g.bid_asks(mk) == o.m[mk].bid_asks()
    assert n_tr > 500


def test_event_times_are_global_across_assets(oracle):
    """Event i of the market's shuffled queue runs at start + i whichever book it belongs to (market_env.rs:116-120):
    the arrival times of one step's orders, taken over all assets, are a permutation of start .. start + n - 1."""
    g = market.MarketEnv(5, 0, [1, 1, 1, 1], 1_000_000, max_orders=1024, max_trades=1024, max_queue=128)
    for k in range(60):
        g.place_order(k % 4, k % 2 == 0, 5, 0, 100 + (k % 7) - (10 if k % 2 == 0 else -10))
    g.step()
    arr = sorted(o[2] for a in range(4) for o in g.get_orders(a))
    assert arr == list(range(60))
    assert sorted(o[2] for o in g.get_orders(0)) != list(range(15))   # and they really are interleaved


def test_markets_reject_builtin_agents():
    from bourse_b200 import core, workloads
    e = core.BatchedEnv(4, 0, 0, 1, 1000, assets=2, max_orders=256, max_trades=256, max_steps=8, max_queue=16)
    e.set_agents(workloads.c3_groups())
    with pytest.raises(ValueError):
        e.run_agents(1, 0)
    with pytest.raises(RuntimeError, match="multiples of assets"):
        core.BatchedEnv(3, 0, 0, 1, 1000, assets=2, max_orders=256, max_trades=256, max_steps=8, max_queue=16)
