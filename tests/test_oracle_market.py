"""CPU suite: the oracle's Market / MarketEnv restatement against the reference's known-answer tests."""
import types

import numpy as np
import pytest

from . import scenarios_market as sm


def _ns(oracle):
    class Env(oracle.MarketEnv):   # scenario-facing getters in the reference's Rust names
        def _cols(self, asset): return self.history(asset)
        def get_prices(self, a): h = self._cols(a); return h[:, 1], h[:, 2]
        def get_volumes(self, a): h = self._cols(a); return h[:, 4], h[:, 3]
        def get_touch_volumes(self, a): h = self._cols(a); return h[:, 5], h[:, 7]
        def get_touch_order_counts(self, a): h = self._cols(a); return h[:, 6], h[:, 8]
        def get_trade_vols(self, a): return self._cols(a)[:, 0]
    return types.SimpleNamespace(Market=oracle.Market, MarketEnv=Env)


@pytest.mark.parametrize("scenario", sm.ALL_MARKET + sm.ALL_MARKET_ENV, ids=lambda f: f.__name__)
def test_reference_known_answers_market(oracle, scenario):
    scenario(_ns(oracle))


def test_market_env_with_one_asset_is_env(oracle):
    """A one-asset MarketEnv is an Env: same shuffle stream, same times (market_env.rs:108-121 vs env.rs:116-135)."""
    rng = np.random.default_rng(5)
    me, e = oracle.MarketEnv(9, 0, [1], 1000), oracle.StepEnv(9, 0, 1, 1000)
    ids = []
    for step in range(30):
        for _ in range(int(rng.integers(0, 12))):
            u = rng.random()
            if u < 0.7 or not ids:
                bid, vol, price = bool(rng.random() < 0.5), int(rng.integers(1, 30)), int(rng.integers(90, 111))
                a = me.place_order(0, bid, vol, 1, price)
                b = e.place_order(bid, vol, 1, price)
                assert a == (0, b)
                ids.append(b)
            elif u < 0.85:
                i = ids[int(rng.integers(len(ids)))]
                me.cancel_order((0, i)); e.cancel_order(i)
            else:
                i = ids[int(rng.integers(len(ids)))]
                p, v = int(rng.integers(90, 111)), int(rng.integers(1, 30))
                me.modify_order((0, i), p, v); e.modify_order(i, p, v)
        me.step(); e.step()
    assert me.get_orders(0) == e.get_orders() and me.get_trades(0) == e.get_trades()
    assert np.array_equal(me.history(0), e._history())
