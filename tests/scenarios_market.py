"""Known-answer scenarios for the multi-asset types, restated from the reference's own tests:
crates/order_book/src/market.rs:397-574 (Market) and crates/step_sim/src/market_env.rs:333-407 (MarketEnv).

`m` exposes `Market(start_time, tick_size, trading)` and `MarketEnv(seed, start_time, tick_sizes, step_size, trading)`
with the reference's Rust method names, so the same numbers are asserted against the CPU oracle and the CUDA path.
Sides are booleans (True = Bid) and order ids `(asset, id)` tuples as everywhere in this repo."""
MAX_PRICE = 2**32 - 1
BID, ASK = True, False
ST_ACTIVE, ST_FILLED, ST_CANCELLED = 1, 2, 3


def market_init(m):  # market.rs:397-409
    market = m.Market(101, [1, 2], True)
    assert market.get_time() == 101
    assert market.bid_vols() == [0, 0] and market.ask_vols() == [0, 0]
    assert market.bid_best_vols() == [0, 0] and market.ask_best_vols() == [0, 0]
    assert market.bid_best_vol_and_orders() == [(0, 0), (0, 0)]
    assert market.ask_best_vol_and_orders() == [(0, 0), (0, 0)]
    assert market.bid_asks() == [(0, MAX_PRICE), (0, MAX_PRICE)]


def market_insert_order(m):  # market.rs:412-487 (no set_time between orders: equal-key collisions)
    market = m.Market(101, [1, 2], True)
    market.create_and_place_order(0, ASK, 10, 0, 100)
    market.create_and_place_order(0, BID, 10, 0, 50)
    assert market.bid_asks() == [(50, 100), (0, MAX_PRICE)]
    assert market.ask_vols() == [10, 0] and market.bid_vols() == [10, 0]
    assert market.bid_best_vols() == [10, 0] and market.ask_best_vols() == [10, 0]
    assert market.bid_best_vol_and_orders() == [(10, 1), (0, 0)]
    assert market.ask_best_vol_and_orders() == [(10, 1), (0, 0)]
    market.create_and_place_order(1, ASK, 20, 0, 20)
    market.create_and_place_order(1, BID, 20, 0, 10)
    assert market.bid_asks() == [(50, 100), (10, 20)]
    assert market.ask_vols() == [10, 20] and market.bid_vols() == [10, 20]
    assert market.bid_best_vol_and_orders() == [(10, 1), (20, 1)]
    assert market.ask_best_vol_and_orders() == [(10, 1), (20, 1)]
    market.create_and_place_order(0, ASK, 10, 0, 90)
    market.create_and_place_order(0, BID, 10, 0, 60)
    assert market.bid_asks() == [(60, 90), (10, 20)]
    assert market.ask_vols() == [20, 20] and market.bid_vols() == [20, 20]
    assert market.bid_best_vol_and_orders() == [(10, 1), (20, 1)]
    assert market.ask_best_vol_and_orders() == [(10, 1), (20, 1)]
    market.create_and_place_order(1, ASK, 10, 0, 20)
    market.create_and_place_order(1, BID, 10, 0, 12)
    assert market.bid_asks() == [(60, 90), (12, 20)]
    assert market.ask_vols() == [20, 30] and market.bid_vols() == [20, 30]
    assert market.bid_best_vols() == [10, 10] and market.ask_best_vols() == [10, 30]
    assert market.bid_best_vol_and_orders() == [(10, 1), (10, 1)]
    assert market.ask_best_vol_and_orders() == [(10, 1), (30, 2)]
    market.create_and_place_order(0, ASK, 10, 0, 110)
    market.create_and_place_order(0, BID, 10, 0, 40)
    assert market.bid_asks() == [(60, 90), (12, 20)]
    assert market.ask_vols() == [30, 30] and market.bid_vols() == [30, 30]
    assert market.bid_best_vol_and_orders() == [(10, 1), (10, 1)]
    assert market.ask_best_vol_and_orders() == [(10, 1), (30, 2)]


def market_cancel_order(m):  # market.rs:490-546
    market = m.Market(0, [1, 2], True)
    market.create_and_place_order(0, ASK, 10, 0, 100)
    market.create_and_place_order(0, BID, 10, 0, 50)
    market.create_and_place_order(0, ASK, 10, 0, 90)
    market.create_and_place_order(0, BID, 10, 0, 60)
    market.create_and_place_order(1, ASK, 50, 0, 20)
    market.create_and_place_order(1, BID, 50, 0, 10)
    assert market.bid_asks() == [(60, 90), (10, 20)]
    assert market.ask_vols() == [20, 50] and market.bid_vols() == [20, 50]
    assert market.bid_best_vol_and_orders() == [(10, 1), (50, 1)]
    assert market.ask_best_vol_and_orders() == [(10, 1), (50, 1)]
    market.cancel_order((0, 0))
    market.cancel_order((0, 3))
    assert market.bid_asks() == [(50, 90), (10, 20)]
    assert market.ask_vols() == [10, 50] and market.bid_vols() == [10, 50]
    assert market.bid_best_vol_and_orders() == [(10, 1), (50, 1)]
    assert market.ask_best_vol_and_orders() == [(10, 1), (50, 1)]
    market.cancel_order((0, 1))
    market.cancel_order((0, 2))
    assert market.bid_asks() == [(0, MAX_PRICE), (10, 20)]
    assert market.ask_vols() == [0, 50] and market.bid_vols() == [0, 50]
    assert market.bid_best_vol_and_orders() == [(0, 0), (50, 1)]
    assert market.ask_best_vol_and_orders() == [(0, 0), (50, 1)]
    assert [market.order((0, i))[1] for i in range(4)] == [ST_CANCELLED] * 4


def market_mod_order_vol(m):  # market.rs:549-573
    market = m.Market(0, [1, 2], True)
    market.create_and_place_order(0, ASK, 10, 0, 100)
    market.create_and_place_order(0, BID, 10, 0, 50)
    market.modify_order((0, 0), None, 8)
    market.modify_order((0, 1), None, 5)
    assert market.ask_vols() == [8, 0] and market.ask_best_vols() == [8, 0]
    assert market.ask_best_vol_and_orders() == [(8, 1), (0, 0)]
    assert market.bid_vols() == [5, 0] and market.bid_best_vols() == [5, 0]
    assert market.bid_best_vol_and_orders() == [(5, 1), (0, 0)]
    assert market.order((0, 0))[4] == 8 and market.order((0, 1))[4] == 5


def market_env_rust(m):  # market_env.rs:333-407
    step_size = 1000
    env = m.MarketEnv(101, 0, [1, 1], step_size, True)
    env.place_order(0, BID, 10, 101, 10)
    env.place_order(0, ASK, 20, 101, 20)
    env.step()
    assert env.bid_asks() == [(10, 20), (0, MAX_PRICE)]
    assert [o[1] for o in env.get_orders(0)] == [ST_ACTIVE, ST_ACTIVE]
    assert env.time() == step_size
    env.place_order(0, BID, 10, 101, 11)
    env.place_order(0, ASK, 20, 101, 21)
    env.step()
    assert env.bid_asks() == [(11, 20), (0, MAX_PRICE)]
    assert len(env.get_orders(0)) == 4 and env.time() == 2 * step_size
    env.place_order(0, BID, 30, 101, None)
    env.step()
    assert env.bid_asks() == [(11, 21), (0, MAX_PRICE)]
    orders = env.get_orders(0)
    assert len(orders) == 5 and orders[1][1] == ST_FILLED and orders[4][1] == ST_FILLED
    assert len(env.get_trades(0)) == 2 and env.time() == 3 * step_size
    assert len(env.get_orders(1)) == 0 and len(env.get_trades(1)) == 0
    bp, ap = env.get_prices(0)
    assert list(bp) == [10, 11, 11] and list(ap) == [20, 20, 21]
    bv, av = env.get_volumes(0)
    assert list(bv) == [10, 20, 20] and list(av) == [20, 40, 10]
    tb, ta = env.get_touch_volumes(0)
    assert list(tb) == [10, 10, 10] and list(ta) == [20, 20, 10]
    nb, na = env.get_touch_order_counts(0)
    assert list(nb) == [1, 1, 1] and list(na) == [1, 1, 1]
    assert list(env.get_trade_vols(0)) == [0, 0, 30]
    assert list(env.get_trade_vols(1)) == [0, 0, 0]   # every asset records every step (market_env.rs:127-130)


def market_env_price_error(m):  # Market::create_order propagates OrderError::PriceError (market.rs:244-254)
    import pytest
    env = m.MarketEnv(0, 0, [2, 2], 1000, True)
    with pytest.raises(ValueError):
        env.place_order(1, BID, 10, 0, 11)
    assert env.place_order(1, BID, 10, 0, 12) == (1, 0)   # the failed call created nothing


ALL_MARKET = [market_init, market_insert_order, market_cancel_order, market_mod_order_vol]
ALL_MARKET_ENV = [market_env_rust, market_env_price_error]
