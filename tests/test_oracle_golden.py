"""CPU suite: pins the oracle against the reference's own known-answer tests (see scenarios.py)."""
import numpy as np
import pytest

from . import scenarios


@pytest.mark.parametrize("scenario", scenarios.ALL_BOOK + scenarios.ALL_ENV + scenarios.ALL_NUMPY,
                         ids=lambda f: f.__name__)
def test_reference_known_answers(oracle, scenario):
    scenario(oracle)


def test_rounding_known_answers(oracle):
    """crates/step_sim/src/agents/common.rs:268-305"""
    L = oracle.lib()
    assert L.orc_round_price(5.0, 2.0, 1) == 6
    assert L.orc_round_price(2.1, 2.0, 1) == 4
    assert L.orc_round_price(3.9, 4.0, 1) == 4
    assert L.orc_round_price(-2.2, 4.0, 1) == 0
    assert L.orc_round_price(1.0 + 2.0**32, 4.0, 1) == 2**32 - 1
    assert L.orc_round_price(5.0, 2.0, 0) == 4
    assert L.orc_round_price(2.1, 2.0, 0) == 2
    assert L.orc_round_price(3.9, 4.0, 0) == 0
    assert L.orc_round_price(-2.2, 4.0, 0) == 0
    assert L.orc_round_price(1.0 + 2.0**32, 4.0, 0) == 2**32 - 1


def test_random_agents_activity(oracle):
    """crates/step_sim/src/agents/random_agent.rs:255-296 (activity 0 / 1; place -> cancel -> place)"""
    for keyed in (False, True):
        env = oracle.StepEnv(101, 0, 1, 1000)
        env.set_groups([oracle.random_group(2, (10, 20), (20, 30), 1, 0.0)])
        env.run_agents(3, 101, keyed=keyed)
        assert env.n_instructions() == 0 and env.get_orders() == []

        env = oracle.StepEnv(101, 0, 1, 1000)
        env.set_groups([oracle.random_group(1, (10, 20), (20, 30), 1, 1.0)])
        env.run_agents(1, 101, keyed=keyed)
        orders = env.get_orders()
        assert len(orders) == 1 and orders[0][1] == 1 and 10 <= orders[0][6] < 20 and 20 <= orders[0][4] < 30
        env.run_agents(1, 101, keyed=keyed)
        assert env.get_orders()[0][1] == 3 and len(env.get_orders()) == 1
        env.run_agents(1, 101, keyed=keyed)
        assert len(env.get_orders()) == 2 and env.get_orders()[1][1] == 1
        assert env.n_instructions() == 3


def test_momentum_first_update_emits_nothing(oracle):
    """crates/step_sim/src/agents/momentum_agent.rs:411-444"""
    for keyed in (False, True):
        env = oracle.StepEnv(101, 0, 1, 1_000_000)
        env.place_order(True, 100, 0, price=1000)
        env.place_order(False, 100, 0, price=1020)
        env.step()
        env.set_groups([oracle.momentum_group(10, 100, 2, 0.1, 100, 1.0, 5.0, 0.5, 1.0, 0.0, 10.0)])
        env.run_agents(1, 101, keyed=keyed)
        assert env.n_instructions() == 0 and len(env.get_orders()) == 2


def test_momentum_only_buys_fire(oracle):
    """momentum_agent.rs:152-202: p_market carries the sign of the momentum, so only m>0 buys fire."""
    env = oracle.StepEnv(5, 0, 1, 1_000_000)
    env.place_order(True, 1000, 0, price=1000)
    env.place_order(False, 1000, 0, price=1020)
    env.step()
    env.set_groups([oracle.momentum_group(10, 50, 1, 0.1, 10, 1.0, 5.0, 0.5, 1.0, 0.0, 1.0)])
    env.run_agents(1, 9)                      # primes last_price = 1010
    env.place_order(True, 5, 0, price=1010)   # mid moves up to 1015 -> m = +5
    env.run_agents(20, 9)
    new = env.get_orders()[3:]
    assert len(new) > 0 and all(o[0] for o in new)      # bids only
    env2 = oracle.StepEnv(5, 0, 1, 1_000_000)
    env2.place_order(True, 1000, 0, price=1000)
    env2.place_order(False, 1000, 0, price=1020)
    env2.step()
    env2.set_groups([oracle.momentum_group(10, 50, 1, 0.1, 10, 1.0, 5.0, 0.5, 1.0, 0.0, 1.0)])
    env2.run_agents(1, 9)
    env2.place_order(False, 5, 0, price=1010)  # mid moves down -> m < 0 -> p < 0 -> nothing fires
    env2.run_agents(5, 9)
    assert len(env2.get_orders()) == 3


def test_n1_same_key_ghost(oracle):
    """two live orders with the same (price, t) share one queue key
    (side.rs:55 insert overwrites); totals count both, only the later one is reachable."""
    ob = oracle.OrderBook(0, 1)
    a = ob.place_order(False, 10, 0, price=100)
    b = ob.place_order(False, 7, 0, price=100)      # same t=0 -> overwrites a's queue slot
    assert ob.best_ask_vol_and_orders() == (17, 2)
    ob.set_time(1)
    c = ob.place_order(True, 12, 1, price=100)      # matches b (7) only; a is a ghost
    assert ob.get_trades() == [(1, False, 100, 7, c, b)]
    assert ob.order_status(a) == 1 and ob.order_status(b) == 2 and ob.order_status(c) == 1
    # the book is now crossed: bid 100 rests against the unreachable ask total
    assert ob.bid_ask() == (100, 2**32 - 1)
    assert ob.best_ask_vol_and_orders() == (10, 1) and ob.ask_vol() == 10
    ob.cancel_order(a)
    assert ob.ask_vol() == 0 and ob.best_ask_vol_and_orders() == (0, 0)


def test_zero_volume_orders(oracle):
    """a zero-volume limit rests (count+1) and yields a zero-volume trade when hit."""
    ob = oracle.OrderBook(0, 1)
    z = ob.place_order(False, 0, 0, price=50)
    assert ob.best_ask_vol_and_orders() == (0, 1) and ob.order_status(z) == 1
    ob.set_time(1)
    ob.place_order(False, 5, 0, price=50)
    ob.set_time(2)
    agg = ob.place_order(True, 3, 0, price=50)
    assert ob.get_trades() == [(2, False, 50, 0, agg, z), (2, False, 50, 3, agg, 1)]
    assert ob.order_status(z) == 2 and ob.order_status(agg) == 2


def test_modify_rules(oracle):
    """SURVEY.md N4 (orderbook.rs:743-772): equal volume or any price => replace (loses priority)."""
    ob = oracle.OrderBook(0, 1)
    a = ob.place_order(True, 10, 0, price=50)
    ob.set_time(1)
    b = ob.place_order(True, 10, 0, price=50)
    ob.set_time(2)
    ob.modify_order(a, new_vol=10)          # equal volume -> replace -> a goes behind b
    ob.set_time(3)
    s = ob.place_order(False, 10, 0, price=50)
    assert ob.get_trades() == [(3, True, 50, 10, s, b)]
    ob.modify_order(a, new_vol=4)           # reduce keeps priority and start_vol
    assert ob.get_orders()[a][4:6] == (4, 10)
    ob.modify_order(b, new_vol=99)          # dead order: no-op
    assert ob.get_orders()[b][1] == 2 and ob.get_orders()[b][4] == 0
    ob.modify_order(a)                      # (None, None): no-op
    assert ob.bid_vol() == 4
    ob.modify_order(a, new_price=51)        # off-tick impossible at tick 1; price-only replace keeps vol
    assert ob.bid_ask()[0] == 51 and ob.bid_vol() == 4


def test_l2_wrapping_levels(oracle):
    """orderbook.rs:229-236, 257-264: level prices wrap (empty ask side starts at u32::MAX)."""
    ob = oracle.OrderBook(0, 1)
    ob.place_order(False, 9, 0, price=2)      # ask at 2: best ask 2
    ob.place_order(True, 4, 0, price=0)       # bid resting at price 0 looks like an empty side's touch
    l2 = ob.level_2_data()
    assert list(l2[:5]) == [0, 0, 2, 9, 4]
    assert tuple(l2[5:9]) == (4, 1, 9, 1)
    ob2 = oracle.OrderBook(0, 1)
    ob2.place_order(False, 3, 0, price=1)
    ob2.place_order(False, 5, 0, price=2**32 - 1)     # limit ask at u32::MAX: allowed, not a market order
    ob2.cancel_order(0)
    l2 = ob2.level_2_data()
    assert l2[2] == 2**32 - 1 and tuple(l2[7:9]) == (5, 1)
    assert ob2.mid_price() == pytest.approx((2**32 - 1) / 2)


def test_noise_agent_known_answers(oracle):
    """crates/step_sim/src/agents/noise_agent.rs:360-425: trader ids agent_id_start.., all-limit then all-cancel."""
    for keyed in (False, True):
        env = oracle.StepEnv(101, 0, 1, 1_000_000)
        env.set_groups([oracle.noise_group(10, 10, 2, 1.0, 0.0, 1.0, 100, 0.0, 10.0)])
        env.run_agents(1, 101, keyed=keyed)
        orders = env.get_orders()
        assert len(orders) == 10 and env.n_instructions() == 10
        assert [o[7] for o in orders] == list(range(10, 20))          # trader ids
        mid = (2**32 - 1) / 2                                         # empty book: bid 0, ask u32::MAX
        for o in orders:
            assert o[1] == 1 and o[4] == 100 and (o[6] % 2 == 0 or o[6] == 2**32 - 1)   # clamp to u32::MAX (common.rs:24)
            assert (o[6] <= mid) if o[0] else (o[6] >= mid)
        env.run_agents(1, 101, keyed=keyed)                           # p_cancel = 1: every live order is cancelled
        orders = env.get_orders()
        assert len(orders) == 20 and all(o[1] == 3 for o in orders[:10]) and all(o[1] == 1 for o in orders[10:])
        env2 = oracle.StepEnv(101, 0, 1, 1_000_000)
        env2.set_groups([oracle.noise_group(0, 8, 1, 0.0, 1.0, 0.1, 7, 0.0, 1.0)])
        env2.run_agents(3, 5, keyed=keyed)                            # market orders only, empty book: all cancelled
        assert len(env2.get_orders()) == 24 and all(o[1] == 3 and o[6] in (0, 2**32 - 1) for o in env2.get_orders())
