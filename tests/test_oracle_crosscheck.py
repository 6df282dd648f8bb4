"""CPU suite: the C++ oracle against the independent pure-Python restatement (oracle/pybook.py) on
random instruction streams, plus invariants the domain offers.  See SURVEY.md 8c(ii)/(iii)."""
import numpy as np
import pytest

from bourse_b200 import abi, workloads
from oracle.pybook import PyBook


def run_pybook(stream, tick_size):
    pb = PyBook(0, tick_size)
    obs = []
    for x in stream:
        of = int(x["op_flags"])
        op = of & 0xFF
        pb.t = int(x["t"])
        if op == abi.OP_NEW:
            pb.place(bool(of & abi.F_BID), int(x["vol"]), int(x["trader"]), None if of & abi.F_MARKET else int(x["price"]))
        elif op == abi.OP_CANCEL:
            pb.cancel(int(x["order_id"]))
        elif op == abi.OP_MODIFY:
            pb.modify(int(x["order_id"]), int(x["price"]) if of & abi.F_HAS_PRICE else None,
                      int(x["vol"]) if of & abi.F_HAS_VOL else None)
        elif op == abi.OP_SET_TRADING:
            pb.trading = bool(x["vol"])
        if of & abi.F_EMIT:
            obs.append(pb.l2())
    return pb, np.array(obs, dtype=np.uint32).reshape(-1, 45)


@pytest.mark.parametrize("time_mode", ["strict", "flat", "jitter"])
@pytest.mark.parametrize("tick_size,seed", [(1, 0), (2, 1), (1, 2)])
def test_oracle_matches_python_restatement(oracle, time_mode, tick_size, seed):
    n = 4000
    s = workloads.replay_stream(n, seed, tick_size=tick_size, half_width=12, time_mode=time_mode, emit_every=16,
                                min_vol=0 if seed == 2 else 1)
    ob = oracle.OrderBook(0, tick_size)
    obs = ob.replay(s, obs_cap=n)
    pb, pobs = run_pybook(s, tick_size)
    assert ob.get_trades() == pb.trades
    assert ob.get_orders() == pb.order_tuples()
    assert np.array_equal(obs, pobs)
    assert list(ob._l1()) == pb.l1()


def test_book_invariants(oracle):
    """Totals equal the sums over Active orders; never crossed while trading with unique keys."""
    s = workloads.replay_stream(20000, 7, tick_size=1, half_width=20, trading_windows=False)
    ob = oracle.OrderBook(0, 1)
    ob.replay(s)
    orders = ob.get_orders()
    act = [o for o in orders if o[1] == 1]
    assert ob.bid_vol() == sum(o[4] for o in act if o[0])
    assert ob.ask_vol() == sum(o[4] for o in act if not o[0])
    bid, ask = ob.bid_ask()
    assert bid < ask
    assert bid == max(o[6] for o in act if o[0]) and ask == min(o[6] for o in act if not o[0])
    trades = ob.get_trades()
    assert len(trades) > 1000
    for t, pbid, price, vol, a, p in trades[:2000]:
        assert orders[p][0] == pbid and orders[a][0] != pbid
    # statuses are consistent with end times
    for o in orders:
        assert (o[3] == 2**64 - 1) == (o[1] in (0, 1))


def test_stream_generator_is_deterministic():
    a = workloads.replay_stream(5000, 3)
    b = workloads.replay_stream(5000, 3)
    assert a.tobytes() == b.tobytes()
    assert np.all(np.diff(a["t"].astype(np.int64)) > 0)
    ops = a["op_flags"] & 0xFF
    assert 0.5 < np.mean(ops == abi.OP_NEW) < 0.7
