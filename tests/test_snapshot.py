"""JSON snapshot: schema of the writer on CPU, round trips on the GPU.
Reference: crates/order_book/src/orderbook.rs:804-832, 891-918; tests/test_order_book.py:190-210."""
import json

import numpy as np
import pytest

from bourse_b200 import snapshot, workloads


def test_json_schema_matches_serde_layout():
    """Hand-written expected text following serde_json's rules for the reference's structs."""
    orders = [(True, 1, 0, 2**64 - 1, 10, 10, 50, 11, 0), (False, 2, 3, 7, 0, 20, 60, 12, 1)]
    trades = [(7, False, 60, 20, 2, 1)]
    d = snapshot.book_to_dict(9, 1, 20, True, orders, [0, 3], trades)
    assert snapshot.dumps(d) == (
        '{"t":9,"tick_size":1,"trade_vol":20,"orders":['
        '{"order":{"side":"Bid","status":"Active","arr_time":0,"end_time":18446744073709551615,"vol":10,"start_vol":10,'
        '"price":50,"trader_id":11,"order_id":0},"key":["Bid",4294967245,0]},'
        '{"order":{"side":"Ask","status":"Filled","arr_time":3,"end_time":7,"vol":0,"start_vol":20,"price":60,'
        '"trader_id":12,"order_id":1},"key":["Ask",60,3]}],'
        '"trades":[{"t":7,"side":"Ask","price":60,"vol":20,"active_order_id":2,"passive_order_id":1}],"trading":true}')
    assert json.loads(snapshot.dumps(d, pretty=True)) == d
    c = snapshot.dict_to_columns(d)
    assert list(c["side"]) == [1, 0] and list(c["status"]) == [1, 2] and list(c["key_time"]) == [0, 3]
    assert c["end_time"][0] == 2**64 - 1 and c["tr_passive"][0] == 1
    bad = json.loads(snapshot.dumps(d))
    bad["orders"][0]["key"][1] = 5
    with pytest.raises(ValueError):
        snapshot.dict_to_columns(bad)


@pytest.mark.gpu
def test_read_write_snapshot(core, tmp_path):
    """tests/test_order_book.py:190-210"""
    ob = core.OrderBook(0, 1)
    ob.place_order(True, 10, 11, price=50)
    ob.place_order(False, 20, 12, price=60)
    ob.place_order(True, 10, 11, price=55)
    ob.place_order(False, 20, 12, price=65)
    path = str(tmp_path / "foo.json")
    ob.save_json_snapshot(path)
    loaded = core.order_book_from_json(path)
    assert ob.bid_ask() == loaded.bid_ask()
    assert ob.best_ask_vol_and_orders() == loaded.best_ask_vol_and_orders()
    assert ob.best_bid_vol_and_orders() == loaded.best_bid_vol_and_orders()
    assert ob.get_orders() == loaded.get_orders() and ob.get_trades() == loaded.get_trades()
    ob.save_json_snapshot(str(tmp_path / "pretty.json"), pretty=True)
    assert json.load(open(tmp_path / "pretty.json")) == json.load(open(path))
    with pytest.raises(OSError):
        ob.save_json_snapshot(str(tmp_path / "no_such_dir" / "x.json"))


@pytest.mark.gpu
def test_snapshot_midstream_continues_identically(core, oracle, tmp_path):
    """Save after half of a C2 stream, reload, apply the second half to both books: identical to never saving."""
    s = workloads.replay_stream(6000, 21, tick_size=1, time_mode="strict", half_width=24)
    a = core.OrderBook(0, 1)
    a.replay(s[:3000])
    path = str(tmp_path / "mid.json")
    a.save_json_snapshot(path)
    b = core.order_book_from_json(path)
    assert a.get_orders() == b.get_orders() and a.get_trades() == b.get_trades()
    assert np.array_equal(a.level_2_data(), b.level_2_data()) and a._l1() == b._l1()
    oa, ob_ = a.replay(s[3000:]), b.replay(s[3000:])
    assert np.array_equal(oa, ob_)
    assert a.get_orders() == b.get_orders() and a.get_trades() == b.get_trades()
    ref = oracle.OrderBook(0, 1)
    ref.replay(s)
    assert b.get_trades() == ref.get_trades() and b.get_orders() == ref.get_orders()
    assert np.array_equal(b.level_2_data(), ref.level_2_data())


@pytest.mark.gpu
def test_snapshot_with_key_collisions(core, tmp_path):
    """Equal-(price,time) keys (N1): like the reference's TryFrom (orderbook.rs:898-905) the load re-inserts every
    Active order, so totals are preserved exactly while a former ghost may become reachable again."""
    s = workloads.replay_stream(3000, 5, tick_size=1, time_mode="flat", half_width=10)
    a = core.OrderBook(0, 1)
    a.replay(s)
    path = str(tmp_path / "flat.json")
    a.save_json_snapshot(path)
    b = core.order_book_from_json(path)
    assert a.get_orders() == b.get_orders() and a.get_trades() == b.get_trades()
    la, lb = a.level_2_data(), b.level_2_data()
    assert la[0] == lb[0] and la[3] == lb[3] and la[4] == lb[4]          # trade_vol and both side totals
    act = [o for o in b.get_orders() if o[1] == 1]
    assert lb[4] == sum(o[4] for o in act if o[0]) and lb[3] == sum(o[4] for o in act if not o[0])
    b.save_json_snapshot(str(tmp_path / "again.json"))
    assert json.load(open(tmp_path / "again.json")) == json.load(open(path))
