"""JSON snapshot of an order book, schema-compatible with the reference's serde output.

Reference: `OrderBook::save_json / load_json` (crates/order_book/src/orderbook.rs:804-832), the serialised
fields (`:93-121`: t, tick_size, trade_vol, orders, trades, trading — the two sides are skipped and rebuilt
from the Active orders on load, `:891-918`), `OrderEntry{order, key}` (`:36-44`), `Order` / `Trade` / `Side` /
`Status` (crates/order_book/src/types.rs:26-118), Python entry points rust/src/order_book.rs:377-398.

serde_json conventions followed: structs are objects with fields in declaration order, unit enum variants
are strings ("Bid", "Active", ...), the key tuple `(Side, u32, u64)` is a 3-element array whose price
component is `u32::MAX - price` for bids (side.rs:300-313); compact output has no whitespace, pretty
output uses two-space indentation.  Keys of orders that are not Active are reconstructed from their
current price (they play no role on load).
"""
from __future__ import annotations

import json
import typing

import numpy as np

SIDE = {True: "Bid", False: "Ask"}
STATUS = ["New", "Active", "Filled", "Cancelled", "Rejected"]
U32_MAX = 2**32 - 1


def book_to_dict(t: int, tick_size: int, trade_vol: int, trading: bool, orders: typing.Sequence[tuple],
                 key_times: typing.Sequence[int], trades: typing.Sequence[tuple]) -> dict:
    """orders: PyOrder tuples (rust/src/types.rs:19-31); trades: PyTrade tuples (rust/src/types.rs:4-17)."""
    o_list = []
    for o, kt in zip(orders, key_times):
        bid, status, arr, end, vol, start_vol, price, trader, oid = o
        o_list.append({
            "order": {"side": SIDE[bool(bid)], "status": STATUS[status], "arr_time": int(arr), "end_time": int(end),
                      "vol": int(vol), "start_vol": int(start_vol), "price": int(price), "trader_id": int(trader),
                      "order_id": int(oid)},
            "key": [SIDE[bool(bid)], int(U32_MAX - price if bid else price), int(kt)],
        })
    t_list = [{"t": int(x[0]), "side": SIDE[bool(x[1])], "price": int(x[2]), "vol": int(x[3]),
               "active_order_id": int(x[4]), "passive_order_id": int(x[5])} for x in trades]
    return {"t": int(t), "tick_size": int(tick_size), "trade_vol": int(trade_vol), "orders": o_list, "trades": t_list,
            "trading": bool(trading)}


def dumps(d: dict, pretty: bool = False) -> str:
    return json.dumps(d, indent=2) if pretty else json.dumps(d, separators=(",", ":"))


def dict_to_columns(d: dict):
    """Inverse of book_to_dict: numpy columns in the bb_load_book layout."""
    orders, trades = d["orders"], d["trades"]
    n, m = len(orders), len(trades)
    side = np.array([o["order"]["side"] == "Bid" for o in orders], dtype=np.uint8).reshape(n)
    status = np.array([STATUS.index(o["order"]["status"]) for o in orders], dtype=np.uint8).reshape(n)
    col = lambda k, dt: np.array([o["order"][k] for o in orders], dtype=dt).reshape(n)  # noqa: E731
    key_time = np.array([o["key"][2] for o in orders], dtype=np.uint64).reshape(n)
    # The reference's TryFrom (orderbook.rs:898-905) re-inserts every ACTIVE order under its stored key, whatever that key says;
    # keys of orders that are no longer on the book are carried along unread.  A key that disagrees with its Active order
    # (other side, other price) would put the order where this engine cannot represent it: refused.  Stale keys of ended
    # orders (e.g. filled during a replace: the key keeps the price from before the modify) are accepted as the reference does.
    for o in orders:
        bid = o["order"]["side"] == "Bid"
        if o["order"]["status"] == "Active" and (o["key"][0] != o["order"]["side"] or
                                                  o["key"][1] != (U32_MAX - o["order"]["price"] if bid else o["order"]["price"])):
            raise ValueError("Failed to convert OrderBookState to an OrderBook")
    tcol = lambda k, dt: np.array([x[k] for x in trades], dtype=dt).reshape(m)  # noqa: E731
    return dict(
        side=side, status=status, arr_time=col("arr_time", np.uint64), end_time=col("end_time", np.uint64),
        vol=col("vol", np.uint32), start_vol=col("start_vol", np.uint32), price=col("price", np.uint32),
        trader=col("trader_id", np.uint32), key_time=key_time,
        tr_t=tcol("t", np.uint64), tr_side=np.array([x["side"] == "Bid" for x in trades], dtype=np.uint8).reshape(m),
        tr_price=tcol("price", np.uint32), tr_vol=tcol("vol", np.uint32), tr_active=tcol("active_order_id", np.uint64),
        tr_passive=tcol("passive_order_id", np.uint64))


def save_json(ob, path: str, pretty: bool = False) -> None:
    """`OrderBook.save_json_snapshot` (rust/src/order_book.rs:377-388)."""
    env = ob._env
    d = book_to_dict(env.time(0), env.tick_size, int(env.book_level_2(0)[0]), ob._trading, env.get_orders(0),
                     env.order_keys(0), env.get_trades(0))
    with open(path, "w") as f:   # OSError propagates like the reference's io::Error
        f.write(dumps(d, pretty))


def load_json(path: str, **kw):
    """`order_book_from_json` (rust/src/order_book.rs:390-398).  Keyword arguments (capacities, engine selection such
    as `price_window`) are forwarded to the OrderBook constructor."""
    from .core import OrderBook

    with open(path) as f:
        d = json.load(f)
    c = dict_to_columns(d)
    n, m = len(c["side"]), len(c["tr_t"])
    kw["max_orders"] = max(kw.get("max_orders", 1 << 18), 2 * n)
    kw["max_trades"] = max(kw.get("max_trades", 1 << 18), 2 * m)
    ob = OrderBook(d["t"], d["tick_size"], d["trading"], **kw)
    ob._env.load_book(0, d["t"], d["trade_vol"], d["trading"], c)
    ob._t = d["t"]
    return ob


order_book_from_json = load_json
