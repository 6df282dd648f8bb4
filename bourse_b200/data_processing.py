"""Pandas views of the order and trade tuples (the reference's `bourse.data_processing`, src/bourse/data_processing.py:1-105):
same function names, column names and value mappings, so analysis code written against the reference keeps working on what
`OrderBook.get_trades / get_orders` and `StepEnv.get_trades / get_orders` return here."""
from __future__ import annotations

import typing


def trades_to_dataframe(trades: typing.List[typing.Tuple]):
    """Columns `time, side, price, vol, active_id, passive_id`; `side` mapped to "bid" / "ask" (the passive order's side)."""
    import pandas as pd

    df = pd.DataFrame.from_records(trades, columns=["time", "side", "price", "vol", "active_id", "passive_id"])
    df["side"] = df["side"].map({True: "bid", False: "ask"})
    return df


def orders_to_dataframe(order_history: typing.List[typing.Tuple]):
    """Columns as in the reference (including its `arr time` spelling); `side` and `status` mapped to their names."""
    import pandas as pd

    columns = ["side", "status", "arr time", "end_time", "vol", "start_vol", "price", "trader_id", "order_id"]
    df = pd.DataFrame.from_records(order_history, columns=columns)
    df["side"] = df["side"].map({True: "bid", False: "ask"})
    df["status"] = df["status"].map({0: "new", 1: "active", 2: "filled", 3: "cancelled", 4: "rejected"})
    return df
