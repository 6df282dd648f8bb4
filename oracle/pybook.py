"""ORACLE — TEST INFRASTRUCTURE ONLY.

A second, independently written restatement of the reference order book in pure Python, used only
to cross-check the C++ oracle on random instruction streams (the Rust reference
cannot be built here, so two independently written restatements agreeing on many events is the
strongest offline evidence available).  Deliberately naive: the price-time queue is a plain dict
keyed ``(side_price, t)`` scanned with ``min`` — O(n) per operation, so small cases only.

Follows crates/order_book/src/side.rs:54-143 and crates/order_book/src/orderbook.rs:356-870.
"""
from __future__ import annotations

U32 = 0xFFFFFFFF
U64 = 0xFFFFFFFFFFFFFFFF
NEW, ACTIVE, FILLED, CANCELLED, REJECTED = range(5)


class PySide:
    def __init__(self, is_bid: bool):
        self.is_bid = is_bid
        self.total = 0
        self.levels = {}   # side_price -> [vol, count]
        self.queue = {}    # (side_price, t) -> order id

    def sp(self, price):  # side.rs:300-313
        return (U32 - price) if self.is_bid else price

    def insert(self, sprice, t, oid, vol):
        self.queue[(sprice, t)] = oid
        if sprice in self.levels:
            self.levels[sprice][0] = (self.levels[sprice][0] + vol) & U32
            self.levels[sprice][1] += 1
        else:
            self.levels[sprice] = [vol, 1]
        self.total = (self.total + vol) & U32

    def remove(self, sprice, t, vol):
        self.queue.pop((sprice, t), None)
        lv = self.levels[sprice]
        lv[0] = (lv[0] - vol) & U32
        lv[1] -= 1
        if lv[1] == 0:
            del self.levels[sprice]
        self.total = (self.total - vol) & U32

    def remove_vol(self, sprice, vol):
        self.levels[sprice][0] = (self.levels[sprice][0] - vol) & U32
        self.total = (self.total - vol) & U32

    def best_key(self):
        return min(self.queue) if self.queue else None

    def best_price(self):
        k = self.best_key()
        raw = k[0] if k is not None else U32
        return (U32 - raw) if self.is_bid else raw

    def best_level(self):
        if not self.levels:
            return (0, 0)
        v = self.levels[min(self.levels)]
        return (v[0], v[1])

    def at(self, price):
        v = self.levels.get(self.sp(price & U32))
        return (v[0], v[1]) if v else (0, 0)


class PyBook:
    def __init__(self, start_time, tick_size, trading=True):
        self.t, self.tick, self.trading = start_time, tick_size, trading
        self.trade_vol = 0
        self.sides = {True: PySide(True), False: PySide(False)}
        self.orders = []   # dicts
        self.trades = []   # tuples (t, passive_is_bid, price, vol, active, passive)

    # ----------------------------------------------------------------- queries
    def bid_ask(self):
        return (self.sides[True].best_price(), self.sides[False].best_price())

    def l2(self):
        bid, ask = self.bid_ask()
        out = [self.trade_vol, bid, ask, self.sides[False].total, self.sides[True].total]
        for i in range(10):
            out += list(self.sides[True].at(bid - i * self.tick)) + list(self.sides[False].at(ask + i * self.tick))
        return out

    def l1(self):
        bid, ask = self.bid_ask()
        b, a = self.sides[True].best_level(), self.sides[False].best_level()
        return [bid, ask, self.sides[True].total, self.sides[False].total, b[0], a[0], b[1], a[1]]

    # ----------------------------------------------------------------- mutations
    def place(self, bid, vol, trader, price=None):
        if price is not None and price % self.tick != 0:
            raise ValueError(price)
        if price is None:
            price = U32 if bid else 0
        oid = len(self.orders)
        o = dict(bid=bid, status=ACTIVE, arr=self.t, end=U64, vol=vol, start_vol=vol, price=price, trader=trader,
                 kt=0, id=oid)
        self.orders.append(o)
        is_market = (price == U32) if bid else (price == 0)
        if is_market:
            if self.trading:
                self._match(o)
                if o["status"] != FILLED:
                    o["status"], o["end"] = CANCELLED, self.t
            else:
                o["status"], o["end"] = REJECTED, self.t
        else:
            if self.trading:
                self._match(o)
            if o["status"] != FILLED:
                self._rest(o)
        return oid

    def _rest(self, o):
        o["kt"] = self.t
        side = self.sides[o["bid"]]
        side.insert(side.sp(o["price"]), self.t, o["id"], o["vol"])

    def _match(self, o):
        opp = self.sides[not o["bid"]]
        while o["vol"] > 0:
            best = opp.best_price()
            if o["bid"] and not (o["price"] >= best):
                break
            if (not o["bid"]) and not (o["price"] <= best):
                break
            k = opp.best_key()
            if k is None:
                break
            p = self.orders[opp.queue[k]]
            tv = min(o["vol"], p["vol"])
            o["vol"] -= tv
            p["vol"] -= tv
            self.trades.append((self.t, p["bid"], p["price"], tv, o["id"], p["id"]))
            if p["vol"] == 0:
                p["end"], p["status"] = self.t, FILLED
            if o["vol"] == 0:
                o["end"], o["status"] = self.t, FILLED
            self.trade_vol = (self.trade_vol + tv) & U32
            psp = opp.sp(p["price"])
            if p["status"] == FILLED:
                opp.remove(psp, p["kt"], tv)
            else:
                opp.remove_vol(psp, tv)

    def cancel(self, oid):
        o = self.orders[oid]
        if o["status"] == ACTIVE:
            o["status"], o["end"] = CANCELLED, self.t
            side = self.sides[o["bid"]]
            side.remove(side.sp(o["price"]), o["kt"], o["vol"])

    def modify(self, oid, new_price=None, new_vol=None):
        o = self.orders[oid]
        if o["status"] != ACTIVE or (new_price is None and new_vol is None):
            return
        side = self.sides[o["bid"]]
        if new_price is None and new_vol < o["vol"]:
            d = o["vol"] - new_vol
            o["vol"] = new_vol
            side.remove_vol(side.sp(o["price"]), d)
            return
        side.remove(side.sp(o["price"]), o["kt"], o["vol"])
        o["vol"] = o["vol"] if new_vol is None else new_vol
        o["price"] = o["price"] if new_price is None else new_price
        if self.trading:
            self._match(o)
        if o["status"] != FILLED:
            self._rest(o)

    def order_tuples(self):
        return [(o["bid"], o["status"], o["arr"], o["end"], o["vol"], o["start_vol"], o["price"], o["trader"], o["id"])
                for o in self.orders]
