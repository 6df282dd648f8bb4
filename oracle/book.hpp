// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/rng.hpp header).  PARITY PINNED for everything in
// this file by the reference's own known-answer tests (tests/test_oracle_golden.py restates them):
// crates/order_book/src/side.rs:320-469, crates/order_book/src/orderbook.rs:925-1274,
// tests/test_order_book.py.
//
// CPU restatement of the reference limit order book.  Follows, function by function:
//   crates/order_book/src/types.rs:5-285      value types, market-order sentinels
//   crates/order_book/src/side.rs:36-313      OrderBookSide / BidSide / AskSide
//   crates/order_book/src/orderbook.rs:144-870 OrderBook (create/place/cancel/modify/match, L1/L2)
// std::map stands where the reference uses BTreeMap so that ordering, first_key_value and the
// insert-overwrites-on-equal-key behaviour (SURVEY.md note N1) hold by construction.
// Arithmetic is u32/u64 wrapping, i.e. the behaviour of the reference built in release mode.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <utility>
#include <vector>

namespace oracle {

typedef uint64_t OrderId;  // types.rs:6  (usize)
typedef uint64_t Nanos;    // types.rs:10
typedef uint32_t Price;    // types.rs:12
typedef uint32_t Vol;      // types.rs:14
typedef uint32_t TraderId; // types.rs:16
typedef uint32_t OrderCount;

static const Price PRICE_MAX = 0xFFFFFFFFu;
static const Nanos NANOS_MAX = ~0ULL;
static const int LEVELS = 10;  // default const generic of OrderBook / Env; hard-wired in the PyO3 layer

enum Side : uint8_t { BID = 0, ASK = 1 };  // types.rs:26-47 (bool true <-> Bid)
enum Status : uint8_t { NEW = 0, ACTIVE = 1, FILLED = 2, CANCELLED = 3, REJECTED = 4 };  // types.rs:51-75

struct OrderKey {  // types.rs:8  (Side, u32, u64)
    Side side;
    uint32_t price;  // ask: price; bid: PRICE_MAX - price   (side.rs:300-313)
    Nanos t;
};

struct Order {  // types.rs:79-101
    Side side;
    Status status;
    Nanos arr_time;
    Nanos end_time;
    Vol vol;
    Vol start_vol;
    Price price;
    TraderId trader_id;
    OrderId order_id;
};

struct Trade {  // types.rs:105-118
    Nanos t;
    Side side;
    Price price;
    Vol vol;
    OrderId active_order_id;
    OrderId passive_order_id;
};

struct OrderEntry {  // orderbook.rs:36-44
    Order order;
    OrderKey key;
};

struct PriceError : public std::runtime_error {  // orderbook.rs:127-142
    Price price, tick_size;
    PriceError(Price p, Price t)
        : std::runtime_error("Price " + std::to_string(p) + " was not a multiple of tick-size " +
                             std::to_string(t)),
          price(p), tick_size(t) {}
};

struct Level2Data {  // types.rs:272-285
    Price bid_price, ask_price;
    Vol bid_vol, ask_vol;
    std::array<std::pair<Vol, OrderCount>, LEVELS> bid_price_levels, ask_price_levels;
};

struct Level1Data {  // types.rs:252-269
    Price bid_price, ask_price;
    Vol bid_vol, ask_vol, bid_touch_vol, ask_touch_vol;
    OrderCount bid_touch_orders, ask_touch_orders;
};

static inline OrderKey get_bid_key(Nanos t, Price price) { return OrderKey{BID, PRICE_MAX - price, t}; }  // side.rs:300
static inline OrderKey get_ask_key(Nanos t, Price price) { return OrderKey{ASK, price, t}; }              // side.rs:311

// side.rs:36-144.  Both sides are stored min-first; the bid wrapper flips prices.
struct OrderBookSide {
    Vol vol_ = 0;
    std::map<Price, std::pair<Vol, OrderCount>> volumes;
    std::map<std::pair<Price, Nanos>, OrderId> orders;

    void insert_order(const OrderKey& key, OrderId idx, Vol vol) {  // side.rs:54-66
        orders[std::make_pair(key.price, key.t)] = idx;               // BTreeMap::insert overwrites
        auto it = volumes.find(key.price);
        if (it != volumes.end()) {
            it->second.first += vol;
            it->second.second += 1;
        } else {
            volumes.emplace(key.price, std::make_pair(vol, (OrderCount)1));
        }
        vol_ += vol;
    }
    void remove_order(const OrderKey& key, Vol vol) {  // side.rs:75-84
        orders.erase(std::make_pair(key.price, key.t));
        auto it = volumes.find(key.price);
        if (it == volumes.end()) throw std::logic_error("remove_order: no level (reference would panic)");
        it->second.first -= vol;
        it->second.second -= 1;
        if (it->second.second == 0) volumes.erase(it);
        vol_ -= vol;
    }
    void remove_vol(Price price, Vol vol) {  // side.rs:93-96
        auto it = volumes.find(price);
        if (it == volumes.end()) throw std::logic_error("remove_vol: no level (reference would panic)");
        it->second.first -= vol;
        vol_ -= vol;
    }
    Price best_price() const { return orders.empty() ? PRICE_MAX : orders.begin()->first.first; }  // side.rs:99-104
    std::pair<Vol, OrderCount> best_vol_and_orders() const {                                        // side.rs:107-112
        return volumes.empty() ? std::make_pair((Vol)0, (OrderCount)0) : volumes.begin()->second;
    }
    Vol best_vol() const { return best_vol_and_orders().first; }  // side.rs:115-120
    Vol vol() const { return vol_; }
    bool best_order_idx(OrderId* out) const {  // side.rs:128-130
        if (orders.empty()) return false;
        *out = orders.begin()->second;
        return true;
    }
    std::pair<Vol, OrderCount> vol_and_orders_at_price(Price price) const {  // side.rs:138-143
        auto it = volumes.find(price);
        return it == volumes.end() ? std::make_pair((Vol)0, (OrderCount)0) : it->second;
    }
};

struct BidSide : OrderBookSide {  // side.rs:154-222
    Price best_price() const { return PRICE_MAX - OrderBookSide::best_price(); }
    std::pair<Vol, OrderCount> vol_and_orders_at_price(Price price) const {
        return OrderBookSide::vol_and_orders_at_price(PRICE_MAX - price);
    }
};
struct AskSide : OrderBookSide {};  // side.rs:224-291

class OrderBook {  // orderbook.rs:95-121
public:
    Nanos t;
    Price tick_size;
    Vol trade_vol;
    AskSide ask_side;
    BidSide bid_side;
    std::vector<OrderEntry> orders;
    std::vector<Trade> trades;
    bool trading;

    OrderBook(Nanos start_time, Price tick, bool trading_)  // orderbook.rs:158-171
        : t(start_time), tick_size(tick), trade_vol(0), trading(trading_) {
        if (tick == 0) throw std::invalid_argument("tick_size must be > 0");
    }

    Nanos get_time() const { return t; }
    void set_time(Nanos x) { t = x; }
    void enable_trading() { trading = true; }
    void disable_trading() { trading = false; }
    Vol get_trade_vol() const { return trade_vol; }
    void reset_trade_vol() { trade_vol = 0; }

    Vol ask_vol() const { return ask_side.vol(); }
    Vol bid_vol() const { return bid_side.vol(); }
    std::pair<Vol, OrderCount> ask_best_vol_and_orders() const { return ask_side.best_vol_and_orders(); }
    std::pair<Vol, OrderCount> bid_best_vol_and_orders() const { return bid_side.best_vol_and_orders(); }
    std::pair<Price, Price> bid_ask() const { return {bid_side.best_price(), ask_side.best_price()}; }  // :267

    double mid_price() const {  // orderbook.rs:272-276 (u32 subtraction wraps in release builds)
        auto ba = bid_ask();
        const uint32_t spread = ba.second - ba.first;
        return (double)ba.first + 0.5 * (double)spread;
    }

    // orderbook.rs:229-236 / 257-264: levels at FIXED tick offsets from the touch, with wrapping
    std::array<std::pair<Vol, OrderCount>, LEVELS> ask_levels() const {
        std::array<std::pair<Vol, OrderCount>, LEVELS> out;
        const Price start = bid_ask().second;
        for (int i = 0; i < LEVELS; ++i)
            out[i] = ask_side.vol_and_orders_at_price(start + (Price)i * tick_size);
        return out;
    }
    std::array<std::pair<Vol, OrderCount>, LEVELS> bid_levels() const {
        std::array<std::pair<Vol, OrderCount>, LEVELS> out;
        const Price start = bid_ask().first;
        for (int i = 0; i < LEVELS; ++i)
            out[i] = bid_side.vol_and_orders_at_price(start - (Price)i * tick_size);
        return out;
    }
    Level1Data level_1_data() const {  // orderbook.rs:287-301
        auto ba = bid_ask();
        auto b = bid_best_vol_and_orders();
        auto a = ask_best_vol_and_orders();
        return Level1Data{ba.first, ba.second, bid_vol(), ask_vol(), b.first, a.first, b.second, a.second};
    }
    Level2Data level_2_data() const {  // orderbook.rs:314-324
        auto ba = bid_ask();
        return Level2Data{ba.first, ba.second, bid_vol(), ask_vol(), bid_levels(), ask_levels()};
    }

    const Order& order(OrderId id) const { return orders.at(id).order; }  // orderbook.rs:337-339 (index panic)

    // orderbook.rs:356-396.  price_is_some=false => market order (types.rs:160-172, 213-225)
    OrderId create_order(Side side, Vol vol, TraderId trader_id, bool price_is_some, Price price) {
        const OrderId order_id = orders.size();
        Order o;
        o.side = side;
        o.status = NEW;
        o.arr_time = t;
        o.end_time = NANOS_MAX;
        o.vol = vol;
        o.start_vol = vol;
        o.trader_id = trader_id;
        o.order_id = order_id;
        if (price_is_some) {
            if (price % tick_size != 0) throw PriceError(price, tick_size);
            o.price = price;
        } else {
            o.price = (side == BID) ? PRICE_MAX : 0;
        }
        const OrderKey key = (side == BID) ? get_bid_key(0, o.price) : get_ask_key(0, o.price);
        orders.push_back(OrderEntry{o, key});
        return order_id;
    }

    OrderId create_and_place_order(Side side, Vol vol, TraderId trader_id, bool price_is_some, Price price) {
        const OrderId id = create_order(side, vol, trader_id, price_is_some, price);  // orderbook.rs:411-421
        place_order(id);
        return id;
    }

    void place_order(OrderId order_id) {  // orderbook.rs:583-611
        OrderEntry e = orders.at(order_id);
        if (e.order.status != NEW) return;
        e.order.status = ACTIVE;
        e.order.arr_time = t;
        if (e.order.side == BID) {
            if (e.order.price == PRICE_MAX) place_market(e); else place_limit(e);
        } else {
            if (e.order.price == 0) place_market(e); else place_limit(e);
        }
        orders[order_id] = e;
    }

    void cancel_order(OrderId order_id) {  // orderbook.rs:622-644
        if (order_id >= orders.size()) throw std::out_of_range("No order with id exists");
        OrderEntry& e = orders[order_id];
        if (e.order.status == ACTIVE) {
            e.order.status = CANCELLED;
            e.order.end_time = t;
            side_of(e.key.side).remove_order(e.key, e.order.vol);
        }
    }

    // orderbook.rs:743-772
    void modify_order(OrderId order_id, bool has_price, Price new_price, bool has_vol, Vol new_vol) {
        OrderEntry e = orders.at(order_id);
        if (e.order.status == ACTIVE) {
            if (!has_price && !has_vol) {
            } else if (!has_price && has_vol) {
                if (new_vol < e.order.vol) {
                    reduce_order_vol(e, e.order.vol - new_vol);
                } else {
                    replace_order(e, e.order.price, new_vol);
                }
            } else if (has_price && !has_vol) {
                replace_order(e, new_price, e.order.vol);
            } else {
                replace_order(e, new_price, new_vol);
            }
        }
        orders[order_id] = e;
    }

private:
    OrderBookSide& side_of(Side s) { return s == BID ? (OrderBookSide&)bid_side : (OrderBookSide&)ask_side; }

    // orderbook.rs:843-870
    Vol match_orders(Order& agg, Order& pass) {
        const Vol tv = std::min(agg.vol, pass.vol);
        agg.vol -= tv;
        pass.vol -= tv;
        trades.push_back(Trade{t, pass.side, pass.price, tv, agg.order_id, pass.order_id});
        if (pass.vol == 0) { pass.end_time = t; pass.status = FILLED; }
        if (agg.vol == 0) { agg.end_time = t; agg.status = FILLED; }
        return tv;
    }

    // orderbook.rs:429-454 (match_bid) and :462-487 (match_ask), folded over the opposite side
    void match(OrderEntry& e, Side agg_side) {
        for (;;) {
            if (!(e.order.vol > 0)) break;
            if (agg_side == BID) { if (!(e.order.price >= ask_side.best_price())) break; }
            else                 { if (!(e.order.price <= bid_side.best_price())) break; }
            OrderBookSide& opp = (agg_side == BID) ? (OrderBookSide&)ask_side : (OrderBookSide&)bid_side;
            OrderId id;
            if (!opp.best_order_idx(&id)) break;
            OrderEntry& m = orders.at(id);
            const Vol tv = match_orders(e.order, m.order);
            trade_vol += tv;
            if (m.order.status == FILLED) opp.remove_order(m.key, tv);
            else opp.remove_vol(m.key.price, tv);
        }
    }

    void place_limit(OrderEntry& e) {  // orderbook.rs:495-505 / 538-548
        if (trading) match(e, e.order.side);
        if (e.order.status != FILLED) {
            e.key = OrderKey{e.order.side, e.key.price, t};
            side_of(e.order.side).insert_order(e.key, e.order.order_id, e.order.vol);
        }
    }
    void place_market(OrderEntry& e) {  // orderbook.rs:517-531 / 560-574
        if (trading) {
            match(e, e.order.side);
            if (e.order.status != FILLED) { e.order.status = CANCELLED; e.order.end_time = t; }
        } else {
            e.order.status = REJECTED;
            e.order.end_time = t;
        }
    }
    void reduce_order_vol(OrderEntry& e, Vol reduce_vol) {  // orderbook.rs:656-667
        e.order.vol -= reduce_vol;
        side_of(e.key.side).remove_vol(e.key.price, reduce_vol);
    }
    void replace_order(OrderEntry& e, Price new_price, Vol new_vol) {  // orderbook.rs:679-723
        side_of(e.key.side).remove_order(e.key, e.order.vol);
        e.order.vol = new_vol;
        e.order.price = new_price;
        if (trading) match(e, e.key.side);
        if (e.order.status != FILLED) {
            e.key = (e.key.side == BID) ? get_bid_key(t, new_price) : get_ask_key(t, new_price);
            side_of(e.key.side).insert_order(e.key, e.order.order_id, e.order.vol);
        }
    }
};

}  // namespace oracle
