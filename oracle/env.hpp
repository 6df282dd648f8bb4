// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/rng.hpp header).
//
// CPU restatement of the discrete-event environment and the built-in agents:
//   crates/step_sim/src/env.rs:84-295              Env::{new, step, place/cancel/modify_order, getters}
//   crates/step_sim/src/data.rs:9-57               Level2DataRecords
//   crates/step_sim/src/runner.rs:46-69            sim_runner
//   crates/step_sim/src/agents/random_agent.rs:48-120   RandomAgents
//   crates/step_sim/src/agents/momentum_agent.rs:16-209 MomentumAgent
//   crates/step_sim/src/agents/common.rs:21-141    round_price_up/down, cancel_live_orders, limit helpers
// PARITY PINNED (book/env results): crates/step_sim/src/env.rs:311-368, tests/test_step_sim/*.py,
// common.rs:268-305 — restated in tests/test_oracle_golden.py.
// PARITY UNPINNED (shuffle order, agent draw sequences): they go through rand/rand_xoshiro/
// rand_distr, see oracle/rng.hpp.  The `Stream` RNG policy below follows the reference's
// one-generator-threaded-through-everything structure; the `Keyed` policy is the new framework's
// Philox contract (DESIGN.md "RNG contract") and is what the CUDA path is checked against.
#pragma once
#include "book.hpp"
#include "rng.hpp"

namespace oracle {

enum EventKind : uint8_t { EV_NEW = 0, EV_CANCEL = 1, EV_MODIFY = 2 };
struct Event {  // types.rs:229-249
    EventKind kind;
    OrderId order_id;
    bool has_price, has_vol;
    Price new_price;
    Vol new_vol;
};

struct Level2DataRecords {  // data.rs:9-57 — one column per field, appended once per step
    std::vector<Price> bid_price, ask_price;
    std::vector<Vol> bid_vol, ask_vol;
    std::array<std::vector<Vol>, LEVELS> bid_vol_at, ask_vol_at;
    std::array<std::vector<OrderCount>, LEVELS> bid_n_at, ask_n_at;
    void append_record(const Level2Data& r) {
        bid_price.push_back(r.bid_price);
        ask_price.push_back(r.ask_price);
        bid_vol.push_back(r.bid_vol);
        ask_vol.push_back(r.ask_vol);
        for (int i = 0; i < LEVELS; ++i) {
            bid_vol_at[i].push_back(r.bid_price_levels[i].first);
            bid_n_at[i].push_back(r.bid_price_levels[i].second);
            ask_vol_at[i].push_back(r.ask_price_levels[i].first);
            ask_n_at[i].push_back(r.ask_price_levels[i].second);
        }
    }
};

class Env {  // env.rs:58-71
public:
    Nanos step_size;
    OrderBook book;
    std::vector<Vol> trade_vols;
    std::vector<Event> transactions;
    Level2Data level_2_data;
    Level2DataRecords records;
    bool keep_records = true;  // benchmarking knob only (not in the reference)

    Env(Nanos start_time, Price tick_size, Nanos step_size_, bool trading)  // env.rs:84-95
        : step_size(step_size_), book(start_time, tick_size, trading) {
        level_2_data = book.level_2_data();
    }

    // env.rs:116-135 with the shuffle supplied by the caller's RNG policy
    template <class ShuffleFn> void step_with(ShuffleFn&& shuffle_fn) {
        const Nanos start_time = book.get_time();
        book.reset_trade_vol();
        std::vector<Event> tx;
        tx.swap(transactions);
        shuffle_fn(tx);
        for (size_t i = 0; i < tx.size(); ++i) {
            book.set_time(start_time + (Nanos)i);
            process_event(tx[i]);
        }
        book.set_time(start_time + step_size);
        level_2_data = book.level_2_data();
        if (keep_records) {
            records.append_record(level_2_data);
            trade_vols.push_back(book.get_trade_vol());
        }
    }
    void step(Xoroshiro128StarStar& rng) {
        step_with([&](std::vector<Event>& tx) { shuffle(rng, tx); });
    }

    void process_event(const Event& ev) {  // orderbook.rs:782-792
        switch (ev.kind) {
            case EV_NEW: book.place_order(ev.order_id); break;
            case EV_CANCEL: book.cancel_order(ev.order_id); break;
            case EV_MODIFY: book.modify_order(ev.order_id, ev.has_price, ev.new_price, ev.has_vol, ev.new_vol); break;
        }
    }

    OrderId place_order(Side side, Vol vol, TraderId trader, bool price_is_some, Price price) {  // env.rs:166-176
        const OrderId id = book.create_order(side, vol, trader, price_is_some, price);
        transactions.push_back(Event{EV_NEW, id, false, false, 0, 0});
        return id;
    }
    void cancel_order(OrderId id) { transactions.push_back(Event{EV_CANCEL, id, false, false, 0, 0}); }  // :189
    void modify_order(OrderId id, bool has_price, Price p, bool has_vol, Vol v) {                         // :208
        transactions.push_back(Event{EV_MODIFY, id, has_price, has_vol, p, v});
    }
    Status order_status(OrderId id) const { return book.order(id).status; }  // env.rs:288-290
};

// crates/order_book/src/market.rs:59-95 (Market: an array of independent order books) and
// crates/step_sim/src/market_env.rs:47-135 (MarketEnv): one transaction queue for all assets, shuffled as a whole;
// event i of the shuffled queue is processed at start_time + i on its asset's book.
// PARITY PINNED by crates/step_sim/src/market_env.rs:333-407 and crates/order_book/src/market.rs:397-574 (restated in
// tests/test_oracle_market.py); the shuffle ORDER is unpinned like Env's (rand / rand_xoshiro, see oracle/rng.hpp).
class MarketEnv {
public:
    struct MarketEvent {  // Event<MarketOrderId>, types.rs:229-249 with order_id = (asset, id)
        uint32_t asset;
        Event ev;
    };
    Nanos step_size;
    std::vector<OrderBook> books;  // Market::order_books
    std::vector<std::vector<Vol>> trade_vols;
    std::vector<MarketEvent> transactions;
    std::vector<Level2Data> level_2_data;
    std::vector<Level2DataRecords> records;

    MarketEnv(Nanos start_time, const std::vector<Price>& tick_sizes, Nanos step_size_, bool trading)  // market_env.rs:73-90
        : step_size(step_size_) {
        for (Price t : tick_sizes) books.emplace_back(start_time, t, trading);
        trade_vols.resize(books.size());
        records.resize(books.size());
        for (auto& b : books) level_2_data.push_back(b.level_2_data());
    }

    void step(Xoroshiro128StarStar& rng) {
        step_with([&](std::vector<MarketEvent>& tx) { shuffle(rng, tx); });
    }
    template <class ShuffleFn> void step_with(ShuffleFn&& shuffle_fn) {  // market_env.rs:108-135
        const Nanos start_time = books[0].get_time();
        for (auto& b : books) b.reset_trade_vol();
        std::vector<MarketEvent> tx;
        tx.swap(transactions);
        shuffle_fn(tx);
        for (size_t i = 0; i < tx.size(); ++i) {
            for (auto& b : books) b.set_time(start_time + (Nanos)i);  // Market::set_time, market.rs:113-117
            OrderBook& b = books[tx[i].asset];
            const Event& ev = tx[i].ev;
            switch (ev.kind) {  // Market::process_event, market.rs:343-353
                case EV_NEW: b.place_order(ev.order_id); break;
                case EV_CANCEL: b.cancel_order(ev.order_id); break;
                case EV_MODIFY: b.modify_order(ev.order_id, ev.has_price, ev.new_price, ev.has_vol, ev.new_vol); break;
            }
        }
        for (auto& b : books) b.set_time(start_time + step_size);
        for (size_t a = 0; a < books.size(); ++a) {
            level_2_data[a] = books[a].level_2_data();
            records[a].append_record(level_2_data[a]);
            trade_vols[a].push_back(books[a].get_trade_vol());
        }
    }

    OrderId place_order(uint32_t asset, Side side, Vol vol, TraderId trader, bool price_is_some, Price price) {  // :163-176
        const OrderId id = books[asset].create_order(side, vol, trader, price_is_some, price);
        transactions.push_back(MarketEvent{asset, Event{EV_NEW, id, false, false, 0, 0}});
        return id;
    }
    void cancel_order(uint32_t asset, OrderId id) { transactions.push_back(MarketEvent{asset, Event{EV_CANCEL, id, false, false, 0, 0}}); }
    void modify_order(uint32_t asset, OrderId id, bool has_price, Price p, bool has_vol, Vol v) {
        transactions.push_back(MarketEvent{asset, Event{EV_MODIFY, id, has_price, has_vol, p, v}});
    }
};

// common.rs:21-41
static inline Price round_price_up(double p, double tick_size) {
    p = std::ceil(p / tick_size) * tick_size;
    p = std::min(std::max(p, 0.0), (double)PRICE_MAX);
    return (Price)p;
}
static inline Price round_price_down(double p, double tick_size) {
    p = std::floor(p / tick_size) * tick_size;
    p = std::min(std::max(p, 0.0), (double)PRICE_MAX);
    return (Price)p;
}

// ------------------------------------------------------------------------------------------------
// Agent groups.  A run is an ordered list of groups, updated in declaration order each step
// (crates/macros/src/lib.rs:57-72), followed by Env::step (runner.rs:63-66).
struct RandomAgentsParams {  // random_agent.rs:48-82
    uint32_t n_agents;
    Price tick_lo, tick_hi;
    Vol vol_lo, vol_hi;
    Price tick_size;
    float activity_rate;
};

struct MomentumParams {  // momentum_agent.rs:16-35 (+ ctor args :118-134)
    TraderId agent_id_start;
    uint32_t n_agents;
    Price tick_size;
    float p_cancel;
    Vol trade_vol;
    double decay, demand, scale, order_ratio, price_dist_mu, price_dist_sigma;
};

static const OrderId NO_ORDER = ~0ULL;

struct RandomAgents {
    RandomAgentsParams p;
    std::vector<OrderId> orders;  // NO_ORDER == None
    explicit RandomAgents(const RandomAgentsParams& p_) : p(p_), orders(p_.n_agents, NO_ORDER) {}

    // random_agent.rs:85-119 with the reference's single shared stream
    template <class E> void update_stream(E& env, Xoroshiro128StarStar& rng) {
        for (uint32_t n = 0; n < p.n_agents; ++n) {
            const float u = gen_f32(rng);
            if (u < p.activity_rate) {
                if (orders[n] != NO_ORDER && env.order_status(orders[n]) == ACTIVE) {
                    env.cancel_order(orders[n]);
                    orders[n] = NO_ORDER;
                } else {
                    const Side side = gen_range_u32(rng, 0, 2) == 0 ? ASK : BID;  // [Ask, Bid].choose
                    const Price tick = gen_range_u32(rng, p.tick_lo, p.tick_hi);
                    const Vol vol = gen_range_u32(rng, p.vol_lo, p.vol_hi);
                    orders[n] = env.place_order(side, vol, n, true, tick * p.tick_size);
                }
            }
        }
    }
    // Same decision logic, draws taken from the Philox block keyed (env, step, agent slot, 0)
    template <class E> void update_keyed(E& env, uint32_t env_id, uint32_t step, uint32_t slot_base, uint32_t k0, uint32_t k1) {
        for (uint32_t n = 0; n < p.n_agents; ++n) {
            const Philox4 r = philox4x32_10(env_id, step, slot_base + n, 0, k0, k1);
            const float u = u32_to_f32_unit(r.v[0]);
            if (u < p.activity_rate) {
                if (orders[n] != NO_ORDER && env.order_status(orders[n]) == ACTIVE) {
                    env.cancel_order(orders[n]);
                    orders[n] = NO_ORDER;
                } else {
                    const Side side = (r.v[1] >> 31) ? BID : ASK;
                    const Price tick = p.tick_lo + mulhi_range(r.v[2], p.tick_hi - p.tick_lo);
                    const Vol vol = p.vol_lo + mulhi_range(r.v[3], p.vol_hi - p.vol_lo);
                    orders[n] = env.place_order(side, vol, n, true, tick * p.tick_size);
                }
            }
        }
    }
};

struct MomentumAgent {
    MomentumParams p;
    std::vector<OrderId> orders;
    bool has_last = false;
    double last_price = 0.0, momentum = 0.0;
    explicit MomentumAgent(const MomentumParams& p_) : p(p_) {}

    // Box-Muller in both policies: rand_distr's ziggurat tables are not reproducible offline.
    static double normal_from(double u1, double u2) {
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925 * u2);
    }

    // momentum_agent.rs:145-209; `draw_*` lambdas abstract where the uniforms come from
    template <class E, class DrawCancel, class DrawTrader>
    void update_impl(E& env, DrawCancel&& draw_cancel, DrawTrader&& draw_trader) {
        // common.rs:56-75 cancel_live_orders
        std::vector<OrderId> live;
        uint32_t k = 0;
        for (OrderId id : orders) {
            if (env.order_status(id) != ACTIVE) { ++k; continue; }
            const float u = draw_cancel(k++);
            if (u > p.p_cancel) live.push_back(id); else env.cancel_order(id);
        }
        const double mid = env.book.mid_price();
        double m = 0.0, p_market = 0.0;
        if (has_last) {
            m = momentum * (1.0 - p.decay) + p.decay * (mid - last_price);
            p_market = p.demand * std::tanh(p.scale * m) / (double)p.n_agents;
        }
        const double p_limit = p.order_ratio * p_market;
        const double tick = (double)p.tick_size;
        for (uint32_t j = 0; j < p.n_agents; ++j) {
            double u_limit, u_market, n1, n2;
            draw_trader(j, &u_limit, &u_market, &n1, &n2);
            const TraderId trader = p.agent_id_start + j;
            if (u_limit < p_limit) {
                if (m > 0.0) {  // common.rs:95-107
                    const double dist = std::fabs(std::exp(p.price_dist_mu + p.price_dist_sigma * normal_from(n1, n2)));
                    live.push_back(env.place_order(BID, p.trade_vol, trader, true, round_price_down(mid - dist, tick)));
                } else if (m < 0.0) {  // common.rs:128-141
                    const double dist = std::fabs(std::exp(p.price_dist_mu + p.price_dist_sigma * normal_from(n1, n2)));
                    live.push_back(env.place_order(ASK, p.trade_vol, trader, true, round_price_up(mid + dist, tick)));
                }
            }
            if (u_market < p_market) {
                if (m > 0.0) env.place_order(BID, p.trade_vol, trader, false, 0);
                else if (m < 0.0) env.place_order(ASK, p.trade_vol, trader, false, 0);
            }
        }
        momentum = m;
        last_price = mid;
        has_last = true;
        orders.swap(live);
    }

    template <class E> void update_stream(E& env, Xoroshiro128StarStar& rng) {
        // NB the reference consumes draws lazily (the normal only when a limit order fires, and
        // through rand_distr's ziggurat); this restatement draws all four uniforms per trader up
        // front.  Only statistical equivalence is claimed for the stream policy (parity unpinned).
        update_impl(
            env, [&](uint32_t) { return gen_f32(rng); },
            [&](uint32_t, double* ul, double* um, double* n1, double* n2) {
                *ul = gen_f64(rng);
                *n1 = gen_f64(rng);
                *n2 = gen_f64(rng);
                *um = gen_f64(rng);
            });
    }
    template <class E>
    void update_keyed(E& env, uint32_t env_id, uint32_t step, uint32_t group, uint32_t slot_base, uint32_t k0, uint32_t k1) {
        update_impl(
            env,
            [&](uint32_t k) {
                const Philox4 r = philox4x32_10(env_id, step, PHILOX_SLOT_CANCEL | group, k >> 2, k0, k1);
                return u32_to_f32_unit(r.v[k & 3]);
            },
            [&](uint32_t j, double* ul, double* um, double* n1, double* n2) {
                const Philox4 a = philox4x32_10(env_id, step, slot_base + j, 0, k0, k1);
                const Philox4 b = philox4x32_10(env_id, step, slot_base + j, 1, k0, k1);
                *ul = u64_to_f64_unit(a.v[0], a.v[1]);
                *um = u64_to_f64_unit(a.v[2], a.v[3]);
                *n1 = u64_to_f64_unit(b.v[0], b.v[1]);
                *n2 = u64_to_f64_unit(b.v[2], b.v[3]);
            });
    }
};

// crates/step_sim/src/agents/noise_agent.rs:14-177
struct NoiseParams {
    TraderId agent_id_start;
    uint32_t n_agents;
    Price tick_size;
    float p_limit, p_market, p_cancel;
    Vol trade_vol;
    double price_dist_mu, price_dist_sigma;
};

struct NoiseAgent {
    NoiseParams p;
    std::vector<OrderId> orders;
    explicit NoiseAgent(const NoiseParams& p_) : p(p_) {}

    // noise_agent.rs:126-177; draws abstracted as in MomentumAgent
    template <class E, class DrawCancel, class DrawTrader>
    void update_impl(E& env, DrawCancel&& draw_cancel, DrawTrader&& draw_trader) {
        std::vector<OrderId> live;
        uint32_t k = 0;
        for (OrderId id : orders) {  // common.rs:56-75
            if (env.order_status(id) != ACTIVE) { ++k; continue; }
            const float u = draw_cancel(k++);
            if (u > p.p_cancel) live.push_back(id); else env.cancel_order(id);
        }
        const double mid = env.book.mid_price();
        const double tick = (double)p.tick_size;
        for (uint32_t j = 0; j < p.n_agents; ++j) {
            float u_limit, u_market;
            bool limit_bid, market_bid;
            double n1, n2;
            draw_trader(j, &u_limit, &limit_bid, &u_market, &market_bid, &n1, &n2);
            const TraderId trader = p.agent_id_start + j;
            if (u_limit < p.p_limit) {
                const double dist = std::fabs(std::exp(p.price_dist_mu + p.price_dist_sigma * MomentumAgent::normal_from(n1, n2)));
                if (limit_bid) live.push_back(env.place_order(BID, p.trade_vol, trader, true, round_price_down(mid - dist, tick)));
                else live.push_back(env.place_order(ASK, p.trade_vol, trader, true, round_price_up(mid + dist, tick)));
            }
            if (u_market < p.p_market) env.place_order(market_bid ? BID : ASK, p.trade_vol, trader, false, 0);
        }
        orders.swap(live);
    }
    template <class E> void update_stream(E& env, Xoroshiro128StarStar& rng) {
        update_impl(
            env, [&](uint32_t) { return gen_f32(rng); },
            [&](uint32_t, float* ul, bool* lb, float* um, bool* mb, double* n1, double* n2) {
                *ul = gen_f32(rng);
                *lb = (rng.next_u64() >> 63) == 0;  // gen_bool(0.5): next_u64 < 2^63
                *n1 = gen_f64(rng);
                *n2 = gen_f64(rng);
                *um = gen_f32(rng);
                *mb = (rng.next_u64() >> 63) == 0;
            });
    }
    template <class E>
    void update_keyed(E& env, uint32_t env_id, uint32_t step, uint32_t group, uint32_t slot_base, uint32_t k0, uint32_t k1) {
        update_impl(
            env,
            [&](uint32_t k) {
                const Philox4 r = philox4x32_10(env_id, step, PHILOX_SLOT_CANCEL | group, k >> 2, k0, k1);
                return u32_to_f32_unit(r.v[k & 3]);
            },
            [&](uint32_t j, float* ul, bool* lb, float* um, bool* mb, double* n1, double* n2) {
                const Philox4 a = philox4x32_10(env_id, step, slot_base + j, 0, k0, k1);
                const Philox4 b = philox4x32_10(env_id, step, slot_base + j, 1, k0, k1);
                *ul = u32_to_f32_unit(a.v[0]);
                *lb = (a.v[1] >> 31) != 0;
                *um = u32_to_f32_unit(a.v[2]);
                *mb = (a.v[3] >> 31) != 0;
                *n1 = u64_to_f64_unit(b.v[0], b.v[1]);
                *n2 = u64_to_f64_unit(b.v[2], b.v[3]);
            });
    }
};

enum GroupKind : uint32_t { GROUP_RANDOM = 0, GROUP_MOMENTUM = 1, GROUP_NOISE = 2 };

struct AgentGroup {
    GroupKind kind;
    RandomAgents* random = nullptr;
    MomentumAgent* momentum = nullptr;
};

// One simulated market: Env + ordered agent groups.  `run_keyed` is what the CUDA path mirrors.
class Sim {
public:
    Env env;
    std::vector<RandomAgents> randoms;
    std::vector<MomentumAgent> momentums;
    std::vector<NoiseAgent> noises;
    std::vector<std::pair<GroupKind, size_t>> order;  // declaration order -> (kind, index)
    uint32_t step_counter = 0;
    uint64_t n_instructions = 0;

    Sim(Nanos start_time, Price tick_size, Nanos step_size, bool trading) : env(start_time, tick_size, step_size, trading) {}
    void add_random(const RandomAgentsParams& p) { order.emplace_back(GROUP_RANDOM, randoms.size()); randoms.emplace_back(p); }
    void add_momentum(const MomentumParams& p) { order.emplace_back(GROUP_MOMENTUM, momentums.size()); momentums.emplace_back(p); }
    void add_noise(const NoiseParams& p) { order.emplace_back(GROUP_NOISE, noises.size()); noises.emplace_back(p); }

    // runner.rs:46-69
    void run_stream(uint64_t seed, uint64_t n_steps) {
        Xoroshiro128StarStar rng = Xoroshiro128StarStar::seed_from_u64(seed);
        for (uint64_t s = 0; s < n_steps; ++s) {
            for (auto& g : order) {
                if (g.first == GROUP_RANDOM) randoms[g.second].update_stream(env, rng);
                else if (g.first == GROUP_MOMENTUM) momentums[g.second].update_stream(env, rng);
                else noises[g.second].update_stream(env, rng);
            }
            n_instructions += env.transactions.size();
            env.step(rng);
            ++step_counter;
        }
    }

    // The two halves of one keyed step, separately callable so that a caller can queue its own instructions between
    // them (agents.update(env) ... env.place_order(..) ... env.step(): the shape of an RL loop with background agents,
    // bb_run_agents_with_rows on the CUDA side).
    void agents_update_keyed(uint64_t seed, uint32_t env_id) {
        const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
        const uint32_t step = step_counter;
        uint32_t slot_base = 0, gi = 0;
        for (auto& g : order) {
            if (g.first == GROUP_RANDOM) {
                randoms[g.second].update_keyed(env, env_id, step, slot_base, k0, k1);
                slot_base += randoms[g.second].p.n_agents;
            } else if (g.first == GROUP_MOMENTUM) {
                momentums[g.second].update_keyed(env, env_id, step, gi, slot_base, k0, k1);
                slot_base += momentums[g.second].p.n_agents;
            } else {
                noises[g.second].update_keyed(env, env_id, step, gi, slot_base, k0, k1);
                slot_base += noises[g.second].p.n_agents;
            }
            ++gi;
        }
    }
    void step_keyed(uint64_t seed, uint32_t env_id) {
        const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
        const uint32_t step = step_counter;
        n_instructions += env.transactions.size();
        env.step_with([&](std::vector<Event>& tx) {
            // Fisher-Yates from the back, one Philox word per position
            for (size_t i = tx.size(); i > 1; --i) {
                const uint32_t idx = (uint32_t)(i - 1);
                const Philox4 r = philox4x32_10(env_id, step, PHILOX_SLOT_SHUFFLE, idx >> 2, k0, k1);
                const uint32_t j = mulhi_range(r.v[idx & 3], (uint32_t)i);
                std::swap(tx[i - 1], tx[j]);
            }
        });
        ++step_counter;
    }
    void run_keyed(uint64_t seed, uint32_t env_id, uint64_t n_steps) {
        for (uint64_t s = 0; s < n_steps; ++s) {
            agents_update_keyed(seed, env_id);
            step_keyed(seed, env_id);
        }
    }
};

// One asset of a MarketEnv seen through the interface the agents use.  The *Market twins of the built-in agents
// (RandomMarketAgents random_agent.rs:165-247, MomentumMarketAgent momentum_agent.rs:282-409, NoiseMarketAgent
// noise_agent.rs:226-345, helpers common.rs:156-261) are the single-asset agents with every call routed through
// (asset, id): statuses and the mid price come from the asset's live book (market_env.rs:328-330,
// `env.get_market().get_order_book(asset).mid_price()`), instructions go to the market's shared queue.
struct AssetEnv {
    MarketEnv& m;
    uint32_t asset;
    OrderBook& book;
    AssetEnv(MarketEnv& m_, uint32_t a) : m(m_), asset(a), book(m_.books[a]) {}
    Status order_status(OrderId id) const { return book.order(id).status; }
    void cancel_order(OrderId id) { m.cancel_order(asset, id); }
    OrderId place_order(Side side, Vol vol, TraderId trader, bool price_is_some, Price price) {
        return m.place_order(asset, side, vol, trader, price_is_some, price);
    }
};

// MarketEnv + ordered agent groups, each bound to one asset: market_sim_runner (runner.rs:107-131) over a
// #[derive(MarketAgentSet)] struct (crates/macros/src/lib.rs).  Keyed contract: the MARKET is the RNG unit — agent
// draws use (market id, step, slot) with slots counted over all groups of the market in declaration order, and one
// shuffle per market and step over the whole queue.
class MarketSim {
public:
    MarketEnv env;
    std::vector<RandomAgents> randoms;
    std::vector<MomentumAgent> momentums;
    std::vector<NoiseAgent> noises;
    struct Slot { GroupKind kind; size_t idx; uint32_t asset; };
    std::vector<Slot> order;
    uint32_t step_counter = 0;
    uint64_t n_instructions = 0;

    MarketSim(Nanos start_time, const std::vector<Price>& ticks, Nanos step_size, bool trading) : env(start_time, ticks, step_size, trading) {}
    void add_random(uint32_t asset, const RandomAgentsParams& p) { order.push_back({GROUP_RANDOM, randoms.size(), asset}); randoms.emplace_back(p); }
    void add_momentum(uint32_t asset, const MomentumParams& p) { order.push_back({GROUP_MOMENTUM, momentums.size(), asset}); momentums.emplace_back(p); }
    void add_noise(uint32_t asset, const NoiseParams& p) { order.push_back({GROUP_NOISE, noises.size(), asset}); noises.emplace_back(p); }

    void run_stream(uint64_t seed, uint64_t n_steps) {  // runner.rs:107-131
        Xoroshiro128StarStar rng = Xoroshiro128StarStar::seed_from_u64(seed);
        for (uint64_t s = 0; s < n_steps; ++s) {
            for (auto& g : order) {
                AssetEnv ae(env, g.asset);
                if (g.kind == GROUP_RANDOM) randoms[g.idx].update_stream(ae, rng);
                else if (g.kind == GROUP_MOMENTUM) momentums[g.idx].update_stream(ae, rng);
                else noises[g.idx].update_stream(ae, rng);
            }
            n_instructions += env.transactions.size();
            env.step(rng);
            ++step_counter;
        }
    }

    void run_keyed(uint64_t seed, uint32_t market_id, uint64_t n_steps) {
        const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
        for (uint64_t s = 0; s < n_steps; ++s) {
            const uint32_t step = step_counter;
            uint32_t slot_base = 0, gi = 0;
            for (auto& g : order) {
                AssetEnv ae(env, g.asset);
                if (g.kind == GROUP_RANDOM) {
                    randoms[g.idx].update_keyed(ae, market_id, step, slot_base, k0, k1);
                    slot_base += randoms[g.idx].p.n_agents;
                } else if (g.kind == GROUP_MOMENTUM) {
                    momentums[g.idx].update_keyed(ae, market_id, step, gi, slot_base, k0, k1);
                    slot_base += momentums[g.idx].p.n_agents;
                } else {
                    noises[g.idx].update_keyed(ae, market_id, step, gi, slot_base, k0, k1);
                    slot_base += noises[g.idx].p.n_agents;
                }
                ++gi;
            }
            n_instructions += env.transactions.size();
            env.step_with([&](std::vector<MarketEnv::MarketEvent>& tx) {
                for (size_t i = tx.size(); i > 1; --i) {  // Fisher-Yates from the back, one Philox word per position
                    const uint32_t idx = (uint32_t)(i - 1);
                    const Philox4 r = philox4x32_10(market_id, step, PHILOX_SLOT_SHUFFLE, idx >> 2, k0, k1);
                    std::swap(tx[i - 1], tx[mulhi_range(r.v[idx & 3], (uint32_t)i)]);
                }
            });
            ++step_counter;
        }
    }
};

}  // namespace oracle
