// Device-side limit order book: one warp owns one book.
//
// Replaces, for the batched path, the reference's per-book data structures and matching engine:
//   crates/order_book/src/side.rs:36-143      OrderBookSide{vol, volumes: BTreeMap, orders: BTreeMap}
//   crates/order_book/src/orderbook.rs:429-772 match_bid/ask, place_*, cancel, reduce, replace, modify
//   crates/order_book/src/orderbook.rs:229-324 bid_ask, *_levels, level_1_data, level_2_data
//
// B200-first layout instead of two B-trees per side:
//   * price ladder = dense 32-level "pages" (vol[32], cnt[32], head[32], tail[32] = 512 B) tagged
//     (price / granule) >> 5 | side.  A page exists only while it holds a level, so any u32 price
//     is representable; the first `p_smem` pages of a book live in shared memory, the rest in HBM.
//   * one lane per page tag: find-page is one __ballot_sync, best bid/ask is one
//     __reduce_min/max_sync over (page key * 32 + __ffs/__clz of the page's 32-bit non-empty bitmap).
//   * two bitmaps per page mirror the reference's two maps: `vmap` <=> key present in `volumes`
//     (count > 0), `qmap` <=> some key of that price present in `orders` (queue non-empty).  They
//     differ only under the reference's equal-(price,time) key collisions.
//   * price-time queue per level = intrusive doubly linked list through the 32-byte hot order
//     records in HBM, kept sorted by key time; the common case (time moved forward) appends without
//     reading the tail record.
// All book code is warp-uniform: every lane executes it with identical values (loads broadcast, stores
// coalesce to one transaction); lanes diverge only in the explicitly lane-parallel helpers.
#pragma once
#include <cstdint>

namespace bb {

typedef uint32_t u32;
typedef uint64_t u64;

#define BB_NIL 0xFFFFFFFFu
#define BB_TAG_FREE 0xFFFFFFFFu
#define BB_FULL 0xFFFFFFFFu

enum { ST_NEW = 0, ST_ACTIVE = 1, ST_FILLED = 2, ST_CANCELLED = 3, ST_REJECTED = 4 };  // types.rs:51-75
#define META_STATUS_MASK 7u
#define META_BID 8u     // side bit (types.rs:26-47: true <-> Bid)
#define META_GHOST 16u  // Active but no longer owning a queue key (N1)

// device error bits == BB_ERR_* of include/bourse_b200.h
#define ERR_CAP_ORDERS 0x01u
#define ERR_CAP_TRADES 0x02u
#define ERR_CAP_PAGES 0x04u
#define ERR_CAP_QUEUE 0x08u
#define ERR_BAD_ID 0x10u
#define ERR_GRANULE 0x20u
#define ERR_CAP_STEPS 0x40u
#define ERR_CAP_LIVE 0x80u

// Order record split by access pattern (types.rs:79-101 `Order` + orderbook.rs:36-44 key):
// hot = what matching / cancel / modify read; cold = write-only timestamps and the trader id.
struct __align__(16) OrderHot {
    u32 price, vol, next, prev;
    u64 key_time;  // time component of the queue key (arrival, or the last replace)
    u32 meta, start_vol;
};
struct __align__(16) OrderCold {
    u64 arr_time, end_time;
    u32 trader, pad0, pad1, pad2;
};
struct __align__(16) TradeRec {  // types.rs:105-118
    u64 t;
    u32 price, vol, active, passive, side_bid, pad;
};

struct __align__(16) BookHdr {  // 128 bytes, head of every book blob
    u64 t, max_key_time, rng_s0, rng_s1;
    u32 n_orders, n_trades, trade_vol, trading;
    u32 side_vol[2];  // [0] ask, [1] bid
    u32 best_q[2];    // best level index (price / granule) by queue, valid iff has_best
    u32 has_best[2];
    u32 err, n_steps;
    u32 step_counter, pad0;
    u64 n_instr, n_transitions, traded_volume, n_trades_total, n_created;
};
static_assert(sizeof(BookHdr) == 128, "BookHdr must stay 128 bytes");
static_assert(sizeof(OrderHot) == 32 && sizeof(OrderCold) == 32 && sizeof(TradeRec) == 32, "record sizes");

// Everything a warp needs to operate on its book.
struct Ctx {
    BookHdr* h;     // shared memory
    u32* tag;       // [p_total] shared
    u32* vmap;      // [p_total] shared
    u32* qmap;      // [p_total] shared
    u32* pg_smem;   // [p_smem][128] shared
    u32* pg_glob;   // [p_total][128] global (slots >= p_smem are live there)
    OrderHot* oh;   // [max_orders] global
    OrderCold* oc;  // [max_orders] global
    TradeRec* tr;   // [max_trades] global
    u32 p_total, p_smem;
    u32 granule, tick;
    u32 max_orders, max_trades;
    u32 lane;
};

__device__ __forceinline__ u32* page_ptr(const Ctx& c, u32 slot) {
    return slot < c.p_smem ? c.pg_smem + slot * 128u : c.pg_glob + (size_t)slot * 128u;
}

// ---- page directory -----------------------------------------------------------------------------
__device__ __forceinline__ u32 find_page(const Ctx& c, u32 side, u32 pkey) {
    const u32 want = (pkey << 1) | side;
    for (u32 b = 0; b < c.p_total; b += 32) {
        const u32 m = __ballot_sync(BB_FULL, c.tag[b + c.lane] == want);
        if (m) return b + __ffs(m) - 1;
    }
    return BB_NIL;
}

__device__ __forceinline__ u32 alloc_page(const Ctx& c, u32 side, u32 pkey) {
    for (u32 b = 0; b < c.p_total; b += 32) {
        const u32 m = __ballot_sync(BB_FULL, c.tag[b + c.lane] == BB_TAG_FREE);
        if (m) {
            const u32 slot = b + __ffs(m) - 1;
            c.tag[slot] = (pkey << 1) | side;
            c.vmap[slot] = 0;
            c.qmap[slot] = 0;
            __syncwarp();
            return slot;
        }
    }
    c.h->err |= ERR_CAP_PAGES;
    return BB_NIL;
}

// best level by queue (== first key of the reference's `orders` map, side.rs:99-104 / 128-130)
__device__ __forceinline__ void recompute_best(const Ctx& c, u32 side) {
    u32 best = side ? 0u : 0xFFFFFFFFu;
    u32 any = 0;
    for (u32 b = 0; b < c.p_total; b += 32) {
        const u32 tg = c.tag[b + c.lane];
        const u32 qm = c.qmap[b + c.lane];
        const bool ok = (tg != BB_TAG_FREE) && ((tg & 1u) == side) && (qm != 0);
        any |= __ballot_sync(BB_FULL, ok);
        if (side) {
            const u32 v = ok ? (((tg >> 1) << 5) + (31u - __clz(qm))) : 0u;
            best = max(best, __reduce_max_sync(BB_FULL, v));
        } else {
            const u32 v = ok ? (((tg >> 1) << 5) + (__ffs(qm) - 1u)) : 0xFFFFFFFFu;
            best = min(best, __reduce_min_sync(BB_FULL, v));
        }
    }
    c.h->best_q[side] = best;
    c.h->has_best[side] = any != 0;
}

// first level of the reference's `volumes` map (side.rs:107-120): (vol, count) at the best price by
// count>0, which can differ from the queue's best only under N1
__device__ __forceinline__ void best_by_volumes(const Ctx& c, u32 side, u32* vol, u32* cnt) {
    u32 best = side ? 0u : 0xFFFFFFFFu;
    u32 any = 0;
    for (u32 b = 0; b < c.p_total; b += 32) {
        const u32 tg = c.tag[b + c.lane];
        const u32 vm = c.vmap[b + c.lane];
        const bool ok = (tg != BB_TAG_FREE) && ((tg & 1u) == side) && (vm != 0);
        any |= __ballot_sync(BB_FULL, ok);
        if (side) {
            const u32 v = ok ? (((tg >> 1) << 5) + (31u - __clz(vm))) : 0u;
            best = max(best, __reduce_max_sync(BB_FULL, v));
        } else {
            const u32 v = ok ? (((tg >> 1) << 5) + (__ffs(vm) - 1u)) : 0xFFFFFFFFu;
            best = min(best, __reduce_min_sync(BB_FULL, v));
        }
    }
    *vol = 0;
    *cnt = 0;
    if (any) {
        const u32 slot = find_page(c, side, best >> 5);
        const u32* pg = page_ptr(c, slot);
        *vol = pg[best & 31u];
        *cnt = pg[32u + (best & 31u)];
    }
}

// side.rs:99-104 + 194-196: empty ask => u32::MAX, empty bid => 0
__device__ __forceinline__ u32 best_price(const Ctx& c, u32 side) {
    if (!c.h->has_best[side]) return side ? 0u : 0xFFFFFFFFu;
    return c.h->best_q[side] * c.granule;
}

// side.rs:138-143 through the bid/ask wrappers: (vol, count) at an arbitrary price; warp-cooperative
__device__ __forceinline__ void level_at(const Ctx& c, u32 side, u32 price, u32* vol, u32* cnt) {
    *vol = 0;
    *cnt = 0;
    const u32 q = price / c.granule;
    if (q * c.granule != price) return;
    const u32 slot = find_page(c, side, q >> 5);
    if (slot == BB_NIL) return;
    if (!((c.vmap[slot] >> (q & 31u)) & 1u)) return;
    const u32* pg = page_ptr(c, slot);
    *vol = pg[q & 31u];
    *cnt = pg[32u + (q & 31u)];
}

// same lookup done independently by each lane (divergent prices) — used by the level-2 emitter
__device__ __forceinline__ void level_at_lane(const Ctx& c, u32 side, u32 price, u32* vol, u32* cnt) {
    *vol = 0;
    *cnt = 0;
    const u32 q = price / c.granule;
    if (q * c.granule != price) return;
    const u32 want = ((q >> 5) << 1) | side;
    u32 slot = BB_NIL;
    for (u32 j = 0; j < c.p_total; ++j)
        if (c.tag[j] == want) slot = j;
    if (slot == BB_NIL) return;
    if (!((c.vmap[slot] >> (q & 31u)) & 1u)) return;
    const u32* pg = page_ptr(c, slot);
    *vol = pg[q & 31u];
    *cnt = pg[32u + (q & 31u)];
}

// ---- volumes-map half of insert_order / remove_order / remove_vol (side.rs:54-96) ------------------
__device__ __forceinline__ void level_add(const Ctx& c, u32 side, u32 slot, u32 l, u32 vol) {
    u32* pg = page_ptr(c, slot);
    const u32 bit = 1u << l;
    if (c.vmap[slot] & bit) {
        pg[l] += vol;
        pg[32u + l] += 1u;
    } else {
        pg[l] = vol;
        pg[32u + l] = 1u;
        c.vmap[slot] |= bit;
    }
    c.h->side_vol[side] += vol;
}

// returns true when the page was released
__device__ __forceinline__ bool level_remove(const Ctx& c, u32 side, u32 slot, u32 l, u32 vol) {
    u32* pg = page_ptr(c, slot);
    const u32 bit = 1u << l;
    pg[l] -= vol;
    const u32 cnt = pg[32u + l] - 1u;
    pg[32u + l] = cnt;
    c.h->side_vol[side] -= vol;
    if (cnt == 0) {
        const u32 vm = c.vmap[slot] & ~bit;
        c.vmap[slot] = vm;
        if (vm == 0 && c.qmap[slot] == 0) {
            c.tag[slot] = BB_TAG_FREE;
            __syncwarp();
            return true;
        }
    }
    return false;
}

// ---- orders-map half: the price-time queue ----------------------------------------------------------
// unlink record `id` (with links prev/next) from level (slot,l); maintains qmap and the cached best
__device__ __forceinline__ void queue_unlink(const Ctx& c, u32 side, u32 slot, u32 l, u32 q, u32 prev, u32 next) {
    u32* pg = page_ptr(c, slot);
    if (prev == BB_NIL) pg[64u + l] = next; else c.oh[prev].next = next;
    if (next == BB_NIL) pg[96u + l] = prev; else c.oh[next].prev = prev;
    if (prev == BB_NIL && next == BB_NIL) {
        c.qmap[slot] &= ~(1u << l);
        __syncwarp();
        if (c.h->has_best[side] && c.h->best_q[side] == q) recompute_best(c, side);
    }
}

// orders.remove(&(price', key_time)) for an order that does NOT own its key any more (N1 ghost):
// whoever owns that key now loses it and becomes a ghost itself.
__device__ __noinline__ void queue_remove_key_slow(const Ctx& c, u32 side, u32 slot, u32 l, u32 q, u64 key_time) {
    const u32* pg = page_ptr(c, slot);
    if (!((c.qmap[slot] >> l) & 1u)) return;
    u32 cur = pg[64u + l];
    while (cur != BB_NIL) {
        const OrderHot r = c.oh[cur];
        if (r.key_time == key_time) {
            c.oh[cur].meta = r.meta | META_GHOST;
            queue_unlink(c, side, slot, l, q, r.prev, r.next);
            return;
        }
        if (r.key_time > key_time) return;
        cur = r.next;
    }
}

// orders.insert((price', t), id) when time did not move strictly forward: sorted position, or take
// over an existing equal key.  Returns the (prev,next) links the new record must carry.
__device__ __noinline__ void queue_insert_slow(const Ctx& c, u32 slot, u32 l, u32 id, u64 t, u32* out_prev, u32* out_next) {
    u32* pg = page_ptr(c, slot);
    u32 cur = pg[96u + l];  // walk back from the tail
    u32 after = BB_NIL;     // node that will follow the new one
    while (cur != BB_NIL) {
        const OrderHot r = c.oh[cur];
        if (r.key_time < t) break;
        if (r.key_time == t) {
            // BTreeMap::insert on an existing key: value replaced, position kept (side.rs:55)
            c.oh[cur].meta = r.meta | META_GHOST;
            if (r.prev == BB_NIL) pg[64u + l] = id; else c.oh[r.prev].next = id;
            if (r.next == BB_NIL) pg[96u + l] = id; else c.oh[r.next].prev = id;
            *out_prev = r.prev;
            *out_next = r.next;
            return;
        }
        after = cur;
        cur = r.prev;
    }
    // insert between cur (may be NIL => new head) and after (may be NIL => new tail)
    if (cur == BB_NIL) pg[64u + l] = id; else c.oh[cur].next = id;
    if (after == BB_NIL) pg[96u + l] = id; else c.oh[after].prev = id;
    *out_prev = cur;
    *out_next = after;
}

// side.rs:54-66 insert_order for a resting order.  Returns false when the order could not rest.
__device__ __forceinline__ bool book_insert(const Ctx& c, u32 side, u32 price, u64 t, u32 id, u32 vol, u32* out_prev,
                                            u32* out_next) {
    *out_prev = BB_NIL;
    *out_next = BB_NIL;
    const u32 q = price / c.granule;
    if (q * c.granule != price) {
        c.h->err |= ERR_GRANULE;
        return false;
    }
    u32 slot = find_page(c, side, q >> 5);
    if (slot == BB_NIL) {
        slot = alloc_page(c, side, q >> 5);
        if (slot == BB_NIL) return false;
    }
    const u32 l = q & 31u;
    level_add(c, side, slot, l, vol);
    u32* pg = page_ptr(c, slot);
    const u32 bit = 1u << l;
    if (!(c.qmap[slot] & bit)) {
        pg[64u + l] = id;
        pg[96u + l] = id;
        c.qmap[slot] |= bit;
        const bool better = !c.h->has_best[side] || (side ? q > c.h->best_q[side] : q < c.h->best_q[side]);
        if (better) {
            c.h->best_q[side] = q;
            c.h->has_best[side] = 1;
        }
    } else if (t > c.h->max_key_time) {
        const u32 tail = pg[96u + l];
        c.oh[tail].next = id;
        pg[96u + l] = id;
        *out_prev = tail;
    } else {
        queue_insert_slow(c, slot, l, id, t, out_prev, out_next);
    }
    if (t > c.h->max_key_time) c.h->max_key_time = t;
    __syncwarp();
    return true;
}

// side.rs:75-84 remove_order(key, vol) for the order described by `r`
__device__ __forceinline__ void book_remove(const Ctx& c, u32 side, const OrderHot& r, u32 vol) {
    const u32 q = r.price / c.granule;
    const u32 slot = find_page(c, side, q >> 5);
    if (slot == BB_NIL) return;  // unreachable for Active orders
    const u32 l = q & 31u;
    if (r.meta & META_GHOST) queue_remove_key_slow(c, side, slot, l, q, r.key_time);
    else queue_unlink(c, side, slot, l, q, r.prev, r.next);
    level_remove(c, side, slot, l, vol);
}

__device__ __forceinline__ void log_trade(const Ctx& c, u64 t, u32 passive_bid, u32 price, u32 vol, u32 active, u32 passive) {
    const u32 n = c.h->n_trades;
    if (n < c.max_trades) {
        TradeRec rec;
        rec.t = t; rec.price = price; rec.vol = vol; rec.active = active; rec.passive = passive;
        rec.side_bid = passive_bid; rec.pad = 0;
        c.tr[n] = rec;
        c.h->n_trades = n + 1;
    } else if (c.max_trades) {
        c.h->err |= ERR_CAP_TRADES;
    }
    c.h->n_trades_total += 1;
}

// orderbook.rs:429-487 match_bid / match_ask + :843-870 match_orders.  Sweeps the opposite side in
// price-time order; returns the aggressor's remaining volume, *filled as the reference's Status::Filled.
__device__ __forceinline__ u32 book_match(const Ctx& c, u32 side, u32 price, u32 vol, u32 id, u64 t, bool* filled) {
    const u32 o = side ^ 1u;
    *filled = false;
    u32 slot = BB_NIL, slot_key = BB_NIL;
    while (vol > 0 && c.h->has_best[o]) {
        const u32 bq = c.h->best_q[o];
        const u32 bprice = bq * c.granule;
        if (side ? (price < bprice) : (price > bprice)) break;
        if (slot_key != (bq >> 5)) {
            slot = find_page(c, o, bq >> 5);
            slot_key = bq >> 5;
        }
        if (slot == BB_NIL) {  // only reachable after a capacity error left the book inconsistent
            c.h->err |= ERR_CAP_PAGES;
            break;
        }
        const u32 l = bq & 31u;
        u32* pg = page_ptr(c, slot);
        const u32 hid = pg[64u + l];
        if (hid >= c.max_orders) {
            c.h->err |= ERR_BAD_ID;
            break;
        }
        const uint4 ph = *reinterpret_cast<const uint4*>(&c.oh[hid]);  // price, vol, next, prev
        const u32 tv = min(vol, ph.y);
        vol -= tv;
        const u32 pvol = ph.y - tv;
        log_trade(c, t, o, ph.x, tv, id, hid);
        c.h->trade_vol += tv;
        c.h->traded_volume += tv;
        c.h->n_transitions += 1;
        if (pvol == 0) {
            c.oh[hid].vol = 0;
            c.oh[hid].meta = ST_FILLED | (o ? META_BID : 0u);
            c.oc[hid].end_time = t;
            // side.remove_order(match.key, tv): the head always owns its key
            const u32 nxt = ph.z;
            pg[64u + l] = nxt;
            bool emptied = false;
            if (nxt == BB_NIL) {
                pg[96u + l] = BB_NIL;
                c.qmap[slot] &= ~(1u << l);
                emptied = true;
            } else {
                c.oh[nxt].prev = BB_NIL;
            }
            const bool released = level_remove(c, o, slot, l, tv);
            if (released) slot_key = BB_NIL;
            if (emptied) {
                __syncwarp();
                recompute_best(c, o);
            }
        } else {
            c.oh[hid].vol = pvol;
            pg[l] -= tv;  // side.remove_vol(price, tv)
            c.h->side_vol[o] -= tv;
        }
        if (vol == 0) *filled = true;
    }
    return vol;
}

// orderbook.rs:583-611 place_order for a freshly created order whose fields are all known to the
// caller (create_order :356-396 happened at submission).  Writes the complete record.
__device__ __forceinline__ void book_place(const Ctx& c, u32 id, u32 side, u32 price, u32 vol, u32 trader, u64 t) {
    if (id >= c.max_orders) {
        c.h->err |= ERR_CAP_ORDERS;
        return;
    }
    const bool market = side ? (price == 0xFFFFFFFFu) : (price == 0u);  // N3: decided by the price value
    u32 rem = vol, status = ST_ACTIVE, prev = BB_NIL, next = BB_NIL;
    u64 end_time = ~0ULL;
    bool filled = false;
    if (market) {
        if (c.h->trading) {
            rem = book_match(c, side, price, vol, id, t, &filled);
            status = filled ? ST_FILLED : ST_CANCELLED;  // orderbook.rs:517-531
        } else {
            status = ST_REJECTED;
        }
        end_time = t;
    } else {
        if (c.h->trading) rem = book_match(c, side, price, vol, id, t, &filled);
        if (filled) {
            status = ST_FILLED;
            end_time = t;
        } else {
            book_insert(c, side, price, t, id, rem, &prev, &next);  // orderbook.rs:499-504
        }
    }
    OrderHot r;
    r.price = price; r.vol = rem; r.next = next; r.prev = prev;
    r.key_time = t; r.meta = status | (side ? META_BID : 0u); r.start_vol = vol;
    OrderCold cr;
    cr.arr_time = t; cr.end_time = end_time; cr.trader = trader; cr.pad0 = cr.pad1 = cr.pad2 = 0;
    c.oh[id] = r;
    c.oc[id] = cr;
    c.h->n_transitions += 1;
}

// orderbook.rs:622-644
__device__ __forceinline__ void book_cancel(const Ctx& c, u32 id, u64 t) {
    if (id >= c.h->n_orders || id >= c.max_orders) {
        c.h->err |= ERR_BAD_ID;
        return;
    }
    const OrderHot r = c.oh[id];
    if ((r.meta & META_STATUS_MASK) != ST_ACTIVE) return;
    const u32 side = (r.meta & META_BID) ? 1u : 0u;
    c.oh[id].meta = ST_CANCELLED | (r.meta & META_BID);
    c.oc[id].end_time = t;
    book_remove(c, side, r, r.vol);
    c.h->n_transitions += 1;
}

// orderbook.rs:743-772 (+ reduce_order_vol :656-667, replace_order :679-723)
__device__ __forceinline__ void book_modify(const Ctx& c, u32 id, bool has_p, u32 new_p, bool has_v, u32 new_v, u64 t) {
    if (id >= c.h->n_orders || id >= c.max_orders) {
        c.h->err |= ERR_BAD_ID;
        return;
    }
    const OrderHot r = c.oh[id];
    if ((r.meta & META_STATUS_MASK) != ST_ACTIVE) return;
    if (!has_p && !has_v) return;
    const u32 side = (r.meta & META_BID) ? 1u : 0u;
    if (!has_p && new_v < r.vol) {
        const u32 d = r.vol - new_v;
        c.oh[id].vol = new_v;
        const u32 q = r.price / c.granule;
        const u32 slot = find_page(c, side, q >> 5);
        if (slot != BB_NIL) {
            page_ptr(c, slot)[q & 31u] -= d;
            c.h->side_vol[side] -= d;
        }
        c.h->n_transitions += 1;
        return;
    }
    const u32 price = has_p ? new_p : r.price;
    const u32 vol = has_v ? new_v : r.vol;
    book_remove(c, side, r, r.vol);
    u32 rem = vol, prev = BB_NIL, next = BB_NIL;
    bool filled = false;
    if (c.h->trading) rem = book_match(c, side, price, vol, id, t, &filled);
    OrderHot w;
    w.price = price; w.vol = rem; w.start_vol = r.start_vol;
    if (filled) {
        w.meta = ST_FILLED | (r.meta & META_BID);
        w.key_time = r.key_time;
        c.oc[id].end_time = t;
    } else {
        book_insert(c, side, price, t, id, rem, &prev, &next);
        w.meta = ST_ACTIVE | (r.meta & META_BID);
        w.key_time = t;
    }
    w.next = next; w.prev = prev;
    c.oh[id] = w;
    c.h->n_transitions += 1;
}

// ---- observation emission ------------------------------------------------------------------------
// Level-1 / level-2 record in the array layout of rust/src/step_sim_numpy.rs:300-368:
// [trade_vol, bid_price, ask_price, ask_vol, bid_vol, then per level i: bid_vol_i, n_bid_i, ask_vol_i, n_ask_i]
// Levels sit at FIXED tick offsets from the touch with wrapping arithmetic (orderbook.rs:229-264).
// Each lane returns the word(s) it owns: word index = lane (and lane + 32 for the 45-word record).
__device__ __forceinline__ void book_obs(const Ctx& c, u32 n_words, u32* w0, u32* w1) {
    const u32 bid = best_price(c, 1), ask = best_price(c, 0);
    u32 a = 0, b = 0;
    if (n_words <= 9u) {
        u32 bv, bc, av, ac;
        level_at(c, 1, bid, &bv, &bc);
        level_at(c, 0, ask, &av, &ac);
        const u32 l = c.lane;
        a = l == 0 ? c.h->trade_vol : l == 1 ? bid : l == 2 ? ask : l == 3 ? c.h->side_vol[0] : l == 4 ? c.h->side_vol[1]
          : l == 5 ? bv : l == 6 ? bc : l == 7 ? av : l == 8 ? ac : 0u;
    } else {
        // words 5..44: level i = (w-5)/4, field f = (w-5)%4 -> f<2 bid (vol,cnt), else ask (vol,cnt)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const u32 w = c.lane + 32u * half;
            u32 val = 0;
            if (w >= 5u && w < 45u) {
                const u32 i = (w - 5u) >> 2, f = (w - 5u) & 3u;
                u32 v, n;
                if (f < 2u) level_at_lane(c, 1, bid - i * c.tick, &v, &n);
                else level_at_lane(c, 0, ask + i * c.tick, &v, &n);
                val = (f & 1u) ? n : v;
            } else if (w < 5u) {
                val = w == 0 ? c.h->trade_vol : w == 1 ? bid : w == 2 ? ask : w == 3 ? c.h->side_vol[0] : c.h->side_vol[1];
            }
            if (half == 0) a = val; else b = val;
        }
    }
    *w0 = a;
    *w1 = b;
}

}  // namespace bb
