// Deep-book engine: ONE CTA PER BOOK, three specialised warps (included by kernels.cuh).
//
// Same reference semantics as the other engines (side.rs:36-143, orderbook.rs:429-772), built for books with ~10^6
// resting orders (BASELINE config C5) where one warp per book left an SM running a single dependent chain:
//
//   fetch warp   walks the instruction stream 32 events ahead of the matcher: the batch itself arrives by one bulk
//                (TMA) copy, and every lane pulls the order record its cancel / modify names (the only DRAM-cold read of
//                the path) into a shared-memory ring with cp.async, so the matcher never waits on HBM;
//   match warp   the serial part: lane 0 runs the price-time logic out of shared memory only (dense tick-indexed ladder,
//                per-side non-empty bitmaps with a summary word, a 64-chunk cache of queue heads), and calls the whole
//                warp in for what is parallel inside one book: sweeping a level's FIFO with an inclusive prefix sum
//                over the resting volumes (one trade per passive order, in queue order, orderbook.rs:429-454, 843-870),
//                chunk loads and the 45-word observation record;
//   retire warp  drains a ring of order-record / trade-log writes (one lane per entry) so that the matcher's chain holds
//                no global store but the 8-byte queue entries.
//
// Price-time queues are ARRAYS, not linked lists: a level's FIFO is a chain of 256-byte chunks in HBM (31 entries of
// {order id, remaining volume} + a link).  An order remembers the position of its entry, so a cancel writes a tombstone
// there and never touches its neighbours, and a sweep reads consecutive entries with one coalesced load.
//
// Preconditions (violations set a per-env error bit and never change results silently):
//   * resting prices inside [d_win_lo, d_win_lo + d_levels)                                     else ERR_CAP_PAGES
//   * strictly increasing time between resting inserts (no equal-(price, time) keys, N1)         else ERR_TIME_ORDER
//   * a price level holds one side at a time (true while trading is enabled and volumes are > 0)  else ERR_LOCKED
//   * chunk pool not exhausted                                                                   else ERR_CAP_PAGES
#pragma once

namespace bb {

#define ERR_LOCKED 0x800u  // deep engine: both sides resting at one price (only possible with trading disabled)

// ---- shared-memory image of a deep book (persisted in the book blob between launches) -------------------------------
//   0     BookHdr (128 B)
//   128   chunk_bump u32 (next never-used chunk id; chunk 0 is a sink), n_free u32 (entries on the free-chunk stack)
//   192   free-chunk stack u32 [DP_FS_CAP]
//   1216  bitmap ask [NW], bitmap bid [NW]   bit q set <=> level q holds resting orders of that side
//         summary ask [8], summary bid [8]   bit w set <=> bitmap word w is non-zero
//         levels [W] x 16 B: {vol, cnt, head, tail}; head / tail = chunk id << 5 | entry index (tail: next free entry)
#define DP_OFF_BUMP 128u
#define DP_OFF_NFREE 132u
#define DP_OFF_FS 192u
#define DP_FS_CAP 256u
#define DP_OFF_BM (DP_OFF_FS + 4u * DP_FS_CAP)
#define DP_NS 8u  // summary words per side: up to 8192 levels
#define DP_CHUNK_BYTES 256u
#define DP_CHUNK_ENTRIES 31u
// scratch (not persisted), relative to the end of the image
#define DP_NC 64u            // chunk cache entries (direct mapped by chunk id)
#define DP_RB 2u             // event-ring depth in batches of 32 (one batch of look-ahead hides the record fetch)
#define DP_RCAP 128u         // retire-ring entries (32 B each)
#define DP_RENT 32u
#define DP_DIRTY 2048u       // dirty-filter buckets
struct DeepOff {             // byte offsets from the CTA's shared-memory base, filled by the host
    u32 bm, sm, lv, image_bytes;
    u32 ctag, cdat, ev_ins, ev_rec, ev_rf, ret, dirty, ctl, bar, total;
};
// control words (u32 each, at DeepOff::ctl)
#define CT_EV_READY 0u      // batches the fetch warp has published
#define CT_EV_CONSUMED 4u   // batches the match warp is done with
#define CT_RET_TAIL 8u      // retire entries produced
#define CT_RET_DONE 12u     // retire entries whose stores are performed and fenced
#define CT_FIN 16u          // match warp: no more retire entries will come
#define CT_ABORT 20u        // any warp: a bounded wait ran out (the launch flags 0x80000000 and unwinds)
#define CT_RERR 24u         // retire warp's error bits
#define CT_NTR 28u          // retire warp's trade-log cursor at exit
#define CT_MERR 32u         // match warp's error bits
#define CT_WORDS 16u

// retire entry kinds
#define RK_NEW 1u
#define RK_REPLACE 2u
#define RK_FILL 3u
#define RK_CANCEL 4u
#define RK_REDUCE 5u

__device__ __forceinline__ u32 ld_acq(u32 a) {
    u32 v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void st_rel(u32 a, u32 v) { asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ u64 ldg64_cg(u64 a) {
    u64 v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ldg128_cg(u64 a) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a) : "memory");
    return v;
}
__device__ __forceinline__ void cp_async16(u32 smem_dst, u64 gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
// bounded spin on a control word; returns false when the wait ran out or another warp aborted
#define DP_SPIN_MAX (1u << 25)
template <class F> __device__ __forceinline__ bool dp_wait(u32 ctl, F cond) {
    for (u32 spin = 0; spin < DP_SPIN_MAX; ++spin) {
        if (cond()) return true;
        if ((spin & 255u) == 255u && ld_acq(ctl + CT_ABORT)) return false;
    }
    st_rel(ctl + CT_ABORT, 1u);
    return false;
}

// Launch-invariant addresses and sizes of the match warp, pinned in registers (an opaque move keeps the optimiser from
// rebuilding `base + constant-bank offset` at every use: ~40 % of the first version's instructions were such rebuilds).
struct DeepReg {
    u32 lv, bma, bmb, sma, smb, ctag, cdat, ret, dirty, ctl, fs;
    u32 win_lo, W, n_chunks, max_orders;
    u64 chunks, oh;
};
__device__ __forceinline__ u32 dp_keep32(u32 v) {
    u32 r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ u64 dp_keep64(u64 v) {
    u64 r;
    asm volatile("mov.u64 %0, %1;" : "=l"(r) : "l"(v));
    return r;
}

// Lane-0 state of the match warp.  Everything the serial chain touches between two events lives here or in shared memory.
struct DeepSt {
    u64 t, max_key_time;
    u32 n_orders, n_trades, trade_vol, vol_ask, vol_bid, bq_ask, bq_bid, flags, err;
    u32 d_instr, d_applied;
    u32 bump, n_free;
    u32 ret_tail;   // retire entries produced so far
    u32 ret_room;   // ... and the value of ret_tail up to which ring space is known to be free
};

__device__ __forceinline__ u32 dp_bm(const DeepReg& r, u32 side, u32 w) { return (side ? r.bmb : r.bma) + 4u * w; }
__device__ __forceinline__ u32 dp_sm(const DeepReg& r, u32 side, u32 w) { return (side ? r.smb : r.sma) + 4u * w; }
__device__ __forceinline__ bool dp_has_best(const DeepSt& s, u32 side) { return (s.flags >> (1u + side)) & 1u; }
__device__ __forceinline__ u32 dp_best_q(const DeepSt& s, u32 side) { return side ? s.bq_bid : s.bq_ask; }
__device__ __forceinline__ void dp_add_side(DeepSt& s, u32 side, u32 dv) {
    if (side) s.vol_bid += dv; else s.vol_ask += dv;
}

// level q of `side` just became empty (its bitmap bit is still set): clear it and, if it was the touch, find the next one
__device__ __forceinline__ void dp_level_gone(const DeepReg& r, DeepSt& s, u32 side, u32 q) {
    const u32 w = q >> 5;
    const u32 ba = dp_bm(r, side, w);
    u32 m = lds(ba) & ~(1u << (q & 31u));
    sts(ba, m);
    if (m == 0u) {
        const u32 sa = dp_sm(r, side, w >> 5);
        sts(sa, lds(sa) & ~(1u << (w & 31u)));
    }
    if (!(dp_has_best(s, side) && dp_best_q(s, side) == q)) return;
    // q was the best level: every other level of this side lies on the far side of it
    if (side == 0u) {
        if (m) { s.bq_ask = (w << 5) + (u32)__ffs(m) - 1u; return; }
        u32 sw = w >> 5;
        u32 ms = lds(dp_sm(r, 0u, sw)) & ~((2u << (w & 31u)) - 1u);  // summary bits above word w
        while (ms == 0u) {
            if (++sw >= DP_NS) { s.flags &= ~FL_HAS_ASK; return; }
            ms = lds(dp_sm(r, 0u, sw));
        }
        const u32 w2 = (sw << 5) + (u32)__ffs(ms) - 1u;
        s.bq_ask = (w2 << 5) + (u32)__ffs(lds(dp_bm(r, 0u, w2))) - 1u;
    } else {
        if (m) { s.bq_bid = (w << 5) + 31u - (u32)__clz(m); return; }
        u32 sw = w >> 5;
        u32 ms = lds(dp_sm(r, 1u, sw)) & ((1u << (w & 31u)) - 1u);  // summary bits below word w
        while (ms == 0u) {
            if (sw == 0u) { s.flags &= ~FL_HAS_BID; return; }
            --sw;
            ms = lds(dp_sm(r, 1u, sw));
        }
        const u32 w2 = (sw << 5) + 31u - (u32)__clz(ms);
        s.bq_bid = (w2 << 5) + 31u - (u32)__clz(lds(dp_bm(r, 1u, w2)));
    }
}

__device__ __forceinline__ u32 dp_alloc_chunk(const DeepReg& r, DeepSt& s) {
    u32 c;
    if (s.n_free) {
        s.n_free -= 1;
        c = lds(r.fs + 4u * s.n_free);
    } else if (s.bump < r.n_chunks) {
        c = s.bump++;
    } else {
        s.err |= ERR_CAP_PAGES;
        return 0u;  // chunk 0 is never handed out: a safe sink
    }
    // a fresh chunk is born in the cache (write-through): no load when a sweep reaches it while it is still resident
    const u32 slot = c & (DP_NC - 1u);
    sts(r.ctag + 4u * slot, c);
    sts(r.cdat + DP_CHUNK_BYTES * slot + 8u * DP_CHUNK_ENTRIES, BB_NIL);
    stg32(r.chunks + (u64)c * DP_CHUNK_BYTES + 8u * DP_CHUNK_ENTRIES, BB_NIL);
    return c;
}
__device__ __forceinline__ void dp_free_chunk(const DeepReg& r, DeepSt& s, u32 c) {
    if (c != 0u && s.n_free < DP_FS_CAP) {
        sts(r.fs + 4u * s.n_free, c);
        s.n_free += 1;
    }
}
// write-through store of one word of a queue entry / chunk link
__device__ __forceinline__ void dp_chunk_st32(const DeepReg& r, u32 c, u32 byte_off, u32 v) {
    stg32(r.chunks + (u64)c * DP_CHUNK_BYTES + byte_off, v);
    const u32 slot = c & (DP_NC - 1u);
    if (lds(r.ctag + 4u * slot) == c) sts(r.cdat + DP_CHUNK_BYTES * slot + byte_off, v);
}

// retire ring ---------------------------------------------------------------------------------------------------------
// room for n more entries?  The consumer's counter is re-read only when the cached bound is used up.
__device__ __forceinline__ bool dp_ret_space(const DeepReg& r, DeepSt& s, u32 n) {
    if (s.ret_tail + n <= s.ret_room) return true;
    st_rel(r.ctl + CT_RET_TAIL, s.ret_tail);  // whatever is waiting to be drained must be visible to the drainer
    const bool ok = dp_wait(r.ctl, [&] {
        s.ret_room = ld_acq(r.ctl + CT_RET_DONE) + DP_RCAP;
        return s.ret_tail + n <= s.ret_room;
    });
    return ok;
}
__device__ __forceinline__ void dp_ret_write(const DeepReg& r, u32 idx, uint4 a, uint4 b) {
    const u32 ea = r.ret + DP_RENT * (idx & (DP_RCAP - 1u));
    sts128(ea, a);
    sts128(ea + 16u, b);
}
__device__ __forceinline__ void dp_mark_dirty(const DeepReg& r, u32 id, u32 upto) { sts(r.dirty + 4u * (id & (DP_DIRTY - 1u)), upto); }

// insert_order (side.rs:54-66) for an order that rests: append to its level's chunk chain.  Returns the entry position.
__device__ __forceinline__ u32 dp_insert(const DeepReg& r, DeepSt& s, u32 side, u32 price, u64 t, u32 id, u32 vol) {
    const u32 q = price - r.win_lo;
    if (q >= r.W) {
        s.err |= ERR_CAP_PAGES;
        return 0u;
    }
    const u32 ba = dp_bm(r, side, q >> 5), bit = 1u << (q & 31u);
    const u32 la = r.lv + 16u * q;
    const u32 bw = lds(ba);
    const uint4 lv = lds128(la);  // (issued before the branch: its latency overlaps the bitmap test)
    u32 pos;
    if (!(bw & bit)) {
        // the level record is shared by the two sides.  While trading is enabled an order only rests where the other
        // side is empty, with one exception the reference allows: a ZERO-volume order never matches (orderbook.rs:436,
        // `while order.vol > 0`) and rests wherever its price says, even inside the other side
        if (lds(dp_bm(r, side ^ 1u, q >> 5)) & bit) {
            s.err |= ERR_LOCKED;
            return 0u;
        }
        const u32 c = dp_alloc_chunk(r, s);
        pos = c << 5;
        sts128(la, make_uint4(vol, 1u, pos, pos + 1u));
        sts(ba, bw | bit);
        if (bw == 0u) {
            const u32 sa = dp_sm(r, side, q >> 10);
            sts(sa, lds(sa) | (1u << ((q >> 5) & 31u)));
        }
        const bool better = !dp_has_best(s, side) || (side ? q > s.bq_bid : q < s.bq_ask);
        if (better) {
            if (side) s.bq_bid = q; else s.bq_ask = q;
            s.flags |= FL_HAS_ASK << side;
        }
    } else {
        if (t <= s.max_key_time) s.err |= ERR_TIME_ORDER;
        u32 c = lv.w >> 5, idx = lv.w & 31u;
        if (idx == DP_CHUNK_ENTRIES) {  // tail chunk full: link a new one
            const u32 c2 = dp_alloc_chunk(r, s);
            dp_chunk_st32(r, c, 8u * DP_CHUNK_ENTRIES, c2);
            c = c2;
            idx = 0u;
        }
        pos = (c << 5) | idx;
        sts128(la, make_uint4(lv.x + vol, lv.y + 1u, lv.z, pos + 1u));
    }
    // the queue entry itself, write-through
    const u32 c = pos >> 5, idx = pos & 31u;
    const u32 slot = c & (DP_NC - 1u);
    const u32 tag = lds(r.ctag + 4u * slot);
    stg64(r.chunks + (u64)c * DP_CHUNK_BYTES + 8u * idx, ((u64)vol << 32) | id);
    if (tag == c) sts64(r.cdat + DP_CHUNK_BYTES * slot + 8u * idx, ((u64)vol << 32) | id);
    dp_add_side(s, side, vol);
    if (t > s.max_key_time) s.max_key_time = t;
    return pos;
}

// remove_order (side.rs:75-84): tombstone the entry, settle the level
__device__ __forceinline__ void dp_remove(const DeepReg& r, DeepSt& s, u32 side, u32 price, u32 vol, u32 pos) {
    const u32 q = price - r.win_lo;
    if (q >= r.W) return;  // never rested (flagged when it was placed)
    const u32 la = r.lv + 16u * q;
    const uint4 lv = lds128(la);
    dp_chunk_st32(r, pos >> 5, 8u * (pos & 31u), BB_NIL);
    dp_add_side(s, side, 0u - vol);
    if (lv.y <= 1u) {
        if ((lv.z >> 5) == (lv.w >> 5)) dp_free_chunk(r, s, lv.z >> 5);  // a longer all-dead chain is left to the pool
        dp_level_gone(r, s, side, q);
    } else {
        sts128(la, make_uint4(lv.x - vol, lv.y - 1u, lv.z, lv.w));
    }
}

// ---- matching, lane 0 alone: match_bid / match_ask (orderbook.rs:429-487) + match_orders (:843-870) over queue chunks that
// are resident in the chunk cache.  Most aggressive orders end after one to three fills, and for those walking the entries
// serially is cheaper than calling the warp in.  Returns true when the rest of the sweep needs the whole warp: the head
// chunk is not resident (one coalesced load), or this order has already filled DP_SERIAL_FILLS passive orders (a real
// sweep: the prefix-sum path takes it from there).  The book is consistent at every return.
#define DP_SERIAL_FILLS 4u
__device__ __forceinline__ bool dp_match_serial(const DeepReg& r, DeepSt& s, u32 side, u32 price, u32& rem, u32 id, bool& aborted) {
    const u32 opp = side ^ 1u;
    u32 fills = 0u;
    while (rem > 0u && dp_has_best(s, opp)) {
        const u32 bq = dp_best_q(s, opp);
        const u32 bprice = r.win_lo + bq;
        if (side ? (price < bprice) : (price > bprice)) break;
        const u32 la = r.lv + 16u * bq;
        const uint4 lv = lds128(la);
        const u32 c = lv.z >> 5, tc = lv.w >> 5;
        const u32 slot = c & (DP_NC - 1u);
        if (lds(r.ctag + 4u * slot) != c || fills >= DP_SERIAL_FILLS) return true;
        const u32 ca = r.cdat + DP_CHUNK_BYTES * slot;
        const u32 end = (c == tc) ? (lv.w & 31u) : DP_CHUNK_ENTRIES;
        u32 idx = lv.z & 31u, lvol = lv.x, cnt = lv.y;
        while (idx < end) {
            const u64 e = lds64(ca + 8u * idx);
            const u32 pid = (u32)e, pvol = (u32)(e >> 32);
            if (pid == BB_NIL) {  // tombstone
                ++idx;
                continue;
            }
            const u32 tv = min(rem, pvol), pv = pvol - tv;
            if (!dp_ret_space(r, s, 1u)) {
                aborted = true;
                return false;
            }
            dp_ret_write(r, s.ret_tail, make_uint4(RK_FILL | (opp << 8) | (pv == 0u ? 0x10000u : 0u), pid, tv, pv),
                         make_uint4((u32)s.t, (u32)(s.t >> 32), bprice, id));
            s.ret_tail += 1;
            dp_mark_dirty(r, pid, s.ret_tail);
            rem -= tv;
            lvol -= tv;
            s.trade_vol += tv;
            s.n_trades += 1;
            ++fills;
            if (pv != 0u) {  // partially filled: stays at the head; the aggressor is done
                sts(ca + 8u * idx + 4u, pv);
                stg32(r.chunks + (u64)c * DP_CHUNK_BYTES + 8u * idx + 4u, pv);
                break;
            }
            ++idx;
            --cnt;
            if (cnt == 0u || rem == 0u || fills >= DP_SERIAL_FILLS) break;
        }
        dp_add_side(s, opp, lvol - lv.x);
        if (cnt == 0u) {  // the level is gone
            if (c == tc) dp_free_chunk(r, s, c);
            dp_level_gone(r, s, opp, bq);
        } else if (idx >= end && c != tc) {  // this chunk is used up: follow the link
            const u32 nc = lds(ca + 8u * DP_CHUNK_ENTRIES);
            dp_free_chunk(r, s, c);
            if (nc < r.n_chunks) {
                sts128(la, make_uint4(lvol, cnt, nc << 5, lv.w));
            } else {
                s.err |= ERR_CAP_PAGES;
                dp_level_gone(r, s, opp, bq);
            }
        } else if (idx >= end) {  // live orders counted but none found: only after an earlier capacity error
            s.err |= ERR_CAP_PAGES;
            dp_level_gone(r, s, opp, bq);
        } else {
            sts128(la, make_uint4(lvol, cnt, (c << 5) | idx, lv.w));
        }
    }
    return false;
}

// ---- the warp-cooperative sweep: the same loop with a prefix sum over the FIFO --------------------------------------------
// Lanes load consecutive queue entries of the head chunk, an inclusive scan over the resting volumes tells every lane
// whether its order is reached and whether it is filled completely, and the trades are emitted in queue order (one per
// passive order, orderbook.rs:843-870).  All 32 lanes call it; the book state lives in lane 0 (`s` is only meaningful
// there).  Returns the aggressor's remaining volume in every lane.
__device__ __forceinline__ u32 dp_sweep_warp(const DeepReg& r, DeepSt& s, u32 lane, u32 side, u32 price, u32 rem, u32 id, u64 t) {
    const u32 opp = side ^ 1u;
    for (u32 guard = 0; guard < (1u << 22); ++guard) {
        u32 go = 0u, bq = 0u;
        if (lane == 0u && rem > 0u && dp_has_best(s, opp)) {
            bq = dp_best_q(s, opp);
            const u32 bprice = r.win_lo + bq;
            go = (side ? (price < bprice) : (price > bprice)) ? 0u : 1u;
        }
        go = __shfl_sync(BB_FULL, go, 0);
        if (!go) break;
        bq = __shfl_sync(BB_FULL, bq, 0);
        const u32 la = r.lv + 16u * bq;
        const uint4 lv = lds128(la);
        const u32 c = lv.z >> 5, idx = lv.z & 31u, tc = lv.w >> 5, tidx = lv.w & 31u;
        const u32 end = (c == tc) ? tidx : DP_CHUNK_ENTRIES;
        const u32 slot = c & (DP_NC - 1u);
        const u32 ca = r.cdat + DP_CHUNK_BYTES * slot;
        if (lds(r.ctag + 4u * slot) != c) {  // chunk not resident: one coalesced 256-byte load
            __syncwarp();
            const u64 v = ldg64_cg(r.chunks + (u64)c * DP_CHUNK_BYTES + 8u * lane);
            sts64(ca + 8u * lane, v);
            if (lane == 0u) sts(r.ctag + 4u * slot, c);
            __syncwarp();
        }
        const u64 e = lds64(ca + 8u * lane);
        const u32 pid = (u32)e, pvol = (u32)(e >> 32);
        const bool live = lane >= idx && lane < end && pid != BB_NIL;
        const u32 v = live ? pvol : 0u;
        u32 incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 y = __shfl_up_sync(BB_FULL, incl, d);
            if (lane >= (u32)d) incl += y;
        }
        const u32 excl = incl - v;
        const bool reached = live && excl < rem;
        const bool full = reached && incl <= rem;
        const u32 mr = __ballot_sync(BB_FULL, reached), mf = __ballot_sync(BB_FULL, full);
        const u32 mp = mr & ~mf;
        const u32 nr = __popc(mr);
        const u32 total = __shfl_sync(BB_FULL, incl, 31);
        const u32 traded = min(rem, total);
        u32 base = 0u, ok = 1u;
        if (lane == 0u) {
            ok = dp_ret_space(r, s, nr) ? 1u : 0u;
            base = s.ret_tail;
        }
        ok = __shfl_sync(BB_FULL, ok, 0);
        if (!ok) break;
        base = __shfl_sync(BB_FULL, base, 0);
        if (reached) {
            const u32 tv = full ? pvol : rem - excl;
            const u32 pv = pvol - tv;
            // trade: side / price are the passive order's (orderbook.rs:853-862)
            dp_ret_write(r, base + __popc(mr & ((1u << lane) - 1u)), make_uint4(RK_FILL | (opp << 8) | (pv == 0u ? 0x10000u : 0u), pid, tv, pv),
                         make_uint4((u32)t, (u32)(t >> 32), r.win_lo + bq, id));
            dp_mark_dirty(r, pid, base + nr);
            if (!full) {  // the partially filled order stays at the head of its level
                sts(ca + 8u * lane + 4u, pv);
                stg32(r.chunks + (u64)c * DP_CHUNK_BYTES + 8u * lane + 4u, pv);
            }
        }
        rem -= traded;
        const u32 cnt = lv.y - __popc(mf), lvol = lv.x - traded;
        __syncwarp();
        if (lane == 0u) {
            s.ret_tail = base + nr;
            st_rel(r.ctl + CT_RET_TAIL, s.ret_tail);
            s.trade_vol += traded;
            s.n_trades += nr;
            dp_add_side(s, opp, 0u - traded);
            if (cnt == 0u) {  // the level is gone
                if (c == tc) dp_free_chunk(r, s, c);
                dp_level_gone(r, s, opp, bq);
            } else if (mp) {
                sts128(la, make_uint4(lvol, cnt, (c << 5) | ((u32)__ffs(mp) - 1u), lv.w));
            } else if (rem == 0u) {  // the aggressor ended exactly on an order boundary
                sts128(la, make_uint4(lvol, cnt, (c << 5) | (32u - (u32)__clz(mf)), lv.w));
            } else if (c != tc) {  // this chunk is used up: follow the link
                const u32 nc = lds(ca + 8u * DP_CHUNK_ENTRIES);
                dp_free_chunk(r, s, c);
                if (nc < r.n_chunks) {
                    sts128(la, make_uint4(lvol, cnt, nc << 5, lv.w));
                } else {  // broken chain: only after an earlier capacity error
                    s.err |= ERR_CAP_PAGES;
                    dp_level_gone(r, s, opp, bq);
                }
            } else {  // live orders counted but none found: only after an earlier capacity error
                s.err |= ERR_CAP_PAGES;
                dp_level_gone(r, s, opp, bq);
            }
        }
        __syncwarp();
    }
    return __shfl_sync(BB_FULL, rem, 0);
}

// (vol, count) of `side` at an arbitrary price: per-lane
__device__ __forceinline__ void dp_level_at(const DeepReg& r, u32 side, u32 price, u32* vol, u32* cnt) {
    *vol = 0;
    *cnt = 0;
    const u32 q = price - r.win_lo;
    if (q >= r.W) return;
    if (!((lds(dp_bm(r, side, q >> 5)) >> (q & 31u)) & 1u)) return;
    const u64 lv = lds64(r.lv + 16u * q);
    *vol = (u32)lv;
    *cnt = (u32)(lv >> 32);
}

// observation words of the book: lane l owns words l and l + 32 (layout: book_obs in book.cuh)
__device__ __forceinline__ void dp_obs(const DeepReg& r, u32 tick, u32 lane, u32 trade_vol, u32 bid, u32 ask, u32 vol_ask, u32 vol_bid,
                                       u32* w0, u32* w1) {
    u32 out[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const u32 w = lane + 32u * half;
        u32 val = 0;
        if (w >= 5u && w < 45u) {
            const u32 i = (w - 5u) >> 2, f = (w - 5u) & 3u;
            u32 v, n;
            if (f < 2u) dp_level_at(r, 1u, bid - i * tick, &v, &n);
            else dp_level_at(r, 0u, ask + i * tick, &v, &n);
            val = (f & 1u) ? n : v;
        } else if (w < 5u) {
            val = w == 0 ? trade_vol : w == 1 ? bid : w == 2 ? ask : w == 3 ? vol_ask : vol_bid;
        }
        out[half] = val;
    }
    *w0 = out[0];
    *w1 = out[1];
}

}  // namespace bb
