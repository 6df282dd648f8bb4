// Deep-book engine, batch-parallel book warp (included by kernels.cuh after deep.cuh; kernel k_deepw).
//
// Same data structures and reference semantics as deep.cuh (side.rs:36-143, orderbook.rs:429-772): dense tick-indexed level
// arrays + per-side bitmaps in shared memory, chunked array queues in HBM, fetch warp in front, retire warp behind.  What
// changes is how the book itself advances.  deep.cuh runs one event at a time on lane 0 of two warps (ladder, queue):
// ~200 dependent instructions each per event, 1250 cycles per event on a B200 SM (profiles/r02_summary.md) — a single
// thread of a GPU is a slow CPU.  Here ONE warp owns the book and takes the events of a batch (32) ONE LANE EACH:
//
//   classify   every lane decides, under the hypothesis that best bid / best ask stay where they are, what its event does:
//              rest on a level behind the touch (ADD), leave a level (REM: cancel, or the first half of a replace), shrink in
//              place (RED), or trade against the touch level of the other side (X) without emptying it.  Anything else —
//              an event that moves the touch or empties a level, a zero-volume order, a doubtful prefetched record, two
//              events naming one order, a market-data record, trading switched off — is COMPLEX;
//   parallel   the events before the first complex one are applied together: level volumes / counts with shared-memory
//              atomics, queue positions of the ADDs by __match_any_sync over the level index, and the trades of all X
//              events by a warp prefix sum over the resting volumes of the touch level's head chunk, in FIFO order, one
//              trade per passive order (orderbook.rs:429-454, 843-870);
//   serial     the complex event runs through the whole reference algorithm warp-uniformly (all lanes the same values,
//              lane 0 stores), after which the rest of the batch is classified again against the new touch.
//
// Trades go to the trade log straight from the book warp (their order in the log is the event order, which only this
// warp knows); order-record updates go through the retire ring as before.
#pragma once

namespace bb {

#define DW_RB 4u        // event-ring depth in batches of 32
#define DW_RCAP 256u    // retire-ring entries
#define DW_NC 32u       // chunk cache entries (direct mapped by chunk id)
#define DW_DIRTY 4096u  // touched-order filter buckets (by order id)

__device__ __forceinline__ void reds_add(u32 a, u32 v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ u32 atoms_or(u32 a, u32 v) {
    u32 o;
    asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory");
    return o;
}

struct BkReg {  // launch-invariant addresses and limits (pinned in registers)
    u32 lvol, lcnt, lht, bma, bmb, sma, smb, ctag, cdat, ret, dirty, ctl, fs;
    u32 win_lo, W, max_orders, n_chunks, max_trades;
    u64 oh, chunks, tr;
};
struct BkSt {  // the book's scalar state: warp-uniform (every lane holds the same values)
    u64 t, max_key_time;
    u32 n_orders, trade_vol, vol_ask, vol_bid, bq_ask, bq_bid, flags, err;
    u32 d_instr, d_applied, zv;
    u32 bump, n_free;       // chunk allocator
    u32 n_tr;               // next trade-log index
    u32 ret_tail, ret_room, ret_pub;
#ifdef DP_PROF
    u32 pf_pass, pf_par, pf_ser, pf_reason[8];
#endif
};
__device__ __forceinline__ u32 bk_bm(const BkReg& r, u32 side, u32 w) { return (side ? r.bmb : r.bma) + 4u * w; }
__device__ __forceinline__ u32 bk_sm(const BkReg& r, u32 side, u32 w) { return (side ? r.smb : r.sma) + 4u * w; }
__device__ __forceinline__ bool bk_has_best(const BkSt& s, u32 side) { return (s.flags >> (1u + side)) & 1u; }
__device__ __forceinline__ u32 bk_best_q(const BkSt& s, u32 side) { return side ? s.bq_bid : s.bq_ask; }
__device__ __forceinline__ void bk_add_side(BkSt& s, u32 side, u32 dv) {
    if (side) s.vol_bid += dv; else s.vol_ask += dv;
}

// lane 0 polls a control word until cond() holds; the outcome is broadcast so that the warp stays uniform
template <class F> __device__ __forceinline__ bool bk_wait(const BkReg& r, u32 lane, F cond, int prof_slot) {
    u32 ok = 1u;
    if (lane == 0u) ok = dp_wait(r.ctl, cond, prof_slot) ? 1u : 0u;
    return __shfl_sync(BB_FULL, ok, 0) != 0u;
}

// ---- retire ring ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool bk_ret_space(const BkReg& r, BkSt& s, u32 lane, u32 n) {
    if (s.ret_tail + n <= s.ret_room) return true;
    u32 room = 0u;
    __syncwarp();  // (the other lanes' ring entries are ordered before lane 0's release)
    if (lane == 0u && s.ret_pub != s.ret_tail) st_rel(r.ctl + CT_RET_TAIL, s.ret_tail);  // what waits to be drained must be visible
    s.ret_pub = s.ret_tail;
    const u32 tail = s.ret_tail;
    const bool ok = bk_wait(r, lane, [&] {
        room = ld_acq(r.ctl + CT_RET_DONE) + DW_RCAP;
        return tail + n <= room;
    }, 8);
    s.ret_room = __shfl_sync(BB_FULL, room, 0);
    return ok;
}
__device__ __forceinline__ void bk_ret_write(const BkReg& r, u32 idx, uint4 a, uint4 b) {
    const u32 ea = r.ret + DP_RENT * (idx & (DW_RCAP - 1u));
    sts128(ea, a);
    sts128(ea + 16u, b);
}
// one entry, written by lane 0
__device__ __forceinline__ bool bk_ret1(const BkReg& r, BkSt& s, u32 lane, uint4 a, uint4 b) {
    if (!bk_ret_space(r, s, lane, 1u)) return false;
    if (lane == 0u) bk_ret_write(r, s.ret_tail, a, b);
    s.ret_tail += 1;
    return true;
}
// everything produced so far becomes visible to the retire warp; `ev_done` events are complete
__device__ __forceinline__ void bk_publish(const BkReg& r, BkSt& s, u32 lane, u32 ev_done) {
    __syncwarp();
    if (lane == 0u) {
        if (s.ret_pub != s.ret_tail) st_rel(r.ctl + CT_RET_TAIL, s.ret_tail);
        st_rel(r.ctl + CT_Q_EV, ev_done);
    }
    s.ret_pub = s.ret_tail;
}

// ---- trade log (types.rs:105-118), written by the lane that found the fill ---------------------------------------------
__device__ __forceinline__ void bk_trade(const BkReg& r, BkSt& s, u32 ti, u32 t_lo, u32 t_hi, u32 price, u32 vol, u32 active, u32 passive,
                                         u32 passive_bid, u32& err) {
    if (ti < r.max_trades) {
        const u64 ta = r.tr + (u64)ti * 32u;
        stg128_cs(ta, t_lo, t_hi, price, vol);
        stg128_cs(ta + 16u, active, passive, passive_bid, 0u);
    } else if (r.max_trades) {
        err |= ERR_CAP_TRADES;
    }
}

// ---- chunk pool (warp-uniform; lane 0 stores) --------------------------------------------------------------------------
__device__ __forceinline__ void bk_chunk_st32(const BkReg& r, u32 c, u32 byte_off, u32 v) {  // write-through store, ONE lane
    stg32(r.chunks + (u64)c * DP_CHUNK_BYTES + byte_off, v);
    const u32 slot = c & (DW_NC - 1u);
    if (lds(r.ctag + 4u * slot) == c) sts(r.cdat + DP_CHUNK_BYTES * slot + byte_off, v);
}
__device__ __forceinline__ u32 bk_alloc(const BkReg& r, BkSt& s, u32 lane) {
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
    if (lane == 0u) {
        const u32 slot = c & (DW_NC - 1u);
        sts(r.ctag + 4u * slot, c);
        sts(r.cdat + DP_CHUNK_BYTES * slot + 8u * DP_CHUNK_ENTRIES, BB_NIL);
        stg32(r.chunks + (u64)c * DP_CHUNK_BYTES + 8u * DP_CHUNK_ENTRIES, BB_NIL);
    }
    __syncwarp();
    return c;
}
__device__ __forceinline__ void bk_free_chunk(const BkReg& r, BkSt& s, u32 lane, u32 c) {
    if (c != 0u && s.n_free < DP_FS_CAP) {
        if (lane == 0u) sts(r.fs + 4u * s.n_free, c);
        s.n_free += 1;
        __syncwarp();
    }
}

// level q of `side` just became empty (its bitmap bit is still set): clear it and, if it was the touch, find the next one
__device__ __forceinline__ void bk_level_gone(const BkReg& r, BkSt& s, u32 lane, u32 side, u32 q) {
    const u32 w = q >> 5;
    const u32 ba = bk_bm(r, side, w);
    const u32 m = lds(ba) & ~(1u << (q & 31u));
    const u32 sa = bk_sm(r, side, w >> 5);
    const u32 sv = lds(sa);
    __syncwarp();
    if (lane == 0u) {
        sts(ba, m);
        if (m == 0u) sts(sa, sv & ~(1u << (w & 31u)));
    }
    __syncwarp();
    if (!(bk_has_best(s, side) && bk_best_q(s, side) == q)) return;
    // q was the best level: every other level of this side lies on the far side of it
    if (side == 0u) {
        if (m) { s.bq_ask = (w << 5) + (u32)__ffs(m) - 1u; return; }
        u32 sw = w >> 5;
        u32 ms = lds(bk_sm(r, 0u, sw)) & ~((2u << (w & 31u)) - 1u);  // summary bits above word w
        while (ms == 0u) {
            if (++sw >= DP_NS) { s.flags &= ~FL_HAS_ASK; return; }
            ms = lds(bk_sm(r, 0u, sw));
        }
        const u32 w2 = (sw << 5) + (u32)__ffs(ms) - 1u;
        s.bq_ask = (w2 << 5) + (u32)__ffs(lds(bk_bm(r, 0u, w2))) - 1u;
    } else {
        if (m) { s.bq_bid = (w << 5) + 31u - (u32)__clz(m); return; }
        u32 sw = w >> 5;
        u32 ms = lds(bk_sm(r, 1u, sw)) & ((1u << (w & 31u)) - 1u);  // summary bits below word w
        while (ms == 0u) {
            if (sw == 0u) { s.flags &= ~FL_HAS_BID; return; }
            --sw;
            ms = lds(bk_sm(r, 1u, sw));
        }
        const u32 w2 = (sw << 5) + 31u - (u32)__clz(ms);
        s.bq_bid = (w2 << 5) + 31u - (u32)__clz(lds(bk_bm(r, 1u, w2)));
    }
}

// insert_order's queue half (side.rs:54-66), one order, warp-uniform: append to the level's chunk chain -> entry position
__device__ __forceinline__ u32 bk_append(const BkReg& r, BkSt& s, u32 lane, u32 q, u32 id, u32 vol) {
    const u32 cnt = lds(r.lcnt + 4u * q);
    const u32 tail = lds(r.lht + 8u * q + 4u);
    __syncwarp();
    u32 pos;
    if (cnt == 0u) {
        const u32 c = bk_alloc(r, s, lane);
        pos = c << 5;
        if (lane == 0u) sts64(r.lht + 8u * q, ((u64)(pos + 1u) << 32) | pos);
    } else {
        u32 c = tail >> 5, idx = tail & 31u;
        if (idx == DP_CHUNK_ENTRIES) {  // tail chunk full: link a new one
            const u32 c2 = bk_alloc(r, s, lane);
            if (lane == 0u) bk_chunk_st32(r, c, 8u * DP_CHUNK_ENTRIES, c2);
            c = c2;
            idx = 0u;
        }
        pos = (c << 5) | idx;
        if (lane == 0u) sts(r.lht + 8u * q + 4u, pos + 1u);
    }
    if (lane == 0u) {
        sts(r.lcnt + 4u * q, cnt + 1u);
        const u32 c = pos >> 5, idx = pos & 31u, slot = c & (DW_NC - 1u);
        stg64(r.chunks + (u64)c * DP_CHUNK_BYTES + 8u * idx, ((u64)vol << 32) | id);
        if (lds(r.ctag + 4u * slot) == c) sts64(r.cdat + DP_CHUNK_BYTES * slot + 8u * idx, ((u64)vol << 32) | id);
    }
    __syncwarp();
    return pos;
}
// remove_order's queue half (side.rs:75-84): tombstone the entry
__device__ __forceinline__ void bk_remove(const BkReg& r, BkSt& s, u32 lane, u32 q, u32 pos) {
    const u32 cnt = lds(r.lcnt + 4u * q);
    const u64 ht = lds64(r.lht + 8u * q);
    __syncwarp();
    if (lane == 0u) {
        bk_chunk_st32(r, pos >> 5, 8u * (pos & 31u), BB_NIL);
        sts(r.lcnt + 4u * q, cnt <= 1u ? 0u : cnt - 1u);
    }
    if (cnt <= 1u && ((u32)ht >> 5) == ((u32)(ht >> 32) >> 5)) bk_free_chunk(r, s, lane, (u32)ht >> 5);  // a longer all-dead chain is left to the pool
    __syncwarp();
}

// ---- match_orders over one level (orderbook.rs:843-870), the whole warp: a prefix sum over the FIFO ------------------------
// Takes `take` volume from the head of level q's queue — and, with `exhaust`, every order left there afterwards (the
// aggressor still had volume, so it also trades, at volume 0, with the zero-volume orders behind: `while order.vol > 0`,
// orderbook.rs:436).  Lanes load consecutive queue entries of the head chunk, an inclusive scan over the resting volumes
// tells every lane whether its order is reached and whether it is filled completely; one trade per passive order, in queue
// order.  `mark`: the value the touched orders' filter buckets get.
__device__ __forceinline__ bool bk_sweep(const BkReg& r, BkSt& s, u32 lane, u32 q, u32 opp, u32 take, bool exhaust, u32 id, u32 t_lo, u32 t_hi,
                                         u32 price, u32 mark, u32& lane_err) {
    for (u32 guard = 0; guard < (1u << 22); ++guard) {
        const u32 cnt0 = lds(r.lcnt + 4u * q);
        if (!((take > 0u || exhaust) && cnt0 > 0u)) break;
        const u64 ht = lds64(r.lht + 8u * q);
        const u32 head = (u32)ht, tail = (u32)(ht >> 32);
        const u32 c = head >> 5, idx = head & 31u, tc = tail >> 5;
        const u32 end = (c == tc) ? (tail & 31u) : DP_CHUNK_ENTRIES;
        const u32 slot = c & (DW_NC - 1u);
        const u32 ca = r.cdat + DP_CHUNK_BYTES * slot;
        if (lds(r.ctag + 4u * slot) != c) {  // chunk not resident: one coalesced 256-byte load
            __syncwarp();
            const u64 v = ldg64_cg(r.chunks + (u64)c * DP_CHUNK_BYTES + 8u * lane);
            sts64(ca + 8u * lane, v);
            if (lane == 0u) sts(r.ctag + 4u * slot, c);
            __syncwarp();
        }
        const u64 e = lds64(ca + 8u * lane);
        const u32 nc = lds(ca + 8u * DP_CHUNK_ENTRIES);
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
        const bool reached = live && (excl < take || exhaust);
        const bool full = reached && incl <= take;
        const u32 mr = __ballot_sync(BB_FULL, reached), mf = __ballot_sync(BB_FULL, full);
        const u32 mp = mr & ~mf;
        const u32 nr = __popc(mr);
        const u32 total = __shfl_sync(BB_FULL, incl, 31);
        const u32 traded = min(take, total);
        if (!bk_ret_space(r, s, lane, nr)) return false;
        if (reached) {
            const u32 tv = full ? pvol : take - excl;
            const u32 pv = pvol - tv;
            const u32 k = __popc(mr & ((1u << lane) - 1u));
            // trade: side / price are the passive order's (orderbook.rs:853-862)
            bk_ret_write(r, s.ret_tail + k, make_uint4(RK_FILL | (opp << 8) | (pv == 0u ? 0x10000u : 0u), pid, tv, pv), make_uint4(t_lo, t_hi, price, id));
            bk_trade(r, s, s.n_tr + k, t_lo, t_hi, price, tv, id, pid, opp, lane_err);
            sts(r.dirty + 4u * (pid & (DW_DIRTY - 1u)), mark);
            if (!full) {  // the partially filled order stays at the head of its level
                sts(ca + 8u * lane + 4u, pv);
                stg32(r.chunks + (u64)c * DP_CHUNK_BYTES + 8u * lane + 4u, pv);
            }
        }
        take -= traded;
        s.ret_tail += nr;
        s.n_tr += nr;
        const u32 cnt = cnt0 - __popc(mf);
        __syncwarp();
        bool done = false;
        if (cnt == 0u) {  // the level is gone
            if (c == tc) bk_free_chunk(r, s, lane, c);
            if (lane == 0u) sts(r.lcnt + 4u * q, 0u);
        } else {
            u32 nh = 0u;
            bool bad = false;
            if (mp) {
                nh = (c << 5) | ((u32)__ffs(mp) - 1u);
                done = true;
            } else if (take == 0u && !exhaust) {  // ended exactly on an order boundary
                nh = (c << 5) | (mf ? 32u - (u32)__clz(mf) : idx);
                done = true;
            } else if (c != tc) {  // this chunk is used up: follow the link
                bk_free_chunk(r, s, lane, c);
                if (nc < r.n_chunks) nh = nc << 5; else bad = true;  // broken chain: only after an earlier capacity error
            } else {  // live orders counted but none found: only after an earlier capacity error
                bad = true;
            }
            if (bad) s.err |= ERR_CAP_PAGES;
            if (lane == 0u) {
                sts(r.lcnt + 4u * q, bad ? 0u : cnt);
                if (!bad) sts(r.lht + 8u * q, nh);
            }
        }
        __syncwarp();
        if (done) break;
    }
    return true;
}

// (vol, count) of `side` at an arbitrary price: per-lane
__device__ __forceinline__ void bk_level_at(const BkReg& r, u32 side, u32 price, u32* vol, u32* cnt) {
    *vol = 0;
    *cnt = 0;
    const u32 q = price - r.win_lo;
    if (q >= r.W) return;
    if (!((lds(bk_bm(r, side, q >> 5)) >> (q & 31u)) & 1u)) return;
    *vol = lds(r.lvol + 4u * q);
    *cnt = lds(r.lcnt + 4u * q);
}
// observation words of the book: lane l owns words l and l + 32 (layout: book_obs in book.cuh)
__device__ __forceinline__ void bk_obs(const BkReg& r, u32 tick, u32 lane, u32 trade_vol, u32 bid, u32 ask, u32 vol_ask, u32 vol_bid, u32* w0,
                                       u32* w1) {
    u32 out[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const u32 w = lane + 32u * half;
        u32 val = 0;
        if (w >= 5u && w < 45u) {
            const u32 i = (w - 5u) >> 2, f = (w - 5u) & 3u;
            u32 v, n;
            if (f < 2u) bk_level_at(r, 1u, bid - i * tick, &v, &n);
            else bk_level_at(r, 0u, ask + i * tick, &v, &n);
            val = (f & 1u) ? n : v;
        } else if (w < 5u) {
            val = w == 0 ? trade_vol : w == 1 ? bid : w == 2 ? ask : w == 3 ? vol_ask : vol_bid;
        }
        out[half] = val;
    }
    *w0 = out[0];
    *w1 = out[1];
}

// =====================================================================================================================
// One event through the whole reference algorithm, warp-uniformly (x, y: the pre-decoded instruction; a, c: the prefetched
// order record of a cancel / modify).  `ev_done`: events complete once this one is; `rf`: events whose record writes were
// in HBM when the record was fetched.  Returns false when a bounded wait ran out.
__device__ __forceinline__ bool bk_serial(const BkReg& r, BkSt& s, u32 lane, uint4 x, uint4 y, uint4 a, uint4 c, u32 ev_done, u32 rf, u32& lane_err) {
    s.t = ((u64)x.y << 32) | x.x;
    const u32 op = x.z & BB_OP_MASK;
    // the order that goes through matching (NEW, or the second half of a replace)
    u32 kind = 0u, side = 0u, price = 0u, vol = 0u, id = 0u;
    bool market = false;
    if (op == BB_OP_NEW) {  // (id, sentinel price and the market flag were settled by the fetch warp)
        s.d_instr += 1;
        if (x.z & DPF_CAP_ORDERS) {
            s.err |= ERR_CAP_ORDERS;
        } else {
            id = x.w;
            s.n_orders = id + 1u;
            kind = RK_NEW;
            side = (x.z >> 8) & 1u;  // BB_F_BID
            price = y.x;
            market = (x.z & DPF_MARKET) != 0u;
            vol = y.y;
        }
    } else if (op == BB_OP_CANCEL || op == BB_OP_MODIFY) {
        s.d_instr += 1;
        id = x.w;
        if (id >= s.n_orders || id >= r.max_orders) {
            s.err |= ERR_BAD_ID;  // the reference panics (orderbook.rs:642, :749)
        } else {
            // The order's record as the fetch warp saw it: usable iff every write to it was in HBM by then, i.e. no event that
            // touched the order (its own, or a sweep that filled it) was still in the pipeline.
            const u32 dv = lds(r.dirty + 4u * (id & (DW_DIRTY - 1u)));
            if (dv > rf) {  // doubtful: let the pipeline drain up to that event, read again
                bk_publish(r, s, lane, ev_done - 1u);
                if (!bk_wait(r, lane, [&] { return ld_acq(r.ctl + CT_EV_RETIRED) >= dv; }, 13)) return false;
                a = ldg128_cg(r.oh + (u64)id * ORD_STRIDE);
                c = ldg128_cg(r.oh + (u64)id * ORD_STRIDE + 16u);
            }
            if ((c.z & META_STATUS_MASK) == ST_ACTIVE) {
                const u32 oside = (c.z & META_BID) ? 1u : 0u;
                const bool has_p = (x.z & BB_F_HAS_PRICE) != 0u, has_v = (x.z & BB_F_HAS_VOL) != 0u;
                const u32 q = a.x - r.win_lo;
                const bool reduce = op == BB_OP_MODIFY && !has_p && has_v && y.y < a.y;
                if (op == BB_OP_MODIFY && !has_p && !has_v) {
                } else if (q >= r.W) {  // never rested (flagged when it was placed)
                } else if (reduce) {  // reduce in place: priority kept (orderbook.rs:755-757)
                    const u32 la = r.lvol + 4u * q;
                    const u32 lv = lds(la);
                    __syncwarp();
                    if (lane == 0u) {
                        sts(la, lv - (a.y - y.y));
                        bk_chunk_st32(r, a.z >> 5, 8u * (a.z & 31u) + 4u, y.y);
                    }
                    bk_add_side(s, oside, y.y - a.y);
                    if (y.y == 0u) s.zv = 1u;
                    if (!bk_ret1(r, s, lane, make_uint4(RK_REDUCE, id, y.y, 0u), make_uint4(x.x, x.y, 0u, 0u))) return false;
                    if (lane == 0u) sts(r.dirty + 4u * (id & (DW_DIRTY - 1u)), ev_done);
                    s.d_applied += 1;
                } else {  // cancel_order (orderbook.rs:622-644), or the remove half of replace_order (:679-723)
                    const bool cancel = op == BB_OP_CANCEL;
                    const u32 la = r.lvol + 4u * q;
                    const u32 nv = lds(la) - a.y;
                    __syncwarp();
                    if (lane == 0u) sts(la, nv);
                    bk_add_side(s, oside, 0u - a.y);
                    bk_remove(r, s, lane, q, a.z);
                    if (cancel && !bk_ret1(r, s, lane, make_uint4(RK_CANCEL | (oside << 8), id, 0u, 0u), make_uint4(x.x, x.y, 0u, 0u))) return false;
                    if (nv == 0u && (!s.zv || lds(r.lcnt + 4u * q) == 0u)) bk_level_gone(r, s, lane, oside, q);
                    if (lane == 0u) sts(r.dirty + 4u * (id & (DW_DIRTY - 1u)), ev_done);
                    if (cancel) {
                        s.d_applied += 1;
                    } else {  // never a market order (N4)
                        kind = RK_REPLACE;
                        side = oside;
                        price = has_p ? y.x : a.x;
                        vol = has_v ? y.y : a.y;
                    }
                }
            }
        }
    } else if (op == BB_OP_SET_TRADING) {
        s.flags = y.y ? (s.flags | FL_TRADING) : (s.flags & ~FL_TRADING);
    } else if (op == BB_OP_RESTORE) {
        s.err |= ERR_ROW_OP;  // bb_load_book is not available on the deep engine
    }
    if (kind) {
        // ---- match_bid / match_ask (orderbook.rs:429-487): level by level from the touch
        u32 rem = vol;
        const u32 opp = side ^ 1u;
        if (s.flags & FL_TRADING) {
            while (rem > 0u && bk_has_best(s, opp)) {
                const u32 bq = bk_best_q(s, opp);
                const u32 bprice = r.win_lo + bq;
                if (side ? (price < bprice) : (price > bprice)) break;
                const u32 la = r.lvol + 4u * bq;
                const u32 lv = lds(la);
                const u32 take = min(rem, lv), nv = lv - take;
                rem -= take;
                const bool exhaust = rem > 0u;  // the aggressor goes on: it takes every order left on this level
                __syncwarp();
                if (lane == 0u) sts(la, nv);
                s.trade_vol += take;
                bk_add_side(s, opp, 0u - take);
                if (!bk_sweep(r, s, lane, bq, opp, take, exhaust, id, x.x, x.y, bprice, ev_done, lane_err)) return false;
                if (nv == 0u && (exhaust || !s.zv || lds(r.lcnt + 4u * bq) == 0u)) bk_level_gone(r, s, lane, opp, bq);
            }
        }
        // ---- rest or finish (orderbook.rs:495-531, 699-722)
        const bool filled = vol != 0u && rem == 0u;
        u32 status = filled ? ST_FILLED : market ? ((s.flags & FL_TRADING) ? ST_CANCELLED : ST_REJECTED) : ST_ACTIVE;
        u32 pos = 0u;
        bool rests = false;
        if (status == ST_ACTIVE) {  // insert_order (side.rs:54-66)
            const u32 q = price - r.win_lo;
            if (q >= r.W) {
                s.err |= ERR_CAP_PAGES;
            } else {
                const u32 ba = bk_bm(r, side, q >> 5), bit = 1u << (q & 31u);
                const u32 la = r.lvol + 4u * q;
                const u32 bw = lds(ba);
                const u32 lv = lds(la);
                bool ok = true;
                if (!(bw & bit)) {
                    // the level array is shared by the two sides.  While trading is enabled an order only rests where the other
                    // side is empty, with one exception the reference allows: a ZERO-volume order never matches
                    // (orderbook.rs:436) and rests wherever its price says
                    if (lds(bk_bm(r, opp, q >> 5)) & bit) {
                        s.err |= ERR_LOCKED;
                        ok = false;
                    } else {
                        const u32 sa = bk_sm(r, side, q >> 10);
                        const u32 sv = lds(sa);
                        __syncwarp();
                        if (lane == 0u) {
                            sts(la, rem);
                            sts(ba, bw | bit);
                            if (bw == 0u) sts(sa, sv | (1u << ((q >> 5) & 31u)));
                        }
                        const bool better = !bk_has_best(s, side) || (side ? q > s.bq_bid : q < s.bq_ask);
                        if (better) {
                            if (side) s.bq_bid = q; else s.bq_ask = q;
                            s.flags |= FL_HAS_ASK << side;
                        }
                    }
                } else {
                    if (s.t <= s.max_key_time) s.err |= ERR_TIME_ORDER;
                    __syncwarp();
                    if (lane == 0u) sts(la, lv + rem);
                }
                __syncwarp();
                if (ok) {
                    bk_add_side(s, side, rem);
                    if (s.t > s.max_key_time) s.max_key_time = s.t;
                    if (rem == 0u) s.zv = 1u;
                    pos = bk_append(r, s, lane, q, id, rem);
                    rests = true;
                }
            }
        }
        // the order's record: resting -> position and key time; ended -> final status and end time (an order that could not
        // rest — flagged above — keeps status Active in its record and is not on the book)
        if (!bk_ret1(r, s, lane, make_uint4(kind | (side << 8) | (status << 12), id, rem, rests ? pos : 0u), make_uint4(x.x, x.y, price, 0u)))
            return false;
        if (lane == 0u) sts(r.dirty + 4u * (id & (DW_DIRTY - 1u)), ev_done);
        s.d_applied += 1;
    }
    bk_publish(r, s, lane, ev_done);
    return true;
}

// =====================================================================================================================
// The batch-parallel step.  One side's head chunk while the X events of a pass are applied to it:
struct SweepSide {
    u32 L, c, idx, end, tc, ca, nc;  // level, head chunk, first live index, end index, tail chunk, cache address, link (uniform)
    u32 S, total;                // volume consumed from / resting in this chunk (uniform)
    u32 nfull, taken;            // orders filled completely in chunks already left behind, volume taken in this pass (uniform)
    u32 pid, pin, pex;           // this lane's entry: order id, inclusive / exclusive prefix of the resting volumes
    bool live, ok;
};
__device__ __forceinline__ void bk_side_load(const BkReg& r, u32 lane, SweepSide& d) {
    const u64 ht = lds64(r.lht + 8u * d.L);
    const u32 head = (u32)ht, tail = (u32)(ht >> 32);
    d.c = head >> 5;
    d.idx = head & 31u;
    d.tc = tail >> 5;
    d.end = (d.c == d.tc) ? (tail & 31u) : DP_CHUNK_ENTRIES;
    const u32 slot = d.c & (DW_NC - 1u);
    d.ca = r.cdat + DP_CHUNK_BYTES * slot;
    if (lds(r.ctag + 4u * slot) != d.c) {  // chunk not resident: one coalesced 256-byte load
        __syncwarp();
        const u64 v = ldg64_cg(r.chunks + (u64)d.c * DP_CHUNK_BYTES + 8u * lane);
        sts64(d.ca + 8u * lane, v);
        if (lane == 0u) sts(r.ctag + 4u * slot, d.c);
        __syncwarp();
    }
    // (everything the pass needs from the cached copy is taken now: the other side's chunk may claim the same cache slot)
    const u64 e = lds64(d.ca + 8u * lane);
    d.nc = lds(d.ca + 8u * DP_CHUNK_ENTRIES);
    d.pid = (u32)e;
    const u32 pvol = (u32)(e >> 32);
    d.live = lane >= d.idx && lane < d.end && d.pid != BB_NIL;
    const u32 v = d.live ? pvol : 0u;
    u32 incl = v;
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) {
        const u32 y = __shfl_up_sync(BB_FULL, incl, k);
        if (lane >= (u32)k) incl += y;
    }
    d.pin = incl;
    d.pex = incl - v;
    d.total = __shfl_sync(BB_FULL, incl, 31);
    d.S = 0u;
}
// one aggressor (volume `rem`, all of which the level can give) against the head of side d
__device__ __forceinline__ bool bk_side_take(const BkReg& r, BkSt& s, u32 lane, SweepSide& d, u32 opp, u32 rem, u32 aid, u32 t_lo, u32 t_hi,
                                             u32 mark, u32& lane_err) {
    const u32 price = r.win_lo + d.L;
    d.taken += rem;
    for (u32 guard = 0; guard < (1u << 22); ++guard) {
        const u32 hi = d.S + rem;
        const bool ov = d.live && d.pex < hi && d.pin > d.S;
        const u32 m = __ballot_sync(BB_FULL, ov);
        const u32 nf = __popc(m);
        if (!bk_ret_space(r, s, lane, nf + 1u)) return false;
        if (ov) {
            const u32 top = min(d.pin, hi);
            const u32 tv = top - max(d.pex, d.S), pv = d.pin - top;
            const u32 k = __popc(m & ((1u << lane) - 1u));
            bk_ret_write(r, s.ret_tail + k, make_uint4(RK_FILL | (opp << 8) | (pv == 0u ? 0x10000u : 0u), d.pid, tv, pv),
                         make_uint4(t_lo, t_hi, price, aid));
            bk_trade(r, s, s.n_tr + k, t_lo, t_hi, price, tv, aid, d.pid, opp, lane_err);
            sts(r.dirty + 4u * (d.pid & (DW_DIRTY - 1u)), mark);
        }
        s.ret_tail += nf;
        s.n_tr += nf;
        if (hi <= d.total) {
            d.S = hi;
            return true;
        }
        // the chunk is used up and the aggressor wants more: follow the link (the level holds more than this pass takes)
        rem = hi - d.total;
        const u32 nc = d.nc;
        d.nfull += __popc(__ballot_sync(BB_FULL, d.live));
        if (d.c == d.tc || nc >= r.n_chunks) {  // broken chain: only after an earlier capacity error
            s.err |= ERR_CAP_PAGES;
            d.ok = false;
            return true;
        }
        bk_free_chunk(r, s, lane, d.c);
        __syncwarp();
        if (lane == 0u) sts(r.lht + 8u * d.L, nc << 5);
        __syncwarp();
        bk_side_load(r, lane, d);
    }
    return true;
}
// the pass is over: head, volumes and counts of the level go back to shared memory
__device__ __forceinline__ void bk_side_done(const BkReg& r, u32 lane, SweepSide& d) {
    const u32 fm = __ballot_sync(BB_FULL, d.live && d.pin <= d.S);
    const u32 rm = __ballot_sync(BB_FULL, d.live && d.pin > d.S);
    if (d.live && d.pex < d.S && d.S < d.pin) {  // the partially filled order stays at the head of its level
        bk_chunk_st32(r, d.c, 8u * lane + 4u, d.pin - d.S);
    }
    const u32 lc = lds(r.lcnt + 4u * d.L), lv = lds(r.lvol + 4u * d.L);
    __syncwarp();
    if (lane == 0u) {
        if (d.ok) {
            sts(r.lht + 8u * d.L, (d.c << 5) | (rm ? (u32)__ffs(rm) - 1u : d.end));
            sts(r.lcnt + 4u * d.L, lc - (d.nfull + __popc(fm)));
        } else {
            sts(r.lcnt + 4u * d.L, 0u);
        }
        sts(r.lvol + 4u * d.L, lv - d.taken);
    }
    __syncwarp();
}

#define CXR_EMIT 0
#define CXR_STATE 1
#define CXR_TOUCH 2
#define CXR_DOUBT 3
#define CXR_SAMEID 4
#define CXR_EMPTY 5
#define CXR_TAKE 6
#define CXR_OTHER 7

// The events of one batch still to do (`pending`: a contiguous run of lanes; lane i holds event i: x, y pre-decoded
// instruction, a, c prefetched record).  `ev0`: events complete before this batch.  Comes back early (obs_lane < 32) after an
// event that asks for a market-data record.  Returns false when a bounded wait ran out.
__device__ __forceinline__ bool bk_batch(const BkReg& r, BkSt& s, u32 lane, u32& pending, uint4 x, uint4 y, uint4 a, uint4 c, u32 ev0, u32 rf,
                                         u32 fast, u32& lane_err, u32& obs_lane) {
    const u32 lt = (1u << lane) - 1u;
    const u32 op = x.z & BB_OP_MASK;
    const u32 t_lo = x.x, t_hi = x.y;
    const u64 t = ((u64)t_hi << 32) | t_lo;
    obs_lane = 32u;
    while (pending) {
        const u32 first = (u32)__ffs(pending) - 1u;
        const bool valid = (pending >> lane) & 1u;
        // ---- classify under the hypothesis "the touch does not move" ---------------------------------------------------
        bool cx = false;
        u32 why = CXR_OTHER;
        bool do_rem = false, do_red = false, do_place = false, is_x = false, is_add = false;
        u32 q1 = 0u, v1 = 0u, pos1 = 0u, red_vol = 0u, oside = 0u;
        u32 pside = 0u, pprice = 0u, pvol = 0u, pkind = 0u, q2 = 0u;
        const u32 id = x.w;
        const bool is_new = valid && op == BB_OP_NEW, is_cm = valid && (op == BB_OP_CANCEL || op == BB_OP_MODIFY);
        const u32 newm = __ballot_sync(BB_FULL, is_new);
        if (valid) {
            if (!fast || s.zv || !(s.flags & FL_TRADING)) { cx = true; why = CXR_STATE; }
            if (x.z & BB_F_EMIT) { cx = true; why = CXR_EMIT; }
            if (is_new) {
                if (x.z & DPF_CAP_ORDERS) cx = true;
                do_place = true;
                pkind = RK_NEW;
                pside = (x.z >> 8) & 1u;
                pprice = y.x;
                pvol = y.y;
            } else if (is_cm) {
                if (id >= s.n_orders + __popc(newm & lt) || id >= r.max_orders) {
                    cx = true;  // unknown id
                } else {
                    if (lds(r.dirty + 4u * (id & (DW_DIRTY - 1u))) > rf) { cx = true; why = CXR_DOUBT; }
                    const bool has_p = (x.z & BB_F_HAS_PRICE) != 0u, has_v = (x.z & BB_F_HAS_VOL) != 0u;
                    q1 = a.x - r.win_lo;
                    if ((c.z & META_STATUS_MASK) != ST_ACTIVE || (op == BB_OP_MODIFY && !has_p && !has_v) || q1 >= r.W) {
                        // nothing to do (cancel / modify of an order that is not on the book are no-ops)
                    } else {
                        oside = (c.z & META_BID) ? 1u : 0u;
                        v1 = a.y;
                        pos1 = a.z;
                        if (op == BB_OP_MODIFY && !has_p && has_v && y.y < a.y) {
                            do_red = true;
                            red_vol = y.y;
                            if (red_vol == 0u) cx = true;
                        } else {
                            do_rem = true;
                            if (op == BB_OP_MODIFY) {
                                do_place = true;
                                pkind = RK_REPLACE;
                                pside = oside;
                                pprice = has_p ? y.x : a.x;
                                pvol = has_v ? y.y : a.y;
                            }
                        }
                        if (bk_has_best(s, oside) && q1 == bk_best_q(s, oside)) { cx = true; why = CXR_TOUCH; }  // the touch level is being traded
                    }
                }
            } else {
                cx = true;  // SET_TRADING, RESTORE, no-ops: the serial path knows
            }
            if (do_place && !cx) {
                const u32 opp = pside ^ 1u;
                if (pvol == 0u) {
                    cx = true;
                } else {
                    bool crosses = false;
                    if (bk_has_best(s, opp)) {
                        const u32 bprice = r.win_lo + bk_best_q(s, opp);
                        crosses = pside ? (pprice >= bprice) : (pprice <= bprice);
                    }
                    const bool market = pkind == RK_NEW && (x.z & DPF_MARKET) != 0u;
                    if (crosses) {
                        is_x = true;
                    } else if (market) {
                        cx = true;  // a market order that finds no other side
                    } else {
                        q2 = pprice - r.win_lo;
                        if (q2 >= r.W) {
                            cx = true;
                        } else if (!bk_has_best(s, pside) || (pside ? q2 > s.bq_bid : q2 < s.bq_ask)) {
                            cx = true;  // a new best price
                            why = CXR_TOUCH;
                        } else if ((lds(bk_bm(r, opp, q2 >> 5)) >> (q2 & 31u)) & 1u) {
                            cx = true;  // the other side rests there (locked level)
                        } else {
                            is_add = true;
                        }
                    }
                }
            }
        }
        {   // time moves strictly forward through the batch and past every resting order's key
            const u32 pl = __shfl_up_sync(BB_FULL, t_lo, 1), ph = __shfl_up_sync(BB_FULL, t_hi, 1);
            const u64 tp = lane == first ? s.max_key_time : (((u64)ph << 32) | pl);
            if (valid && t <= tp) cx = true;
        }
        // two events of the batch naming one order: the later one cannot trust its prefetched record
        {
            const u32 key = (is_new || is_cm) ? id : (0xFFFFFF00u | lane);
            const u32 mg = __match_any_sync(BB_FULL, key);
            if ((is_new || is_cm) && (mg & lt)) { cx = true; why = CXR_SAMEID; }
        }
        // a level must not run empty inside a pass (its chunk chain and bitmap bit would change hands)
        {
            const u32 key = (valid && do_rem) ? q1 : (0xFFFFFF00u | lane);
            const u32 mg = __match_any_sync(BB_FULL, key);
            if (valid && do_rem && lds(r.lcnt + 4u * q1) <= (u32)__popc(mg & lt) + 1u) { cx = true; why = CXR_EMPTY; }
        }
        // the touch level must hold more than the pass takes from it
        {
            const u32 ta = (valid && is_x && pside == 1u) ? pvol : 0u, tb = (valid && is_x && pside == 0u) ? pvol : 0u;
            u32 ca_ = ta, cb_ = tb;
#pragma unroll
            for (int k = 1; k < 32; k <<= 1) {
                const u32 ya = __shfl_up_sync(BB_FULL, ca_, k), yb = __shfl_up_sync(BB_FULL, cb_, k);
                if (lane >= (u32)k) { ca_ += ya; cb_ += yb; }
            }
            if (valid && is_x) {
                const u32 avail = lds(r.lvol + 4u * bk_best_q(s, pside ^ 1u));
                if ((pside ? ca_ : cb_) >= avail) { cx = true; why = CXR_TAKE; }
            }
        }
        const u32 cm = __ballot_sync(BB_FULL, cx && valid);
        const u32 clean = cm ? (pending & (((u32)1u << ((u32)__ffs(cm) - 1u)) - 1u)) : pending;
#ifdef DP_PROF
        s.pf_pass += 1;
        s.pf_par += __popc(clean);
#endif
        if (clean) {
            const bool act = (clean >> lane) & 1u;
            const u32 last = 31u - (u32)__clz(clean);
            const u32 mark = ev0 + last + 1u;
            // ---- X: the trades, aggressors in event order, each against the head of its touch level ----------------------
            u32 xm = __ballot_sync(BB_FULL, act && is_x);
            if (xm) {
                SweepSide sa, sb;  // ask side (taken by bids), bid side (taken by asks)
                sa.L = s.bq_ask; sb.L = s.bq_bid;
                sa.nfull = sb.nfull = sa.taken = sb.taken = 0u;
                sa.ok = sb.ok = true;
                const u32 xa = __ballot_sync(BB_FULL, act && is_x && pside == 1u);
                if (xa) bk_side_load(r, lane, sa);
                if (xm & ~xa) bk_side_load(r, lane, sb);
                while (xm) {
                    const u32 k = (u32)__ffs(xm) - 1u;
                    xm &= xm - 1u;
                    const u32 rem = __shfl_sync(BB_FULL, pvol, k), aid = __shfl_sync(BB_FULL, id, k);
                    const u32 kl = __shfl_sync(BB_FULL, t_lo, k), kh = __shfl_sync(BB_FULL, t_hi, k);
                    bool okk;
                    if ((xa >> k) & 1u) okk = sa.ok ? bk_side_take(r, s, lane, sa, 0u, rem, aid, kl, kh, mark, lane_err) : true;
                    else okk = sb.ok ? bk_side_take(r, s, lane, sb, 1u, rem, aid, kl, kh, mark, lane_err) : true;
                    if (!okk || !bk_ret_space(r, s, lane, 1u)) return false;
                    // the aggressor's own record: Filled
                    if (lane == k)
                        bk_ret_write(r, s.ret_tail, make_uint4(pkind | (pside << 8) | (ST_FILLED << 12), id, 0u, 0u), make_uint4(t_lo, t_hi, pprice, 0u));
                    s.ret_tail += 1;
                }
                if (xa) {
                    bk_side_done(r, lane, sa);
                    s.trade_vol += sa.taken;
                    s.vol_ask -= sa.taken;
                }
                if (sb.taken || !sb.ok) {
                    bk_side_done(r, lane, sb);
                    s.trade_vol += sb.taken;
                    s.vol_bid -= sb.taken;
                }
            }
            // ---- ADD: queue positions, one group per level ------------------------------------------------------------------
            const bool addl = act && is_add, reml = act && do_rem, redl = act && do_red;
            u32 pos2 = 0u;
            if (__ballot_sync(BB_FULL, addl)) {
                const u32 mg = __match_any_sync(BB_FULL, addl ? q2 : (0xFFFFFF00u | lane));
                const u32 leader = (u32)__ffs(mg) - 1u, rank = __popc(mg & lt), gsize = __popc(mg);
                const bool lead = addl && lane == leader;
                u32 cnt0 = 1u, tail = 0u;
                if (addl) {
                    cnt0 = lds(r.lcnt + 4u * q2);
                    tail = lds(r.lht + 8u * q2 + 4u);
                }
                const bool empty = cnt0 == 0u;
                const u32 tidx = empty ? 0u : (tail & 31u);
                const u32 g = tidx + rank, kk = g / DP_CHUNK_ENTRIES, idx = g - DP_CHUNK_ENTRIES * kk;
                const u32 glast = tidx + gsize - 1u, kmax = glast / DP_CHUNK_ENTRIES;
                u32 c0 = empty ? 0u : (tail >> 5), c1 = 0u, c2 = 0u;
                // the leaders' new chunks, handed out one at a time (the allocator is warp-uniform state)
                const u32 need = lead ? kmax + (empty ? 1u : 0u) : 0u;
                u32 nm = __ballot_sync(BB_FULL, need > 0u);
                while (nm) {
                    const u32 L = (u32)__ffs(nm) - 1u;
                    nm &= nm - 1u;
                    const u32 n = __shfl_sync(BB_FULL, need, L);
                    const u32 o0 = __shfl_sync(BB_FULL, empty ? 0u : 1u, L);
                    for (u32 j = 0; j < n; ++j) {
                        const u32 cc = bk_alloc(r, s, lane);
                        if (lane == L) {
                            const u32 ord = o0 + j;
                            if (ord == 0u) c0 = cc; else if (ord == 1u) c1 = cc; else c2 = cc;
                        }
                    }
                }
                if (lead) {
                    if (kmax >= 1u) bk_chunk_st32(r, c0, 8u * DP_CHUNK_ENTRIES, c1);
                    if (kmax >= 2u) bk_chunk_st32(r, c1, 8u * DP_CHUNK_ENTRIES, c2);
                }
                const u32 src = addl ? leader : lane;
                const u32 l0 = __shfl_sync(BB_FULL, c0, src), l1 = __shfl_sync(BB_FULL, c1, src), l2 = __shfl_sync(BB_FULL, c2, src);
                if (addl) {
                    const u32 myc = kk == 0u ? l0 : kk == 1u ? l1 : l2;
                    pos2 = (myc << 5) | idx;
                    const u64 ent = ((u64)pvol << 32) | id;
                    stg64(r.chunks + (u64)myc * DP_CHUNK_BYTES + 8u * idx, ent);
                    const u32 slot = myc & (DW_NC - 1u);
                    if (lds(r.ctag + 4u * slot) == myc) sts64(r.cdat + DP_CHUNK_BYTES * slot + 8u * idx, ent);
                }
                if (lead) {
                    const u32 il = glast - DP_CHUNK_ENTRIES * kmax;
                    const u32 cl = kmax == 0u ? c0 : kmax == 1u ? c1 : c2;
                    const u32 ntail = ((cl << 5) | il) + 1u;
                    if (empty) {
                        sts64(r.lht + 8u * q2, ((u64)ntail << 32) | (c0 << 5));
                        sts(r.lvol + 4u * q2, 0u);
                        const u32 w = q2 >> 5;
                        if (atoms_or(bk_bm(r, pside, w), 1u << (q2 & 31u)) == 0u) atoms_or(bk_sm(r, pside, w >> 5), 1u << (w & 31u));
                    } else {
                        sts(r.lht + 8u * q2 + 4u, ntail);
                    }
                }
                __syncwarp();
            }
            // ---- level volumes and counts; tombstones; volume rewrites -------------------------------------------------------
            if (addl) {
                reds_add(r.lvol + 4u * q2, pvol);
                reds_add(r.lcnt + 4u * q2, 1u);
            }
            if (reml) {
                bk_chunk_st32(r, pos1 >> 5, 8u * (pos1 & 31u), BB_NIL);
                reds_add(r.lvol + 4u * q1, 0u - v1);
                reds_add(r.lcnt + 4u * q1, 0xFFFFFFFFu);
            }
            if (redl) {
                bk_chunk_st32(r, pos1 >> 5, 8u * (pos1 & 31u) + 4u, red_vol);
                reds_add(r.lvol + 4u * q1, red_vol - v1);
            }
            // ---- order-record updates (X events wrote theirs above) ----------------------------------------------------------
            const bool ent = addl || (reml && !do_place) || redl;
            const u32 em = __ballot_sync(BB_FULL, ent);
            if (em) {
                if (!bk_ret_space(r, s, lane, __popc(em))) return false;
                if (ent) {
                    uint4 ea, eb = make_uint4(t_lo, t_hi, 0u, 0u);
                    if (addl) {
                        ea = make_uint4(pkind | (pside << 8) | (ST_ACTIVE << 12), id, pvol, pos2);
                        eb.z = pprice;
                    } else if (reml) {
                        ea = make_uint4(RK_CANCEL | (oside << 8), id, 0u, 0u);
                    } else {
                        ea = make_uint4(RK_REDUCE, id, red_vol, 0u);
                    }
                    bk_ret_write(r, s.ret_tail + __popc(em & lt), ea, eb);
                }
                s.ret_tail += __popc(em);
            }
            if (act && (is_new || is_cm)) sts(r.dirty + 4u * (id & (DW_DIRTY - 1u)), mark);
            // ---- the book's scalars ------------------------------------------------------------------------------------------
            const u32 nnew = __popc(__ballot_sync(BB_FULL, act && is_new));
            s.n_orders += nnew;
            s.d_instr += __popc(__ballot_sync(BB_FULL, act && (is_new || is_cm)));
            s.d_applied += __popc(__ballot_sync(BB_FULL, act && (do_place || do_rem || do_red)));
            int da = 0, db = 0;
            if (addl) { if (pside) db += (int)pvol; else da += (int)pvol; }
            if (reml) { if (oside) db -= (int)v1; else da -= (int)v1; }
            if (redl) { if (oside) db -= (int)(v1 - red_vol); else da -= (int)(v1 - red_vol); }
            s.vol_ask += (u32)__reduce_add_sync(BB_FULL, da);
            s.vol_bid += (u32)__reduce_add_sync(BB_FULL, db);
            s.t = ((u64)__shfl_sync(BB_FULL, t_hi, last) << 32) | __shfl_sync(BB_FULL, t_lo, last);
            const u32 am = __ballot_sync(BB_FULL, addl);
            if (am) {
                const u32 la = 31u - (u32)__clz(am);
                s.max_key_time = ((u64)__shfl_sync(BB_FULL, t_hi, la) << 32) | __shfl_sync(BB_FULL, t_lo, la);
            }
            bk_publish(r, s, lane, mark);
            pending &= ~clean;
        }
        if (cm) {
            const u32 k = (u32)__ffs(cm) - 1u;
#ifdef DP_PROF
            s.pf_ser += 1;
            s.pf_reason[__shfl_sync(BB_FULL, why, k) & 7u] += 1;
#endif
            uint4 kx, ky, ka, kc;
            kx.x = __shfl_sync(BB_FULL, x.x, k); kx.y = __shfl_sync(BB_FULL, x.y, k); kx.z = __shfl_sync(BB_FULL, x.z, k); kx.w = __shfl_sync(BB_FULL, x.w, k);
            ky.x = __shfl_sync(BB_FULL, y.x, k); ky.y = __shfl_sync(BB_FULL, y.y, k); ky.z = __shfl_sync(BB_FULL, y.z, k); ky.w = __shfl_sync(BB_FULL, y.w, k);
            ka.x = __shfl_sync(BB_FULL, a.x, k); ka.y = __shfl_sync(BB_FULL, a.y, k); ka.z = __shfl_sync(BB_FULL, a.z, k); ka.w = __shfl_sync(BB_FULL, a.w, k);
            kc.x = __shfl_sync(BB_FULL, c.x, k); kc.y = __shfl_sync(BB_FULL, c.y, k); kc.z = __shfl_sync(BB_FULL, c.z, k); kc.w = __shfl_sync(BB_FULL, c.w, k);
            if (!bk_serial(r, s, lane, kx, ky, ka, kc, ev0 + k + 1u, rf, lane_err)) return false;
            pending &= ~(1u << k);
            if (kx.z & BB_F_EMIT) {  // the caller writes the market-data record and comes back for the rest of the batch
                obs_lane = k;
                return true;
            }
        }
        (void)why;
    }
    return true;
}

}  // namespace bb
