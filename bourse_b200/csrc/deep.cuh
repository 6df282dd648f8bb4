// Deep-book engine: ONE CTA PER BOOK, four specialised warps in a pipeline (included by kernels.cuh).
//
// Same reference semantics as the other engines (side.rs:36-143, orderbook.rs:429-772), built for books with ~10^6
// resting orders (BASELINE config C5), where one warp per book leaves an SM running a single dependent chain.  An event's
// work is cut ALONG ITS DATA instead: what the event does to the price ladder is decided by one warp, how that lands on
// the orders of a level by another, and they run one behind the other on different events.
//
//   fetch warp   walks the instruction stream a batch (32 events) ahead: the batch arrives by one bulk (TMA) copy; every
//                lane settles what does not depend on the book (order ids, market sentinels; the write-once half of a new
//                order's record goes straight to HBM) and pulls the order record its cancel / modify names — the only
//                DRAM-cold read of the path — into shared memory with cp.async;
//   ladder warp  (lane 0) the price ladder: a dense tick-indexed array of level volumes with per-side non-empty bitmaps and
//                a summary word, best bid / ask, side totals.  It decides how much volume an aggressive order takes from
//                each level it crosses and where an order rests, and says so in commands; it never touches a queue;
//   queue warp   the price-time queues: lane 0 walks a level's FIFO out of a 64-chunk cache for the one-to-three-fill sweeps
//                that make up most of the flow; the whole warp is called in for chunk loads and for real sweeps, which
//                take a level's FIFO with an inclusive prefix sum over the resting volumes (one trade per passive order, in
//                queue order, orderbook.rs:429-454, 843-870);
//   retire warp  drains a ring of order-record / trade-log writes (one lane per entry), so neither chain above holds a
//                global store but the 8-byte queue entries.
//
// Price-time queues are ARRAYS, not linked lists: a level's FIFO is a chain of 256-byte chunks in HBM (31 entries of
// {order id, remaining volume} + a link).  An order remembers the position of its entry, so a cancel writes a tombstone
// there and never touches its neighbours, and a sweep reads consecutive entries with one coalesced load.
//
// The ladder warp runs ahead of the queue warp, so it cannot see which orders a sweep it has just ordered will fill.  A
// cancel / modify may therefore use its prefetched order record only if (a) no event that names the order and (b) no
// sweep of the order's price level was still on its way to HBM when the record was fetched; two small filters (by order
// id, by level) answer that, and the rare doubtful case waits for the pipeline to drain and reads the record again.
//
// Preconditions (violations set a per-env error bit and never change results silently):
//   * resting prices inside [d_win_lo, d_win_lo + d_levels)                                     else ERR_CAP_PAGES
//   * strictly increasing time between resting inserts (no equal-(price, time) keys, N1)         else ERR_TIME_ORDER
//   * a price level holds one side at a time (true while trading is enabled and volumes are > 0)  else ERR_LOCKED
//   * chunk pool not exhausted                                                                   else ERR_CAP_PAGES
#pragma once

namespace bb {

#define ERR_LOCKED 0x800u  // deep engine: both sides resting at one price (trading disabled, or a crossing zero-volume order)

// ---- shared-memory image of a deep book (persisted in the book blob between launches) -------------------------------
//   0     BookHdr (128 B; free_top doubles as the "a zero-volume order has rested" flag)
//   128   chunk_bump u32 (next never-used chunk id; chunk 0 is a sink), n_free u32 (entries on the free-chunk stack)
//   192   free-chunk stack u32 [DP_FS_CAP]                                                     (queue warp)
//   1216  bitmap ask [NW], bitmap bid [NW]   bit q set <=> level q holds resting orders of that side   (ladder warp)
//         summary ask [8], summary bid [8]   bit w set <=> bitmap word w is non-zero                   (ladder warp)
//         lvol [W] u32   resting volume per price level                                                (ladder warp)
//         lcnt [W] u32   resting orders per price level                                                (queue warp)
//         lht  [W] u64   {head, tail} of the level's queue: chunk id << 5 | entry index (tail: next free entry)  (queue warp)
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
#define DP_CCAP 64u          // ladder -> queue command ring entries (32 B each)
#define DP_RCAP 128u         // retire-ring entries (32 B each)
#define DP_RENT 32u
#define DP_DIRTY 2048u       // touched-order filter buckets (by order id)
#define DP_SWEPT 256u        // swept-level filter buckets (by level index)
struct DeepOff {             // byte offsets from the CTA's shared-memory base, filled by the host
    u32 bm, sm, lvol, lcnt, lht, image_bytes;
    u32 smem_image;  // bytes of the image that live in shared memory during a launch (k_deepw: up to lcnt; its queues' counts / heads / tails stay in the blob)
    u32 ctag, cdat, ev_ins, ev_rec, ev_rf, cmd, ret, dirty, swept, ctl, bar, total;
};
// control words (u32 each, at DeepOff::ctl)
#define CT_EV_READY 0u      // fetch: batches published
#define CT_EV_CONSUMED 4u   // ladder: batches it is done with
#define CT_CMD_TAIL 8u      // ladder: commands produced
#define CT_CMD_DONE 12u     // queue: commands fully processed
#define CT_RET_TAIL 16u     // queue: retire entries produced
#define CT_RET_DONE 20u     // retire: entries whose stores are performed and fenced
#define CT_Q_EV 24u         // queue: events whose commands are all processed (published after CT_RET_TAIL)
#define CT_EV_RETIRED 28u   // retire: events whose record writes are all performed and fenced
#define CT_FIN_L 32u        // ladder: no more commands will come
#define CT_FIN 36u          // queue: no more retire entries will come
#define CT_ABORT 40u        // any warp: a bounded wait ran out (the launch flags 0x80000000 and unwinds)
#define CT_RERR 44u         // retire warp's error bits
#define CT_QERR 48u         // queue warp's error bits
#define CT_LERR 52u         // ladder warp's error bits
#define CT_QNTR 56u         // queue warp: trades made in this launch
// k_deepw (deepw.cuh): chain warp <-> replay warp
#define CT_MOP_TAIL 64u     // chain: micro-ops published
#define CT_MOP_DONE 68u     // replay: micro-ops replayed
#define CT_DRAIN 72u        // chain: odd = "finish what is published, then park" (even: run)
#define CT_PARKED 76u       // replay: the CT_DRAIN value it has parked for
#define RS_BUMP 80u         // the replay warp's scalars, at rest here between rounds (the chain warp borrows them while the
#define RS_NFREE 84u        // replay warp is parked: complex events run on the chain warp)
#define RS_NTR 88u
#define RS_RET_TAIL 92u
#define RS_ERR 96u
#define CT_WORDS 32u

// ladder -> queue commands: a = {op | side << 8 | kind-or-flag << 12 | last << 16 | status << 20, order id, level / position, volume},
//                            b = {t lo, t hi, price / level, events completed once this command is done}
#define QC_APPEND 1u   // the order rests: append to level a.z (kind = RK_NEW / RK_REPLACE), b.z = price
#define QC_END 2u      // the order ended without resting (Filled / Cancelled / Rejected = status), a.w = volume left
#define QC_REMOVE 3u   // tombstone the entry at position a.z of level b.z; flag: a cancel (else the first half of a replace)
#define QC_REDUCE 4u   // entry at position a.z: volume becomes a.w
#define QC_SWEEP 5u    // aggressor a.y takes a.w volume from level a.z of side `side`; flag: and every order left there
#define QCF_LAST (1u << 16)

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
// Developer profiling (-DDP_PROF): cycles book 0's warps spend in each kind of wait, printed at the end of the launch.
#ifdef DP_PROF
__device__ unsigned long long g_dp_prof[64];
#define DP_T0 const long long dp_t0_ = clock64();
#define DP_ADD(slot)                                                                  \
    if (blockIdx.x == 0) {                                                            \
        atomicAdd(&g_dp_prof[slot], (unsigned long long)(clock64() - dp_t0_));        \
        atomicAdd(&g_dp_prof[32 + (slot)], 1ull);                                     \
    }
#else
#define DP_T0
#define DP_ADD(slot)
#endif
// bounded spin on a control word; returns false when the wait ran out or another warp aborted
#define DP_SPIN_MAX (1u << 25)
// NAP: the waiting warp sleeps between polls (producers / consumers that are far ahead: their polling would otherwise take
// issue slots and shared-memory bandwidth from the book warps of the SM)
template <bool NAP = false, class F> __device__ __forceinline__ bool dp_wait(u32 ctl, F cond, int prof_slot = 0) {
    DP_T0
    for (u32 spin = 0; spin < DP_SPIN_MAX; ++spin) {
        if (cond()) {
            DP_ADD(prof_slot)
            return true;
        }
        if (NAP && spin >= 4u) __nanosleep(spin < 64u ? 100u : 400u);
        if ((spin & 255u) == 255u && ld_acq(ctl + CT_ABORT)) return false;
    }
    st_rel(ctl + CT_ABORT, 1u);
    return false;
}
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

// =====================================================================================================================
// LADDER warp (lane 0): price levels, volumes, best bid / ask — everything that decides WHAT an event does to the book.
// It never touches a queue: how a level's volume is spread over its orders is the queue warp's business, told through the
// command ring.  Launch-invariant addresses are pinned in registers (an opaque move keeps the optimiser from rebuilding
// `base + constant-bank offset` at every use).
struct LadReg {
    u32 lvol, lcnt, bma, bmb, sma, smb, cmd, dirty, swept, ctl;
    u32 win_lo, W, max_orders;
    u64 oh;
};
struct LadSt {
    u64 t, max_key_time;
    u32 n_orders, trade_vol, vol_ask, vol_bid, bq_ask, bq_bid, flags, err;
    u32 d_instr, d_applied;
    u32 zv;         // a zero-volume order has rested: "volume 0" no longer means "level empty"
    u32 cmd_tail;   // commands produced so far
    u32 cmd_room;   // ... and the value of cmd_tail up to which ring space is known to be free
    u32 cmd_pub;    // ... and the value last published to the queue warp
};
__device__ __forceinline__ u32 ld_bm(const LadReg& r, u32 side, u32 w) { return (side ? r.bmb : r.bma) + 4u * w; }
__device__ __forceinline__ u32 ld_sm(const LadReg& r, u32 side, u32 w) { return (side ? r.smb : r.sma) + 4u * w; }
__device__ __forceinline__ bool ld_has_best(const LadSt& s, u32 side) { return (s.flags >> (1u + side)) & 1u; }
__device__ __forceinline__ u32 ld_best_q(const LadSt& s, u32 side) { return side ? s.bq_bid : s.bq_ask; }
__device__ __forceinline__ void ld_add_side(LadSt& s, u32 side, u32 dv) {
    if (side) s.vol_bid += dv; else s.vol_ask += dv;
}

// level q of `side` just became empty (its bitmap bit is still set): clear it and, if it was the touch, find the next one
__device__ __forceinline__ void ld_level_gone(const LadReg& r, LadSt& s, u32 side, u32 q) {
    const u32 w = q >> 5;
    const u32 ba = ld_bm(r, side, w);
    u32 m = lds(ba) & ~(1u << (q & 31u));
    sts(ba, m);
    if (m == 0u) {
        const u32 sa = ld_sm(r, side, w >> 5);
        sts(sa, lds(sa) & ~(1u << (w & 31u)));
    }
    if (!(ld_has_best(s, side) && ld_best_q(s, side) == q)) return;
    // q was the best level: every other level of this side lies on the far side of it
    if (side == 0u) {
        if (m) { s.bq_ask = (w << 5) + (u32)__ffs(m) - 1u; return; }
        u32 sw = w >> 5;
        u32 ms = lds(ld_sm(r, 0u, sw)) & ~((2u << (w & 31u)) - 1u);  // summary bits above word w
        while (ms == 0u) {
            if (++sw >= DP_NS) { s.flags &= ~FL_HAS_ASK; return; }
            ms = lds(ld_sm(r, 0u, sw));
        }
        const u32 w2 = (sw << 5) + (u32)__ffs(ms) - 1u;
        s.bq_ask = (w2 << 5) + (u32)__ffs(lds(ld_bm(r, 0u, w2))) - 1u;
    } else {
        if (m) { s.bq_bid = (w << 5) + 31u - (u32)__clz(m); return; }
        u32 sw = w >> 5;
        u32 ms = lds(ld_sm(r, 1u, sw)) & ((1u << (w & 31u)) - 1u);  // summary bits below word w
        while (ms == 0u) {
            if (sw == 0u) { s.flags &= ~FL_HAS_BID; return; }
            --sw;
            ms = lds(ld_sm(r, 1u, sw));
        }
        const u32 w2 = (sw << 5) + 31u - (u32)__clz(ms);
        s.bq_bid = (w2 << 5) + 31u - (u32)__clz(lds(ld_bm(r, 1u, w2)));
    }
}

// command ring: room for one more?  The consumer's counter is re-read only when the cached bound is used up.
__device__ __forceinline__ bool ld_cmd(const LadReg& r, LadSt& s, uint4 a, uint4 b) {
    if (s.cmd_tail >= s.cmd_room) {
        if (s.cmd_pub != s.cmd_tail) {
            s.cmd_pub = s.cmd_tail;
            st_rel(r.ctl + CT_CMD_TAIL, s.cmd_pub);
        }
        if (!dp_wait(r.ctl, [&] {
                s.cmd_room = ld_acq(r.ctl + CT_CMD_DONE) + DP_CCAP;
                return s.cmd_tail < s.cmd_room;
            }, 4))
            return false;
    }
    const u32 ea = r.cmd + 32u * (s.cmd_tail & (DP_CCAP - 1u));
    sts128(ea, a);
    sts128(ea + 16u, b);
    s.cmd_tail += 1;
    return true;
}
__device__ __forceinline__ void ld_publish(const LadReg& r, LadSt& s) {
    if (s.cmd_pub != s.cmd_tail) {
        s.cmd_pub = s.cmd_tail;
        st_rel(r.ctl + CT_CMD_TAIL, s.cmd_pub);
    }
}
// wait until the queue warp has processed every command produced so far (its per-level order counts are then current)
__device__ __forceinline__ bool ld_sync_queue(const LadReg& r, LadSt& s) {
    ld_publish(r, s);
    if (ld_acq(r.ctl + CT_CMD_DONE) == s.cmd_tail) return true;
    return dp_wait(r.ctl, [&] { return ld_acq(r.ctl + CT_CMD_DONE) == s.cmd_tail; }, 5);
}
// the volume of level q of `side` just reached zero: is the level empty?  Without zero-volume orders: yes.
__device__ __forceinline__ bool ld_emptied(const LadReg& r, LadSt& s, u32 q, bool& aborted) {
    if (!s.zv) return true;
    if (!ld_sync_queue(r, s)) {
        aborted = true;
        return false;
    }
    return lds(r.lcnt + 4u * q) == 0u;
}

// =====================================================================================================================
// QUEUE warp: the price-time queues.  Lane 0 walks them serially out of the chunk cache; the whole warp is called in for
// chunk loads and for real sweeps (prefix sum over the FIFO).
struct QueReg {
    u32 lcnt, lht, ctag, cdat, cmd, ret, ctl, fs;
    u32 n_chunks, win_lo;
    u64 chunks;
};
struct QueSt {
    u32 bump, n_free, err, n_trades;
    u32 ret_tail, ret_room, ret_pub;
    u32 cmd_head;
};
__device__ __forceinline__ u32 qu_alloc_chunk(const QueReg& r, QueSt& s) {
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
__device__ __forceinline__ void qu_free_chunk(const QueReg& r, QueSt& s, u32 c) {
    if (c != 0u && s.n_free < DP_FS_CAP) {
        sts(r.fs + 4u * s.n_free, c);
        s.n_free += 1;
    }
}
// write-through store of one word of a queue entry / chunk link
__device__ __forceinline__ void qu_chunk_st32(const QueReg& r, u32 c, u32 byte_off, u32 v) {
    stg32(r.chunks + (u64)c * DP_CHUNK_BYTES + byte_off, v);
    const u32 slot = c & (DP_NC - 1u);
    if (lds(r.ctag + 4u * slot) == c) sts(r.cdat + DP_CHUNK_BYTES * slot + byte_off, v);
}
__device__ __forceinline__ bool qu_ret_space(const QueReg& r, QueSt& s, u32 n) {
    if (s.ret_tail + n <= s.ret_room) return true;
    if (s.ret_pub != s.ret_tail) {
        s.ret_pub = s.ret_tail;
        st_rel(r.ctl + CT_RET_TAIL, s.ret_pub);  // whatever is waiting to be drained must be visible to the drainer
    }
    return dp_wait(r.ctl, [&] {
        s.ret_room = ld_acq(r.ctl + CT_RET_DONE) + DP_RCAP;
        return s.ret_tail + n <= s.ret_room;
    }, 8);
}
__device__ __forceinline__ void qu_ret_write(const QueReg& r, u32 idx, uint4 a, uint4 b) {
    const u32 ea = r.ret + DP_RENT * (idx & (DP_RCAP - 1u));
    sts128(ea, a);
    sts128(ea + 16u, b);
}

// insert_order's queue half (side.rs:54-66): append to the level's chunk chain.  Returns the entry position.
__device__ __forceinline__ u32 qu_append(const QueReg& r, QueSt& s, u32 q, u32 id, u32 vol) {
    const u32 cnt = lds(r.lcnt + 4u * q);
    u32 pos;
    if (cnt == 0u) {
        const u32 c = qu_alloc_chunk(r, s);
        pos = c << 5;
        sts64(r.lht + 8u * q, ((u64)(pos + 1u) << 32) | pos);
    } else {
        const u32 tail = lds(r.lht + 8u * q + 4u);
        u32 c = tail >> 5, idx = tail & 31u;
        if (idx == DP_CHUNK_ENTRIES) {  // tail chunk full: link a new one
            const u32 c2 = qu_alloc_chunk(r, s);
            qu_chunk_st32(r, c, 8u * DP_CHUNK_ENTRIES, c2);
            c = c2;
            idx = 0u;
        }
        pos = (c << 5) | idx;
        sts(r.lht + 8u * q + 4u, pos + 1u);
    }
    sts(r.lcnt + 4u * q, cnt + 1u);
    const u32 c = pos >> 5, idx = pos & 31u, slot = c & (DP_NC - 1u);
    const u32 tag = lds(r.ctag + 4u * slot);
    stg64(r.chunks + (u64)c * DP_CHUNK_BYTES + 8u * idx, ((u64)vol << 32) | id);
    if (tag == c) sts64(r.cdat + DP_CHUNK_BYTES * slot + 8u * idx, ((u64)vol << 32) | id);
    return pos;
}
// remove_order's queue half (side.rs:75-84): tombstone the entry
__device__ __forceinline__ void qu_remove(const QueReg& r, QueSt& s, u32 q, u32 pos) {
    qu_chunk_st32(r, pos >> 5, 8u * (pos & 31u), BB_NIL);
    const u32 cnt = lds(r.lcnt + 4u * q);
    if (cnt <= 1u) {
        const u64 ht = lds64(r.lht + 8u * q);
        if (((u32)ht >> 5) == ((u32)(ht >> 32) >> 5)) qu_free_chunk(r, s, (u32)ht >> 5);  // a longer all-dead chain is left to the pool
        sts(r.lcnt + 4u * q, 0u);
    } else {
        sts(r.lcnt + 4u * q, cnt - 1u);
    }
}

// ---- match_orders over one level (orderbook.rs:843-870), lane 0 alone, out of chunks resident in the cache ----------------
// Takes `take` volume from the head of level q's queue — and, with `exhaust`, every order left there afterwards (the
// aggressor still had volume, so it also trades, at volume 0, with the zero-volume orders behind: `while order.vol > 0`,
// orderbook.rs:436).  One trade per passive order, in queue order.  Returns true when the rest needs the whole warp: the
// head chunk is not resident, or DP_SERIAL_FILLS fills are done (a real sweep: the prefix-sum path takes over).
#define DP_SERIAL_FILLS 4u
__device__ __forceinline__ bool qu_sweep_serial(const QueReg& r, QueSt& s, u32 q, u32 opp, u32& take, bool exhaust, u32 id, u32 t_lo, u32 t_hi,
                                                u32 price, bool& aborted) {
    u32 fills = 0u;
    u32 cnt = lds(r.lcnt + 4u * q);
    while ((take > 0u || exhaust) && cnt > 0u) {
        const u64 ht = lds64(r.lht + 8u * q);
        const u32 head = (u32)ht, tail = (u32)(ht >> 32);
        const u32 c = head >> 5, tc = tail >> 5;
        const u32 slot = c & (DP_NC - 1u);
        if (lds(r.ctag + 4u * slot) != c || fills >= DP_SERIAL_FILLS) return true;
        const u32 ca = r.cdat + DP_CHUNK_BYTES * slot;
        const u32 end = (c == tc) ? (tail & 31u) : DP_CHUNK_ENTRIES;
        u32 idx = head & 31u;
        bool stop = false;
        while (idx < end) {
            const u64 e = lds64(ca + 8u * idx);
            const u32 pid = (u32)e, pvol = (u32)(e >> 32);
            if (pid == BB_NIL) {  // tombstone
                ++idx;
                continue;
            }
            if (take == 0u && !exhaust) {
                stop = true;
                break;
            }
            const u32 tv = min(take, pvol), pv = pvol - tv;
            if (!qu_ret_space(r, s, 1u)) {
                aborted = true;
                return false;
            }
            qu_ret_write(r, s.ret_tail, make_uint4(RK_FILL | (opp << 8) | (pv == 0u ? 0x10000u : 0u), pid, tv, pv), make_uint4(t_lo, t_hi, price, id));
            s.ret_tail += 1;
            s.n_trades += 1;
            take -= tv;
            ++fills;
            if (pv != 0u) {  // partially filled: stays at the head; the aggressor is done
                sts(ca + 8u * idx + 4u, pv);
                stg32(r.chunks + (u64)c * DP_CHUNK_BYTES + 8u * idx + 4u, pv);
                stop = true;
                break;
            }
            ++idx;
            --cnt;
            if (cnt == 0u || fills >= DP_SERIAL_FILLS) break;
        }
        if (cnt == 0u) {  // the level is gone
            if (c == tc) qu_free_chunk(r, s, c);
            sts(r.lcnt + 4u * q, 0u);
            return false;
        }
        sts(r.lcnt + 4u * q, cnt);
        if (idx >= end && c != tc) {  // this chunk is used up: follow the link
            const u32 nc = lds(ca + 8u * DP_CHUNK_ENTRIES);
            qu_free_chunk(r, s, c);
            if (nc >= r.n_chunks) {  // broken chain: only after an earlier capacity error
                s.err |= ERR_CAP_PAGES;
                sts(r.lcnt + 4u * q, 0u);
                return false;
            }
            sts(r.lht + 8u * q, nc << 5);
        } else if (idx >= end) {  // live orders counted but none found: only after an earlier capacity error
            s.err |= ERR_CAP_PAGES;
            sts(r.lcnt + 4u * q, 0u);
            return false;
        } else {
            sts(r.lht + 8u * q, (c << 5) | idx);
        }
        if (stop) return false;
    }
    return false;
}

// ---- the warp-cooperative version: a prefix sum over the FIFO ---------------------------------------------------------------
// Lanes load consecutive queue entries of the head chunk (one coalesced load when it is not resident), an inclusive scan
// over the resting volumes tells every lane whether its order is reached and whether it is filled completely, and the
// trades are emitted in queue order.  All 32 lanes call it; the queue state lives in lane 0 (`s` is only meaningful there).
// Returns (in every lane) the volume still to take.
__device__ __forceinline__ u32 qu_sweep_warp(const QueReg& r, QueSt& s, u32 lane, u32 q, u32 opp, u32 take, bool exhaust, u32 id, u32 t_lo,
                                              u32 t_hi, u32 price) {
    for (u32 guard = 0; guard < (1u << 22); ++guard) {
        const u32 cnt0 = lds(r.lcnt + 4u * q);
        if (!((take > 0u || exhaust) && cnt0 > 0u)) break;
        const u64 ht = lds64(r.lht + 8u * q);
        const u32 head = (u32)ht, tail = (u32)(ht >> 32);
        const u32 c = head >> 5, idx = head & 31u, tc = tail >> 5;
        const u32 end = (c == tc) ? (tail & 31u) : DP_CHUNK_ENTRIES;
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
        const bool reached = live && (excl < take || exhaust);
        const bool full = reached && incl <= take;
        const u32 mr = __ballot_sync(BB_FULL, reached), mf = __ballot_sync(BB_FULL, full);
        const u32 mp = mr & ~mf;
        const u32 nr = __popc(mr);
        const u32 total = __shfl_sync(BB_FULL, incl, 31);
        const u32 traded = min(take, total);
        u32 base = 0u, ok = 1u;
        if (lane == 0u) {
            ok = qu_ret_space(r, s, nr) ? 1u : 0u;
            base = s.ret_tail;
        }
        ok = __shfl_sync(BB_FULL, ok, 0);
        if (!ok) break;
        base = __shfl_sync(BB_FULL, base, 0);
        if (reached) {
            const u32 tv = full ? pvol : take - excl;
            const u32 pv = pvol - tv;
            // trade: side / price are the passive order's (orderbook.rs:853-862)
            qu_ret_write(r, base + __popc(mr & ((1u << lane) - 1u)), make_uint4(RK_FILL | (opp << 8) | (pv == 0u ? 0x10000u : 0u), pid, tv, pv),
                         make_uint4(t_lo, t_hi, price, id));
            if (!full) {  // the partially filled order stays at the head of its level
                sts(ca + 8u * lane + 4u, pv);
                stg32(r.chunks + (u64)c * DP_CHUNK_BYTES + 8u * lane + 4u, pv);
            }
        }
        take -= traded;
        const u32 cnt = cnt0 - __popc(mf);
        __syncwarp();
        bool done = false;
        if (lane == 0u) {
            s.ret_tail = base + nr;
            s.n_trades += nr;
            if (cnt == 0u) {  // the level is gone
                if (c == tc) qu_free_chunk(r, s, c);
                sts(r.lcnt + 4u * q, 0u);
            } else {
                sts(r.lcnt + 4u * q, cnt);
                if (mp) {
                    sts(r.lht + 8u * q, (c << 5) | ((u32)__ffs(mp) - 1u));
                    done = true;
                } else if (take == 0u && !exhaust) {  // ended exactly on an order boundary
                    sts(r.lht + 8u * q, (c << 5) | (mf ? 32u - (u32)__clz(mf) : idx));
                    done = true;
                } else if (c != tc) {  // this chunk is used up: follow the link
                    const u32 nc = lds(ca + 8u * DP_CHUNK_ENTRIES);
                    qu_free_chunk(r, s, c);
                    if (nc < r.n_chunks) {
                        sts(r.lht + 8u * q, nc << 5);
                    } else {  // broken chain: only after an earlier capacity error
                        s.err |= ERR_CAP_PAGES;
                        sts(r.lcnt + 4u * q, 0u);
                    }
                } else {  // live orders counted but none found: only after an earlier capacity error
                    s.err |= ERR_CAP_PAGES;
                    sts(r.lcnt + 4u * q, 0u);
                }
            }
        }
        done = __shfl_sync(BB_FULL, done ? 1u : 0u, 0) != 0u;
        __syncwarp();
        if (done) break;
    }
    return take;
}

// (vol, count) of `side` at an arbitrary price: per-lane (after the ladder warp has synchronised with the queue warp)
__device__ __forceinline__ void dp_level_at(const LadReg& r, u32 side, u32 price, u32* vol, u32* cnt) {
    *vol = 0;
    *cnt = 0;
    const u32 q = price - r.win_lo;
    if (q >= r.W) return;
    if (!((lds(ld_bm(r, side, q >> 5)) >> (q & 31u)) & 1u)) return;
    *vol = lds(r.lvol + 4u * q);
    *cnt = lds(r.lcnt + 4u * q);
}

// observation words of the book: lane l owns words l and l + 32 (layout: book_obs in book.cuh)
__device__ __forceinline__ void dp_obs(const LadReg& r, u32 tick, u32 lane, u32 trade_vol, u32 bid, u32 ask, u32 vol_ask, u32 vol_bid,
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
