/* A compiled host driving libbourse_b200.so through include/bourse_b200.h alone — what a Rust `extern "C"` binding (or
 * any other FFI) would do; no Python, no torch.  It restates three of the reference's own known-answer tests:
 *   tests/test_order_book.py:91-135            (immediate-mode OrderBook: trades, statuses, touch data)
 *   crates/step_sim/src/env.rs:311-368          (Env: queued instructions, step, per-step records)
 *   tests/test_step_sim/test_numpy_api.py:7-34  (StepEnvNumpy.submit_limit_orders + level_1_data / level_2_data)
 * Exit code 0 = all assertions hold; 3 = no usable CUDA device (bb_create returned BB_ECUDA: there is no CPU fallback);
 * 1 = an assertion failed.  Built and run by tests/test_c_host.py. */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bourse_b200.h"

#define CHECK(cond)                                                                                     \
    do {                                                                                                \
        if (!(cond)) {                                                                                  \
            fprintf(stderr, "%s:%d: check failed: %s (%s)\n", __FILE__, __LINE__, #cond, bb_last_error(h)); \
            return 1;                                                                                   \
        }                                                                                               \
    } while (0)
#define OK(call) CHECK((call) == BB_OK)

static bb_config base_config(uint64_t seed, uint64_t step_size) {
    bb_config c;
    memset(&c, 0, sizeof c);
    c.struct_size = sizeof c;
    c.n_envs = 1;
    c.seed = seed;
    c.step_size = step_size;
    c.tick_size = 1;
    c.trading = 1;
    c.obs_words = BB_OBS_L2;
    c.max_orders = 1024;
    c.max_trades = 1024;
    c.max_steps = 64;
    c.max_queue = 64;
    return c;
}

/* OrderBook.place_order at time t: one NEW instruction; the id is the book's order count before the call */
static int place(bb_handle* h, uint64_t t, int bid, uint32_t vol, uint32_t trader, int market, uint32_t price, uint64_t* id) {
    bb_instr x;
    uint64_t offs[2] = {0, 1};
    memset(&x, 0, sizeof x);
    OK(bb_n_orders(h, 0, id));
    x.t = t;
    x.op_flags = BB_OP_NEW | (bid ? BB_F_BID : 0u) | (market ? BB_F_MARKET : 0u);
    x.price = price;
    x.vol = vol;
    x.trader = trader;
    OK(bb_replay(h, &x, offs));
    return 0;
}

static int test_order_book_trades(void) {
    bb_handle* h = NULL;
    bb_config c = base_config(0, 1);
    int rc = bb_create(&c, &h);
    if (rc == BB_ECUDA) return 3;
    CHECK(rc == BB_OK);
    uint64_t id0, id1, id2, id3, id4, id5;
    uint8_t st;
    uint32_t l1[8];
    if (place(h, 0, 1, 10, 11, 0, 50, &id0) || place(h, 0, 0, 20, 12, 0, 60, &id1) || place(h, 0, 1, 10, 11, 0, 55, &id2) ||
        place(h, 0, 0, 20, 12, 0, 65, &id3) || place(h, 10, 1, 30, 11, 1, 0, &id4))
        return 1;
    OK(bb_order_status(h, 0, id4, &st)); CHECK(st == 2);
    OK(bb_order_status(h, 0, id1, &st)); CHECK(st == 2);
    OK(bb_book_level1(h, 0, l1));  /* bid, ask, bid_vol, ask_vol, best_bid_vol, best_ask_vol, n_bid, n_ask */
    CHECK(l1[0] == 55 && l1[1] == 65 && l1[2] == 20 && l1[3] == 10 && l1[4] == 10 && l1[6] == 1 && l1[5] == 10 && l1[7] == 1);
    if (place(h, 20, 0, 20, 12, 0, 55, &id5)) return 1;
    OK(bb_order_status(h, 0, id5, &st)); CHECK(st == 1);
    OK(bb_order_status(h, 0, id2, &st)); CHECK(st == 2);
    OK(bb_book_level1(h, 0, l1));
    CHECK(l1[0] == 50 && l1[1] == 55 && l1[2] == 10 && l1[3] == 20);
    uint64_t n = 0, t[3], active[3], passive[3];
    uint8_t side[3];
    uint32_t price[3], vol[3];
    OK(bb_n_trades(h, 0, &n)); CHECK(n == 3);
    OK(bb_trades(h, 0, 0, 3, t, side, price, vol, active, passive));
    CHECK(t[0] == 10 && t[1] == 10 && t[2] == 20);
    CHECK(side[0] == 0 && side[1] == 0 && side[2] == 1);          /* side of the passive order */
    CHECK(price[0] == 60 && price[1] == 65 && price[2] == 55);
    CHECK(vol[0] == 20 && vol[1] == 10 && vol[2] == 10);
    CHECK(active[0] == id4 && active[1] == id4 && active[2] == id5);
    CHECK(passive[0] == id1 && passive[1] == id3 && passive[2] == id2);
    /* a price off the tick grid is the reference's PriceError, an unknown id its panic */
    OK(bb_destroy(h));
    c.tick_size = 2;
    OK(bb_create(&c, &h));
    uint32_t act = BB_ACT_NEW, v = 5, tr = 0, p = 51;
    uint8_t sb = 1;
    uint64_t out = 0, done = 99;
    CHECK(bb_submit(h, 1, NULL, &act, &sb, &v, &tr, &p, NULL, NULL, &out, &done) == BB_EPRICE && done == 0);
    CHECK(strstr(bb_last_error(h), "Price 51 was not a multiple of tick-size 2") != NULL);
    CHECK(bb_order_status(h, 0, 7, &st) == BB_EBADID);
    OK(bb_destroy(h));
    return 0;
}

static int test_env(void) {
    bb_handle* h = NULL;
    bb_config c = base_config(101, 1000);
    int rc = bb_create(&c, &h);
    if (rc == BB_ECUDA) return 3;
    CHECK(rc == BB_OK);
    const uint32_t act2[2] = {BB_ACT_NEW, BB_ACT_NEW}, tr2[2] = {101, 101};
    const uint8_t side2[2] = {1, 0};
    uint64_t ids[2], done, t;
    uint32_t vol2[2] = {10, 20}, price2[2] = {10, 20}, l2[45];
    OK(bb_submit(h, 2, NULL, act2, side2, vol2, tr2, price2, NULL, NULL, ids, &done)); CHECK(ids[0] == 0 && ids[1] == 1);
    OK(bb_step(h, 1));
    OK(bb_level2(h, l2)); CHECK(l2[1] == 10 && l2[2] == 20);
    OK(bb_time(h, 0, &t)); CHECK(t == 1000);
    price2[0] = 11; price2[1] = 21;
    OK(bb_submit(h, 2, NULL, act2, side2, vol2, tr2, price2, NULL, NULL, ids, &done)); CHECK(ids[0] == 2 && ids[1] == 3);
    OK(bb_step(h, 1));
    OK(bb_level2(h, l2)); CHECK(l2[1] == 11 && l2[2] == 20);
    const uint32_t act1 = BB_ACT_NEW, v30 = 30, tr1 = 101, zero = 0, fl = BB_F_MARKET;
    const uint8_t bid = 1;
    OK(bb_submit(h, 1, NULL, &act1, &bid, &v30, &tr1, &zero, NULL, &fl, ids, &done));   /* market bid for 30 */
    OK(bb_step(h, 1));
    OK(bb_level2(h, l2)); CHECK(l2[1] == 11 && l2[2] == 21 && l2[3] == 10 /* ask_vol */ && l2[0] == 30 /* trade_vol */);
    OK(bb_time(h, 0, &t)); CHECK(t == 3000);
    uint8_t status[5];
    OK(bb_orders(h, 0, 0, 5, NULL, status, NULL, NULL, NULL, NULL, NULL, NULL));
    CHECK(status[0] == 1 && status[1] == 2 && status[2] == 1 && status[3] == 1 && status[4] == 2);
    uint64_t n_tr; uint32_t n_steps, hist[3 * 45];
    OK(bb_n_trades(h, 0, &n_tr)); CHECK(n_tr == 2);
    OK(bb_n_steps(h, 0, &n_steps)); CHECK(n_steps == 3);
    OK(bb_history(h, 0, 0, 3, hist));
    const uint32_t want[3][9] = {{0, 10, 20, 20, 10, 10, 1, 20, 1}, {0, 11, 20, 40, 20, 10, 1, 20, 1}, {30, 11, 21, 10, 20, 10, 1, 10, 1}};
    for (int s = 0; s < 3; ++s)   /* trade_vol, bid, ask, ask_vol, bid_vol, touch bid vol / n, touch ask vol / n */
        for (int k = 0; k < 9; ++k) CHECK(hist[45 * s + k] == want[s][k]);
    OK(bb_destroy(h));
    return 0;
}

static int test_numpy_arrays(void) {
    bb_handle* h = NULL;
    bb_config c = base_config(101, 100000);
    int rc = bb_create(&c, &h);
    if (rc == BB_ECUDA) return 3;
    CHECK(rc == BB_OK);
    const uint32_t act[6] = {1, 1, 1, 1, 1, 1}, vol[6] = {10, 11, 12, 10, 11, 12}, tr[6] = {1, 1, 1, 2, 2, 2}, price[6] = {20, 20, 19, 22, 22, 23};
    const uint8_t side[6] = {1, 1, 1, 0, 0, 0};
    uint64_t ids[6], done;
    uint32_t l1[9], l2[45];
    OK(bb_submit(h, 6, NULL, act, side, vol, tr, price, NULL, NULL, ids, &done));
    for (int i = 0; i < 6; ++i) CHECK(ids[i] == (uint64_t)i);
    OK(bb_step(h, 1));
    const uint32_t want[13] = {0, 20, 22, 33, 33, 21, 2, 21, 2, 12, 1, 12, 1};
    OK(bb_level1(h, l1)); OK(bb_level2(h, l2));
    for (int k = 0; k < 9; ++k) CHECK(l1[k] == want[k]);
    for (int k = 0; k < 45; ++k) CHECK(l2[k] == (k < 13 ? want[k] : 0u));
    bb_stats_t st;
    OK(bb_stats(h, &st)); CHECK(st.instructions == 6 && st.orders_created == 6 && st.trades == 0 && st.env_steps == 1 && st.error_envs == 0);
    OK(bb_destroy(h));
    return 0;
}

/* Multi-GPU behind the C ABI: `n_shards` handles, one per device (all on device 0 when the box has a single
 * GPU: NCCL then refuses duplicate devices, so the 1-GPU case runs one shard through the same calls), each owning a
 * contiguous block of the global env ids; in-kernel RandomAgents keyed by GLOBAL env id; bb_gather_stats all-gathers the
 * shards' statistics over NCCL.  The sharded job must equal the unsharded one: same total counters, and shard k's level-1
 * checksum equals the checksum of the same envs taken from ... a handle created with the same env_id_base. */
static int run_shard(int device, uint32_t base, uint32_t n_envs, bb_handle** out) {
    bb_handle* h = NULL;
    bb_config c = base_config(0, 1000000);
    c.device = device;
    c.n_envs = n_envs;
    c.env_id_base = base;
    c.obs_words = BB_OBS_L1;
    c.max_orders = 4096;
    c.max_trades = 8192;
    c.max_steps = 32;
    c.max_queue = 128;
    int rc = bb_create(&c, &h);
    if (rc == BB_ECUDA) return 3;
    CHECK(rc == BB_OK);
    bb_agent_group g[2];
    memset(g, 0, sizeof g);
    g[0].kind = BB_GROUP_RANDOM; g[0].n_agents = 50; g[0].tick_lo = 40; g[0].tick_hi = 60; g[0].vol_lo = 10; g[0].vol_hi = 20;
    g[0].tick_size = 2; g[0].rate = 0.8f;
    g[1] = g[0]; g[1].tick_lo = 10; g[1].tick_hi = 90; g[1].vol_lo = 50; g[1].vol_hi = 70; g[1].rate = 0.2f;
    OK(bb_set_agents(h, g, 2));
    OK(bb_run_agents(h, 101, 32));
    OK(bb_synchronize(h));
    *out = h;
    return 0;
}

static int test_multi_gpu_gather(int n_dev) {
    const uint32_t total = 24;
    const int n_shards = n_dev >= 2 ? 2 : 1;
    bb_handle* h = NULL;      /* (for CHECK's error text) */
    bb_handle* whole = NULL;
    bb_handle* shard[2] = {NULL, NULL};
    int rc = run_shard(0, 0, total, &whole);
    if (rc) return rc;
    h = whole;
    bb_stats_t ref;
    OK(bb_stats(whole, &ref));
    CHECK(ref.instructions > 10000 && ref.trades > 1000 && ref.error_envs == 0);
    for (int k = 0; k < n_shards; ++k) {
        const uint32_t base = k * (total / n_shards), cnt = (k == n_shards - 1) ? total - base : total / n_shards;
        rc = run_shard(k, base, cnt, &shard[k]);
        if (rc) return rc;
    }
    bb_comm* comm = NULL;
    rc = bb_comm_init_all(n_shards, NULL, &comm);
    if (rc != BB_OK) {
        fprintf(stderr, "bb_comm_init_all: %s\n", bb_comm_last_error());
        return 1;
    }
    CHECK(bb_comm_n_ranks(comm) == n_shards);
    bb_stats_t all[2];
    double ms_in[2] = {12.5, 40.25}, ms_out[2] = {0, 0};
    rc = bb_gather_stats(comm, shard, (uint32_t)n_shards, ms_in, all, ms_out);
    if (rc != BB_OK) {
        fprintf(stderr, "bb_gather_stats: %s\n", bb_comm_last_error());
        return 1;
    }
    uint64_t instr = 0, trades = 0, vol = 0, steps = 0;
    for (int k = 0; k < n_shards; ++k) {
        instr += all[k].instructions; trades += all[k].trades; vol += all[k].traded_volume; steps += all[k].env_steps;
        CHECK(ms_out[k] == ms_in[k]);
        bb_stats_t own;
        h = shard[k];
        OK(bb_stats(shard[k], &own));
        CHECK(own.l1_checksum == all[k].l1_checksum && own.instructions == all[k].instructions);
    }
    h = whole;
    CHECK(instr == ref.instructions && trades == ref.trades && vol == ref.traded_volume && steps == ref.env_steps);
    if (n_shards == 1) CHECK(all[0].l1_checksum == ref.l1_checksum);
    OK(bb_comm_destroy(comm));
    for (int k = 0; k < n_shards; ++k) OK(bb_destroy(shard[k]));
    OK(bb_destroy(whole));
    printf("ok multi_gpu_gather_%d_shards\n", n_shards);
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && strcmp(argv[1], "--multi-gpu") == 0) {   /* tests/test_c_host.py passes the number of visible devices */
        const int rc = test_multi_gpu_gather(argc > 2 ? atoi(argv[2]) : 1);
        if (rc == 3) fprintf(stderr, "no usable CUDA device: %s\n", bb_last_error(NULL));
        return rc;
    }
    if (bb_abi_version() != BB_ABI_VERSION) {
        fprintf(stderr, "ABI version mismatch\n");
        return 1;
    }
    int (*tests[])(void) = {test_order_book_trades, test_env, test_numpy_arrays};
    const char* names[] = {"order_book_trades", "env", "numpy_arrays"};
    for (int i = 0; i < 3; ++i) {
        const int rc = tests[i]();
        if (rc == 3) {
            fprintf(stderr, "no usable CUDA device: %s\n", bb_last_error(NULL));
            return 3;
        }
        if (rc) return 1;
        printf("ok %s\n", names[i]);
    }
    return 0;
}
