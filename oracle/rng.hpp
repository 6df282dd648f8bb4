// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped product path;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// Random number generation used by the CPU restatement.
//
// Two families live here:
//
//  1. "rand-compatible" stream RNG: a restatement of the *published algorithms* of the third-party
//     crates the reference depends on (NOT present under /root/reference; pinned in its
//     Cargo.lock:623-665): rand 0.8.5 (Rng::gen::<f32/f64>, gen_range for u32, SliceRandom::shuffle /
//     choose), rand_xoshiro 0.6.0 (Xoroshiro128StarStar, SplitMix64 seeding).  Call sites in the
//     reference: crates/step_sim/src/env.rs:121 (shuffle), crates/step_sim/src/runner.rs:53 and
//     rust/src/step_sim.rs:73 (seed_from_u64), crates/step_sim/src/agents/random_agent.rs:91-101.
//     PARITY UNPINNED: no reference test pins a value produced through these, and the crates'
//     sources are not available offline, so these are restated from the published algorithm only.
//
//  2. Philox4x32-10 counter RNG keyed (env, step, agent, block): this is the *new framework's* RNG
//     (BASELINE.json north_star) restated on the CPU so that agent-driven GPU runs can be checked
//     draw-for-draw against the oracle.  Known-answer vectors from the Random123 distribution are
//     checked in tests/test_oracle_rng.py.
#pragma once
#include <cstdint>
#include <cmath>
#include <utility>
#include <vector>

namespace oracle {

// ------------------------------------------------------------------------------------------------
// SplitMix64 (rand_xoshiro 0.6.0 `SplitMix64`; used by `seed_from_u64` of every xoshiro generator)
struct SplitMix64 {
    uint64_t x;
    explicit SplitMix64(uint64_t seed) : x(seed) {}
    uint64_t next_u64() {
        x += 0x9e3779b97f4a7c15ULL;
        uint64_t z = x;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    }
};

static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

// Xoroshiro128** (rand_xoshiro 0.6.0).  seed_from_u64 fills (s0, s1) from two SplitMix64 outputs.
struct Xoroshiro128StarStar {
    uint64_t s0, s1;
    Xoroshiro128StarStar() : s0(1), s1(2) {}
    static Xoroshiro128StarStar seed_from_u64(uint64_t seed) {
        SplitMix64 sm(seed);
        Xoroshiro128StarStar r;
        r.s0 = sm.next_u64();
        r.s1 = sm.next_u64();
        return r;
    }
    uint64_t next_u64() {
        const uint64_t r = rotl64(s0 * 5, 7) * 9;
        uint64_t t = s1 ^ s0;
        s0 = rotl64(s0, 24) ^ t ^ (t << 16);
        s1 = rotl64(t, 37);
        return r;
    }
    // rand_xoshiro 0.6.0 xoroshiro128starstar.rs: `fn next_u32(&mut self) -> u32 { self.next_u64() as u32 }` — the LOW
    // half (the ** scrambler has no weak low bits; only the `+` variants take `>> 32`).  Recalled from the crate source,
    // which is not available offline: the generator itself is pinned on the crate's own known-answer vector
    // (tests/test_oracle_rng.py::test_xoroshiro128starstar_public_vector), this truncation is NOT pinned.
    uint32_t next_u32() { return (uint32_t)next_u64(); }
};

// rand 0.8.5 `Standard` floats: 24 / 53 random mantissa bits scaled into [0, 1)
template <class R> static inline float gen_f32(R& rng) {
    return (float)(rng.next_u32() >> 8) * (1.0f / 16777216.0f);
}
template <class R> static inline double gen_f64(R& rng) {
    return (double)(rng.next_u64() >> 11) * (1.0 / 9007199254740992.0);
}

// rand 0.8.5 UniformInt<u32>::sample_single(low, high) for the half-open range low..high
// (widening multiply with a conservative rejection zone).
template <class R> static inline uint32_t gen_range_u32(R& rng, uint32_t low, uint32_t high) {
    const uint32_t range = high - low;  // caller guarantees low < high
    if (range == 0) return rng.next_u32();
    int lz = __builtin_clz(range);
    const uint32_t zone = (range << lz) - 1u;
    for (;;) {
        const uint32_t v = rng.next_u32();
        const uint64_t m = (uint64_t)v * (uint64_t)range;
        const uint32_t lo = (uint32_t)m;
        if (lo <= zone) return low + (uint32_t)(m >> 32);
    }
}

// rand 0.8.5 SliceRandom::shuffle: Fisher-Yates from the back, index drawn with gen_range(0..i+1)
template <class R, class T> static inline void shuffle(R& rng, std::vector<T>& v) {
    for (size_t i = v.size(); i > 1; --i) {
        const uint32_t j = gen_range_u32(rng, 0, (uint32_t)i);
        std::swap(v[i - 1], v[j]);
    }
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Mirrors bourse_b200/csrc/philox.cuh bit for bit.
struct Philox4 {
    uint32_t v[4];
};

static inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                    uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0;
        const uint64_t p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return Philox4{{c0, c1, c2, c3}};
}

// Shared conventions of the batched simulator's draws (see DESIGN.md "RNG contract"):
//   counter = (global env id, step, agent slot, block);  key = (seed lo, seed hi)
static const uint32_t PHILOX_SLOT_SHUFFLE = 0xFFFFFFFFu;   // agent slot used by the per-step shuffle
static const uint32_t PHILOX_SLOT_CANCEL = 0x80000000u;    // | group index: momentum cancel sweep

static inline float u32_to_f32_unit(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }
static inline double u64_to_f64_unit(uint32_t hi, uint32_t lo) {
    const uint64_t x = ((uint64_t)hi << 32) | lo;
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}
// unbiased-enough range draw without a rejection loop: floor(r * n / 2^32)
static inline uint32_t mulhi_range(uint32_t r, uint32_t n) { return (uint32_t)(((uint64_t)r * n) >> 32); }

}  // namespace oracle
