// Philox4x32-10 counter-based RNG (Salmon et al., SC'11) and the draw conventions of the batched
// simulator ("RNG contract" in DESIGN.md).  Replaces the reference's single shared
// Xoroshiro128StarStar stream (crates/step_sim/src/runner.rs:53-66) for in-kernel agents: every draw
// is addressed by (seed; global env id, step, agent slot, block) so lanes draw independently and
// results do not depend on how envs are sharded over GPUs.
#pragma once
#include <cstdint>

namespace bb {

#define PHILOX_SLOT_SHUFFLE 0xFFFFFFFFu
#define PHILOX_SLOT_CANCEL 0x80000000u

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ float u32_to_f32_unit(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }
__device__ __forceinline__ double u64_to_f64_unit(uint32_t hi, uint32_t lo) {
    const uint64_t x = ((uint64_t)hi << 32) | lo;
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ uint32_t mulhi_range(uint32_t r, uint32_t n) { return __umulhi(r, n); }

// Xoroshiro128** + the rand-0.8.5 style range draw, used for the per-env shuffle in Env mode so that a
// StepEnv(seed, ...) reproduces the stream the reference seeds at rust/src/step_sim.rs:73 and consumes
// at crates/step_sim/src/env.rs:121 (parity unpinned: restated from the published algorithms).
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
__device__ __forceinline__ uint64_t xoroshiro_next(uint64_t& s0, uint64_t& s1) {
    const uint64_t r = rotl64(s0 * 5ULL, 7) * 9ULL;
    const uint64_t t = s1 ^ s0;
    s0 = rotl64(s0, 24) ^ t ^ (t << 16);
    s1 = rotl64(t, 37);
    return r;
}
__device__ __forceinline__ uint32_t xoroshiro_range(uint64_t& s0, uint64_t& s1, uint32_t range) {
    const uint32_t zone = (range << __clz(range)) - 1u;
    for (;;) {
        const uint32_t v = (uint32_t)xoroshiro_next(s0, s1);  // next_u32 = low half (rand_xoshiro 0.6.0 `next_u64() as u32`)
        const uint32_t lo = v * range;
        if (lo <= zone) return __umulhi(v, range);
    }
}

}  // namespace bb
