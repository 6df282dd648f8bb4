"""CPU suite: known-answer vectors for the RNGs the oracle (and the CUDA path) use."""
import numpy as np


def _philox(oracle, ctr, key):
    out = np.zeros(4, np.uint32)
    oracle.lib().orc_philox(*ctr, *key, out.ctypes.data)
    return [int(x) for x in out]


def test_philox4x32_10_known_answers(oracle):
    """Random123 kat_vectors for philox4x32-10."""
    assert _philox(oracle, (0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xFFFFFFFF
    assert _philox(oracle, (f, f, f, f), (f, f)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _philox(oracle, (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_xoroshiro_seeding_matches_splitmix(oracle):
    """Xoroshiro128** seeded through SplitMix64 (rand_xoshiro 0.6.0 seed_from_u64): recompute in Python."""
    M = (1 << 64) - 1

    def splitmix(x):
        x = (x + 0x9e3779b97f4a7c15) & M
        z = x
        z = ((z ^ (z >> 30)) * 0xbf58476d1ce4e5b9) & M
        z = ((z ^ (z >> 27)) * 0x94d049bb133111eb) & M
        return x, z ^ (z >> 31)

    def rotl(x, k):
        return ((x << k) | (x >> (64 - k))) & M

    for seed in (0, 101, 2**63 + 5):
        x, s0 = splitmix(seed)
        x, s1 = splitmix(x)
        exp = []
        for _ in range(8):
            exp.append((rotl((s0 * 5) & M, 7) * 9) & M)
            t = s1 ^ s0
            s0 = rotl(s0, 24) ^ t ^ ((t << 16) & M)
            s1 = rotl(t, 37)
        out = np.zeros(8, np.uint64)
        oracle.lib().orc_xoroshiro(seed, 8, out.ctypes.data)
        assert [int(v) for v in out] == exp
    # SplitMix64 reference output for seed 1234567 (public test vector of the algorithm)
    x, z = splitmix(1234567)
    assert z == 6457827717110365317


def test_xoroshiro128starstar_public_vector(oracle):
    """The generator pinned on the public known-answer vector of xoroshiro128** for state {1, 2} — the values rand_xoshiro
    0.6.0 checks its own implementation against (xoroshiro128starstar.rs `reference` test, taken from Vigna's
    xoroshiro128starstar.c), hard-coded here rather than recomputed.  What stays UNPINNED is how rand 0.8.5 turns these
    words into draws: `next_u32` = low half (`next_u64() as u32`, as recalled from the crate source; the oracle and the CUDA
    kernels implement the same choice), f32 = 24 high bits of that u32, gen_range = widening multiply with rejection."""
    exp = [5760, 97769243520, 9706862127477703552, 9223447511460779954, 8358291023205304566, 15695619998649302768,
           8517900938696309774, 16586480348202605369, 6959129367028440372, 16822147227405758281]
    n = len(exp)
    o64, o32, of, orr = np.zeros(n, np.uint64), np.zeros(n, np.uint32), np.zeros(n, np.float32), np.zeros(n, np.uint32)
    oracle.lib().orc_xoroshiro_state(1, 2, n, o64.ctypes.data, o32.ctypes.data, of.ctypes.data, 100, orr.ctypes.data)
    assert [int(v) for v in o64] == exp
    # derived draws, stated independently of the C++ code: low half, 24-bit mantissa float, widening-multiply range
    assert [int(v) for v in o32] == [e & 0xFFFFFFFF for e in exp]
    assert [float(v) for v in of] == [((e & 0xFFFFFFFF) >> 8) / 16777216.0 for e in exp]
    # rand 0.8.5 UniformInt::sample_single: zone = (range << lz) - 1, reject while lo(v * range) > zone; the draws share
    # one word stream, so restate the loop over the low halves of a longer run of the generator
    m = 64
    w64, w32, wf, wr = np.zeros(m, np.uint64), np.zeros(m, np.uint32), np.zeros(m, np.float32), np.zeros(m, np.uint32)
    oracle.lib().orc_xoroshiro_state(1, 2, m, w64.ctypes.data, w32.ctypes.data, wf.ctypes.data, 100, wr.ctypes.data)
    zone = ((100 << (32 - (100).bit_length())) & 0xFFFFFFFF) - 1
    # the range draws consume their own copy of the stream; regenerate enough words for m draws with rejections
    M = (1 << 64) - 1
    s0, s1 = 1, 2

    def nxt():
        nonlocal s0, s1
        r = ((((s0 * 5) & M) << 7 | ((s0 * 5) & M) >> 57) & M) * 9 & M
        t = s1 ^ s0
        s0 = ((s0 << 24 | s0 >> 40) & M) ^ t ^ ((t << 16) & M)
        s1 = (t << 37 | t >> 27) & M
        return r & 0xFFFFFFFF

    want = []
    while len(want) < m:
        v = nxt()
        if ((v * 100) & 0xFFFFFFFF) <= zone:
            want.append((v * 100) >> 32)
    assert [int(v) for v in wr] == want and max(want) < 100


def test_shuffle_is_a_permutation(oracle):
    for n in (0, 1, 2, 17, 100):
        out = np.zeros(max(n, 1), np.uint32)
        oracle.lib().orc_shuffle_perm(101, n, out.ctypes.data)
        assert sorted(out[:n]) == list(range(n))
