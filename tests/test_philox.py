"""The on-device generator on the host: Philox4x32-10 (oracle and library) against Random123's
known-answer vectors, the keyed MX2 expansion (library against oracle, avalanche / bit-independence /
sequential-slot statistics).  CPU only (sdb_philox_block / sdb_random_block are host code)."""
import ctypes as C

import numpy as np

from oracle import philox_oracle as po

# Random123 kat_vectors: "philox4x32 10 <ctr x4> <key x2> <expected x4>"
KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_oracle_philox_known_answers():
    for ctr, key, want in KAT:
        got = po.philox4x32_10(*[np.array([c]) for c in ctr], key[0], key[1])
        assert tuple(int(g[0]) for g in got) == want


def test_library_philox_known_answers():
    import sdeint_b200
    lib = sdeint_b200._lib.load()
    for ctr, key, want in KAT:
        out = (C.c_uint32 * 4)()
        seed = key[0] | (key[1] << 32)
        path = ctr[0] | (ctr[1] << 32)
        assert lib.sdb_philox_block(seed, path, ctr[2], ctr[3], C.byref(out)) == 0
        assert tuple(out) == want


def test_oracle_normals_moments():
    z = po.delta_w(400, 5, 1.0, 500, seed=123)
    assert abs(z.mean()) < 5e-3 and abs(z.var() - 1.0) < 1e-2
    assert abs((z ** 4).mean() - 3.0) < 0.05
    # distinct paths / steps / components are distinct streams
    assert len(np.unique(z)) == z.size


def test_library_random_block_matches_oracle():
    """sdb_random_block (the block the kernels turn into Box-Muller pair `slot`) against the numpy
    restatement: slot 0 = the plain Philox block, slot >= 1 = MX2 of the step's key block."""
    import sdeint_b200
    lib = sdeint_b200._lib.load()
    rng = np.random.default_rng(2026)
    for t in range(600):
        seed = int(rng.integers(0, 2 ** 63))
        path = int(rng.integers(0, 2 ** 40))
        step = int(rng.integers(0, 2 ** 31))
        slot = 0 if t % 6 == 0 else int(rng.integers(1, 5000))
        out = (C.c_uint32 * 4)()
        assert lib.sdb_random_block(seed, path, step, slot, C.byref(out)) == 0
        got = po.random_block(seed, np.array([path], dtype=np.uint64), np.array([step]), np.array([slot]))
        assert tuple(out) == tuple(int(g[0]) for g in got)
        if slot == 0:
            raw = (C.c_uint32 * 4)()
            lib.sdb_philox_block(seed, path, step, 0, C.byref(raw))
            assert tuple(raw) == tuple(out)


def _bits(words):
    return np.concatenate([((w[:, None] >> np.arange(32, dtype=np.uint64)) & np.uint64(1))
                           for w in words], axis=1).astype(np.float32)


def test_mx2_avalanche_and_independence():
    """MX2 under random keys: every output bit flips with probability 1/2 when the slot number
    changes (single-bit, double-bit and additive differences), flips of different output bits are
    uncorrelated, the blocks of consecutive slots are uncorrelated bit by bit, and no bit (nor the
    xor of one bit position over the four words -- the linear relation a bare multiply-add leaves)
    is biased.  Thresholds: max over the tested cells of a N(0,1) statistic, 6 sigma."""
    N = 1 << 17
    rng = np.random.default_rng(99)
    K = [rng.integers(0, 2 ** 32, N, dtype=np.uint64) for _ in range(4)]
    s = rng.integers(1, 4096, N, dtype=np.uint64)
    base = _bits(po.mx2_block(K, s))
    rt = np.sqrt(N)
    assert np.abs(base.mean(axis=0) - 0.5).max() * 2 * rt < 6.0
    others = [s ^ np.uint64(1 << b) for b in range(12)]
    others += [s ^ np.uint64((1 << a) | (1 << b)) for a, b in ((0, 1), (0, 5), (3, 7), (2, 11))]
    others += [s + np.uint64(dl) for dl in (1, 2, 3, 5, 10, 55)]
    for i, s2 in enumerate(others):
        flip = np.abs(_bits(po.mx2_block(K, s2)) - base)
        assert np.abs(flip.mean(axis=0) - 0.5).max() * 2 * rt < 6.0, i
        if i < 12 and i % 3 == 0:                       # bit-independence criterion
            f = flip * 2 - 1
            c = (f.T @ f) / N
            np.fill_diagonal(c, 0.0)
            assert np.abs(c).max() * rt < 6.5, i
    a = base * 2 - 1
    b = _bits(po.mx2_block(K, s + np.uint64(1))) * 2 - 1
    assert np.abs(a.T @ b).max() / rt < 6.5             # consecutive slots, all 128 x 128 bit pairs
    c = (a.T @ a) / N
    np.fill_diagonal(c, 0.0)
    assert np.abs(c).max() * rt < 6.5                   # within one block
    # MX2 ends on a multiply-add, whose LOW bits are weak across the four words taken together (bit 0 of
    # the four words xors to zero, bit 1 to a biased quadratic form of four bits, ...).  The normal
    # transform never takes the low bits of w0 / w2 (it uses bits 12.. and 9..), and the low bits of
    # w1 / w3 are the last bits of two 52-bit mantissas -- so the relation is checked where it matters:
    # the high bit positions of all four words, and the low bits of the pair (w1, w3) alone.
    assert np.all((base[:, 0] + base[:, 32] + base[:, 64] + base[:, 96]) % 2 == 0)
    for j in (12, 15, 16, 24, 31):
        x = (base[:, j] + base[:, 32 + j] + base[:, 64 + j] + base[:, 96 + j]) % 2
        assert abs(x.mean() - 0.5) * 2 * rt < 5.0, j
    for j in (0, 1, 2, 5):
        x = (base[:, 32 + j] + base[:, 96 + j]) % 2
        assert abs(x.mean() - 0.5) * 2 * rt < 5.0, j


def test_oracle_normals_across_slots():
    """Normals of ONE (path, step) across 40 slots (the keyed expansion) and across steps: sample
    correlations of z and of z^2 within 5.5 sigma for every pair."""
    P, Q = 40000, 80
    z = po.normals(7, np.arange(P, dtype=np.uint64)[:, None], 3, np.arange(Q)[None, :])
    assert abs(z.mean()) < 5 / np.sqrt(z.size) and abs(z.var() - 1) < 5 * np.sqrt(2 / z.size)
    for v in (z, z * z - 1.0):
        v = (v - v.mean(axis=0)) / v.std(axis=0)
        c = v.T @ v / P
        np.fill_diagonal(c, 0.0)
        assert np.abs(c).max() < 5.5 / np.sqrt(P)
