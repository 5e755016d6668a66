"""Philox4x32-10: oracle and the library's host entry point against Random123's
known-answer vectors.  CPU only (sdb_philox_block is host code)."""
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
