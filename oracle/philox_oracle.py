"""Numpy restatement of the on-device noise source (TEST INFRASTRUCTURE).

Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11; Random123 v1.14 `philox.h`): not part of
the reference (which uses numpy's PCG64) -- it is this repository's own generator, so the
oracle restates the PUBLISHED algorithm and is pinned by Random123's known-answer vectors
(`kat_vectors`: three philox4x32 10-round entries, checked in tests/test_philox.py).

Normal transform (sdeint_b200/csrc/sdb_kernels.cuh, sdb_normal_pair): one 128-bit block
(w0..w3) -> a = (w0 >> 12) 2^32 + w1 (52 bits);  u1 = (2 a + 1) 2^-53 in (0,1);  rad = sqrt(-2 ln u1);
t = (((w2 >> 9) mod 2^20) 2^32 + w3) 2^-52 in [0,1);  (c, s) = (cos, sin)(pi t / 4);  o = w2 >> 29;
z0 = (o&2 ? -1 : 1) rad (o&1 ? s : c),  z1 = (o&4 ? -1 : 1) rad (o&1 ? c : s)
(Box-Muller with the angle folded into one octant and unfolded by three random bits).
Normal number q of (path, step) is element q&1 of random block q>>1 of (seed, path, step).

Random block `slot` of (seed, path, step) (sdb_random_block in csrc/sdb_kernels.cuh):
  slot 0      the Philox4x32-10 block with counter (path_lo, path_hi, step, 0), key (seed_lo, seed_hi);
  slot s >= 1 MX2(K, s), K = the Philox4x32-10 block with counter (path_lo, path_hi, step, 0xFFFFFFFF):
              x_i = K_i + s G_i;  x_i ^= x_i >> 15;  y_i = x_i C_i + x_{i+1};  y_i ^= y_i >> 13;
              out_i = y_i C_{i+1} + y_{i+1}                                  (indices mod 4, 32-bit wrap)
              (the last step is a multiply-add: its high bits are the well-mixed ones, and the transform
              above takes the leading mantissa bits / octant bits from the top of w0, w2)
(the 64-bit products of Philox share the FP64 pipe's issue resource on sm_100a, so only two Philox
blocks are computed per (path, step); the other slots are a 32-bit multiply-xorshift expansion of a
key block that is never output itself).
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy arrays of uint32-valued integers."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)) & MASK
        n1 = p1 & MASK
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)) & MASK
        n3 = p0 & MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


KEYSLOT = 0xFFFFFFFF
MX_G = (0x9E3779B1, 0x85EBCA77, 0xC2B2AE3D, 0x27D4EB2F)
MX_C = (0xED5AD4BB, 0xAC4C1B51, 0x31848BAB, 0x7FEB352D)


def mx2_block(K, s):
    """MX2 expansion (module doc); K = 4 arrays of uint32-valued uint64, s = slot array."""
    s = np.asarray(s, dtype=np.uint64) & MASK
    x = [(K[i] + s * np.uint64(MX_G[i])) & MASK for i in range(4)]
    x = [v ^ (v >> np.uint64(15)) for v in x]
    y = [(x[i] * np.uint64(MX_C[i]) + x[(i + 1) & 3]) & MASK for i in range(4)]
    y = [v ^ (v >> np.uint64(13)) for v in y]
    return tuple((y[i] * np.uint64(MX_C[(i + 1) & 3]) + y[(i + 1) & 3]) & MASK for i in range(4))


def random_block(seed, path, step, slot):
    """The 128-bit block of (seed, path, step, slot) as four uint32-valued arrays."""
    path = np.asarray(path, dtype=np.uint64)
    step = np.asarray(step, dtype=np.uint64)
    slot = np.asarray(slot, dtype=np.uint64)
    path, step, slot = np.broadcast_arrays(path, step, slot)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    first = philox4x32_10(path & MASK, path >> np.uint64(32), step, np.zeros_like(slot), k0, k1)
    if not np.any(slot != 0):
        return first
    K = philox4x32_10(path & MASK, path >> np.uint64(32), step, np.full_like(slot, KEYSLOT), k0, k1)
    rest = mx2_block(K, slot)
    return tuple(np.where(slot == 0, a, b) for a, b in zip(first, rest))


def normal_pairs(seed, path, step, slot):
    """(z0, z1) for broadcastable integer arrays path, step, slot (spec in the module doc)."""
    r0, r1, r2, r3 = random_block(seed, path, step, slot)
    a = ((r0 >> np.uint64(12)) << np.uint64(32)) | r1            # 52 bits
    x = (a << np.uint64(1)) | np.uint64(1)                        # 53-bit odd integer
    u1 = x.astype(np.float64) * 2.0 ** -53                        # exact, in (0, 1)
    rad = np.sqrt(-2.0 * np.log(u1))
    t = ((((r2 >> np.uint64(9)) & np.uint64((1 << 20) - 1)) << np.uint64(32)) | r3).astype(np.float64) * 2.0 ** -52
    c, s = np.cos(np.pi * t / 4), np.sin(np.pi * t / 4)
    o = (r2 >> np.uint64(29)).astype(np.int64)
    swap, neg0, neg1 = (o & 1) == 1, (o & 2) != 0, (o & 4) != 0
    a0 = np.where(swap, s, c)
    b0 = np.where(swap, c, s)
    return np.where(neg0, -rad * a0, rad * a0), np.where(neg1, -rad * b0, rad * b0)


def normals(seed, path, step, q):
    """Standard normal number q (array) of (path, step)."""
    q = np.asarray(q, dtype=np.int64)
    z0, z1 = normal_pairs(seed, path, step, q >> 1)
    return np.where((q & 1) == 0, z0, z1)


def delta_w(steps, m, h, paths, seed, path_id0=0):
    """dW[p, n, j] as sdb_delta_w / the fused kernels generate it."""
    p = (np.arange(paths, dtype=np.uint64) + np.uint64(path_id0))[:, None, None]
    n = np.arange(steps)[None, :, None]
    j = np.arange(m)[None, None, :]
    return np.sqrt(h) * normals(seed, p, n, j)


def kpw_normals(steps, m, n_terms, paths, seed, path_id0=0):
    """X[k], Y[k] (n, P, N, m) as the device draws them for Ikpw/Iwik."""
    p = (np.arange(paths, dtype=np.uint64) + np.uint64(path_id0))[:, None, None]
    n = np.arange(steps)[None, :, None]
    j = np.arange(m)[None, None, :]
    X = np.stack([normals(seed, p, n, m + 2 * k * m + j) for k in range(n_terms)])
    Y = np.stack([normals(seed, p, n, m + (2 * k + 1) * m + j) for k in range(n_terms)])
    return X, Y


def wik_tail_normals(steps, m, n_terms, paths, seed, path_id0=0):
    M = m * (m - 1) // 2
    p = (np.arange(paths, dtype=np.uint64) + np.uint64(path_id0))[:, None, None]
    n = np.arange(steps)[None, :, None]
    r = np.arange(M)[None, None, :]
    return normals(seed, p, n, m + 2 * n_terms * m + r)
