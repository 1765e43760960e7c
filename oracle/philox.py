"""Philox4x32-10 counter-based streams of the product's *generated-RNG* mode, restated
in numpy -- TEST INFRASTRUCTURE ONLY (checker for `podsdf_df_random_slabs_*`).

The reference draws from the unseeded global MT19937 (DF:1344, 1361-1366, 1460-1467), so
there is no reference-defined stream to match; the product defines its own, and this is
their CPU restatement (Salmon et al. 2011, Random123 constants).

Layout (shared with pods-digital-filter_b200/csrc/philox.cuh):
  key     = (seed & 0xffffffff, seed >> 32);  pair = idx // 2, half = idx % 2, idx = p*Kp + q in the padded
            plane (Jp x Kp), s = global slab index, comp in 0..2
  kind "u32" (PODSDF_RNG_U32, the default): one call = ONE point x FOUR consecutive slabs
    counter = (idx & 0xffffffff, idx >> 32, sq & 0xffffffff, comp + 4*(sq >> 32)),  sq = s // 4
    word r[s % 4];  v = 1 + (r + 1/2) 2^-32 built bitwise in the mantissa of [1, 2);
    x = (v - 1.5) * (2 sqrt(3))                    (exact subtraction, one rounding)
  kind "u53" (PODSDF_RNG_U53, round 1's stream): one call = 2 points of one slab
    counter = (pair & 0xffffffff, pair >> 32, s & 0xffffffff, comp + 4*(s >> 32))
    output words r0..r3:  (r0, r1) -> point idx = 2*pair,  (r2, r3) -> point 2*pair + 1
    u = ((a >> 5) * 2^26 + (b >> 6)) / 2^53        (numpy's 53-bit double, random_sample)
    x = -sqrt(3) + (2 sqrt(3)) * u                 (np.random.uniform(low, high): low + (high-low)*u)
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
SQRT3 = np.sqrt(3.0)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over uint64 arrays holding 32-bit values.  Returns r0..r3."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def u32_to_uniform(r):
    """32 random bits (uint64 array holding 32-bit values) -> U(-sqrt3, sqrt3), as philox_u32_to_uniform."""
    bits = np.uint64(0x3FF0000000000000) | (np.asarray(r, dtype=np.uint64) << np.uint64(20)) | np.uint64(0x80000)
    v = bits.view(np.float64)
    return (v - 1.5) * (SQRT3 + SQRT3)


def slabs(seed, s0, n, Jp, Kp, kind="u32"):
    """Random slabs [s0, s0+n) of all three components: (n, 3, Jp, Kp) float64 in U(-sqrt3, sqrt3)."""
    if kind == "u32":
        return _slabs_u32(seed, s0, n, Jp, Kp)
    assert kind == "u53"
    npts = Jp * Kp
    npair = (npts + 1) // 2
    pair = np.arange(npair, dtype=np.uint64)
    out = np.empty((n, 3, Jp, Kp))
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    rng = SQRT3 - (-SQRT3)

    def uni(a, b):
        u = ((a >> np.uint64(5)).astype(np.float64) * 67108864.0
             + (b >> np.uint64(6)).astype(np.float64)) * (1.0 / 9007199254740992.0)
        return -SQRT3 + rng * u

    for i in range(n):
        s = s0 + i
        for comp in range(3):
            r = philox4x32_10(pair & MASK, pair >> np.uint64(32),
                              np.full_like(pair, s & 0xFFFFFFFF),
                              np.full_like(pair, comp + 4 * (s >> 32)), k0, k1)
            flat = np.empty(2 * npair)
            flat[0::2] = uni(r[0], r[1])
            flat[1::2] = uni(r[2], r[3])
            out[i, comp] = flat[:npts].reshape(Jp, Kp)
    return out


def _slabs_u32(seed, s0, n, Jp, Kp):
    npts = Jp * Kp
    idx = np.arange(npts, dtype=np.uint64)
    out = np.empty((n, 3, Jp, Kp))
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    cache = {}
    for i in range(n):
        s = s0 + i
        sq, w = s >> 2, s & 3
        for comp in range(3):
            if (sq, comp) not in cache:
                cache[(sq, comp)] = philox4x32_10(idx & MASK, idx >> np.uint64(32),
                                                  np.full_like(idx, sq & 0xFFFFFFFF),
                                                  np.full_like(idx, comp + 4 * (sq >> 32)), k0, k1)
            out[i, comp] = u32_to_uniform(cache[(sq, comp)][w]).reshape(Jp, Kp)
        cache = {k: v for k, v in cache.items() if k[0] >= sq}
    return out
