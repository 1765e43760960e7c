"""Seeded fuzz of the DF device path against a plain numpy contraction (run on a GPU box; not part of the
test-suite):  random plane sizes, independent filter half-widths per axis (incl. 0 and the wide/tuned tap
counts), coefficient masks (none / diagonal / adapt1d pattern / full), rotation, chunk sizes, plan impls,
injected and generated streams with random time offsets.  Prints one JSON line with the worst error."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import podsdf_loader  # noqa: E402
from oracle import philox as ophilox  # noqa: E402  (checker only)

P = podsdf_loader.load()
df = P.digitalfilters


def reference(slabs, ns, J, K, bx, by, bz, coef, mask, rot):
    """slabs (ns+2nfx, 3, Jp, Kp) -> planes (ns, 3, J, K): sliding-window contraction, then M x + m, then rot."""
    swv = np.lib.stride_tricks.sliding_window_view
    y = swv(slabs, len(bx), axis=0)[:ns] @ bx           # (ns, 3, Jp, Kp)
    y = swv(y, len(by), axis=2) @ by                    # (ns, 3, J, Kp)
    y = swv(y, len(bz), axis=3) @ bz                    # (ns, 3, J, K)
    if mask:
        c = coef.reshape(12, J, K)
        M, m = c[:9].reshape(3, 3, J, K), c[9:]
        y = np.einsum("abjk,tbjk->tajk", M, y) + m[None]
    if rot is not None:
        y = np.einsum("ab,tbjk->tajk", rot, y)
    return y


def one_case(rs, idx):
    taps = [0, 1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 24, 32]
    nfx, nfy, nfz = (int(rs.choice(taps)) for _ in range(3))
    if rs.rand() < 0.4:
        nfy = nfz = nfx                                     # square filters hit the tuned plane kernels
    J = int(rs.choice([1, 2, 7, 15, 16, 17, 31, 33, 48, 64, 65, 100, 130]))
    K = int(rs.choice([1, 3, 8, 31, 63, 64, 65, 90, 127, 128, 129, 200]))
    while (J + 2 * nfy) * (K + 2 * nfz) * (2 * nfx + 40) * 3 > 6e6:      # keep the numpy reference quick
        J, K = max(1, J // 2), max(1, K // 2)
    ns = int(rs.randint(1, 30))
    b = [df._taps(n, max(n / 2.0, 0.6)) for n in (nfx, nfy, nfz)]
    kind = int(rs.randint(0, 4))
    coef = rs.uniform(-1, 1, (12, J * K))
    mask = [0, 0b111000000000 | 0b100010001, 0b111000000000 | 0b101010001 | (1 << 6), 0xFFF][kind]
    if kind in (1, 2):
        coef[[i for i in range(12) if not (mask >> i) & 1]] = 0.0
    rot = None
    if rs.rand() < 0.3:
        q, _ = np.linalg.qr(rs.standard_normal((3, 3)))
        rot = q
    impl = int(rs.choice([0, 0, 1, 2])) if all(n in (4, 6, 8, 10, 12, 16, 24, 32) for n in (nfx,)) else int(rs.choice([0, 1]))
    chunk = int(rs.choice([8, 16, 24, 512]))
    p = df._plan(J, K, nfx, nfy, nfz, *b)
    try:
        p.set_impl(impl)
    except Exception:
        impl = 0
        p.set_impl(0)
    p.set_chunk(chunk)
    p.set_coefficients(coef if mask else None, mask)
    p.set_rotation(rot)
    seeded = rs.rand() < 0.5
    try:
        if seeded:
            seed, t0 = int(rs.randint(0, 2 ** 62)), int(rs.randint(0, 5000))
            slabs = ophilox.slabs(seed, t0, ns + 2 * nfx, J + 2 * nfy, K + 2 * nfz)
            got = p.generate_seeded(seed, t0, ns)
        else:
            slabs = rs.uniform(-1.7, 1.7, (ns + 2 * nfx, 3, J + 2 * nfy, K + 2 * nfz))
            got = p.generate_injected(slabs)
    except P._capi.PodsdfError as e:
        if impl == 2:                                        # tuned-only plan without a tuned kernel: loud, expected
            return None
        raise RuntimeError("case %d %s failed: %s" % (idx, (J, K, nfx, nfy, nfz, kind, impl, chunk, seeded), e))
    ref = reference(slabs, ns, J, K, *b, coef, mask, rot)
    scale = max(np.sqrt(np.mean(ref ** 2)), 1e-300)
    err = float(np.abs(got - ref).max() / scale)
    return err, dict(J=J, K=K, nf=(nfx, nfy, nfz), ns=ns, kind=kind, rot=rot is not None, impl=impl, chunk=chunk, seeded=seeded)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    worst, worst_cfg, ran = 0.0, None, 0
    for i in range(n):
        r = one_case(rs, i)
        if r is None:
            continue
        ran += 1
        if r[0] > worst:
            worst, worst_cfg = r
        if r[0] > 1e-11:
            print(json.dumps({"FAIL": r[1], "err": r[0]}))
    print(json.dumps({"cases": ran, "worst_rel_err": worst, "worst_case": worst_cfg}))


if __name__ == "__main__":
    main()
