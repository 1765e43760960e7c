"""Seeded fuzz of podsdf_eig_sym_top_* (all eigenvalues + the nm leading eigenvectors; this library's two-stage
tridiagonalisation, bulge chasing with mailbox hand-offs, multisection, inverse iteration, back-transformation -- or the
library solver when the route declines) against numpy: random sizes 256 .. 2600 (odd, multiples of 32 +- 1, sizes that
switch the panel-QR variant), random nm, four kinds of spectra.  Run on a GPU box; prints one JSON line."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import podsdf_loader  # noqa: E402

P = podsdf_loader.load()
capi = P._capi


def matrix(rs, N, kind):
    Q, _ = np.linalg.qr(rs.standard_normal((N, N)))
    if kind == "decay":
        lam = 10.0 ** rs.uniform(0, 3) * np.arange(1, N + 1, dtype=float) ** -rs.uniform(0.5, 2.0)
    elif kind == "lowrank":
        r = int(rs.randint(1, max(2, N // 4)))
        lam = np.r_[np.sort(rs.uniform(1.0, 100.0, r))[::-1], 1e-6 * rs.uniform(0.5, 1.0, N - r)]
    elif kind == "clustered":
        lam = np.sort(rs.uniform(0.0, 1.0, N))[::-1]
        lam[:4] = lam[0] * (1.0 + 1e-12 * np.arange(4))
    else:  # indefinite
        lam = rs.standard_normal(N) * 10.0
    C = (Q * lam[None, :]) @ Q.T
    return (C + C.T) / 2


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
    sizes = [256, 257, 287, 288, 289, 300, 511, 512, 513, 777, 1000, 1023, 1025, 1500, 2047, 2049, 2113, 2600]
    worst = {"lam": (0.0, None), "resid": (0.0, None), "orth": (0.0, None)}
    routes = {0: 0, 1: 0}
    for case in range(n):
        N = int(rs.choice(sizes))
        kind = str(rs.choice(["decay", "lowrank", "clustered", "indef"]))
        nm = int(rs.randint(1, N // 4 + 1))
        C = matrix(rs, N, kind)
        lam, V, path = capi.eig_sym_top(C, nm)
        routes[int(path)] += 1
        ref = np.linalg.eigvalsh(C)
        nrm = max(np.abs(ref).max(), 1e-300)
        top = lam[::-1][:nm]
        errs = {"lam": float(np.abs(lam - ref).max() / nrm),
                "resid": float(np.abs(C @ V.T - V.T * top[None, :]).max() / nrm),
                "orth": float(np.abs(V @ V.T - np.eye(nm)).max())}
        for k, v in errs.items():
            if not (v <= worst[k][0]):
                worst[k] = (v, dict(N=N, nm=nm, kind=kind, path=int(path)))
    print(json.dumps({"cases": n, "routes": routes, "worst": {k: {"err": v[0], "case": v[1]} for k, v in worst.items()}}))


if __name__ == "__main__":
    main()
