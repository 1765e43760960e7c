"""Stage-by-stage check of the device tridiagonalisation (csrc/sbr_kernels.cu) against numpy / LAPACK on a GPU box:
band matrix after stage 1, tridiagonal matrix after stage 2 (same spectrum as C), leading eigenpairs through the
back-transformation, and the timings at N = 10^4."""
import json
import os
import sys
import time

import numpy as np
from scipy.linalg import eigvalsh_tridiagonal

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import podsdf_loader  # noqa: E402

P = podsdf_loader.load()
capi = P._capi


def sym_matrix(N, rs, kind):
    if kind == "gram":                     # snapshot-Gram-like: decaying spectrum + noise floor
        r = min(N, 200)
        X = rs.standard_normal((r, N)) * (np.arange(1, r + 1)[:, None] ** -0.8)
        X = np.concatenate([X, 1e-3 * rs.standard_normal((N, N))])
        return X.T @ X / N
    A = rs.standard_normal((N, N))
    return (A + A.T) / 2


def band_dense(band, N, b=32):
    Bm = np.zeros((N, N))
    for dd in range(b + 1):
        v = band[:N - dd, dd]
        Bm[np.arange(dd, N), np.arange(N - dd)] = v
        Bm[np.arange(N - dd), np.arange(dd, N)] = v
    return Bm


def main():
    sizes = [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [300, 515, 1000, 2049]
    rs = np.random.RandomState(5)
    out = []
    for N in sizes:
        for kind in ("gram", "indef"):
            C = sym_matrix(N, rs, kind)
            lam = np.linalg.eigvalsh(C)
            nrm = np.abs(lam).max()
            d, e, band, ms = capi.sym_tridiag(C, want_band=(N <= 2100))
            rec = dict(N=N, kind=kind, ms=ms)
            if band is not None:
                rec["band_eig_err"] = float(np.abs(np.linalg.eigvalsh(band_dense(band, N)) - lam).max() / nrm)
            rec["tri_eig_err"] = float(np.abs(eigvalsh_tridiagonal(d, e) - lam).max() / nrm)
            if kind == "gram":
                nm = min(100, N // 4)
                l2, V, path = capi.eig_sym_top(C, nm)
                R = C @ V.T - V.T * l2[::-1][:nm][None, :]
                rec.update(path=path, top_lam_err=float(np.abs(l2 - lam).max() / nrm),
                           top_resid=float(np.abs(R).max() / nrm), top_orth=float(np.abs(V @ V.T - np.eye(nm)).max()))
            out.append(rec)
            print(json.dumps(rec), flush=True)
    if "--big" in sys.argv:
        N = 10000
        C = sym_matrix(N, rs, "gram")
        for rep in range(2):
            t0 = time.time()
            d, e, _, ms = capi.sym_tridiag(C)
            wall = time.time() - t0
        lam = np.linalg.eigvalsh(C)
        rec = dict(N=N, ms=ms, wall_s=wall, tri_eig_err=float(np.abs(eigvalsh_tridiagonal(d, e) - lam).max() / lam.max()))
        for rep in range(2):
            t0 = time.time()
            l2, V, path = capi.eig_sym_top(C, 100)
            rec["eig_top_wall_s"] = time.time() - t0
        R = C @ V.T - V.T * l2[::-1][:100][None, :]
        rec.update(path=path, top_lam_err=float(np.abs(l2 - lam).max() / lam.max()), top_resid=float(np.abs(R).max() / lam.max()),
                   top_orth=float(np.abs(V @ V.T - np.eye(100)).max()))
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
