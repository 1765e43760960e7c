"""Seeded fuzz of the PODFS device entries against numpy (run on a GPU box; not part of the test-suite):
random (M, N, nm), both snapshot layouts, weighted and unweighted Gram, eigen-solve identities, projection,
DFT of the temporal modes.  Prints one JSON line with the worst errors."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import podsdf_loader  # noqa: E402

P = podsdf_loader.load()
capi = P._capi


def one_case(rs):
    ncomp = int(rs.choice([1, 3]))
    pts = int(rs.choice([1, 5, 16, 43, 128, 129, 1000, 4099]))
    M = ncomp * pts
    N = int(rs.choice([1, 2, 3, 8, 17, 64, 127, 128, 129, 200, 257, 513]))
    nm = int(rs.randint(1, min(N, 130) + 1))
    A = rs.standard_normal((M, N)) * rs.uniform(0.1, 10.0, (M, 1))
    layout = int(rs.randint(0, 2))
    arr = A if layout == 0 else np.ascontiguousarray(A.T)
    w = rs.uniform(0.5, 2.0, pts) if rs.rand() < 0.5 else None
    scale = float(rs.uniform(0.1, 2.0))
    C = capi.gram(arr, layout, w, scale)
    wrow = np.tile(w, ncomp) if w is not None else np.ones(M)
    Cref = scale * (A.T * wrow[None, :]) @ A
    e_gram = np.abs(C - Cref).max() / max(np.abs(Cref).max(), 1e-300)
    sym = float(np.abs(C - C.T).max())
    lam, V = capi.eig_sym(Cref)
    nrm = max(np.abs(lam).max(), 1e-300)
    e_eig = max(np.abs(Cref @ V - V * lam[None, :]).max() / nrm, np.abs(V.T @ V - np.eye(N)).max(),
                np.abs(lam - np.linalg.eigvalsh(Cref)).max() / nrm)
    T = rs.standard_normal((N, nm))
    dinv = rs.uniform(0.5, 2.0, nm)
    Phi = capi.project(arr, T, dinv, nm, scale, layout)
    Pref = scale * (A @ T) * dinv[None, :]
    e_proj = np.abs(Phi - Pref).max() / max(np.abs(Pref).max(), 1e-300)
    c = capi.dft_modes(T, nm)
    cref = np.fft.fftshift(np.fft.fft(T, axis=0), axes=0) / N
    e_dft = np.abs(c - cref).max() / max(np.abs(cref).max(), 1e-300)
    return dict(gram=float(e_gram), sym=sym, eig=float(e_eig), proj=float(e_proj), dft=float(e_dft)), \
        dict(M=M, N=N, nm=nm, layout=layout, weighted=w is not None)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    worst = {}
    for _ in range(n):
        errs, cfg = one_case(rs)
        for k, v in errs.items():
            if v > worst.get(k, (-1.0, None))[0]:
                worst[k] = (v, cfg)
        if max(errs["gram"], errs["eig"], errs["proj"], errs["dft"]) > 1e-11 or errs["sym"] != 0.0:
            print(json.dumps({"FAIL": cfg, "errs": errs}))
    print(json.dumps({"cases": n, "worst": {k: {"err": v[0], "case": v[1]} for k, v in worst.items()}}))


if __name__ == "__main__":
    main()
