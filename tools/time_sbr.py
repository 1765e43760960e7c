"""One tridiagonalisation + leading-pairs solve at size N (default 10^4) for kernel-level timing under ncu."""
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
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rs = np.random.RandomState(5)
r = min(N, 200)
X = rs.standard_normal((r, N)) * (np.arange(1, r + 1)[:, None] ** -0.8)
X = np.concatenate([X, 1e-3 * rs.standard_normal((N, N))])
C = X.T @ X / N
for _ in range(reps):
    d, e, _, ms = capi.sym_tridiag(C)
lam = eigvalsh_tridiagonal(d, e)
t0 = time.time()
l2, V, path = capi.eig_sym_top(C, 100)
wall = time.time() - t0
R = C @ V.T - V.T * l2[::-1][:100][None, :]
print(json.dumps(dict(N=N, stage_ms=ms, eig_top_wall_s=wall, path=path, lam_vs_tri=float(np.abs(l2 - lam).max() / lam.max()),
                      top_resid=float(np.abs(R).max() / lam.max()), top_orth=float(np.abs(V @ V.T - np.eye(100)).max()))))
