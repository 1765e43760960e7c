"""podsdf_eig_sym_dev against podsdf_eig_sym_top_dev (nm = 100) at N = 10^4 on config 4's spectrum."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import podsdf_loader  # noqa: E402

P = podsdf_loader.load()
capi, L = P._capi, P._capi.lib()
N, nm = 10000, 100
dev = torch.device("cuda", 0)
t = torch.arange(N, dtype=torch.float64, device=dev) / N
r = torch.arange(1, 201, dtype=torch.float64, device=dev)
G = (torch.cos(2 * np.pi * r[:, None] * t[None, :] + 0.7 * r[:, None])
     + 0.5 * torch.sin(2 * np.pi * (r[:, None] + 3) * t[None, :])) * (r ** -0.8)[:, None]
M = 32768
p = (torch.arange(M, dtype=torch.float64, device=dev) + 0.5) / M
A = torch.sin(np.pi * r[None, :] * p[:, None] + 0.3 * r[None, :]) @ G
A += 1e-3 * (torch.rand((M, N), dtype=torch.float64, device=dev) - 0.5)
A -= A.mean(dim=1, keepdim=True)
C0 = (A.T @ A) / N
del A
lam = torch.empty(N, dtype=torch.float64, device=dev)
lam2 = torch.empty(N, dtype=torch.float64, device=dev)
out = {"N": N, "nm": nm}
for name in ("full", "top"):
    best = 1e9
    for rep in range(3):
        C = C0.clone()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if name == "full":
            capi.check(L.podsdf_eig_sym_dev(0, C.data_ptr(), N, lam.data_ptr(), 0))
        else:
            path = ctypes.c_int(-1)
            capi.check(L.podsdf_eig_sym_top_dev(0, C.data_ptr(), N, nm, lam2.data_ptr(), ctypes.byref(path), 0))
            out["path"] = path.value
        torch.cuda.synchronize()
        if rep:
            best = min(best, time.perf_counter() - t0)
    out[name + "_s"] = round(best, 4)
    if name == "full":
        Vfull = C[N - nm:].flip(0).clone()
    else:
        V = C[:nm]
        top = lam2.flip(0)[:nm]
        out["eigenvalue_max_rel_diff"] = float((lam2 - lam).abs().max() / lam.max())
        out["residual"] = float((V @ C0 - top[:, None] * V).abs().max() / lam.max())
        out["orthonormality"] = float((V @ V.T - torch.eye(nm, dtype=torch.float64, device=dev)).abs().max())
        out["min_abs_dot_with_full_solver"] = float((V * Vfull).sum(dim=1).abs().min())
print(json.dumps(out))
