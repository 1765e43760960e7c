"""podsdf_eig_sym_dev at N = 10^4 on two spectra: a Wishart matrix (as tools/probe_eig.py) and the separated
low-rank + noise Gram matrix of config 4 (as bench.py builds it)."""
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
N = 10000
dev = torch.device("cuda", 0)
out = {"N": N}


def run(name, C0):
    lam = torch.empty(N, dtype=torch.float64, device=dev)
    best = 1e9
    for r in range(3):
        C = C0.clone()
        torch.cuda.synchronize()
        t = time.perf_counter()
        capi.check(L.podsdf_eig_sym_dev(0, C.data_ptr(), N, lam.data_ptr(), 0))
        torch.cuda.synchronize()
        if r:
            best = min(best, time.perf_counter() - t)
    out[name] = round(best, 4)


torch.manual_seed(0)
B = torch.randn(N, 2 * N, dtype=torch.float64, device=dev)
run("wishart_s", (B @ B.T) / N)
del B
t = torch.arange(N, dtype=torch.float64, device=dev) / N
r = torch.arange(1, 201, dtype=torch.float64, device=dev)
G = (torch.cos(2 * np.pi * r[:, None] * t[None, :] + 0.7 * r[:, None])
     + 0.5 * torch.sin(2 * np.pi * (r[:, None] + 3) * t[None, :])) * (r ** -0.8)[:, None]
M = 32768
p = (torch.arange(M, dtype=torch.float64, device=dev) + 0.5) / M
A = torch.sin(np.pi * r[None, :] * p[:, None] + 0.3 * r[None, :]) @ G
A += 1e-3 * (torch.rand((M, N), dtype=torch.float64, device=dev) - 0.5)
A -= A.mean(dim=1, keepdim=True)
run("lowrank_noise_s", (A.T @ A) / N)
print(json.dumps(out))
