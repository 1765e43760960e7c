"""Gram kernel timing / profiling driver: python tools/time_gram.py M N [layout]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import podsdf_loader
pkg = podsdf_loader.load(); capi = pkg._capi; L = capi.lib()
M, N = int(sys.argv[1]), int(sys.argv[2]); layout = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda", 0)
A = torch.randn((M, N) if layout == 0 else (N, M), dtype=torch.float64, device=dev)
C = torch.empty((N, N), dtype=torch.float64, device=dev)
for _ in range(2):
    capi.check(L.podsdf_gram_dev(0, A.data_ptr(), M, N, layout, None, 0, 1.0 / N, C.data_ptr(), 0))
ms = np.zeros(1); capi.check(L.podsdf_gram_last_ms(capi.ptr(ms)))
print(json.dumps({"M": M, "N": N, "layout": layout, "ms": float(ms[0]), "tflops": M * N * (N + 1.0) / (ms[0] * 1e-3) / 1e12}))
