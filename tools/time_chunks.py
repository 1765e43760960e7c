"""Config-3 generated-mode step time (2500 planes of 512^2 x 3, nf = 10) against the time-loop chunk size."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import podsdf_loader  # noqa: E402

P = podsdf_loader.load()
df = P.digitalfilters
J = K = 512
nf, ln, ns = 10, 5.0, 2500
taps = df._taps(nf, ln)
p = df._plan(J, K, nf, nf, nf, taps, taps, taps)
rs = np.random.RandomState(0)
coef = np.zeros((12, J * K))
coef[0], coef[4], coef[8], coef[9] = (rs.uniform(0.5, 1.0, J * K) for _ in range(4))
p.set_coefficients(coef, 0b001100010001)
out = torch.empty((ns, 3, J, K), dtype=torch.float64, device="cuda")
res = {}
for chunk in [int(c) for c in sys.argv[1:]] or [256, 512, 840, 1256, 2504]:
    p.set_chunk(chunk)
    best = 1e9
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        p.generate_seeded_dev(7, 0, ns, out.data_ptr(), 0)
        e1.record()
        torch.cuda.synchronize()
        if rep:
            best = min(best, e0.elapsed_time(e1))
    res[chunk] = dict(ms=round(best, 3), kernels_ms=[round(x, 3) for x in p.last_kernel_ms()])
print(json.dumps(res))
