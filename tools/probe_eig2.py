"""Probe: cuSOLVER Dsyevdx (index range = the nm largest) and Dsytrd alone at N = 10^4 on two spectra."""
import ctypes, json, sys, time
import numpy as np, torch
N, k = 10000, 100
cs = ctypes.CDLL("libcusolver.so.11")
h = ctypes.c_void_p(); assert cs.cusolverDnCreate(ctypes.byref(h)) == 0
vp, d = ctypes.c_void_p, ctypes.c_double
dev = torch.device("cuda", 0)
W = torch.empty(N, dtype=torch.float64, device=dev)
info = torch.zeros(1, dtype=torch.int32, device=dev)
out = {"N": N, "k": k}

def spectra():
    torch.manual_seed(0)
    B = torch.randn(N, 2 * N, dtype=torch.float64, device=dev)
    yield "wishart", (B @ B.T) / N
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
    yield "lowrank", (A.T @ A) / N

for name, C0 in spectra():
    # syevdx, range I: il..iu are 1-based ascending indices
    lwork = ctypes.c_int(0); meig = ctypes.c_int(0)
    A = C0.clone()
    rc = cs.cusolverDnDsyevdx_bufferSize(h, 1, 1002, 0, N, vp(A.data_ptr()), N, d(0), d(0), N - k + 1, N, ctypes.byref(meig), vp(W.data_ptr()), ctypes.byref(lwork))
    out[name + "_syevdx_buf_rc"] = rc
    work = torch.empty(max(lwork.value, 1), dtype=torch.float64, device=dev)
    best = 1e9
    for r_ in range(2):
        A = C0.clone(); torch.cuda.synchronize(); t0 = time.perf_counter()
        rc = cs.cusolverDnDsyevdx(h, 1, 1002, 0, N, vp(A.data_ptr()), N, d(0), d(0), N - k + 1, N, ctypes.byref(meig), vp(W.data_ptr()), vp(work.data_ptr()), lwork.value, vp(info.data_ptr()))
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    out[name + "_syevdx_top_s"] = round(best, 4); out[name + "_syevdx_rc"] = rc; out[name + "_meig"] = meig.value; out[name + "_info"] = int(info.item())
    # check residual of the returned pairs: columns of A (column-major) = rows of the torch tensor
    V = A[:k]            # first meig columns (column-major) = first k rows here
    lam = W[:k]
    res = (V @ C0 - lam[:, None] * V).abs().max().item() / lam.max().item()
    out[name + "_syevdx_residual"] = res
    out[name + "_syevdx_orth"] = (V @ V.T - torch.eye(k, dtype=torch.float64, device=dev)).abs().max().item()
    # sytrd alone
    dd = torch.empty(N, dtype=torch.float64, device=dev); ee = torch.empty(N, dtype=torch.float64, device=dev); tau = torch.empty(N, dtype=torch.float64, device=dev)
    rc = cs.cusolverDnDsytrd_bufferSize(h, 0, N, vp(A.data_ptr()), N, vp(dd.data_ptr()), vp(ee.data_ptr()), vp(tau.data_ptr()), ctypes.byref(lwork))
    work = torch.empty(max(lwork.value, 1), dtype=torch.float64, device=dev)
    best = 1e9
    for r_ in range(2):
        A = C0.clone(); torch.cuda.synchronize(); t0 = time.perf_counter()
        rc = cs.cusolverDnDsytrd(h, 0, N, vp(A.data_ptr()), N, vp(dd.data_ptr()), vp(ee.data_ptr()), vp(tau.data_ptr()), vp(work.data_ptr()), lwork.value, vp(info.data_ptr()))
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    out[name + "_sytrd_s"] = round(best, 4); out[name + "_sytrd_rc"] = rc
print(json.dumps(out))
