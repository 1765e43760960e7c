"""Probe: cuSOLVER symmetric eigen-solver variants at N = 10^4 (time of Dsyevd with / without vectors, Dsyevdx top-k)."""
import ctypes, json, sys, time
import numpy as np, torch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
cs = ctypes.CDLL("libcusolver.so.11")
h = ctypes.c_void_p(); assert cs.cusolverDnCreate(ctypes.byref(h)) == 0
torch.manual_seed(0)
B = torch.randn(N, 2 * N, dtype=torch.float64, device="cuda")
C0 = (B @ B.T) / N; del B
W = torch.empty(N, dtype=torch.float64, device="cuda")
info = torch.zeros(1, dtype=torch.int32, device="cuda")
vp = ctypes.c_void_p
out = {"N": N}
def run(name, jobz):
    lwork = ctypes.c_int(0)
    A = C0.clone()
    cs.cusolverDnDsyevd_bufferSize(h, jobz, 0, N, vp(A.data_ptr()), N, vp(W.data_ptr()), ctypes.byref(lwork))
    work = torch.empty(lwork.value, dtype=torch.float64, device="cuda")
    best = 1e9
    for r in range(2):
        A = C0.clone(); torch.cuda.synchronize(); t = time.perf_counter()
        rc = cs.cusolverDnDsyevd(h, jobz, 0, N, vp(A.data_ptr()), N, vp(W.data_ptr()), vp(work.data_ptr()), lwork.value, vp(info.data_ptr()))
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    out[name] = best
run("syevd_vectors_s", 1)
run("syevd_novectors_s", 0)
# syevdx: top k by index range
k = 100
lwork = ctypes.c_int(0); meig = ctypes.c_int(0)
A = C0.clone()
d = ctypes.c_double
rc = cs.cusolverDnDsyevdx_bufferSize(h, 1, 2, 0, N, vp(A.data_ptr()), N, d(0), d(0), N - k + 1, N, ctypes.byref(meig), vp(W.data_ptr()), ctypes.byref(lwork))
work = torch.empty(max(lwork.value, 1), dtype=torch.float64, device="cuda")
best = 1e9
for r in range(2):
    A = C0.clone(); torch.cuda.synchronize(); t = time.perf_counter()
    rc = cs.cusolverDnDsyevdx(h, 1, 2, 0, N, vp(A.data_ptr()), N, d(0), d(0), N - k + 1, N, ctypes.byref(meig), vp(W.data_ptr()), vp(work.data_ptr()), lwork.value, vp(info.data_ptr()))
    torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
out["syevdx_top100_s"] = best; out["syevdx_rc"] = rc; out["meig"] = meig.value
t = time.perf_counter(); torch.linalg.eigvalsh(C0); torch.cuda.synchronize(); out["torch_eigvalsh_s"] = time.perf_counter() - t
print(json.dumps(out))
