"""Device-resident and multi-GPU drivers of the two hot paths (one process per GPU).

PyTorch is used here for plumbing only -- device buffers, streams and `torch.distributed` (NCCL on
GPUs, gloo in the CPU tests); every FLOP of the hot paths runs in libpodsdf's kernels through the
C ABI (`_capi`).

Sharding:
  DF    time-sharded: rank r generates steps [t0_r, t1_r).  Step t needs random slabs t..t+2nfx
        (DF:1454-1467), so a shard needs a 2nfx-slab halo from its right neighbour.  In the
        generated-RNG mode the halo is regenerated from the counter-based stream (no traffic); in
        the injected mode it is exchanged over NVLink (`exchange_halo`, NCCL send/recv).
  POD   the Gram matrix is a sum over spatial rows: rank r holds rows [m0_r, m1_r) of the snapshot
        matrix, forms its N x N partial and one all-reduce (sum, float64) completes C on every
        rank; the eigen-solve is replicated, the projection is local to the row shard.
"""
import numpy as np

try:
    from . import _capi
except ImportError:
    import _capi


def shard_range(n, rank, world, align=1):
    """Contiguous block partition of range(n): the first ranks get one extra block of `align` items.  DF time
    shards use align=8: the tensor-path time filter sums in 8-step tiles, so shards that start on a multiple
    of 8 reproduce the single-GPU run bit for bit."""
    nb = (n + align - 1) // align
    base, extra = divmod(nb, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return min(lo * align, n), min(hi * align, n)


# ------------------------------------------------------------------------------------- DF
def exchange_halo(slabs, halo, rank=None, world=None, group=None):
    """Injected mode.  `slabs` (n_local, 3, Jp, Kp) holds this rank's own slabs [t0_r, t1_r) of the global
    random stream; returns the tensor extended by the first `halo` (= 2 nfx) slabs of the right neighbour,
    which the last `halo` steps of this shard need.  The last rank gets no extension: its window is
    truncated (the global stream must therefore carry 2 nfx slabs more than the steps wanted, exactly as
    the reference's initial block does, DF:1361-1366).  Works on CUDA tensors over NCCL and on CPU
    tensors over gloo (the unit tests)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group) if rank is None else rank
    world = dist.get_world_size(group) if world is None else world
    if world == 1 or halo == 0:
        return slabs
    if slabs.shape[0] < halo:
        raise ValueError("a shard must hold at least 2*nfx slabs to serve its neighbour's halo")
    ops = []
    recv = None
    if rank > 0:                     # my first `halo` slabs are the left neighbour's halo
        ops.append(dist.P2POp(dist.isend, slabs[:halo].contiguous(), rank - 1, group))
    if rank < world - 1:
        recv = torch.empty((halo,) + tuple(slabs.shape[1:]), dtype=slabs.dtype, device=slabs.device)
        ops.append(dist.P2POp(dist.irecv, recv, rank + 1, group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return slabs if recv is None else torch.cat([slabs, recv], dim=0)


class DfShard(object):
    """One rank's share of a DF run, resident on its GPU."""

    def __init__(self, plan, nsteps_local, device):
        import torch
        self.plan = plan
        self.n = int(nsteps_local)
        self.device = device
        self.out = torch.empty((self.n, 3, plan.J, plan.K), dtype=torch.float64, device=device)

    def generate_seeded(self, seed, t0, stream=None):
        import torch
        st = torch.cuda.current_stream(self.device) if stream is None else stream
        self.plan.generate_seeded_dev(seed, t0, self.n, self.out.data_ptr(), st.cuda_stream)
        return self.out

    def generate_injected(self, slabs, stream=None):
        import torch
        st = torch.cuda.current_stream(self.device) if stream is None else stream
        n_slabs = slabs.shape[0]
        assert n_slabs - 2 * self.plan.nfx <= self.n
        self.plan.generate_injected_dev(slabs.data_ptr(), n_slabs, self.out.data_ptr(), st.cuda_stream)
        return self.out[:n_slabs - 2 * self.plan.nfx]


def generate_injected_sharded(plan, own, out, rank=None, world=None, group=None, stream=None):
    """Injected mode on a time shard with the halo exchange HIDDEN behind the interior steps: `own`
    (n_own, 3, Jp, Kp) holds this rank's slabs [t0_r, t1_r) (the last rank also owns the 2 nfx tail slabs).  The
    steps whose windows lie inside `own` start at once; meanwhile the first 2 nfx slabs go to the left neighbour and
    the right neighbour's arrive (NCCL send/recv on its own stream); the last 24 steps (a multiple of 8 >= 2 nfx, so
    tile anchors -- and therefore bits -- match the one-GPU run) follow on the received halo.  `out` (n_steps, 3, J, K)
    with n_steps = t1_r - t0_r is filled; returns it."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group) if rank is None else rank
    world = dist.get_world_size(group) if world is None else world
    st = torch.cuda.current_stream(own.device) if stream is None else stream
    halo = 2 * plan.nfx
    n_steps = out.shape[0]
    if world == 1 or rank == world - 1:                  # nothing to receive: the shard owns its whole window
        ops = []
        if world > 1 and rank > 0:
            ops.append(dist.P2POp(dist.isend, own[:halo].contiguous(), rank - 1, group))
        reqs = dist.batch_isend_irecv(ops) if ops else []
        plan.generate_injected_dev(own.data_ptr(), n_steps + halo, out.data_ptr(), st.cuda_stream)
        for r in reqs:
            r.wait()
        return out
    tail = ((halo + 7) // 8) * 8
    if n_steps < tail + 8 or own.shape[0] != n_steps:
        ext = exchange_halo(own, halo, rank, world, group)
        plan.generate_injected_dev(ext.data_ptr(), ext.shape[0], out.data_ptr(), st.cuda_stream)
        return out
    recv = torch.empty((halo,) + tuple(own.shape[1:]), dtype=own.dtype, device=own.device)
    ops = [dist.P2POp(dist.irecv, recv, rank + 1, group)]
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, own[:halo].contiguous(), rank - 1, group))
    reqs = dist.batch_isend_irecv(ops)
    n_in = n_steps - tail                                 # interior steps: windows end at slab n_in + halo - 1 < n_steps
    plan.generate_injected_dev(own.data_ptr(), n_in + halo, out.data_ptr(), st.cuda_stream)
    for r in reqs:
        r.wait()
    last = torch.cat([own[n_in:], recv], dim=0)          # slabs n_in .. n_steps + halo - 1
    plan.generate_injected_dev(last.data_ptr(), last.shape[0], out[n_in:].data_ptr(), st.cuda_stream)
    return out


# ------------------------------------------------------------------------------------ POD
def allreduce_gram(C, group=None, device_index=None):
    """Sum of the per-rank N x N partial Gram matrices (float64) on every rank.  On GPUs the exactly symmetric partials
    travel as their packed upper triangle -- N (N + 1) / 2 doubles instead of N^2 -- through
    podsdf_sym_pack_dev / _unpack_dev; CPU tensors (gloo unit tests) are reduced in full."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return C
    if C.is_cuda and C.dim() == 2 and C.shape[0] == C.shape[1] and C.is_contiguous() and C.shape[0] <= 65535:
        N = C.shape[0]
        L = _capi.lib()
        di = C.device.index if device_index is None else device_index
        st = torch.cuda.current_stream(C.device).cuda_stream
        packed = torch.empty(N * (N + 1) // 2, dtype=torch.float64, device=C.device)
        _capi.check(L.podsdf_sym_pack_dev(di, C.data_ptr(), N, packed.data_ptr(), st))
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        _capi.check(L.podsdf_sym_unpack_dev(di, packed.data_ptr(), N, C.data_ptr(), st))
        return C
    dist.all_reduce(C, op=dist.ReduceOp.SUM, group=group)
    return C


def select_modes(lam_asc, V_rows, ns, nm, tol_CN=1.0e-15):
    """Host-side post-processing of the eigen-solve, restating POD:1310-1327 for the device pipeline:
    descending order, validity count, truncation, and the scale factors a_j = v_j sqrt(lambda_j ns/|v_j|^2).
    lam_asc (N,) ascending eigenvalues; V_rows(j) returns eigenvector j (of the ascending order) as (N,).
    Returns (energy_desc (N,), nm, T (N, nm) scaled temporal modes)."""
    try:
        from . import PODFS as pod
    except ImportError:
        import PODFS as pod
    N = lam_asc.shape[0]
    vals = np.where(np.isnan(lam_asc), -1.0e10, lam_asc)
    order = np.lexsort((-np.arange(N), -vals))
    energy = vals[order]
    nv = pod._count_valid(energy, ns, tol_CN)
    if nm < 0 or nm > nv:
        nm = nv
    T = np.empty((N, nm))
    for j in range(nm):
        v = V_rows(int(order[j]))
        T[:, j] = v * np.sqrt(energy[j] / (np.sum(v * v) / ns))
    return energy, nm, T


def pod_resident(A, M, N, layout, nm, dt, et, device_index=0, group=None, timers=None):
    """PODFS of a device-resident (row shard of a) snapshot matrix: mean removal, Gram (+ all-reduce),
    eigen-solve, projection, Fourier fit.  A: CUDA float64 tensor, modified in place (mean removed).
    Returns dict(mean, energy, nm, T, Phi, c128) with Phi (M, nm) on the device."""
    import torch
    L = _capi.lib()
    dev = A.device
    st = torch.cuda.current_stream(dev).cuda_stream
    ev = lambda: torch.cuda.Event(enable_timing=True)
    marks = [ev() for _ in range(6)]
    mean = torch.empty(M, dtype=torch.float64, device=dev)
    marks[0].record()
    _capi.check(L.podsdf_mean_rows_dev(device_index, A.data_ptr(), M, N, layout, mean.data_ptr(), 1, st))
    C = torch.empty((N, N), dtype=torch.float64, device=dev)
    marks[1].record()
    _capi.check(L.podsdf_gram_dev(device_index, A.data_ptr(), M, N, layout, None, 0, 1.0 / N, C.data_ptr(), st))
    allreduce_gram(C, group)
    marks[2].record()
    lam = torch.empty(N, dtype=torch.float64, device=dev)
    # all eigenvalues + the leading vectors only (rows 0.. of C, descending): PODFS never reads the other N - nm
    # eigenvectors (POD:1329-1333); nm < 0 ("keep every valid mode") needs them all
    n_vec = N if nm < 0 else min(int(nm), N)
    route_val = 0                        # 1 = this library's own tridiagonal route, 0 = full library eigen-solver
    if n_vec < N:
        import ctypes
        route = ctypes.c_int(0)
        _capi.check(L.podsdf_eig_sym_top_dev(device_index, C.data_ptr(), N, n_vec, lam.data_ptr(), ctypes.byref(route), st))
        route_val = int(route.value)
        lead = C[:n_vec].cpu().numpy()                   # one copy of the leading rows instead of one per mode
        vec = lambda j: lead[N - 1 - j]                  # ascending index j -> row of the descending list
    else:
        _capi.check(L.podsdf_eig_sym_dev(device_index, C.data_ptr(), N, lam.data_ptr(), st))   # rows of C = vectors
        vec = lambda j: C[j].cpu().numpy()
    marks[3].record()
    lam_h = lam.cpu().numpy()
    energy, nm, T = select_modes(lam_h, vec, N, nm)
    T_d = torch.from_numpy(np.ascontiguousarray(T)).to(dev)
    dinv = torch.from_numpy(1.0 / energy[:nm]).to(dev)
    Phi = torch.empty((M, nm), dtype=torch.float64, device=dev)
    _capi.check(L.podsdf_project_dev(device_index, A.data_ptr(), M, N, layout, T_d.data_ptr(), nm, dinv.data_ptr(),
                                     nm, 1.0 / N, Phi.data_ptr(), st))
    marks[4].record()
    import time as _time
    h0 = _time.perf_counter()
    c = torch.empty((N, nm, 2), dtype=torch.float64, device=dev)
    _capi.check(L.podsdf_dft_modes_dev(device_index, T_d.data_ptr(), N, nm, nm, c.data_ptr(), st))
    h1 = _time.perf_counter()
    c128 = c.cpu().numpy().view(np.complex128).reshape(N, nm)
    h2 = _time.perf_counter()
    try:
        from . import PODFS as pod
    except ImportError:
        import PODFS as pod
    c64, c_ind, c_count = pod.select_fourier(c128, et)
    h3 = _time.perf_counter()
    marks[5].record()
    if timers is not None:                # host wall clock of the three pieces of the "dft" part (the last one is numpy)
        for name, dtm in (("host_dft_launch", h1 - h0), ("host_dft_readback", h2 - h1), ("host_fourier_ranking", h3 - h2)):
            timers[name] = timers.get(name, 0.0) + 1e3 * dtm
    torch.cuda.synchronize(dev)
    if timers is not None:
        names = ["mean", "gram", "eig", "select_project", "dft"]
        for i, name in enumerate(names):
            timers[name] = timers.get(name, 0.0) + marks[i].elapsed_time(marks[i + 1])
        timers.setdefault("dft_each_step", []).append(marks[4].elapsed_time(marks[5]))
    return dict(mean=mean, energy=energy, nm=nm, T=T, Phi=Phi, c128=c128, c_ind=c_ind, c_count=c_count, route=route_val)


# ----------------------------------------------------------------------------- DF -> PODFS
def df_to_podfs(plan, nsteps, nm, dt, et, seed=None, slabs=None, device_index=0):
    """Config-1 style run entirely on one GPU: generate `nsteps` inlet planes (generated stream `seed`, or the
    injected CUDA tensor `slabs`), keep them in HBM in the generator's own layout [t][c][p] -- which IS the
    snapshot matrix transposed (DF:1471-1473) -- and run PODFS on it without a transpose or a host round
    trip (DF:1492-1510).  Returns the pod_resident dict plus `planes` (mean already removed)."""
    import torch
    dev = torch.device("cuda", device_index)
    shard = DfShard(plan, nsteps, dev)
    if slabs is not None:
        shard.generate_injected(slabs)
    else:
        shard.generate_seeded(seed, 0)
    M = 3 * plan.J * plan.K
    res = pod_resident(shard.out.view(nsteps, M), M, nsteps, _capi.LAYOUT_TIME_BY_ROWS, nm, dt, et, device_index)
    res["planes"] = shard.out
    return res


def reshard_time_to_rows(planes, group=None):
    """DF -> PODFS hand-off across GPUs.  `planes` (n_local, M) holds this rank's time
    shard of the snapshot matrix in the generator's layout a[t][row]; PODFS wants every rank to own a block of
    rows for ALL snapshots.  One all-to-all (NCCL over NVLink; gloo on CPU) moves block (t in shard s, rows of
    rank r) from rank s to rank r; the result (N_total, M_local) is again the time-by-rows layout, so no
    transpose pass is needed.  Returns (a_local, (m0, m1))."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_local, M = planes.shape
    if world == 1:
        return planes, (0, M)
    counts = [torch.zeros(1, dtype=torch.int64, device=planes.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([n_local], dtype=torch.int64, device=planes.device), group=group)
    n_of = [int(c.item()) for c in counts]
    cuts = [shard_range(M, r, world) for r in range(world)]
    m0, m1 = cuts[rank]
    send = [planes[:, a:b].contiguous() for a, b in cuts]                       # my steps, rows of rank r
    recv = [torch.empty((n_of[s], m1 - m0), dtype=planes.dtype, device=planes.device) for s in range(world)]
    if dist.get_backend(group) == "nccl":
        dist.all_to_all(recv, send, group=group)
    else:                                   # gloo (CPU tests) has no all-to-all: same exchange as point-to-point pairs
        recv[rank].copy_(send[rank])
        ops = [dist.P2POp(dist.isend, send[r], r, group) for r in range(world) if r != rank]
        ops += [dist.P2POp(dist.irecv, recv[s], s, group) for s in range(world) if s != rank]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return torch.cat(recv, dim=0), (m0, m1)


def df_to_podfs_sharded(plan, nsteps_total, nm, dt, et, seed, device_index=0, group=None):
    """Config-3/4 style run on all ranks: every rank generates its time shard (generated stream, 8-aligned
    shard starts), the planes are re-sharded by rows with one all-to-all, PODFS runs on the row shard with one
    all-reduce of the Gram matrix.  Returns pod_resident's dict for this rank's rows plus `rows` = (m0, m1)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = torch.device("cuda", device_index)
    t0, t1 = shard_range(nsteps_total, rank, world, 8)
    shard = DfShard(plan, max(t1 - t0, 1), dev)
    shard.generate_seeded(seed, t0)
    M = 3 * plan.J * plan.K
    torch.cuda.synchronize(dev)
    a_local, rows = reshard_time_to_rows(shard.out.view(-1, M)[:t1 - t0], group)
    del shard
    res = pod_resident(a_local, rows[1] - rows[0], nsteps_total, _capi.LAYOUT_TIME_BY_ROWS, nm, dt, et, device_index, group)
    res["rows"] = rows
    res["planes_rows"] = a_local
    return res


# -------------------------------------------------------------------- PODFS reconstruction
def temporal_reconstruction(c64, c_ind, c_count, period, times):
    """a_i(t) = Re sum_n c[c_ind[i,n], i] exp(2 pi i k_n t / period), k_n = c_ind[i,n] - N//2: the series the
    CFD solver evaluates (docs/usage/CFDCodeIntegration.rst:85-91) and the reference's plot-only y2
    (POD:1602-1612).  Host side: it is (modes x kept coefficients x times), tiny.  Returns (len(times), nm)."""
    N, nm = c64.shape
    times = np.asarray(times, dtype=np.float64)
    y2 = np.zeros((times.shape[0], nm))
    for i in range(nm):
        sel = c_ind[i, :c_count[i]]
        k = sel.astype(np.int64) - N // 2
        ph = np.exp(1j * 2 * np.pi * np.outer(times, k) / period)
        y2[:, i] = (ph @ c64[sel, i].astype(np.complex128)).real
    return y2


def reconstruct_field(Phi, mean, y2, device_index=0):
    """U(x, t_j) = mean(x) + sum_i Phi_i(x) a_i(t_j): the PODFS inlet a solver rebuilds, for a-posteriori
    error and statistics checks.  Phi (M, nm) CUDA tensor, mean (M,) CUDA tensor, y2 (Nt, nm) numpy.
    Returns a CUDA tensor (M, Nt).  The product runs in the FP64 tensor-core projection kernel."""
    import torch
    L = _capi.lib()
    M, nm = Phi.shape
    Nt = y2.shape[0]
    dev = Phi.device
    T = torch.from_numpy(np.ascontiguousarray(y2.T)).to(dev)          # (nm, Nt): "temporal modes" of the product
    ones = torch.ones(Nt, dtype=torch.float64, device=dev)
    U = torch.empty((M, Nt), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    _capi.check(L.podsdf_project_dev(device_index, Phi.contiguous().data_ptr(), M, nm, 0, T.data_ptr(), Nt, ones.data_ptr(),
                                     Nt, 1.0, U.data_ptr(), st))
    U += mean[:, None]
    return U


# ------------------------------------------------------------- one- / two-point statistics
def field_moments(F, P, N, layout, device_index=0):
    """mean (3, P) and central second moments (6, P) = uu, vv, ww, uv, uw, vw of a three-component field held on the
    device as a snapshot matrix (layout 0: rows-by-time, 1: time-by-rows).  CUDA tensors back."""
    import torch
    L = _capi.lib()
    dev = F.device
    mean = torch.empty((3, P), dtype=torch.float64, device=dev)
    stress = torch.empty((6, P), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    _capi.check(L.podsdf_field_moments_dev(device_index, F.data_ptr(), P, N, layout, mean.data_ptr(), stress.data_ptr(), st))
    return mean, stress


def lag_covariance(F, J, K, N, layout, comp, lags, mean=None, device_index=0):
    """< a'(t,j,k) a'(t+dt,j+dj,k+dk) > of component `comp` for every (dt, dj, dk) in `lags` (any number; sent in groups
    of 16).  `mean` (3, J*K) CUDA tensor or None for a mean-free field.  Returns a numpy vector."""
    import ctypes
    import torch
    L = _capi.lib()
    lags = np.ascontiguousarray(np.asarray(lags, dtype=np.int32).reshape(-1, 3))
    out = np.empty(lags.shape[0])
    st = torch.cuda.current_stream(F.device).cuda_stream
    for a in range(0, lags.shape[0], 16):
        blk = np.ascontiguousarray(lags[a:a + 16])
        res = np.empty(blk.shape[0])
        _capi.check(L.podsdf_lag_covariance_dev(device_index, F.data_ptr(), J, K, N, layout,
                                                None if mean is None else mean.data_ptr(), int(comp),
                                                blk.ctypes.data_as(ctypes.c_void_p), blk.shape[0],
                                                res.ctypes.data_as(ctypes.c_void_p), st))
        out[a:a + blk.shape[0]] = res
    return out


def reconstruction_statistics(res, planes, J, K, dt, lags=None, device_index=0):
    """A-posteriori check of a PODFS compression on the device: rebuild the inlet the CFD solver would
    see, U~(x, t_j) = mean + sum_i Phi_i(x) a~_i(t_j) with a~_i the truncated Fourier series (docs/usage/
    CFDCodeIntegration.rst:85-91; the reference's y2, POD:1602-1612), and compare with the generated planes:
    relative error energy, mean profile, the six Reynolds stresses per point, and two-point covariances of u along
    time, j and k.  `res` = pod_resident's dict, `planes` (N, 3 J K) the MEAN-REMOVED snapshots on the device in the
    generator's layout.  Every field-sized operation runs on the device (projection kernel, statistics kernels)."""
    import torch
    N = planes.shape[0]
    P = J * K
    c64 = res["c128"].astype(np.complex64)
    y2 = temporal_reconstruction(c64, res["c_ind"], res["c_count"], N * dt, np.arange(N) * dt)      # (N, nm)
    U = reconstruct_field(res["Phi"], torch.zeros_like(res["mean"]), y2, device_index)              # (3P, N) fluctuation
    diff = planes.t() - U
    err = float((diff * diff).sum() / (planes * planes).sum())
    # error energy the truncation to nm modes alone would leave (exact temporal modes): 1 - sum_{i<nm} lambda_i / sum lambda
    lam = np.asarray(res["energy"], dtype=np.float64)
    trunc = float(1.0 - lam[:res["nm"]].sum() / lam[lam > 0].sum())
    m_ref, s_ref = field_moments(planes, P, N, _capi.LAYOUT_TIME_BY_ROWS, device_index)
    m_rec, s_rec = field_moments(U, P, N, _capi.LAYOUT_ROWS_BY_TIME, device_index)
    if lags is None:
        lags = [(0, 0, 0)] + [(r, 0, 0) for r in (1, 2, 4, 8)] + [(0, r, 0) for r in (1, 2, 4)] + [(0, 0, r) for r in (1, 2, 4)]
    lags = [tuple(int(v) for v in l) for l in lags if l[0] < N and l[1] < J and l[2] < K]
    c_ref = lag_covariance(planes, J, K, N, _capi.LAYOUT_TIME_BY_ROWS, 0, lags, None, device_index)
    c_rec = lag_covariance(U, J, K, N, _capi.LAYOUT_ROWS_BY_TIME, 0, lags, None, device_index)
    return dict(error_energy=err, truncation_energy=trunc, mean_generated=m_ref, mean_reconstructed=m_rec,
                stress_generated=s_ref, stress_reconstructed=s_rec, lags=lags, cov_generated=c_ref, cov_reconstructed=c_rec,
                field=U)
