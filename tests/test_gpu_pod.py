"""GPU parity of the PODFS path (through the C ABI) against the CPU oracle and the golden vectors."""
import os

import numpy as np
import pytest

from oracle import pod_oracle as podo
from tests import replay
from tests.util import TOL, fix_signs, golden, pkg, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    p = pkg()
    p._capi.require_device()
    return p


def test_correlation_matrix_matches_reference(P):
    g = golden("pod_parts.npz")
    A, w = g["in_A"], g["in_w"]
    ns, npts = A.shape[1], A.shape[0] // 3
    for layout_view in (A, np.ascontiguousarray(A.T).T):      # reference layout and transposed planes
        C = np.zeros((ns, ns))
        P.PODFS.calculate_correlation_matrix(ns, npts, 3, "false", [], layout_view, C)
        assert rel_err(C, g["out_C_false"]) < 1e-13
        assert np.array_equal(C, C.T)                          # exactly symmetric, like dsyrk
        P.PODFS.calculate_correlation_matrix(ns, npts, 3, "true", w, layout_view, C)
        assert rel_err(C, g["out_C_true"]) < 1e-13
        assert np.array_equal(C, C.T)
    C[:, :] = 7.0
    P.PODFS.calculate_correlation_matrix(ns, npts, 3, "maybe", [], A, C)
    assert not C.any()                                         # unknown flag leaves zeros (POD:1453-1456)


@pytest.mark.parametrize("M,N", [(12288, 1000), (3 * 1111, 257), (300, 129), (48, 5), (4998, 128)])
def test_gram_sizes_and_layouts(P, M, N):
    rs = np.random.RandomState(M + N)
    A = rs.standard_normal((M, N))
    ref = np.zeros((N, N))
    podo.calculate_correlation_matrix(N, M // 3, 3, "false", [], A, ref)
    capi = P._capi
    assert rel_err(capi.gram(A, capi.LAYOUT_ROWS_BY_TIME, None, 1.0 / N), ref) < 1e-12
    assert rel_err(capi.gram(np.ascontiguousarray(A.T), capi.LAYOUT_TIME_BY_ROWS, None, 1.0 / N), ref) < 1e-12
    w = rs.uniform(0.5, 2.0, M // 3)
    refw = (A * np.tile(w, 3)[:, None]).T @ A / N
    assert rel_err(capi.gram(A, capi.LAYOUT_ROWS_BY_TIME, w, 1.0 / N), refw) < 1e-12
    assert rel_err(capi.gram(np.ascontiguousarray(A.T), capi.LAYOUT_TIME_BY_ROWS, w, 1.0 / N), refw) < 1e-12


def test_eigen_solver(P):
    rs = np.random.RandomState(2)
    N = 300
    B = rs.standard_normal((N, 2 * N))
    C = B @ B.T / N
    lam, V = P._capi.eig_sym(C)
    ref = np.linalg.eigvalsh(C)
    assert np.abs(lam - ref).max() < 1e-12 * ref.max()
    assert np.abs(V.T @ V - np.eye(N)).max() < 1e-12
    assert np.abs(C @ V - V * lam[None, :]).max() < 1e-11 * ref.max()


def test_projection_and_dft(P):
    rs = np.random.RandomState(3)
    M, N, nm = 1000, 333, 37
    A = rs.standard_normal((M, N))
    T = rs.standard_normal((N, N))
    lam = rs.uniform(0.5, 2.0, nm)
    ref = (A @ T[:, :nm]) @ np.diag(1.0 / lam) / N
    capi = P._capi
    assert rel_err(capi.project(A, T, 1.0 / lam, nm, 1.0 / N), ref) < 1e-12
    assert rel_err(capi.project(np.ascontiguousarray(A.T), T, 1.0 / lam, nm, 1.0 / N, capi.LAYOUT_TIME_BY_ROWS), ref) < 1e-12
    for n in (333, 1000, 64):
        a = rs.standard_normal((n, 50))
        c = capi.dft_modes(a, 23)
        ref128, _, _ = podo.dft_coefficients(a, n, 23, 0.01, direct=False)
        assert np.abs(c - ref128).max() < 1e-13
    a = rs.standard_normal((40, 6))
    ref128, _, _ = podo.dft_coefficients(a, 40, 6, 0.37, direct=True)      # the reference's own loop
    assert np.abs(capi.dft_modes(a, 6) - ref128).max() < 1e-14


def _check_pod(P, A, ns, npts, nm, dt, et, ref, ref_fc, weights=None, tol=1e-9):
    df, pod = P.digitalfilters, P.PODFS
    i_d = df.obj()
    i_d.verbose, i_d.hdf5, i_d.ns, i_d.dt, i_d.nm, i_d.et = False, True, ns, dt, nm, et
    cwd = os.getcwd()
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            os.mkdir("PODFS")
            pod.POD(A, ns, npts, 3, "true" if weights is not None else "false", weights if weights is not None else [],
                    "PODFS/", "false", 1.0e-15, nm, 0, "false", "false", None, None, dt, "velocity", 1, ns, 1, 1, i_d)
            pod.fourier_coefficients(i_d)
            dat = open("PODFS/PODFS.dat").read()
            assert os.path.exists("PODFS/POD.eigenvalues.dat")
        finally:
            os.chdir(cwd)
    nmr = ref["nm"]
    assert i_d.nm == nmr
    e_ref = ref["energy"].real
    assert np.abs(i_d.energy[:nmr] - e_ref[:nmr]).max() < TOL * e_ref[0]
    tm_ref = ref["temporal_modes"].real[:, :nmr]
    tm = fix_signs(i_d.temporal_modes[:, :nmr], tm_ref)
    sm = fix_signs(i_d.spatial_modes, ref["spatial_modes"])
    # eigenvector accuracy scales with 1/gap: use the gap-aware bound of SURVEY section 7
    gaps = np.minimum(np.abs(np.diff(e_ref[:nmr + 1])), np.abs(np.diff(np.r_[np.inf, e_ref[:nmr]]))) / e_ref[0]
    bound = tol / np.maximum(gaps, 1e-6)
    assert (np.abs(tm - tm_ref).max(axis=0) / np.abs(tm_ref).max(axis=0) < np.maximum(bound, 1e-9)).all()
    assert (np.abs(sm - ref["spatial_modes"]).max(axis=0) / np.abs(ref["spatial_modes"]).max(axis=0)
            < np.maximum(bound, 1e-9)).all()
    # invariants (appendix F-6): orthonormal spatial modes, mean(a_i^2) = lambda_i, rank-nm reconstruction
    if weights is None:
        assert np.abs(i_d.spatial_modes.T @ i_d.spatial_modes - np.eye(nmr)).max() < 1e-9
    assert np.abs(np.mean(i_d.temporal_modes[:, :nmr] ** 2, axis=0) - e_ref[:nmr]).max() < 1e-9 * e_ref[0]
    recon = i_d.spatial_modes @ i_d.temporal_modes[:, :nmr].T
    recon_ref = ref["spatial_modes"] @ tm_ref.T
    assert rel_err(recon, recon_ref) < 1e-8
    return i_d, dat


@pytest.mark.parametrize("tag", ["default", "rotated", "roundjet", "square", "ring", "dt", "profile", "prf"])
def test_pod_and_fourier_match_reference_main(P, tag):
    """POD + Fourier fit of the reference's own snapshot matrix == what the reference's main() produced."""
    g = golden("main_%s.npz" % tag)
    o = replay.parse(g["in_argv"])
    ns, npts = o["n"], g["out_A"].shape[0] // 3          # (a .prf file sets the plane size itself)
    s = dict(nm=o["m"], et=o["e"], dt=float(g["out_dt"]))
    A = g["out_A"]
    ref = dict(nm=int(g["out_nm"]), temporal_modes=g["out_temporal_modes"], spatial_modes=g["out_spatial_modes"],
               energy=np.mean(g["out_temporal_modes"].real ** 2, axis=0))
    i_d, dat = _check_pod(P, A, ns, npts, s["nm"], s["dt"], s["et"], ref, None)
    # Fourier coefficients: |c| is sign-invariant, the coefficients follow the mode's sign
    sign = np.sign(np.sum(i_d.temporal_modes[:, :ref["nm"]] * g["out_temporal_modes"].real[:, :ref["nm"]], axis=0))
    assert np.array_equal(i_d.N_FC, g["out_N_FC"])
    FC, FCr = i_d.FC, g["out_FC"]
    row = 0
    for i in range(ref["nm"]):
        n = int(i_d.N_FC[i])
        assert np.array_equal(FC[row:row + n, 0], FCr[row:row + n, 0])          # same harmonics, same order
        assert np.abs(sign[i] * FC[row:row + n, 1:] - FCr[row:row + n, 1:]).max() < 2e-7 * np.abs(FCr[row:row + n, 1:]).max()
        row += n
    assert i_d.period == float(g["out_period"])
    assert dat.split("\n")[:2 + ref["nm"]] == str(g["out_podfs_dat"]).split("\n")[:2 + ref["nm"]]


def test_pod_config1_shape_against_oracle(P):
    """12 288 x 1000 snapshot matrix (config 1's POD), 20 modes, against the numpy oracle -- both layouts."""
    M, N, nm = 12288, 1000, 20
    A = podo.synthetic_snapshots(M, N, seed=1)
    _, A = (A.mean(1), A - A.mean(1)[:, None])
    ref = podo.pod(A, N, M // 3, 3, "false", [], 1.0e-15, nm, eig=np.linalg.eigh)
    f = podo.fourier_coefficients(ref["temporal_modes"], N, ref["nm"], 0.01, 0.9, direct=False)
    for view in (A, np.ascontiguousarray(A.T).T):
        i_d, dat = _check_pod(P, view, N, M // 3, nm, 0.01, 0.9, ref, f)
        sign = np.sign(np.sum(i_d.temporal_modes[:, :nm] * ref["temporal_modes"].real[:, :nm], axis=0))
        assert np.abs(i_d.c128 * sign[None, :] - f["c128"]).max() < TOL * np.abs(f["c128"]).max() * 10
        assert np.array_equal(i_d.c_count, f["c_count"])


def test_pod_weighted_branch_against_oracle(P):
    M, N, nm = 3 * 700, 40, 6
    A = podo.synthetic_snapshots(M, N, seed=5, rank_r=12)
    A = A - A.mean(1)[:, None]
    w = np.random.RandomState(1).uniform(0.5, 2.0, M // 3)
    ref = podo.pod(A, N, M // 3, 3, "true", w, 1.0e-15, nm)
    _check_pod(P, A, N, M // 3, nm, 0.02, 0.9, ref, None, weights=w)


@pytest.mark.parametrize("M,N", [(37, 2048), (19, 3001), (5, 10240), (11, 10241), (300, 2047)])
def test_row_means_long_rows(P, M, N):
    """podsdf_mean_rows_dev, layout 0 (DF:1492-1495): rows of 2048 .. 10240 entries take the kernel that keeps a row in
    registers between the sum and the subtraction; shorter / longer rows the two-pass kernel.  Both against numpy, with
    and without the in-place subtraction."""
    torch = pytest.importorskip("torch")
    capi, L = P._capi, P._capi.lib()
    rs = np.random.RandomState(M + N)
    A = rs.standard_normal((M, N)) + 5.0 * rs.standard_normal((M, 1))
    for sub in (0, 1):
        d = torch.from_numpy(A.copy()).cuda()
        mean = torch.full((M,), float("nan"), dtype=torch.float64, device="cuda")
        capi.check(L.podsdf_mean_rows_dev(0, d.data_ptr(), M, N, 0, mean.data_ptr(), sub, 0))
        torch.cuda.synchronize()
        assert np.abs(mean.cpu().numpy() - A.mean(1)).max() < 1e-13
        want = A - A.mean(1)[:, None] if sub else A
        assert np.abs(d.cpu().numpy() - want).max() < 1e-13


def test_device_entries_mean_gram_project(P):
    torch = pytest.importorskip("torch")
    capi = P._capi
    L = capi.lib()
    rs = np.random.RandomState(9)
    M, N, nm = 3 * 500, 200, 16
    A = rs.standard_normal((M, N)) + rs.standard_normal((M, 1)) * 3
    for layout, arr in ((0, A), (1, np.ascontiguousarray(A.T))):
        d = torch.from_numpy(arr.copy()).cuda()
        mean = torch.empty(M, dtype=torch.float64, device="cuda")
        capi.check(L.podsdf_mean_rows_dev(0, d.data_ptr(), M, N, layout, mean.data_ptr(), 1, 0))
        torch.cuda.synchronize()
        assert np.abs(mean.cpu().numpy() - A.mean(1)).max() < 1e-13
        Ac = A - A.mean(1)[:, None]
        got = d.cpu().numpy() if layout == 0 else d.cpu().numpy().T
        assert np.abs(got - Ac).max() < 1e-13
        C = torch.empty((N, N), dtype=torch.float64, device="cuda")
        capi.check(L.podsdf_gram_dev(0, d.data_ptr(), M, N, layout, None, 0, 1.0 / N, C.data_ptr(), 0))
        torch.cuda.synchronize()
        assert rel_err(C.cpu().numpy(), Ac.T @ Ac / N) < 1e-12
        lam = torch.empty(N, dtype=torch.float64, device="cuda")
        Cv = C.clone()
        capi.check(L.podsdf_eig_sym_dev(0, Cv.data_ptr(), N, lam.data_ptr(), 0))
        V = Cv.cpu().numpy()          # rows are eigenvectors
        assert np.abs(V @ V.T - np.eye(N)).max() < 1e-12
        assert np.abs(lam.cpu().numpy() - np.linalg.eigvalsh(Ac.T @ Ac / N)).max() < 1e-11


def test_config1_df_to_podfs_on_device(P):
    """BASELINE config 1 end to end on the device: 64x64x3 planes, nf=10, ln=5, hyperbolic-tangent profile,
    20 POD modes -- DF planes stay in HBM and feed PODFS in the generator's own (transposed) layout.
    300 steps here (the oracle's scipy loop takes ~10 s); checked against oracle DF -> mean -> POD -> Fourier."""
    torch = pytest.importorskip("torch")
    from oracle import df_oracle as dfo
    df = P.digitalfilters
    J = K = 64
    nf, ln, ns, nm = 10, 5.0, 300, 20
    U, uu, vv, ww, uw = dfo.build_profile("hyperbolic-tangent", "top-hat", 1.0, 0.02, K)
    _, _, _, _, _, _, dt = dfo.derive_filter_sizes(ln, 2.0, 0.1, U, 0.0)
    X = dfo.reference_stream(0, ns, J, K, nf, nf, nf)
    A = dfo.generate(X, ns, J, K, ln, ln, ln, nf, nf, nf, ("1d", U, uu, vv, ww, uw), rotate=(1.0, 0.0, 0.0))
    mean, A = dfo.remove_mean(A)
    ref = podo.pod(A, ns, J * K, 3, "false", [], 1.0e-15, nm, eig=np.linalg.eigh)
    f = podo.fourier_coefficients(ref["temporal_modes"], ns, ref["nm"], dt, 0.9, direct=False)

    plan = df._plan(J, K, nf, nf, nf, df._taps(nf, ln), df._taps(nf, ln), df._taps(nf, ln))
    plan.set_impl(0)
    plan.set_coefficients(*P._lund.pack(*P._lund.lund_1d(U, uu, vv, ww, uw, J, K), J, K, df.prof_rotation_matrix(1.0, 0.0, 0.0)))
    slabs = torch.from_numpy(np.ascontiguousarray(np.moveaxis(X[:, :ns + 2 * nf], 0, 1))).cuda()
    res = P.pipeline.df_to_podfs(plan, ns, nm, dt, 0.9, slabs=slabs)
    assert rel_err(res["mean"].cpu().numpy(), mean) < 1e-12
    assert res["nm"] == ref["nm"]
    e_ref = ref["energy"].real
    assert np.abs(res["energy"][:nm] - e_ref[:nm]).max() < TOL * e_ref[0]
    tm_ref = ref["temporal_modes"].real[:, :nm]
    sign = np.sign(np.sum(res["T"] * tm_ref, axis=0))
    gaps = np.minimum(np.abs(np.diff(e_ref[:nm + 1])), np.abs(np.diff(np.r_[np.inf, e_ref[:nm]]))) / e_ref[0]
    bound = np.maximum(1e-9 / np.maximum(gaps, 1e-6), 1e-9)
    assert (np.abs(res["T"] * sign - tm_ref).max(axis=0) / np.abs(tm_ref).max(axis=0) < bound).all()
    phi = res["Phi"].cpu().numpy()
    assert (np.abs(phi * sign - ref["spatial_modes"]).max(axis=0) / np.abs(ref["spatial_modes"]).max(axis=0) < bound).all()
    assert np.abs(phi.T @ phi - np.eye(nm)).max() < 1e-9
    assert (np.abs(res["c128"] * sign - f["c128"]).max(axis=0) / np.abs(f["c128"]).max(axis=0) < bound * 10).all()
    assert np.array_equal(res["c_count"], f["c_count"])


def test_podfs_reconstruction_evaluator(P):
    """U(x,t) = mean + sum_i Phi_i a_i(t) with a_i(t) from the kept Fourier coefficients, against
    the oracle's y2 (POD:1602-1612) and a numpy product; with all coefficients kept (et = 1) and all valid modes
    the reconstruction returns the snapshots."""
    torch = pytest.importorskip("torch")
    M, N, nm = 3 * 400, 64, 12
    A = podo.synthetic_snapshots(M, N, seed=2, rank_r=10, noise=0.0)
    mean = A.mean(1)
    A0 = A - mean[:, None]
    dev = torch.device("cuda", 0)
    res = P.pipeline.pod_resident(torch.from_numpy(A.copy()).to(dev), M, N, 0, nm, 0.05, 0.9)
    c64 = res["c128"].astype(np.complex64)
    period = N * 0.05
    times = np.linspace(0, (N - 1) * 0.05, N)
    y2 = P.pipeline.temporal_reconstruction(c64, res["c_ind"], res["c_count"], period, times)
    f = podo.fourier_coefficients(res["T"], N, res["nm"], 0.05, 0.9, reconstruct=True)
    assert np.abs(y2 - f["y2"]).max() < 1e-6 * np.abs(f["y2"]).max()          # complex64 storage on both sides
    U = P.pipeline.reconstruct_field(res["Phi"], res["mean"], y2).cpu().numpy()
    ref = mean[:, None] + res["Phi"].cpu().numpy() @ y2.T
    assert rel_err(U, ref) < 1e-12
    # full series (et = 1 keeps every harmonic): the rank-10 data is reproduced to complex64 accuracy
    c64b, indb, cntb = P.PODFS.select_fourier(res["c128"], 1.0)
    y2b = P.pipeline.temporal_reconstruction(c64b, indb, cntb, period, times)
    Ub = P.pipeline.reconstruct_field(res["Phi"][:, :10].contiguous(), res["mean"], y2b[:, :10]).cpu().numpy()
    assert np.abs(Ub - A).max() < 5e-6 * np.abs(A0).max()


def test_welch_psd_matches_reference(P):
    """`nsigproclib_no_mpi.fct_welch` on the device DFT == the reference's own outputs (SIG:10-68).  The
    reference squares complex64 block spectra, so agreement is to float32 rounding of the largest line."""
    g = golden("welch.npz")
    sp = P.nsigproclib_no_mpi
    for N in (600, 128, 75):
        for iw in (1, 2, 3):
            f, Sxx, M = sp.fct_welch(g["in_x"], float(g["in_fs"]), N, iw)
            ref = g["out_Sxx_%d_%d" % (N, iw)]
            assert M == int(g["out_M_%d_%d" % (N, iw)])
            assert np.array_equal(f, g["out_f_%d_%d" % (N, iw)])
            assert Sxx.dtype == np.complex64 and np.abs(Sxx - ref).max() < 4e-7 * np.abs(ref).max()
    with pytest.raises(SystemExit):
        sp.fct_welch(g["in_x"][:10], 1.0, 11, 1)


@pytest.mark.parametrize("N,nm,kind", [(600, 20, "separated"), (1000, 100, "separated"), (777, 33, "noise_floor"),
                                       (512, 16, "clustered"), (300, 300, "separated"), (100, 5, "separated")])
def test_leading_eigenpairs(P, N, nm, kind):
    """`podsdf_eig_sym_top_*`: all eigenvalues + the nm leading eigenvectors.  Separated spectra take
    the tridiagonal route (bisection + inverse iteration), clustered leading eigenvalues and small/odd sizes the full
    solver; either way eigenvalues, residuals and orthonormality are at rounding level."""
    rs = np.random.RandomState(N + nm)
    Q, _ = np.linalg.qr(rs.standard_normal((N, N)))
    if kind == "separated":
        lam = np.arange(1, N + 1, dtype=np.float64) ** -1.6
    elif kind == "noise_floor":                      # a few strong modes above a flat, dense floor (rank-deficient data)
        lam = np.r_[np.arange(1, 61, dtype=np.float64) ** -1.2, 1e-9 * (1 + 1e-3 * rs.uniform(size=N - 60))]
        lam[700:] = 0.0
    else:                                            # two equal leading eigenvalues
        lam = np.arange(1, N + 1, dtype=np.float64) ** -1.0
        lam[3] = lam[2]
    C = (Q * lam[None, :]) @ Q.T
    C = 0.5 * (C + C.T)
    w, V, path = P._capi.eig_sym_top(C, nm)
    expect_fast = kind in ("separated", "noise_floor") and N >= 256 and 4 * nm <= N
    assert path == (1 if expect_fast else 0), (path, kind)
    ref = np.linalg.eigvalsh(C)
    assert np.abs(w - ref).max() < 2e-14 * ref[-1] * 10
    assert (np.diff(w) >= 0).all()
    top = w[::-1][:nm]
    assert np.abs(V @ C - top[:, None] * V).max() < 1e-13 * ref[-1]
    assert np.abs(V @ V.T - np.eye(nm)).max() < 1e-13
    # same invariant subspaces as the full solver (vectors up to sign where the eigenvalue is simple)
    w2, V2 = P._capi.eig_sym(C)
    full = V2[:, ::-1][:, :nm].T
    if kind != "clustered":
        gap = np.minimum(np.abs(np.diff(np.r_[top, w[::-1][nm] if nm < N else -np.inf])),
                         np.abs(np.diff(np.r_[np.inf, top]))) / ref[-1]
        dots = np.abs(np.sum(V * full, axis=1))
        assert (1.0 - dots < 1e-12 / np.maximum(gap, 1e-12) ** 2 + 1e-13).all()


def test_field_statistics_kernels_and_reconstruction_check(P):
    """one- and two-point statistics on the device.  (a) the moment and lag-covariance kernels against
    numpy on a random three-component field, both layouts; (b) `pipeline.reconstruction_statistics` on a generated inlet:
    with every Fourier coefficient kept (et = 1) the reconstruction error equals the energy of the discarded POD modes,
    and the Reynolds stresses / two-point covariances of the reconstruction approach those of the generated field."""
    torch = pytest.importorskip("torch")
    pipeline, capi = P.pipeline, P._capi
    rs = np.random.RandomState(11)
    J, K, N = 7, 9, 40
    Pn = J * K
    F = rs.standard_normal((N, 3 * Pn)) * rs.uniform(0.5, 2.0, (1, 3 * Pn)) + rs.uniform(-3, 3, (1, 3 * Pn))
    F[:, Pn:2 * Pn] += 0.5 * F[:, :Pn]
    Fd = torch.from_numpy(F).cuda()
    Ft = torch.from_numpy(np.ascontiguousarray(F.T)).cuda()
    m = F.mean(0).reshape(3, Pn)
    a = F.reshape(N, 3, Pn) - m[None]
    want = np.stack([(a[:, x] * a[:, y]).mean(0) for x, y in ((0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2))])
    lags = [(0, 0, 0), (1, 0, 0), (3, 0, 0), (0, 1, 0), (0, 2, 0), (0, 0, 1), (0, 0, 4), (2, 1, 3)] + [(5, 0, 0)] * 10
    for layout, T in ((1, Fd), (0, Ft)):
        mean, stress = pipeline.field_moments(T, Pn, N, layout)
        assert np.abs(mean.cpu().numpy() - m).max() < 1e-13
        assert np.abs(stress.cpu().numpy() - want).max() < 1e-12
        for comp in (0, 2):
            got = pipeline.lag_covariance(T, J, K, N, layout, comp, lags, mean)
            f = a[:, comp].reshape(N, J, K)
            ref = [np.mean(f[:N - dt, :J - dj, :K - dk] * f[dt:, dj:, dk:]) for dt, dj, dk in lags]
            assert np.abs(got - np.array(ref)).max() < 1e-13
    # (b) a generated inlet through PODFS and back
    df = P.digitalfilters
    J = K = 32
    nf, ln, ns, nm = 4, 2.0, 160, 24
    U, uu, vv, ww, uw = df.build_profile("hyperbolic-tangent", "top-hat", 1.0, 0.05, K)
    plan = df._plan(J, K, nf, nf, nf, df._taps(nf, ln), df._taps(nf, ln), df._taps(nf, ln))
    plan.set_impl(0)
    plan.set_coefficients(*P._lund.pack(*P._lund.lund_1d(U, uu, vv, ww, uw, J, K), J, K))
    plan.set_rotation(None)
    res = pipeline.df_to_podfs(plan, ns, nm, 0.01, 1.0, seed=5)
    st = pipeline.reconstruction_statistics(res, res["planes"].view(ns, -1), J, K, 0.01)
    assert abs(st["error_energy"] - st["truncation_energy"]) < 1e-6 * max(st["truncation_energy"], 1e-12) + 1e-9
    assert 0.0 < st["error_energy"] < 0.9
    sg, sr = st["stress_generated"].cpu().numpy(), st["stress_reconstructed"].cpu().numpy()
    assert np.all(sr[:3].sum(1) <= sg[:3].sum(1) * (1 + 1e-12))            # a truncation only removes energy
    assert np.abs(st["mean_reconstructed"].cpu().numpy()).max() < 1e-12    # fluctuation field: zero mean
    # covariances: reconstruction variance = generated variance minus the discarded energy (per unit volume)
    assert abs((st["cov_generated"][0] - st["cov_reconstructed"][0])) <= st["cov_generated"][0]


@pytest.mark.parametrize("N,kind", [(3, "indef"), (5, "gram"), (33, "indef"), (34, "gram"), (64, "indef"), (65, "gram"),
                                    (100, "indef"), (300, "gram"), (515, "indef"), (1029, "gram"), (2500, "gram")])
def test_two_stage_tridiagonalisation(P, N, kind):
    """The library-free reduction behind podsdf_eig_sym_top_* (csrc/sbr_kernels.cu): the band matrix after stage 1 and
    the tridiagonal matrix after stage 2 are orthogonally similar to C -- their spectra equal LAPACK's spectrum of C to
    rounding -- for a Gram-like (decaying spectrum + noise floor) and an indefinite symmetric matrix, sizes that are not
    multiples of the panel width or of the tile size (N = 2500 takes the 128-row panel QR for its first panels, the others
    the 64-row one; the 256-row one runs in test_gpu_fullsize at N = 10^4)."""
    from scipy.linalg import eigvalsh_tridiagonal
    rs = np.random.RandomState(N)
    if kind == "gram":
        r = min(N, 120)
        X = rs.standard_normal((r, N)) * (np.arange(1, r + 1)[:, None] ** -0.8)
        X = np.concatenate([X, 1e-3 * rs.standard_normal((N, N))])
        C = X.T @ X / N
    else:
        A = rs.standard_normal((N, N))
        C = (A + A.T) / 2
    lam = np.linalg.eigvalsh(C)
    nrm = np.abs(lam).max()
    d, e, band, ms = P._capi.sym_tridiag(C, want_band=True)
    Bm = np.zeros((N, N))
    for dd in range(min(33, N)):
        v = band[:N - dd, dd]
        Bm[np.arange(dd, N), np.arange(N - dd)] = v
        Bm[np.arange(N - dd), np.arange(dd, N)] = v
    assert np.abs(np.linalg.eigvalsh(Bm) - lam).max() < 1e-12 * nrm
    assert np.abs(eigvalsh_tridiagonal(d, e) - lam).max() < 1e-12 * nrm
    assert abs(d.sum() - np.trace(C)) < 1e-11 * nrm * N


def test_packed_triangle_roundtrip(P):
    """podsdf_sym_pack_dev / _unpack_dev (the Gram all-reduce moves N (N + 1) / 2 doubles): unpack(pack(C)) == C for an
    exactly symmetric C, and the packed order is row-major over j >= i."""
    torch = pytest.importorskip("torch")
    capi = P._capi
    L = capi.lib()
    N = 301
    rs = np.random.RandomState(3)
    A = rs.standard_normal((N, N))
    C = A + A.T
    Cd = torch.from_numpy(C).cuda()
    packed = torch.empty(N * (N + 1) // 2, dtype=torch.float64, device="cuda")
    out = torch.full((N, N), float("nan"), dtype=torch.float64, device="cuda")
    capi.check(L.podsdf_sym_pack_dev(0, Cd.data_ptr(), N, packed.data_ptr(), 0))
    capi.check(L.podsdf_sym_unpack_dev(0, packed.data_ptr(), N, out.data_ptr(), 0))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), C)
    assert np.array_equal(packed.cpu().numpy(), C[np.triu_indices(N)])
