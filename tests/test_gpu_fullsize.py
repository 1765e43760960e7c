"""Full-size runs (BASELINE.json configs 3 and 4, one GPU's shard) checked through size-independent
properties -- the CPU oracle cannot run these sizes, so the checks are identities the path must satisfy:
shard/chunk reproducibility and ensemble statistics for the generator; a random quadratic form and the trace
(checksums of the whole Gram matrix), eigen-residuals, orthonormality and the captured-energy identity for
PODFS.  Needs ~40 GB of HBM."""
import numpy as np
import pytest

from oracle import df_oracle as dfo
from tests.util import pkg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    return pkg()


def test_config3_full_shard_reproducible_and_statistics(P):
    """Config 3: 512x512x3, nf = 10, ln = 5, round jet, one GPU's shard of 2500 planes, generated stream."""
    torch = pytest.importorskip("torch")
    df = P.digitalfilters
    J = K = 512
    nf, ln, ns = 10, 5.0, 2500
    if torch.cuda.mem_get_info()[0] < 45e9:
        pytest.skip("needs 45 GB of free HBM")
    U, uu, vv, ww, uw = dfo.build_profile("circular-hyperbolic-tangent", "top-hat", 1.0, 0.1, K)
    fac, mean = P._lund.lund_2d(U, uu, vv, ww, uw, J, K, "circular-hyperbolic-tangent", 0.5)
    taps = df._taps(nf, ln)
    p = df._plan(J, K, nf, nf, nf, taps, taps, taps)
    p.set_coefficients(*P._lund.pack(fac, mean, J, K))
    seed = 20261018
    out = torch.empty((ns, 3, J, K), dtype=torch.float64, device="cuda")
    p.generate_seeded_dev(seed, 0, ns, out.data_ptr(), 0)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out).all())
    # (1) a time shard anywhere in the run reproduces its own planes bit for bit (offsets multiple of 8) ...
    for t0, n in ((0, 16), (1240, 24), (2480, 20)):
        part = torch.empty((n, 3, J, K), dtype=torch.float64, device="cuda")
        p.generate_seeded_dev(seed, t0, n, part.data_ptr(), 0)
        torch.cuda.synchronize()
        assert torch.equal(part, out[t0:t0 + n]), t0
    # ... and the result does not depend on the chunking of the time loop
    p.set_chunk(136)
    again = torch.empty((600, 3, J, K), dtype=torch.float64, device="cuda")
    p.generate_seeded_dev(seed, 0, 600, again.data_ptr(), 0)
    torch.cuda.synchronize()
    assert torch.equal(again, out[:600])
    del again
    # (2) ensemble statistics over the 2500 planes: mean = prescribed mean profile, variances = prescribed
    # stresses (top-hat 0.1 U), within the sampling error of ~2500 / (2 ln) independent samples
    m = out.mean(dim=0).cpu().numpy()
    v = out.var(dim=0, unbiased=False).cpu().numpy()
    full = lambda a: np.broadcast_to(np.asarray(a, dtype=np.float64), (J, K))
    target_var = np.stack([full(fac[0]) ** 2, full(fac[2]) ** 2 + full(fac[1]) ** 2,
                           full(fac[5]) ** 2 + full(fac[3]) ** 2 + full(fac[4]) ** 2])
    target_mean = np.stack([full(c) for c in mean])
    n_eff = ns / (2 * ln)
    sd = np.sqrt(target_var)
    assert np.abs(m - target_mean).max() < 7.0 * sd.max() / np.sqrt(n_eff)
    live = target_var > 1e-6
    ratio = v[live] / target_var[live]
    assert abs(ratio.mean() - 1.0) < 0.01 and ratio.min() > 0.5 and ratio.max() < 1.6
    # (3) time correlation of the streamwise component at lag r equals sum_j b_j b_{j+r} (Klein 2003 eq. 9)
    u = out[:, 0] - torch.from_numpy(target_mean[0]).cuda()
    centre = (slice(None), slice(J // 2 - 40, J // 2 + 40), slice(K // 2 - 40, K // 2 + 40))
    uc = u[centre]
    var0 = float((uc * uc).mean())
    for r in (1, 3, 5, 8):
        rho = float((uc[:-r] * uc[r:]).mean()) / var0
        expect = float(np.sum(taps[:-r] * taps[r:]))
        assert abs(rho - expect) < 0.03, (r, rho, expect)


def _synthetic_shard(torch, M, N, dev):
    """Config-4 snapshots: 200 separated modes + noise + a mean field, built in row blocks."""
    g = torch.Generator(device=dev)
    g.manual_seed(99)
    A = torch.empty((M, N), dtype=torch.float64, device=dev)
    t = torch.arange(N, dtype=torch.float64, device=dev) / N
    r = torch.arange(1, 201, dtype=torch.float64, device=dev)
    G = (torch.cos(2 * np.pi * r[:, None] * t[None, :] + 0.7 * r[:, None])
         + 0.5 * torch.sin(2 * np.pi * (r[:, None] + 3) * t[None, :])) * (r ** -0.8)[:, None]
    for m0 in range(0, M, 32768):
        m1 = min(M, m0 + 32768)
        p = (torch.arange(m0, m1, dtype=torch.float64, device=dev) + 0.5) / M
        A[m0:m1] = torch.sin(np.pi * r[None, :] * p[:, None] + 0.3 * r[None, :]) @ G
        A[m0:m1] += 1e-3 * (torch.rand((m1 - m0, N), dtype=torch.float64, device=dev, generator=g) - 0.5)
        A[m0:m1] += 1.0 + p[:, None]
    return A


def test_config4_shard_podfs_identities(P):
    """Config 4, one GPU's row shard: 393216 rows (3 components x 131072 points) x 10^4 snapshots, 100 modes."""
    torch = pytest.importorskip("torch")
    capi, L = P._capi, P._capi.lib()
    M, N, nm = 393216, 10000, 100
    if torch.cuda.mem_get_info()[0] < 60e9:
        pytest.skip("needs 60 GB of free HBM")
    dev = torch.device("cuda", 0)
    A = _synthetic_shard(torch, M, N, dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    rowmean_ref = A.mean(dim=1)
    mean = torch.empty(M, dtype=torch.float64, device=dev)
    capi.check(L.podsdf_mean_rows_dev(0, A.data_ptr(), M, N, 0, mean.data_ptr(), 1, st))
    assert float((mean - rowmean_ref).abs().max()) < 1e-13
    assert float(A.mean(dim=1).abs().max()) < 1e-13                      # mean removed in place
    C = torch.empty((N, N), dtype=torch.float64, device=dev)
    capi.check(L.podsdf_gram_dev(0, A.data_ptr(), M, N, 0, None, 0, 1.0 / N, C.data_ptr(), st))
    torch.cuda.synchronize()
    # Gram checksums: exact symmetry, trace = |A|_F^2 / N, and a random quadratic form x'Cx = |Ax|^2 / N
    assert torch.equal(C, C.T)
    fro = sum(float((A[m0:m0 + 32768] ** 2).sum()) for m0 in range(0, M, 32768))
    assert abs(float(torch.trace(C)) * N - fro) < 1e-11 * fro
    x = torch.randn(N, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
    quad = float(x @ (C @ x)) * N
    ax2 = float(((A @ x) ** 2).sum())
    assert abs(quad - ax2) < 1e-10 * ax2
    # eigen-solve: residuals of ALL pairs, orthonormality, trace
    C0 = C.clone()
    lam = torch.empty(N, dtype=torch.float64, device=dev)
    capi.check(L.podsdf_eig_sym_dev(0, C.data_ptr(), N, lam.data_ptr(), st))           # rows of C = eigenvectors
    torch.cuda.synchronize()
    V = C
    lam0 = float(lam.max())
    assert float((V @ C0 - lam[:, None] * V).abs().max()) < 1e-11 * lam0
    assert float((V @ V.T - torch.eye(N, dtype=torch.float64, device=dev)).abs().max()) < 1e-11
    assert abs(float(lam.sum()) - float(torch.trace(C0))) < 1e-12 * float(torch.trace(C0))
    assert bool((lam[1:] >= lam[:-1]).all())                                            # ascending, as documented
    # the route PODFS takes at this size (all eigenvalues + nm leading vectors, no library call: two-stage
    # tridiagonalisation with the 256-row panel QR, bulge chasing, multisection, inverse iteration, back-transformation)
    import ctypes
    C1 = C0.clone()
    lam1 = torch.empty(N, dtype=torch.float64, device=dev)
    route = ctypes.c_int(-1)
    capi.check(L.podsdf_eig_sym_top_dev(0, C1.data_ptr(), N, nm, lam1.data_ptr(), ctypes.byref(route), st))
    torch.cuda.synchronize()
    assert route.value == 1
    assert float((lam1 - lam).abs().max()) < 1e-12 * lam0
    Vt = C1[:nm]                                                                        # rows = leading vectors, descending
    lt = torch.flip(lam1, dims=[0])[:nm]
    assert float((Vt @ C0 - lt[:, None] * Vt).abs().max()) < 1e-12 * lam0
    assert float((Vt @ Vt.T - torch.eye(nm, dtype=torch.float64, device=dev)).abs().max()) < 1e-12
    del C0, C1
    # projection: orthonormal spatial modes and the captured-energy identity |A - Phi T'|_F^2 / N = sum of the rest
    lam_h = lam.cpu().numpy()
    energy, nm_kept, T = P.pipeline.select_modes(lam_h, lambda j: V[j].cpu().numpy(), N, nm)
    assert nm_kept == nm
    T_d = torch.from_numpy(np.ascontiguousarray(T)).to(dev)
    dinv = torch.from_numpy(1.0 / energy[:nm]).to(dev)
    Phi = torch.empty((M, nm), dtype=torch.float64, device=dev)
    capi.check(L.podsdf_project_dev(0, A.data_ptr(), M, N, 0, T_d.data_ptr(), nm, dinv.data_ptr(), nm, 1.0 / N,
                                    Phi.data_ptr(), st))
    torch.cuda.synchronize()
    assert float((Phi.T @ Phi - torch.eye(nm, dtype=torch.float64, device=dev)).abs().max()) < 1e-9
    rest = 0.0
    for m0 in range(0, M, 32768):
        rest += float(((A[m0:m0 + 32768] - Phi[m0:m0 + 32768] @ T_d.T) ** 2).sum())
    tail = float(energy[nm:].sum())
    assert abs(rest / N - tail) < 1e-8 * float(energy.sum())
    # Fourier fit of the kept temporal modes against torch's FFT of the same columns
    c = torch.empty((N, nm, 2), dtype=torch.float64, device=dev)
    capi.check(L.podsdf_dft_modes_dev(0, T_d.data_ptr(), N, nm, nm, c.data_ptr(), st))
    torch.cuda.synchronize()
    ref = torch.fft.fftshift(torch.fft.fft(T_d, dim=0), dim=0) / N
    got = torch.view_as_complex(c)
    assert float((got - ref).abs().max()) < 1e-12 * float(ref.abs().max())


GUARD = 8192     # doubles either side of every output


def _guarded(torch, n):
    """A device buffer of n doubles between two sentinel-filled guard zones; returns (whole, interior view)."""
    whole = torch.full((n + 2 * GUARD,), -7.25e77, dtype=torch.float64, device="cuda")
    return whole, whole[GUARD:GUARD + n]


def _guards_intact(whole):
    return bool((whole[:GUARD] == -7.25e77).all()) and bool((whole[-GUARD:] == -7.25e77).all())


@pytest.mark.parametrize("J,K,nfx,nfy,nfz", [(7, 5, 2, 2, 2), (33, 65, 4, 4, 4), (64, 64, 10, 10, 10), (17, 130, 16, 8, 4),
                                             (100, 31, 3, 6, 1), (1, 40, 2, 0, 3), (129, 127, 32, 32, 32)])
def test_device_outputs_stay_inside_their_buffers(P, J, K, nfx, nfy, nfz):
    """compute-sanitizer is unavailable on this pool: every device-pointer entry writes into a buffer with
    sentinel guard zones on both sides (ragged sizes, generic and tuned kernels, both modes), the guards must be
    untouched and every interior element overwritten."""
    torch = pytest.importorskip("torch")
    df = P.digitalfilters
    ns = 21
    p = df._plan(J, K, nfx, nfy, nfz, df._taps(nfx, max(nfx / 2, 1.0)), df._taps(nfy, max(nfy / 2, 1.0)),
                 df._taps(nfz, max(nfz / 2, 1.0)))
    rs = np.random.RandomState(1)
    coef = rs.uniform(-1, 1, (12, J * K))
    for impl in (1, 0):
        for mask in (0, 0xFFF):
            p.set_impl(impl)
            p.set_chunk(8)
            p.set_coefficients(coef if mask else None, mask)
            n_out = ns * 3 * J * K
            whole, out = _guarded(torch, n_out)
            p.generate_seeded_dev(11, 5, ns, out.data_ptr(), 0)
            torch.cuda.synchronize()
            assert _guards_intact(whole) and not bool((out == -7.25e77).any()), ("seeded", impl, mask)
            n_sl = (ns + 2 * nfx) * 3 * (J + 2 * nfy) * (K + 2 * nfz)
            wsl, slabs = _guarded(torch, n_sl)
            p.random_slabs_dev(11, 5, ns + 2 * nfx, slabs.data_ptr(), 0)
            torch.cuda.synchronize()
            assert _guards_intact(wsl) and not bool((slabs == -7.25e77).any()), ("slabs", impl, mask)
            whole2, out2 = _guarded(torch, n_out)
            p.generate_injected_dev(slabs.data_ptr(), ns + 2 * nfx, out2.data_ptr(), 0)
            torch.cuda.synchronize()
            assert _guards_intact(whole2) and _guards_intact(wsl), ("injected", impl, mask)
            assert float((out - out2).abs().max()) < 1e-13 * max(1.0, float(out.abs().max()))


@pytest.mark.parametrize("M,N,nm", [(48, 5, 3), (300, 129, 7), (3 * 1111, 257, 33), (4998, 128, 128)])
def test_pod_device_outputs_stay_inside_their_buffers(P, M, N, nm):
    torch = pytest.importorskip("torch")
    capi, L = P._capi, P._capi.lib()
    for layout in (0, 1):
        wa, a = _guarded(torch, M * N)
        a.copy_(torch.randn(M * N, dtype=torch.float64, device="cuda"))
        wm, mean = _guarded(torch, M)
        capi.check(L.podsdf_mean_rows_dev(0, a.data_ptr(), M, N, layout, mean.data_ptr(), 1, 0))
        wc, C = _guarded(torch, N * N)
        capi.check(L.podsdf_gram_dev(0, a.data_ptr(), M, N, layout, None, 0, 1.0 / N, C.data_ptr(), 0))
        wl, lam = _guarded(torch, N)
        wv, V = _guarded(torch, N * N)
        V.copy_(C)
        capi.check(L.podsdf_eig_sym_dev(0, V.data_ptr(), N, lam.data_ptr(), 0))
        wt, T = _guarded(torch, N * nm)
        T.copy_(V.view(N, N)[N - nm:].T.contiguous().view(-1))
        dinv = torch.ones(nm, dtype=torch.float64, device="cuda")
        wp, Phi = _guarded(torch, M * nm)
        capi.check(L.podsdf_project_dev(0, a.data_ptr(), M, N, layout, T.data_ptr(), nm, dinv.data_ptr(), nm, 1.0 / N,
                                        Phi.data_ptr(), 0))
        wf, c = _guarded(torch, N * nm * 2)
        capi.check(L.podsdf_dft_modes_dev(0, T.data_ptr(), N, nm, nm, c.data_ptr(), 0))
        torch.cuda.synchronize()
        for name, w, inner in (("A", wa, a), ("mean", wm, mean), ("C", wc, C), ("lam", wl, lam), ("V", wv, V),
                               ("T", wt, T), ("Phi", wp, Phi), ("c", wf, c)):
            assert _guards_intact(w), (name, layout)
            assert not bool((inner == -7.25e77).any()), (name, layout)
