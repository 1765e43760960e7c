"""GPU parity of the digital-filter path (through the C ABI) against the CPU oracle and the golden
vectors made by the reference's own code.  Tolerance: 1e-10 relative (FP64), the north-star bar;
actual differences are ~1e-15 (re-association of the separable sums only)."""
import os

import numpy as np
import pytest

from oracle import df_oracle as dfo
from oracle import philox as ophilox
from tests import replay
from tests.util import TOL, golden, pkg, rel_err

pytestmark = pytest.mark.gpu

IMPLS = [1, 0]   # generic kernels, auto (tuned where a compile-time tap count exists)


@pytest.fixture(scope="module")
def P():
    p = pkg()
    p._capi.require_device()
    return p


@pytest.mark.parametrize("tag", ["iso", "aniso", "wide"])
def test_filter_block_matches_reference(P, tag):
    g = golden("filter.npz")
    nfx, nfy, nfz, J, K = (int(v) for v in g["in_%s_params" % tag][:5])
    lnx, lny, lnz = g["in_%s_params" % tag][5:]
    x = g["in_%s_x" % tag]
    y = np.zeros((J, K))
    P.digitalfilters.filter3DSciPy1D(x, y, None, J, K, lnx, lny, lnz, nfx, nfy, nfz)
    assert rel_err(y, g["out_%s_sep" % tag]) < 1e-13
    y2 = np.zeros((J, K))
    P.digitalfilters.filter3D(x, y2, g["out_%s_a" % tag], J, K, nfx, nfy, nfz)   # DF:1407-1438 A/B check
    assert rel_err(y2, g["out_%s_dense" % tag]) < 1e-13


def test_adapt_variants_match_reference(P):
    df = P.digitalfilters
    g = golden("lund.npz")
    J, K = 9, 11
    U, uu, vv, ww, uw = g["out_build_profile"]
    P2 = g["in_profile2"]
    y0 = g["in_y"]

    def run(fn, *args):
        y = [y0[c].copy() for c in range(3)]
        fn(y[0], y[1], y[2], *args)
        return np.stack(y)

    assert rel_err(run(df.adapt1d, U, uu, vv, ww, uw, J, K), g["out_adapt1d_tophat"]) < 1e-14
    ref = g["out_adapt1d_shear"]
    assert np.isnan(ref).any()
    assert rel_err(run(df.adapt1d, *P2, J, K), ref) < 1e-14          # NaN where the reference has NaN
    for mp in ("double", "circular", "ring"):
        for tag, prof in (("tophat", (U, uu, vv, ww, uw)), ("shear", tuple(P2))):
            got = run(df.adapt2d, *prof, J, K, mp + "-hyperbolic-tangent", 0.5)
            assert rel_err(got, g["out_adapt2d_%s_%s" % (mp, tag)]) < 1e-14, (mp, tag)
    assert rel_err(run(df.adapt2prf, *g["in_prf"], J, K), g["out_adapt2prf"]) < 1e-14
    for n, rot in zip(g["in_normals"], g["out_rotated"]):
        assert rel_err(df.rotate_velocity(g["in_col"], *n), rot) < 1e-14


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("tag", ["default", "rotated", "roundjet", "square", "ring", "dt", "profile", "prf"])
def test_time_loop_matches_reference_main(P, tag, impl, tmp_path):
    """The fused device loop on the reference's own random draws == the A the reference's main() built."""
    g = golden("main_%s.npz" % tag)
    prof = tmp_path / "profile_in.txt"
    prof.write_text(str(g["in_profile_text"]))
    prf = tmp_path / "inlet.prf"
    prf.write_text(str(g["in_prf_text"]))
    s = replay.setup(g["in_argv"], str(prof), str(prf))
    J, K, ns = s["J"], s["K"], s["ns"]
    X = dfo.reference_stream(int(g["in_seed"]), ns, J, K, s["nfx"], s["nfy"], s["nfz"])
    slabs = np.ascontiguousarray(np.moveaxis(X, 0, 1))
    lund = P._lund
    if s["lund"][0] == "2d":
        fm = lund.lund_2d(*s["profile"], J, K, s["mean_profile"], s["ring"])
    elif s["lund"][0] == "prf":
        fm = lund.lund_prf(*s["profile"])
    else:
        fm = lund.lund_1d(*s["profile"], J, K)
    R = P.digitalfilters.prof_rotation_matrix(*s["normal"]) if s["normal"] is not None else None
    planes = P.digitalfilters.generate(ns, J, K, s["lnx"], s["lny"], s["lnz"], s["nfx"], s["nfy"], s["nfz"],
                                       fm, R, slabs=slabs, impl=impl, chunk=5)
    A = planes.reshape(ns, -1).T
    ref = g["out_A"] + g["out_mean_field"][:, None]
    assert rel_err(A, ref) < 1e-13
    mean, A2 = dfo.remove_mean(np.array(A))
    assert rel_err(A2, g["out_A"]) < 1e-12


def _config(J, K, nfx, nfy, nfz, lnx, lny, lnz, ns, seed, lund):
    X = dfo.reference_stream(seed, ns, J, K, nfx, nfy, nfz)
    return X, dfo.generate(X, ns, J, K, lnx, lny, lnz, nfx, nfy, nfz, lund, rotate=None)


@pytest.mark.parametrize("impl", IMPLS)
def test_config1_isotropic_64(P, impl):
    """BASELINE config 1 shape: 64x64, nf=10, ln=5, hyperbolic-tangent profile (adapt1d) + rotation."""
    J = K = 64
    nf, ln, ns = 10, 5.0, 24
    U, uu, vv, ww, uw = dfo.build_profile("hyperbolic-tangent", "top-hat", 1.0, 0.02, K)
    X = dfo.reference_stream(0, ns, J, K, nf, nf, nf)
    ref = dfo.generate(X, ns, J, K, ln, ln, ln, nf, nf, nf, ("1d", U, uu, vv, ww, uw), rotate=(0.6, -0.48, 0.64))
    fm = P._lund.lund_1d(U, uu, vv, ww, uw, J, K)
    R = P.digitalfilters.prof_rotation_matrix(0.6, -0.48, 0.64)
    planes = P.digitalfilters.generate(ns, J, K, ln, ln, ln, nf, nf, nf, fm, R,
                                       slabs=np.ascontiguousarray(np.moveaxis(X, 0, 1)), impl=impl, chunk=16)
    assert rel_err(planes.reshape(ns, -1).T, ref) < TOL * 1e-2


@pytest.mark.parametrize("impl", IMPLS)
def test_config2_channel_anisotropic(P, impl):
    """BASELINE config 2 shape: 256x128, (nfx,nfy,nfz)=(16,8,4), (ln)=(8,4,2), channel profile, adapt1d."""
    J, K, ns = 256, 128, 6
    nfx, nfy, nfz, lnx, lny, lnz = 16, 8, 4, 8.0, 4.0, 2.0
    prof = dfo.channel_profile(K)
    X, ref = _config(J, K, nfx, nfy, nfz, lnx, lny, lnz, ns, 1, ("1d",) + prof)
    fm = P._lund.lund_1d(*prof, J, K)
    planes = P.digitalfilters.generate(ns, J, K, lnx, lny, lnz, nfx, nfy, nfz, fm, None,
                                       slabs=np.ascontiguousarray(np.moveaxis(X, 0, 1)), impl=impl)
    assert rel_err(planes.reshape(ns, -1).T, ref) < TOL * 1e-2


def test_config3_round_jet_512(P):
    """BASELINE config 3 shape at full plane size: 512x512, nf=10, round jet (adapt2d), injected window."""
    J = K = 512
    nf, ln, ns = 10, 5.0, 3
    U, uu, vv, ww, uw = dfo.build_profile("circular-hyperbolic-tangent", "top-hat", 1.0, 0.05, K)
    lund = ("2d", U, uu, vv, ww, uw, "circular-hyperbolic-tangent", 0.5)
    X, ref = _config(J, K, nf, nf, nf, ln, ln, ln, ns, 2, lund)
    fm = P._lund.lund_2d(U, uu, vv, ww, uw, J, K, "circular-hyperbolic-tangent", 0.5)
    planes = P.digitalfilters.generate(ns, J, K, ln, ln, ln, nf, nf, nf, fm, None,
                                       slabs=np.ascontiguousarray(np.moveaxis(X, 0, 1)))
    assert rel_err(planes.reshape(ns, -1).T, ref) < TOL * 1e-2


@pytest.mark.parametrize("nf", [4, 6, 8, 10, 12, 16])
@pytest.mark.parametrize("transform", ["diagonal", "shear", "rotated"])
def test_tuned_instantiations_match_oracle(P, nf, transform):
    """Every compile-time (2nf+1, 2nf+1) instantiation of the tuned plane kernels -- the per-component (diagonal)
    variant, the coupled variant with a shear stress (adapt1d pattern) and the coupled variant with a rotated plane
    normal -- and of the tuned time kernels, against the oracle port of the reference loop (DF:100-178, 1119-1131)
    on an injected stream.  The plane is two 64-column strips wide with a ragged second strip."""
    J, K, ns = 40, 80, 3
    ln = nf / 2.0
    if transform == "shear":
        prof = dfo.channel_profile(K)
    else:
        prof = dfo.build_profile("hyperbolic-tangent", "top-hat", 1.0, 0.05, K)
    normal = (0.6, -0.48, 0.64) if transform == "rotated" else None
    X = dfo.reference_stream(100 + nf, ns, J, K, nf, nf, nf)
    ref = dfo.generate(X, ns, J, K, ln, ln, ln, nf, nf, nf, ("1d",) + tuple(prof), rotate=normal)
    fm = P._lund.lund_1d(*prof, J, K)
    R = P.digitalfilters.prof_rotation_matrix(*normal) if normal else None
    planes = P.digitalfilters.generate(ns, J, K, ln, ln, ln, nf, nf, nf, fm, R,
                                       slabs=np.ascontiguousarray(np.moveaxis(X, 0, 1)), impl=P._capi.IMPL_TUNED)
    df = P.digitalfilters
    info = df._plan(J, K, nf, nf, nf, df._taps(nf, ln), df._taps(nf, ln), df._taps(nf, ln)).info()
    assert info["tuned_time"] and info["tuned_plane"] and info["diagonal"] == (transform == "diagonal")
    df._plan(J, K, nf, nf, nf, df._taps(nf, ln), df._taps(nf, ln), df._taps(nf, ln)).set_impl(P._capi.IMPL_AUTO)
    assert rel_err(planes.reshape(ns, -1).T, ref) < TOL * 1e-2


@pytest.mark.parametrize("J,K,nfx,nfy,nfz", [(1, 37, 3, 0, 4), (5, 1, 2, 3, 0), (7, 131, 4, 4, 4), (70, 65, 10, 10, 10),
                                             (9, 200, 6, 6, 6), (33, 40, 24, 5, 20), (3, 3, 0, 0, 0)])
def test_ragged_and_degenerate_planes(P, J, K, nfx, nfy, nfz):
    """2-D inflow (jma=1, nfy=0 -- DF:1264), single rows/columns, odd padded sizes, strips that end
    mid-tile, filter widths without a tuned kernel, and the identity filter."""
    ns = 2 * (2 * nfx + 1) + 3
    ln = lambda n: max(n / 2.0, 0.7)
    rs = np.random.RandomState(J * 1000 + K)
    X = rs.uniform(-dfo.SQRT3, dfo.SQRT3, (3, ns + 2 * nfx, J + 2 * nfy, K + 2 * nfz))
    ref = dfo.generate(X, ns, J, K, ln(nfx), ln(nfy), ln(nfz), nfx, nfy, nfz, None)
    for impl in IMPLS:
        planes = P.digitalfilters.generate(ns, J, K, ln(nfx), ln(nfy), ln(nfz), nfx, nfy, nfz, None, None,
                                           slabs=np.ascontiguousarray(np.moveaxis(X, 0, 1)), impl=impl, chunk=7)
        assert rel_err(planes.reshape(ns, -1).T, ref) < 1e-12, impl


def test_filter_is_linear(P):
    """Size-independent property: F(a x1 + b x2) = a F(x1) + b F(x2)."""
    J, K, nf, ns = 96, 80, 8, 20
    rs = np.random.RandomState(4)
    x1 = rs.standard_normal((ns + 2 * nf, 3, J + 2 * nf, K + 2 * nf))
    x2 = rs.standard_normal(x1.shape)
    gen = lambda x: P.digitalfilters.generate(ns, J, K, 4.0, 4.0, 4.0, nf, nf, nf, None, None, slabs=x)
    lhs = gen(2.5 * x1 - 0.75 * x2)
    rhs = 2.5 * gen(x1) - 0.75 * gen(x2)
    assert rel_err(lhs, rhs) < 1e-13


@pytest.mark.parametrize("kind", ["u32", "u53"])
def test_philox_stream_and_time_shards(P, kind):
    """Generated-RNG mode, both streams: the device stream equals its numpy restatement bit for bit; seeded
    generation equals injection of the materialised stream; any time shard (even or odd offset -- the U32 stream
    pairs slabs) reproduces its own window."""
    J, K, nf = 20, 27, 4
    df = P.digitalfilters
    p = df._plan(J, K, nf, nf, nf, df._taps(nf, 2.0), df._taps(nf, 2.0), df._taps(nf, 2.0))
    p.set_rng(P._capi.RNG_U32 if kind == "u32" else P._capi.RNG_U53)
    try:
        seed = 0x1234567890ABCDEF
        for s0, n in ((5, 9), (4, 3), (0, 1)):
            dev = p.random_slabs(seed, s0, n)
            assert np.array_equal(dev, ophilox.slabs(seed, s0, n, J + 2 * nf, K + 2 * nf, kind))
        big = p.random_slabs(seed, (1 << 33) + 1, 2)       # slab index beyond 32 bits
        assert np.array_equal(big, ophilox.slabs(seed, (1 << 33) + 1, 2, J + 2 * nf, K + 2 * nf, kind))
        ns = 30
        p.set_coefficients(None, 0)
        for impl in IMPLS:
            p.set_impl(impl)
            p.set_chunk(11)
            whole = p.generate_seeded(seed, 0, ns)
            inj = p.generate_injected(p.random_slabs(seed, 0, ns + 2 * nf))
            # same numbers; the two modes may run different time-filter kernels (DFMA window / DMMA Toeplitz)
            assert rel_err(whole, inj) < 1e-14
            shard = p.generate_seeded(seed, 16, ns - 16)          # shard offsets that are multiples of 8: bitwise
            assert np.array_equal(shard, whole[16:])
            shard = p.generate_seeded(seed, 17, ns - 17)          # any other offset: same numbers to rounding
            assert rel_err(shard, whole[17:]) < 1e-14
            odd = p.generate_seeded(seed + 1, 3, 9)               # odd start of the stream itself
            assert rel_err(odd, p.generate_injected(p.random_slabs(seed + 1, 3, 9 + 2 * nf))) < 1e-14
        ref = dfo.generate(np.moveaxis(ophilox.slabs(seed, 0, ns + 2 * nf, J + 2 * nf, K + 2 * nf, kind), 1, 0),
                           ns, J, K, 2.0, 2.0, 2.0, nf, nf, nf, None)
        assert rel_err(whole.reshape(ns, -1).T, ref) < 1e-13
    finally:
        p.set_rng(P._capi.RNG_U32)


@pytest.mark.parametrize("nf", [8, 10, 16])
def test_seeded_equals_injected_tuned_sizes(P, nf):
    """Every tuned time-filter size the seeded path can take (DFMA window at 17 / 21 taps, DMMA Toeplitz at 33)
    against injection of the same materialised U32 stream, odd and even chunk starts."""
    J = K = 64
    df = P.digitalfilters
    ln = nf / 2.0
    p = df._plan(J, K, nf, nf, nf, df._taps(nf, ln), df._taps(nf, ln), df._taps(nf, ln))
    p.set_coefficients(None, 0)
    p.set_impl(P._capi.IMPL_TUNED)
    try:
        for t0, ns, chunk in ((0, 50, 1280), (7, 45, 16), (1, 96, 40)):
            p.set_chunk(chunk)
            a = p.generate_seeded(11, t0, ns)
            X = p.random_slabs(11, t0, ns + 2 * nf)
            assert np.array_equal(X, ophilox.slabs(11, t0, ns + 2 * nf, J + 2 * nf, K + 2 * nf))
            assert rel_err(a, p.generate_injected(X)) < 1e-14
        ref = dfo.generate(np.moveaxis(X, 1, 0), ns, J, K, ln, ln, ln, nf, nf, nf, None)
        assert rel_err(a.reshape(ns, -1).T, ref) < 1e-13
    finally:
        p.set_impl(P._capi.IMPL_AUTO)


def test_generated_field_statistics(P):
    """unit variance, two-point correlations sum_j b_j b_{j+r} along y, z and t, and the
    prescribed Reynolds stresses / mean after the Lund transform (k = 4 sigma bands)."""
    J = K = 128
    nf, ln, ns = 6, 3.0, 600
    df = P.digitalfilters
    b = df._taps(nf, ln)
    planes = df.generate(ns, J, K, ln, ln, ln, nf, nf, nf, None, None, seed=11)
    f = planes[:, 0]
    n_eff = ns * J * K / (8.0 * ln ** 3)
    tol = 4.0 * np.sqrt(2.0 / n_eff) + 2e-3
    assert abs(f.mean()) < 4.0 / np.sqrt(n_eff)
    assert abs(f.var() - 1.0) < tol
    for r in (1, 2, 4, 7):
        rho = float(np.sum(b[:-r] * b[r:]))
        assert abs(np.mean(f[:, :-r, :] * f[:, r:, :]) - rho) < tol, ("y", r)
        assert abs(np.mean(f[:, :, :-r] * f[:, :, r:]) - rho) < tol, ("z", r)
        assert abs(np.mean(f[:-r] * f[r:]) - rho) < tol, ("t", r)
    # independent components
    assert abs(np.mean(planes[:, 0] * planes[:, 1])) < tol
    # Lund: channel-like profile with shear stress
    prof = dfo.channel_profile(K)
    fm = P._lund.lund_1d(*prof, J, K)
    pl = df.generate(ns, J, K, ln, ln, ln, nf, nf, nf, fm, None, seed=12)
    U, uu, vv, ww, uw = prof
    up = pl - pl.mean(axis=(0, 2), keepdims=True)
    band = 6.0 / np.sqrt(ns * J / (4.0 * ln * ln))
    assert np.abs(pl[:, 0].mean(axis=(0, 1)) - U).max() < band * np.sqrt(uu.max())
    assert np.abs((up[:, 0] ** 2).mean(axis=(0, 1)) - uu).max() < band * uu.max() * 2
    assert np.abs((up[:, 1] ** 2).mean(axis=(0, 1)) - vv).max() < band * vv.max() * 2
    assert np.abs((up[:, 2] ** 2).mean(axis=(0, 1)) - ww).max() < band * ww.max() * 2
    assert np.abs((up[:, 0] * up[:, 2]).mean(axis=(0, 1)) - uw).max() < band * np.sqrt(uu.max() * ww.max()) * 2


def test_device_pointer_entries_and_kernel_timers(P):
    """The *_dev entries (device pointers + stream) give the same planes as the *_host entries."""
    torch = pytest.importorskip("torch")
    J, K, nf, ns = 64, 96, 10, 40
    df = P.digitalfilters
    p = df._plan(J, K, nf, nf, nf, df._taps(nf, 5.0), df._taps(nf, 5.0), df._taps(nf, 5.0))
    p.set_impl(0)
    p.set_chunk(16)
    p.set_coefficients(None, 0)
    host = p.generate_seeded(5, 3, ns)
    out = torch.empty((ns, 3, J, K), dtype=torch.float64, device="cuda:0")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        p.generate_seeded_dev(5, 3, ns, out.data_ptr(), st.cuda_stream)
    st.synchronize()
    assert np.array_equal(out.cpu().numpy(), host)
    t_time, t_plane = p.last_kernel_ms()
    assert t_time > 0 and t_plane > 0
    assert p.info() == dict(chunk=16, tuned_time=True, tuned_plane=True, diagonal=True, passes=3)
    p.set_rotation(P.digitalfilters.prof_rotation_matrix(1.0, 1.0, 0.5))
    assert p.info()["diagonal"] is False
    p.set_rotation(None)
    slabs = torch.empty((ns + 2 * nf, 3, J + 2 * nf, K + 2 * nf), dtype=torch.float64, device="cuda:0")
    p.random_slabs_dev(5, 3, ns + 2 * nf, slabs.data_ptr(), 0)
    out2 = torch.empty_like(out)
    p.generate_injected_dev(slabs.data_ptr(), ns + 2 * nf, out2.data_ptr(), 0)
    torch.cuda.synchronize()
    assert float((out - out2).abs().max()) < 1e-14     # same stream; the modes may use different time kernels
    assert P._capi.launch_count() > 0


@pytest.mark.parametrize("tag", ["default", "rotated", "roundjet", "dt", "profile", "prf"])
def test_cli_main_is_a_drop_in(P, tag, tmp_path, monkeypatch):
    """`digitalfilters.main()` with the reference's own flags and a seeded global np.random reproduces what the
    reference's main() produced (golden): snapshot matrix, mean field, POD modes up to sign, PODFS.dat."""
    g = golden("main_%s.npz" % tag)
    monkeypatch.chdir(tmp_path)
    (tmp_path / "profile_in.txt").write_text(str(g["in_profile_text"]))
    (tmp_path / "inlet.prf").write_text(str(g["in_prf_text"]))
    argv = [str(tmp_path / "profile_in.txt") if str(a) == "PROFILE" else str(a) for a in g["in_argv"]]
    argv = [str(tmp_path / "inlet.prf") if a == "PRF" else a for a in argv]
    np.random.seed(int(g["in_seed"]))
    i_d = P.digitalfilters.main(argv)
    assert rel_err(i_d.mean_field, g["out_mean_field"]) < 1e-13
    nm = int(g["out_nm"])
    assert i_d.nm == nm
    tm_ref = g["out_temporal_modes"].real[:, :nm]
    sign = np.sign(np.sum(i_d.temporal_modes[:, :nm] * tm_ref, axis=0))
    lam = np.mean(tm_ref ** 2, axis=0)
    # the last kept modes of these tiny cases sit close to each other: bound scaled by the relative gap
    gap = np.minimum(np.abs(np.diff(np.r_[lam, 0.0])), np.abs(np.diff(np.r_[np.inf, lam]))) / lam[0]
    err = np.abs(i_d.temporal_modes[:, :nm] * sign - tm_ref).max(axis=0) / np.abs(tm_ref).max(axis=0)
    assert (err < 1e-10 / np.maximum(gap, 1e-4)).all()
    err = np.abs(i_d.spatial_modes * sign - g["out_spatial_modes"]).max(axis=0) / np.abs(g["out_spatial_modes"]).max(axis=0)
    assert (err < 1e-10 / np.maximum(gap, 1e-4)).all()
    for name in ["PODFS/PODFS.dat", "PODFS/PODFS_mean.prf", "PODFS/POD.eigenvalues.dat", "PODFS/PODFS_mode_%04d.prf" % nm]:
        assert (tmp_path / name).exists(), name
    dat = (tmp_path / "PODFS" / "PODFS.dat").read_text().split("\n")
    ref = str(g["out_podfs_dat"]).split("\n")
    assert dat[:2 + nm] == ref[:2 + nm]                  # mode count, period, coefficients kept per mode
    assert len(dat) == len(ref)
    assert [l.split("\t")[0] for l in dat] == [l.split("\t")[0] for l in ref]     # same harmonics in the same order


@pytest.mark.parametrize("nf", [24, 32, 48, 64])
def test_wide_filters_config5(P, nf):
    """BASELINE config 5's wide end (nf = 24..64): the time filter runs in one or two <=65-tap passes, the plane
    filter in the per-component tuned kernel (diagonal transform) -- against the oracle, tuned and generic."""
    J, K = 40, 72
    ln = nf / 2.0
    ns = 2 * nf + 7
    U, uu, vv, ww, uw = dfo.build_profile("hyperbolic-tangent", "top-hat", 1.0, 0.05, K)
    rs = np.random.RandomState(nf)
    X = rs.uniform(-dfo.SQRT3, dfo.SQRT3, (3, ns + 2 * nf, J + 2 * nf, K + 2 * nf))
    ref = dfo.generate(X, ns, J, K, ln, ln, ln, nf, nf, nf, ("1d", U, uu, vv, ww, uw))
    fm = P._lund.lund_1d(U, uu, vv, ww, uw, J, K)
    slabs = np.ascontiguousarray(np.moveaxis(X, 0, 1))
    for impl in (2, 1):          # 2 = tuned kernels only (fails loudly if one is missing), 1 = generic
        planes = P.digitalfilters.generate(ns, J, K, ln, ln, ln, nf, nf, nf, fm, None, slabs=slabs, impl=impl, chunk=50)
        assert rel_err(planes.reshape(ns, -1).T, ref) < 1e-12, impl
    # generated stream through the wide kernels == injection of the materialised stream
    df = P.digitalfilters
    p = df._plan(J, K, nf, nf, nf, df._taps(nf, ln), df._taps(nf, ln), df._taps(nf, ln))
    p.set_impl(2)
    p.set_chunk(64)
    assert rel_err(p.generate_seeded(9, 5, ns), p.generate_injected(p.random_slabs(9, 5, ns + 2 * nf))) < 1e-14


@pytest.mark.parametrize("impl", IMPLS)
def test_time_loop_with_full_stress_tensor(P, impl):
    """adapt2prf inside the fused loop (DF:1445-1446): full symmetric R(j,k) with every clamp branch, 2-D means
    -> the general 3x3 + means pattern of the coupled plane kernel, plus a rotation on top."""
    g = golden("lund.npz")
    prf = g["in_prf"]                     # U,V,W,uu,vv,ww,uv,uw,vw on a 9 x 11 plane, clamps fire at many points
    J, K, nf, ln, ns = 9, 11, 4, 2.0, 19
    X = dfo.reference_stream(3, ns, J, K, nf, nf, nf)
    ref = dfo.generate(X, ns, J, K, ln, ln, ln, nf, nf, nf, ("prf",) + tuple(prf), rotate=(0.6, -0.48, 0.64))
    fm = P._lund.lund_prf(*prf)
    R = P.digitalfilters.prof_rotation_matrix(0.6, -0.48, 0.64)
    planes = P.digitalfilters.generate(ns, J, K, ln, ln, ln, nf, nf, nf, fm, R,
                                       slabs=np.ascontiguousarray(np.moveaxis(X, 0, 1)), impl=impl)
    assert rel_err(planes.reshape(ns, -1).T, ref) < 1e-13
    # a larger plane so that interior (unguarded) strips of the tuned kernel are exercised too
    J, K, ns = 70, 150, 12
    rs = np.random.RandomState(8)
    f = rs.standard_normal((6, J, K)) * 0.05
    fields = [rs.standard_normal((J, K)) for _ in range(3)] + [f[0] ** 2 + 0.01, f[1] ** 2 + 0.01, f[2] ** 2 + 0.01, f[3], f[4], f[5]]
    X = dfo.reference_stream(4, ns, J, K, nf, nf, nf)
    ref = dfo.generate(X, ns, J, K, ln, ln, ln, nf, nf, nf, ("prf",) + tuple(fields))
    planes = P.digitalfilters.generate(ns, J, K, ln, ln, ln, nf, nf, nf, P._lund.lund_prf(*fields), None,
                                       slabs=np.ascontiguousarray(np.moveaxis(X, 0, 1)), impl=impl)
    assert rel_err(planes.reshape(ns, -1).T, ref) < 1e-13


def test_cli_quickstart_and_verbose_outputs(P, tmp_path, monkeypatch):
    """docs/usage/quickstart.rst:72 (`-n 5`) and the verbose flag: the artefact list of the reference
    (PODFS_mean.prf, PODFS_mode_????.prf, PODFS.dat; with -v the per-step planes and temporal modes)."""
    monkeypatch.chdir(tmp_path)
    np.random.seed(1)
    i_d = P.digitalfilters.main(["-n", "5", "-v", "--rng", "philox", "--seed", "3"])
    names = sorted(p.name for p in (tmp_path / "PODFS").iterdir())
    assert "PODFS.dat" in names and "PODFS_mean.prf" in names and "POD.eigenvalues.dat" in names
    assert sum(n.startswith("PODFS_mode_") for n in names) == i_d.nm
    assert sum(n.endswith("E-01.prf") or n.endswith("E+00.prf") for n in names) == 5          # one plane file per step
    assert sum(n.startswith("POD.temporal_mode_") for n in names) >= i_d.nm
    assert sum(n.startswith("POD_PSD_mode_") for n in names) == i_d.nm and i_d.PSD.shape == (5, i_d.nm)
    first = (tmp_path / "PODFS" / "PODFS_mean.prf").read_text().split("\n")
    assert first[0] == "# PODFS_mean # name of the profile" and first[10] == "data,x,y,z,u,v,w"
    assert first[11].startswith("0.000000000000,-0.500000000000,0.550000011921,")          # CFDCodeIntegration.rst:53
    assert len(first) == 11 + 110 + 1
    tree = P.HDF5.hdf5_tree(_with_hdf5_fields(P, i_d))
    assert tree["main"]["mean"]["data"].shape == (110 * 6,) and tree["main"]["attrs"]["N_POD"] == i_d.nm


def _with_hdf5_fields(P, i_d):
    i_d.hdf5 = True
    P.PODFS.fourier_coefficients(i_d)
    P.PODFS.pod2prf(i_d)
    return i_d


def test_host_path_short_passes_match_device_path(P):
    """The host entries cut the run into passes of <= 256 planes (double-buffered D2H), the device entries into
    passes of the plan's chunk: same bits, pinned or pageable output, generated and injected stream."""
    torch = pytest.importorskip("torch")
    J, K, nf, ns = 48, 80, 6, 700
    df = P.digitalfilters
    p = df._plan(J, K, nf, nf, nf, df._taps(nf, 3.0), df._taps(nf, 3.0), df._taps(nf, 3.0))
    rs = np.random.RandomState(2)
    coef = rs.uniform(0.5, 1.5, (12, J * K))
    p.set_coefficients(coef, 0b111100010001)
    assert p.info()["chunk"] == 1280
    dev = torch.empty((ns, 3, J, K), dtype=torch.float64, device="cuda")
    p.generate_seeded_dev(77, 40, ns, dev.data_ptr(), 0)
    torch.cuda.synchronize()
    assert p.info()["passes"] == 1
    host = p.generate_seeded(77, 40, ns)                       # pageable numpy output
    assert p.info()["passes"] == 3                             # 256 + 256 + 188
    assert np.array_equal(host, dev.cpu().numpy())
    pinned, owner = P._capi.pinned_empty((ns, 3, J, K))
    p.generate_seeded(77, 40, ns, out=pinned)
    assert np.array_equal(pinned, host)
    owner.free()
    slabs = p.random_slabs(77, 40, ns + 2 * nf)
    inj = p.generate_injected(slabs)
    assert rel_err(inj, host) < 1e-14                          # injected mode runs the tensor-path time filter
    p.set_chunk(96)
    assert np.array_equal(p.generate_injected(slabs), inj)     # and is itself independent of the pass length
