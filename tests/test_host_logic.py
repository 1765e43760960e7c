"""Host-side logic of the shims (no GPU): taps, Lund coefficient set-up, eigenvalue sorting, CLI table,
inlet-plane geometry -- each against the oracle / the golden vectors made by the reference."""
import numpy as np
import pytest

from oracle import df_oracle as dfo
from oracle import pod_oracle as podo
from tests.util import golden, pkg


def test_calccoeff_and_coeff3d_bit_exact():
    df = pkg().digitalfilters
    g = golden("taps.npz")
    for key in g.files:
        n, ln = key[len("out_b_n"):].split("_ln")
        assert np.array_equal(df.calccoeff(np.zeros(2 * int(n) + 1), int(n), float(ln)), g[key])
    f = golden("filter.npz")
    nfx, nfy, nfz = (int(v) for v in f["in_aniso_params"][:3])
    a = np.zeros((1, 2 * nfx + 1, 2 * nfy + 1, 2 * nfz + 1))
    df.coeff3D(a, nfx, nfy, nfz, *f["in_aniso_params"][5:])
    assert np.array_equal(a, f["out_aniso_a"])


def _dense(coef, J, K):
    M = coef[:9].reshape(3, 3, J, K)
    m = coef[9:].reshape(3, J, K)
    return M, m


def test_lund_coefficients_reproduce_reference_outputs():
    """y = M x + m evaluated on the host from the packed planes equals the reference's adapt* outputs."""
    lund = pkg()._lund
    g = golden("lund.npz")
    J, K = 9, 11
    y0 = g["in_y"]
    U, uu, vv, ww, uw = g["out_build_profile"]
    P2 = g["in_profile2"]

    def apply(fac_mean, rotation=None):
        coef, mask = lund.pack(fac_mean[0], fac_mean[1], J, K, rotation)
        for i in range(12):
            assert bool(mask >> i & 1) == bool(np.any(coef[i] != 0.0))
        M, m = _dense(coef, J, K)
        with np.errstate(invalid="ignore"):
            return np.einsum("abjk,bjk->ajk", M, y0) + m

    def close(a, b):
        assert np.array_equal(np.isnan(a), np.isnan(b))
        ok = ~np.isnan(b)
        assert np.abs(a[ok] - b[ok]).max() < 1e-14

    close(apply(lund.lund_1d(U, uu, vv, ww, uw, J, K)), g["out_adapt1d_tophat"])
    close(apply(lund.lund_1d(*P2, J, K)), g["out_adapt1d_shear"])
    for mp in ("double", "circular", "ring"):
        for tag, prof in (("tophat", (U, uu, vv, ww, uw)), ("shear", tuple(P2))):
            close(apply(lund.lund_2d(*prof, J, K, mp + "-hyperbolic-tangent", 0.5)), g["out_adapt2d_%s_%s" % (mp, tag)])
    close(apply(lund.lund_prf(*g["in_prf"])), g["out_adapt2prf"])


def test_rotation_matrix_and_profiles_bit_exact():
    df = pkg().digitalfilters
    g = golden("lund.npz")
    for n, R in zip(g["in_normals"], g["out_rotmat"]):
        assert np.array_equal(df.prof_rotation_matrix(*n), R)
    got = df.build_profile("hyperbolic-tangent", "top-hat", 1.3, 0.05, 11)
    assert np.array_equal(np.stack(got), g["out_build_profile"])
    with pytest.raises(Exception):
        df.build_profile("parabola", "top-hat", 1.0, 0.1, 5)


def test_sort_eigenvalues_matches_reference():
    pod = pkg().PODFS
    g = golden("pod_parts.npz")
    en, tm = g["in_energy"].copy(), g["in_tm"].copy()
    pod.sort_eigenvalues(len(en), en, tm)
    assert np.array_equal(en, g["out_energy_sorted"])
    assert np.array_equal(tm, g["out_tm_sorted"])


def test_count_valid_modes_matches_oracle():
    pod = pkg().PODFS
    rs = np.random.RandomState(0)
    for n in (5, 12, 40):
        e = np.sort(rs.uniform(0.1, 2.0, n))[::-1].copy()
        assert pod._count_valid(e, n, 1e-15) == podo.count_valid_modes(e, n, 1e-15)
        e[n // 2:] = -1e-3
        assert pod._count_valid(e, n, 1e-15) == podo.count_valid_modes(e, n, 1e-15)


def test_cli_flag_table_matches_reference_names():
    df = pkg().digitalfilters
    opts, _ = df._parser().parse_args(["-n", "5"])
    # defaults documented in SURVEY section 5 (DF:1143-1237)
    assert (opts.nsteps, opts.lengthscale, opts.fwidth, opts.kma, opts.jma) == (5, 3.0, 2.0, 11, 10)
    assert (opts.dt, opts.nm, opts.et, opts.res, opts.nx, opts.bulk_velocity) == (0.0, 20, 0.9, 0.1, 1.0, 1.0)
    assert (opts.turbulence_intensity, opts.ring, opts.mean_profile) == (0.02, 0.5, "hyperbolic-tangent")
    for flag in ["--u_dash", "--nsteps", "--fwidth", "--nm", "--U0", "--bulk_velocity", "--turb_profile",
                 "--rotate", "--massflow", "--density", "--hdf5", "--test_gradients", "--non_dim", "--ring"]:
        assert df._parser().has_option(flag), flag


def test_inflow_plane_matches_doc_fixture():
    """docs/usage/CFDCodeIntegration.rst:53 -- first cell of the default 10x11 plane is
    (0, -0.5, 0.55) and prints as 0.550000011921 (float32 points, "%0.12f")."""
    p = pkg()
    i_d = p.digitalfilters.obj()
    i_d.jma, i_d.kma, i_d.res, i_d.n, i_d.rot, i_d.t_o = 10, 11, 0.1, [1.0, 0.0, 0.0], 0.0, [0, 0, 0]
    grid = p.PODFS.make_inflow_plane(i_d)
    assert grid.GetNumberOfCells() == 110
    c = grid.centers[0]
    assert [p.PODFS._fmt(v) for v in c] == ["0.000000000000", "-0.500000000000", "0.550000011921"]
    # cell id = k + j*K: the second cell moves along -z
    assert grid.centers[1][2] < grid.centers[0][2] and grid.centers[1][1] == grid.centers[0][1]


def test_reference_stream_draw_order():
    df = pkg().digitalfilters
    np.random.seed(3)
    X = df.reference_stream(5, 4, 6, 2, 1, 3)
    assert np.array_equal(np.moveaxis(X, 1, 0), dfo.reference_stream(3, 5, 4, 6, 2, 1, 3))


def test_read_profile_matches_reference(tmp_path):
    """read_profile (DF:487-522): mirror the half profile, spline onto kma points, zero the walls."""
    df = pkg().digitalfilters
    g = golden("read_profile.npz")
    f = tmp_path / "profile.txt"
    f.write_text(str(g["in_text"]))
    for kma in (11, 32):
        got = np.stack(df.read_profile(str(f), kma))
        assert np.array_equal(got, g["out_kma%d" % kma])


@pytest.mark.parametrize("tag", ["aligned", "massflow", "bulk_sdr", "testgrad"])
def test_read_prf_matches_reference(tag, tmp_path):
    """read_prf (DF:524-1035), the vectorised product version, against what the reference returned for the same
    file and flags: raster size, plane normal and centre, length scale exactly; fields to rounding (the 2x2
    box mean is summed in a different order)."""
    df = pkg().digitalfilters
    g = golden("read_prf.npz")
    f = tmp_path / (tag + ".prf")
    f.write_text(str(g["in_text_" + tag]))
    a = g["in_args_" + tag]
    out = df.read_prf(str(f), a[0], a[1], a[2], a[3], bool(a[4]), bool(a[5]))
    names = ["U", "V", "W", "uu", "vv", "ww", "uv", "uw", "vw", "lnx", "kma", "jma", "xn", "yn", "zn", "xc", "yc", "zc"]
    for name, v in zip(names, out):
        ref = g["out_%s_%s" % (name, tag)]
        if name in ("U", "V", "W", "uu", "vv", "ww", "uv", "uw", "vw"):
            assert v.shape == ref.shape == (int(g["out_jma_" + tag]), int(g["out_kma_" + tag]))
            assert np.abs(v - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max()), (tag, name)
        else:
            assert np.array_equal(np.asarray(v), ref), (tag, name)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs first): exactly one JSON line on stdout with the
    contract's keys; it must not need a GPU."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--no-pod"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.split("\n") if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "df_planes_per_s" and d["unit"] == "planes/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "planes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_fourier_ranking_matches_the_tuple_sort():
    """`PODFS.select_fourier` ranks with one uint64 key sort per mode; the reference sorts (|c|, index) tuples in
    reverse (POD:1584).  Same order, counts and complex64 storage, including ties, zeros and signed zeros."""
    pod = pkg().PODFS
    rs = np.random.RandomState(17)
    for ns, nm, et in ((64, 5, 0.9), (257, 3, 0.5), (40, 4, 0.75), (31, 2, 0.999)):   # (et = 1 can overrun the reference's loop)
        c = rs.standard_normal((ns, nm)) + 1j * rs.standard_normal((ns, nm))
        c[3:9, 0] = c[2, 0]                       # a run of equal magnitudes
        c[10:14, nm - 1] = 0.0                    # exact zeros
        c[20, 0] = complex(-0.0, 0.0)
        c[5, 1] = np.conj(c[6, 1])                # equal magnitude, different value
        got = pod.select_fourier(c, et)
        ref = podo.select_coefficients(c, et, nep50=False)
        for g, r in zip(got, ref):
            assert np.array_equal(g, r)
    c = rs.standard_normal((50, 2)) + 0j
    c[7, 1] = np.nan                              # NaN column: falls back to the two-key sort, must not raise
    pod.select_fourier(c, 0.9)


def test_philox_restatement_known_answers_and_u32_stream():
    """oracle/philox.py: Philox4x32-10 against the Random123 known-answer vectors (kat_vectors of the Random123
    distribution: zero, all-ones and pi-digit inputs), and the properties of the U32 stream the device kernels
    use by default: exactly centred levels, exact end points, slab pairs sharing one call, shard reproducibility."""
    from oracle import philox as ph
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = ph.philox4x32_10(*[np.array([c], dtype=np.uint64) for c in ctr], key[0], key[1])
        assert tuple(int(g[0]) for g in got) == want
    # conversion: word 0 -> lowest level, 2^32-1 -> highest, symmetric, spacing 2 sqrt3 / 2^32
    x = ph.u32_to_uniform(np.array([0, 2 ** 32 - 1, 2 ** 31 - 1, 2 ** 31], dtype=np.uint64))
    s3 = np.sqrt(3.0)
    assert x[0] == -x[1] and x[2] == -x[3]
    assert abs(x[1] - s3 * (1 - 2.0 ** -32)) < 1e-15 and abs(x[3] - s3 * 2.0 ** -32) < 1e-24
    # stream structure
    Jp, Kp = 9, 11          # odd number of points: the last pair is half used
    a = ph.slabs(5, 0, 6, Jp, Kp)
    assert np.array_equal(ph.slabs(5, 3, 2, Jp, Kp), a[3:5])
    assert np.abs(a).max() < s3 and abs(a.mean()) < 0.1 and abs(a.var() - 1.0) < 0.1
    b = ph.slabs(5, 0, 6, Jp, Kp, kind="u53")
    assert not np.array_equal(a, b) and np.abs(b).max() < s3
