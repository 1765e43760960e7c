"""The CPU oracle (oracle/*.py) against golden vectors produced by the reference's own code
(tests/golden/*.npz, made by oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import df_oracle as dfo
from oracle import pod_oracle as podo
from tests import replay

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def test_taps_bit_exact():
    g = load("taps.npz")
    for key in g.files:
        n, ln = key[len("out_b_n"):].split("_ln")
        b = dfo.calccoeff(int(n), float(ln))
        assert np.array_equal(b, g[key]), key
        assert abs(np.sum(b * b) - 1.0) < 1e-14        # DF:80-86 => unit output variance


@pytest.mark.parametrize("tag", ["iso", "aniso", "wide"])
def test_filter_matches_reference(tag):
    g = load("filter.npz")
    nfx, nfy, nfz, J, K = (int(v) for v in g["in_%s_params" % tag][:5])
    lnx, lny, lnz = g["in_%s_params" % tag][5:]
    x = g["in_%s_x" % tag]
    assert np.array_equal(dfo.coeff3d(nfx, nfy, nfz, lnx, lny, lnz), g["out_%s_a" % tag])
    y = dfo.filter3d_separable(x, J, K, lnx, lny, lnz, nfx, nfy, nfz)
    assert np.array_equal(y, g["out_%s_sep" % tag])
    yd = dfo.filter3d_dense(x, dfo.coeff3d(nfx, nfy, nfz, lnx, lny, lnz), J, K, nfy, nfz)
    assert np.array_equal(yd, g["out_%s_dense" % tag])
    # the reference's own commented A/B check (DF:1407-1438): dense == separable to rounding
    assert np.abs(yd - y).max() < 1e-13


def test_lund_variants_match_reference():
    g = load("lund.npz")
    J, K = 9, 11
    U, uu, vv, ww, uw = dfo.build_profile("hyperbolic-tangent", "top-hat", 1.3, 0.05, K)
    assert np.array_equal(np.stack([U, uu, vv, ww, uw]), g["out_build_profile"])
    P2 = g["in_profile2"]
    y0 = g["in_y"]

    def run(fn, *args):
        y = [y0[c].copy() for c in range(3)]
        fn(y[0], y[1], y[2], *args)
        return np.stack(y)

    for loops in (dfo.adapt1d, dfo.adapt1d_loops):
        with np.errstate(invalid="ignore"):
            assert np.array_equal(run(loops, U, uu, vv, ww, uw, J, K), g["out_adapt1d_tophat"])
            got = run(loops, *P2, J, K)
        ref = g["out_adapt1d_shear"]
        assert np.isnan(ref).any()                       # unrealizable point -> NaN, no clamp (DF:169)
        assert np.array_equal(got, ref, equal_nan=True)
    for mp in ("double", "circular", "ring"):
        full = mp + "-hyperbolic-tangent"
        for tag, prof in (("tophat", (U, uu, vv, ww, uw)), ("shear", tuple(P2))):
            got = run(dfo.adapt2d, *prof, J, K, full, 0.5)
            ref = g["out_adapt2d_%s_%s" % (mp, tag)]
            assert np.array_equal(got, ref, equal_nan=True), (mp, tag)
    got = run(dfo.adapt2d_round_loops, U, uu, vv, ww, uw, J, K)
    assert np.array_equal(got, g["out_adapt2d_circular_tophat"], equal_nan=True)
    prf = g["in_prf"]
    got = run(dfo.adapt2prf, *prf, J, K)
    assert np.array_equal(got, g["out_adapt2prf"], equal_nan=True)


def test_rotation_matches_reference():
    g = load("lund.npz")
    for n, R, rot in zip(g["in_normals"], g["out_rotmat"], g["out_rotated"]):
        assert np.array_equal(dfo.prof_rotation_matrix(*n), R)
        assert np.array_equal(dfo.rotate_velocity(g["in_col"], *n, loops=True), rot)
        assert np.allclose(dfo.rotate_velocity(g["in_col"], *n), rot, rtol=0, atol=1e-15)


def test_read_profile_matches_reference(tmp_path):
    g = load("read_profile.npz")
    f = tmp_path / "p.txt"
    f.write_text(str(g["in_text"]))
    for kma in (11, 32):
        assert np.array_equal(np.stack(dfo.read_profile(str(f), kma)), g["out_kma%d" % kma])


def test_pod_parts_match_reference():
    g = load("pod_parts.npz")
    A, w = g["in_A"], g["in_w"]
    ns, P = A.shape[1], A.shape[0] // 3
    C = np.zeros((ns, ns))
    podo.calculate_correlation_matrix(ns, P, 3, "false", [], A, C)
    assert np.array_equal(C, g["out_C_false"])
    podo.calculate_correlation_matrix(ns, P, 3, "true", w, A, C)
    assert np.array_equal(C, g["out_C_true"])
    en, tm = g["in_energy"].copy(), g["in_tm"].copy()
    podo.sort_eigenvalues(ns, en, tm)
    assert np.array_equal(en, g["out_energy_sorted"])
    assert np.array_equal(tm, g["out_tm_sorted"])


@pytest.mark.parametrize("tag", ["aligned", "massflow", "bulk_sdr", "testgrad"])
def test_read_prf_matches_reference(tag, tmp_path):
    """2-D .prf ingestion (DF:524-1035) against the reference's own outputs: axis-aligned raster, tilted and
    jittered face list with mass-flow rescaling, omega instead of epsilon with bulk-velocity rescaling, and
    the --test_gradients switch."""
    g = load("read_prf.npz")
    path = tmp_path / (tag + ".prf")
    path.write_text(str(g["in_text_" + tag]))
    a = g["in_args_" + tag]
    out = dfo.read_prf(str(path), a[0], a[1], a[2], a[3], bool(a[4]), bool(a[5]))
    names = ["U", "V", "W", "uu", "vv", "ww", "uv", "uw", "vw", "lnx", "kma", "jma", "xn", "yn", "zn", "xc", "yc", "zc"]
    for name, v in zip(names, out):
        assert np.array_equal(np.asarray(v), g["out_%s_%s" % (name, tag)]), (tag, name)


def test_welch_psd_matches_reference():
    """fct_welch (SIG:10-68) against the reference's own outputs."""
    g = load("welch.npz")
    for N in (600, 128, 75):
        for iw in (1, 2, 3):
            f, Sxx, M = podo.fct_welch(g["in_x"], float(g["in_fs"]), N, iw)
            assert M == int(g["out_M_%d_%d" % (N, iw)])
            assert np.array_equal(f, g["out_f_%d_%d" % (N, iw)])
            assert np.array_equal(Sxx, g["out_Sxx_%d_%d" % (N, iw)])


MAIN = ["default", "rotated", "roundjet", "square", "ring", "dt", "profile", "prf"]


@pytest.mark.parametrize("tag", MAIN)
def test_main_replay_matches_reference(tag, tmp_path):
    """The whole DF loop + mean removal + POD + Fourier fit, replayed by the oracle on the
    reference's own random draws, equals what the reference's main() produced."""
    g = load("main_%s.npz" % tag)
    prof = tmp_path / "profile_in.txt"
    prof.write_text(str(g["in_profile_text"]))
    prf = tmp_path / "inlet.prf"
    prf.write_text(str(g["in_prf_text"]))
    s = replay.setup(g["in_argv"], str(prof), str(prf))
    J, K, ns = s["J"], s["K"], s["ns"]
    X = dfo.reference_stream(int(g["in_seed"]), ns, J, K, s["nfx"], s["nfy"], s["nfz"])
    A = dfo.generate(X, ns, J, K, s["lnx"], s["lny"], s["lnz"], s["nfx"], s["nfy"], s["nfz"],
                     s["lund"], rotate=s["normal"], loops=True)
    mean, A = dfo.remove_mean(A)
    assert np.array_equal(mean, g["out_mean_field"])
    assert np.array_equal(A, g["out_A"])
    assert float(g["out_dt"]) == s["dt"]
    # vectorised forms agree with the loop forms to rounding
    A2 = dfo.generate(X, ns, J, K, s["lnx"], s["lny"], s["lnz"], s["nfx"], s["nfy"], s["nfz"],
                      s["lund"], rotate=s["normal"], loops=False)
    _, A2 = dfo.remove_mean(A2)
    assert np.abs(A2 - A).max() <= 1e-14 * max(1.0, np.abs(A).max())
    r = podo.pod(A, ns, J * K, 3, "false", [], 1.0e-15, s["nm"])
    assert r["nm"] == int(g["out_nm"])
    assert np.array_equal(r["temporal_modes"], g["out_temporal_modes"])
    assert np.array_equal(r["spatial_modes"], g["out_spatial_modes"])
    f = podo.fourier_coefficients(r["temporal_modes"], ns, r["nm"], s["dt"], s["et"], nep50=True)
    assert np.array_equal(f["c_count"], g["out_N_FC"])
    assert np.array_equal(f["FC"], g["out_FC"])
    assert f["period"] == float(g["out_period"])
    assert f["podfs_dat"] == str(g["out_podfs_dat"])
    # legacy (numpy-1.x) threshold arithmetic gives the same counts on these cases
    f2 = podo.fourier_coefficients(r["temporal_modes"], ns, r["nm"], s["dt"], s["et"], nep50=False)
    assert np.array_equal(f2["c_count"], f["c_count"])
    # FFT form of the same DFT agrees to complex64 rounding
    f3 = podo.fourier_coefficients(r["temporal_modes"], ns, r["nm"], s["dt"], s["et"], direct=False)
    assert np.abs(f3["c128"] - f["c128"]).max() < 1e-13 * np.abs(f["c128"]).max() + 1e-16
