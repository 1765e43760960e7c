"""Generate tests/golden/*.npz from the reference's OWN code (TEST INFRASTRUCTURE ONLY).

Runs the real `/root/reference/digitalfilters.py` / `PODFS.py` under Python 3 through
the mechanical shim `oracle/ref_py3.py` (only possible in the build container, where
`/root/reference` exists) and stores small input/output vectors.  The fixtures travel
to the GPU box; the reference does not.  Re-run with `python -m oracle.make_golden`.

Every array stored under key `out_*` was produced by reference code; `in_*` are the
inputs it was given.  Nothing here calls the restatement in df_oracle.py / pod_oracle.py.
"""
import io
import contextlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        return fn(*a, **k)


def golden_taps(DF):
    out = {}
    for n, ln in [(3, 1.5), (10, 5.0), (6, 3.0), (16, 8.0), (4, 2.0), (8, 4.0), (5, 1.7)]:
        out["out_b_n%d_ln%s" % (n, ln)] = DF.calccoeff(np.zeros(2 * n + 1), n, ln)
    np.savez(os.path.join(GOLD, "taps.npz"), **out)


def golden_filter(DF):
    rs = np.random.RandomState(11)
    out = {}
    for tag, (nfx, nfy, nfz, J, K, lnx, lny, lnz) in {
        "iso": (3, 3, 3, 6, 7, 1.5, 1.5, 1.5),
        "aniso": (4, 2, 3, 9, 5, 2.0, 1.0, 1.6),
        "wide": (10, 10, 10, 8, 12, 5.0, 5.0, 5.0),
    }.items():
        x = rs.uniform(-np.sqrt(3.0), np.sqrt(3.0), size=(2 * nfx + 1, J + 2 * nfy, K + 2 * nfz))
        a = np.zeros((1, 2 * nfx + 1, 2 * nfy + 1, 2 * nfz + 1))
        DF.coeff3D(a, nfx, nfy, nfz, lnx, lny, lnz)
        y_dense = np.zeros((J, K))
        DF.filter3D(x, y_dense, a, J, K, nfx, nfy, nfz)
        y_sep = np.zeros((J, K))
        DF.filter3DSciPy1D(x, y_sep, a, J, K, lnx, lny, lnz, nfx, nfy, nfz)
        out["in_%s_params" % tag] = np.array([nfx, nfy, nfz, J, K, lnx, lny, lnz], dtype=np.float64)
        out["in_%s_x" % tag] = x
        out["out_%s_a" % tag] = a
        out["out_%s_dense" % tag] = y_dense
        out["out_%s_sep" % tag] = y_sep
    np.savez(os.path.join(GOLD, "filter.npz"), **out)


def golden_lund(DF):
    rs = np.random.RandomState(5)
    out = {}
    J, K = 9, 11
    U, uu, vv, ww, uw = DF.build_profile("hyperbolic-tangent", "top-hat", 1.3, 0.05, K)
    out["out_build_profile"] = np.stack([U, uu, vv, ww, uw])
    # a profile with shear stress; realizable in the middle, deliberately unrealizable at
    # one point so the raw path (adapt1d) yields NaN and the clamped paths clamp
    z = np.linspace(0, 1, K)
    s = 16 * z ** 2 * (1 - z) ** 2
    U2 = 1 - (2 * z - 1) ** 8
    uu2, vv2, ww2 = 0.01 * s, 0.004 * s, 0.003 * s
    uw2 = -0.002 * s * (1 - 2 * z)
    ww2[3] = 0.2 * uw2[3] ** 2 / uu2[3]          # negative radicand for a22
    prof = np.stack([U2, uu2, vv2, ww2, uw2])
    out["in_profile2"] = prof
    y0 = rs.standard_normal((3, J, K))
    out["in_y"] = y0
    yu, yv, yw = (y0[c].copy() for c in range(3))
    with np.errstate(invalid="ignore"):
        DF.adapt1d(yu, yv, yw, U, uu, vv, ww, uw, J, K)
    out["out_adapt1d_tophat"] = np.stack([yu, yv, yw])
    yu, yv, yw = (y0[c].copy() for c in range(3))
    with np.errstate(invalid="ignore"):
        DF.adapt1d(yu, yv, yw, U2, uu2, vv2, ww2, uw2, J, K)
    out["out_adapt1d_shear"] = np.stack([yu, yv, yw])
    for mp in ("double-hyperbolic-tangent", "circular-hyperbolic-tangent", "ring-hyperbolic-tangent"):
        for tag, P in (("tophat", (U, uu, vv, ww, uw)), ("shear", (U2, uu2, vv2, ww2, uw2))):
            if mp.startswith("double") and J != K and False:
                continue
            yu, yv, yw = (y0[c].copy() for c in range(3))
            with np.errstate(invalid="ignore"):
                DF.adapt2d(yu, yv, yw, P[0], P[1], P[2], P[3], P[4], J, K, mp, 0.5)
            out["out_adapt2d_%s_%s" % (mp.split("-")[0], tag)] = np.stack([yu, yv, yw])
    # adapt2prf: full 2-D stress fields, some realizable, some forcing every clamp
    f = rs.standard_normal((6, J, K)) * 0.05
    uu3 = f[0] ** 2 + 0.01
    vv3 = f[1] ** 2 + 0.01
    ww3 = f[2] ** 2 + 0.01
    uv3, uw3, vw3 = f[3], f[4], f[5]          # large off-diagonals: clamps fire at many points
    uu3[0, 0] = 0.0                            # a00 = 0 branch
    uu3[1, 1] = 1e-6; uv3[1, 1] = 0.5          # a10^2 > R11 branch
    m3 = rs.standard_normal((3, J, K))
    out["in_prf"] = np.stack([m3[0], m3[1], m3[2], uu3, vv3, ww3, uv3, uw3, vw3])
    yu, yv, yw = (y0[c].copy() for c in range(3))
    with np.errstate(invalid="ignore"):
        DF.adapt2prf(yu, yv, yw, m3[0], m3[1], m3[2], uu3, vv3, ww3, uv3, uw3, vw3, J, K)
    out["out_adapt2prf"] = np.stack([yu, yv, yw])
    # rotation
    normals = [(1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (-1.0, 0.0, 0.0), (0.6, -0.48, 0.64), (0.0, 0.0, -1.0),
               (1 / np.sqrt(3),) * 3]
    out["in_normals"] = np.array(normals)
    out["out_rotmat"] = np.stack([DF.prof_rotation_matrix(*n) for n in normals])
    col = rs.standard_normal(3 * J * K)
    out["in_col"] = col
    out["out_rotated"] = np.stack([DF.rotate_velocity(col, *n) for n in normals])
    np.savez(os.path.join(GOLD, "lund.npz"), **out)


def golden_main(DF, POD, tag, argv, seed):
    """Run the reference's real main() (DF:1134-1510) with a seeded global RNG and capture
    what it hands to / gets back from PODFS."""
    cap = {}
    real_POD = POD.POD

    def spy_POD(A, *rest):
        cap["A"] = np.array(A)                 # mean-removed snapshot matrix (DF:1492-1495)
        i_d = rest[-1]
        cap["mean_field"] = np.array(i_d.mean_field)
        real_POD(A, *rest)
        cap["temporal_modes"] = np.array(i_d.temporal_modes)
        cap["spatial_modes"] = np.array(i_d.spatial_modes)
        cap["nm"] = i_d.nm

    POD.make_inflow_plane = lambda i_d: None    # VTK geometry only (POD:1243-1290)
    POD.pod2prf = lambda i_d: None              # writers (POD:1662-1800)
    POD.save_plane = lambda *a: None
    DF.pod.POD = spy_POD
    real_fc = POD.fourier_coefficients

    def spy_fc(i_d):
        i_d.hdf5 = True                         # so FC / N_FC / period are published (POD:1627-1639)
        real_fc(i_d)
        cap["N_FC"] = np.array(i_d.N_FC)
        cap["FC"] = np.array(i_d.FC)
        cap["period"] = i_d.period
        cap["dt"] = i_d.dt
        cap["ns"] = i_d.ns
        i_d.hdf5 = False

    DF.pod.fourier_coefficients = spy_fc
    tmp = os.path.join(os.path.dirname(HERE), "build", "golden_tmp_" + tag)
    os.makedirs(tmp, exist_ok=True)
    cwd = os.getcwd()
    os.chdir(tmp)
    old_argv = sys.argv
    try:
        if "PROFILE" in argv:
            with open("profile_in.txt", "w") as f:
                f.write(PROFILE_TXT)
            argv = [os.path.abspath("profile_in.txt") if a == "PROFILE" else a for a in argv]
        if "PRF" in argv:
            DF.plt = _NoPlot()
            with open("inlet.prf", "w") as f:
                f.write(MAIN_PRF_TXT)
            argv = [os.path.abspath("inlet.prf") if a == "PRF" else a for a in argv]
        sys.argv = ["digitalfilters.py"] + argv
        np.random.seed(seed)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            _quiet(DF.main)
        cap["podfs_dat"] = open(os.path.join("PODFS", "PODFS.dat")).read()
    finally:
        sys.argv = old_argv
        os.chdir(cwd)
        DF.pod.POD = real_POD
        DF.pod.fourier_coefficients = real_fc
    np.savez(os.path.join(GOLD, "main_%s.npz" % tag),
             in_argv=np.array(MAIN_CASES[tag][0]), in_seed=seed, in_profile_text=np.array(PROFILE_TXT),
             in_prf_text=np.array(MAIN_PRF_TXT),
             out_A=cap["A"], out_mean_field=cap["mean_field"],
             out_temporal_modes=cap["temporal_modes"], out_spatial_modes=cap["spatial_modes"],
             out_nm=cap["nm"], out_N_FC=cap["N_FC"], out_FC=cap["FC"], out_period=cap["period"],
             out_dt=cap["dt"], out_ns=cap["ns"], out_podfs_dat=np.array(cap["podfs_dat"]))
    return cap


# half-channel profile: wall at y=0, centre at y=1 (read_profile mirrors it)
PROFILE_TXT = """y U uu vv ww uv
0.00 0.000 0.0000 0.0000 0.0000 0.00000
0.05 0.420 0.0061 0.0009 0.0021 -0.00110
0.12 0.650 0.0083 0.0019 0.0030 -0.00160
0.25 0.800 0.0064 0.0022 0.0029 -0.00140
0.45 0.910 0.0041 0.0019 0.0023 -0.00095
0.70 0.975 0.0024 0.0014 0.0016 -0.00048
1.00 1.000 0.0016 0.0011 0.0012 0.00000
"""


def golden_read_profile(DF):
    """read_profile (DF:487-522) on a small 1-D profile file (text stored in the fixture)."""
    tmp = os.path.join(os.path.dirname(HERE), "build", "golden_profile.txt")
    os.makedirs(os.path.dirname(tmp), exist_ok=True)
    with open(tmp, "w") as f:
        f.write(PROFILE_TXT)
    out = {"in_text": np.array(PROFILE_TXT)}
    for kma in (11, 32):
        out["out_kma%d" % kma] = np.stack(_quiet(DF.read_profile, tmp, kma))
    np.savez(os.path.join(GOLD, "read_profile.npz"), **out)


def prf_text(J, K, res, normal=(1, 0, 0), centre=(0, 0, 0), jitter=0.0, seed=0, cols="k,e"):
    """A synthetic 2-D inlet profile in the `.prf` layout the reference reads (DF:524-575) and writes
    (POD:854-887): header lines, a `data,<names>` line, comma-separated rows.  J x K points on a plane with the
    given normal/centre, optionally jittered like an unstructured CFD face list; a jet-like velocity along the
    normal plus small in-plane components, k and epsilon (or omega) fields."""
    rs = np.random.RandomState(seed)
    n = np.array(normal, float)
    n /= np.linalg.norm(n)
    a = np.cross(n, [0, 0, 1.0])
    a /= np.linalg.norm(a)
    b = np.cross(n, a)
    Y, Z = np.meshgrid((np.arange(J) - (J - 1) / 2) * res, (np.arange(K) - (K - 1) / 2) * res, indexing="ij")
    Y = Y + jitter * res * rs.uniform(-1, 1, Y.shape)
    Z = Z + jitter * res * rs.uniform(-1, 1, Z.shape)
    Y, Z = Y.reshape(-1), Z.reshape(-1)
    pts = np.array(centre)[None, :] + Y[:, None] * a[None, :] + Z[:, None] * b[None, :]
    r = np.sqrt((Y / (J * res / 2)) ** 2 + (Z / (K * res / 2)) ** 2)
    vel = (2.0 * (1 - 0.8 * r ** 2))[:, None] * n[None, :] + (0.1 * np.sin(3 * Y))[:, None] * a[None, :] \
        - (0.05 * np.cos(2 * Z))[:, None] * b[None, :]
    k = 0.02 * (1 + r) ** 2
    second = 0.5 * (0.3 + r) if cols.endswith("e") else 40.0 * (0.3 + r)
    lines = ["# synthetic inlet profile", "units: SI", "data,x,y,z,u,v,w,%s" % cols]
    for i in range(len(r)):
        lines.append(",".join("%.12f" % x for x in (*pts[i], *vel[i], k[i], second[i])))
    return "\n".join(lines) + "\n"


class _NoPlot:
    """read_prf draws 16 contour plots through nplotlib (DF:846-867, 1012-1023); not part of its result."""

    def contourf(self, *a, **k):
        pass

    def close(self, *a, **k):
        pass


PRF_CASES = {
    # tag: (prf_text kwargs, (res, mdot, den, bulk_velocity, non_dim, TestGrad))
    "aligned": (dict(J=9, K=12, res=0.1), (0.1, 0.0, 0.0, 1.0, False, False)),
    "massflow": (dict(J=10, K=8, res=0.05, normal=(1, 0.5, 0.3), centre=(1, 2, 3), jitter=0.2, seed=1),
                 (0.05, 0.3, 1.2, 1.0, False, False)),
    "bulk_sdr": (dict(J=8, K=11, res=0.05, normal=(0.2, 1, -0.4), centre=(0, -1, 0.5), jitter=0.15, seed=2,
                      cols="k,sdr"), (0.04, 0.0, 0.0, 2.5, True, False)),
    "testgrad": (dict(J=9, K=12, res=0.1), (0.1, 0.0, 0.0, 1.0, False, True)),
}
PRF_NAMES = ["U", "V", "W", "uu", "vv", "ww", "uv", "uw", "vw", "lnx", "kma", "jma", "xn", "yn", "zn", "xc", "yc", "zc"]


def golden_read_prf(DF):
    """read_prf (DF:524-1035) on synthetic .prf files (texts stored in the fixture)."""
    DF.plt = _NoPlot()
    tmp = os.path.join(os.path.dirname(HERE), "build", "golden_prf")
    os.makedirs(tmp, exist_ok=True)
    out = {}
    for tag, (kw, args) in PRF_CASES.items():
        text = prf_text(**kw)
        path = os.path.join(tmp, tag + ".prf")
        with open(path, "w") as f:
            f.write(text)
        res = _quiet(DF.read_prf, path, *args)
        out["in_text_" + tag] = np.array(text)
        out["in_args_" + tag] = np.array(args, dtype=np.float64)
        for name, v in zip(PRF_NAMES, res):
            out["out_%s_%s" % (name, tag)] = np.asarray(v)
    np.savez(os.path.join(GOLD, "read_prf.npz"), **out)


def golden_welch(SIG):
    """fct_welch (SIG:10-68): Welch PSD of a two-tone test signal, the three windows, even and odd block
    sizes, one and several 50 %-overlap blocks."""
    rs = np.random.RandomState(11)
    t = np.arange(600) * 0.01
    x = np.sin(2 * np.pi * 7.0 * t) + 0.4 * np.cos(2 * np.pi * 19.0 * t + 0.3) + 0.05 * rs.standard_normal(t.size)
    out = {"in_x": x, "in_fs": 100.0}
    for N in (600, 128, 75):
        for iw in (1, 2, 3):
            f, Sxx, M = _quiet(SIG.fct_welch, x, 100.0, N, iw)
            out["out_f_%d_%d" % (N, iw)] = np.array(f)
            out["out_Sxx_%d_%d" % (N, iw)] = np.array(Sxx)
            out["out_M_%d_%d" % (N, iw)] = M
    np.savez(os.path.join(GOLD, "welch.npz"), **out)


def golden_pod(POD):
    """calculate_correlation_matrix both branches, sort_eigenvalues with a NaN and a tie."""
    rs = np.random.RandomState(3)
    P, ns = 20, 9
    A = rs.standard_normal((3 * P, ns))
    w = rs.uniform(0.5, 2.0, P)
    out = {"in_A": A, "in_w": w}
    C = np.zeros((ns, ns)); _quiet(POD.calculate_correlation_matrix, ns, P, 3, "false", [], A, C)
    out["out_C_false"] = C.copy()
    C = np.zeros((ns, ns)); _quiet(POD.calculate_correlation_matrix, ns, P, 3, "true", w, A, C)
    out["out_C_true"] = C.copy()
    en = np.array([3.0, np.nan, 5.0, 5.0, -1.0, 0.5, 7.0, 0.0, 2.0])
    tm = rs.standard_normal((ns, ns))
    out["in_energy"] = en.copy(); out["in_tm"] = tm.copy()
    POD.sort_eigenvalues(ns, en, tm)
    out["out_energy_sorted"] = en; out["out_tm_sorted"] = tm
    np.savez(os.path.join(GOLD, "pod_parts.npz"), **out)


MAIN_PRF_TXT = prf_text(J=9, K=8, res=0.1, normal=(1, 0.3, -0.2), centre=(0.5, 0, 1), jitter=0.1, seed=5)

MAIN_CASES = {
    # tag: (argv, seed)
    "default": (["-n", "14", "-l", "1.5", "-f", "2", "-k", "7", "-j", "6", "-m", "5"], 0),
    "rotated": (["-n", "12", "-l", "1.5", "-f", "2", "-k", "7", "-j", "6", "-m", "4",
                 "--nx", "1", "--ny", "1", "--nz", "0.5", "--U0", "2.0", "--u_dash", "0.1"], 1),
    "roundjet": (["-n", "12", "-l", "1.5", "-f", "2", "-k", "9", "-j", "8", "-m", "4",
                  "-p", "circular-hyperbolic-tangent", "--u_dash", "0.1"], 2),
    "square": (["-n", "10", "-l", "1.0", "-f", "2", "-k", "7", "-j", "7", "-m", "3",
                "-p", "double-hyperbolic-tangent", "--u_dash", "0.1"], 3),
    "ring": (["-n", "10", "-l", "1.0", "-f", "2", "-k", "9", "-j", "9", "-m", "3",
              "-p", "ring-hyperbolic-tangent", "--ring", "0.4", "--u_dash", "0.1"], 4),
    "dt": (["-n", "16", "-l", "1.5", "-f", "2", "-k", "7", "-j", "6", "-m", "6", "-t", "0.08",
            "-e", "0.8"], 5),
    # 1-D profile file (read_profile, shear stress, no rotation); "PROFILE" is replaced by a file holding PROFILE_TXT
    "profile": (["-n", "12", "-l", "1.5", "-f", "2", "-k", "9", "-j", "6", "-m", "4", "-i", "PROFILE"], 6),
    # 2-D .prf file (read_prf + adapt2prf; plane size, normal and origin come from the file); "PRF" -> MAIN_PRF_TXT
    "prf": (["-n", "10", "-l", "1.0", "-f", "2", "-m", "3", "-r", "0.1", "--U0", "1.5", "-i", "PRF"], 7),
}


def main():
    from oracle import ref_py3
    DF, POD = ref_py3.load()
    os.makedirs(GOLD, exist_ok=True)
    golden_taps(DF)
    golden_filter(DF)
    golden_lund(DF)
    golden_pod(POD)
    golden_read_profile(DF)
    golden_read_prf(DF)
    golden_welch(ref_py3._loaded["SIG"])
    for tag, (argv, seed) in MAIN_CASES.items():
        cap = golden_main(DF, POD, tag, argv, seed)
        print("main_%s: A %s nm %d N_FC %s" % (tag, cap["A"].shape, cap["nm"], cap["N_FC"]))
    print("golden vectors written to", GOLD)


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(HERE))
    main()
