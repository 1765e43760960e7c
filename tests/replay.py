"""Replay of the reference CLI's parameter derivation (DF:1244-1354) for the golden `main_*`
cases, shared by the oracle tests (CPU) and the product parity tests (GPU).  Test helper only."""
import numpy as np

from oracle import df_oracle as dfo

DEFAULTS = dict(n=20, l=3.0, f=2.0, k=11, j=10, t=0.0, m=20, e=0.9, r=0.1, nx=1.0, ny=0.0, nz=0.0,
                U0=1.0, u_dash=0.02, ring=0.5, p="hyperbolic-tangent")


def parse(argv):
    o = dict(DEFAULTS)
    argv = [str(a) for a in argv]
    i = 0
    while i < len(argv):
        key = argv[i].lstrip("-")
        o[key] = argv[i + 1]
        i += 2
    for k in ("n", "k", "j", "m"):
        o[k] = int(o[k])
    for k in ("l", "f", "t", "e", "r", "nx", "ny", "nz", "U0", "u_dash", "ring"):
        o[k] = float(o[k])
    return o


def setup(argv, profile_path=None, prf_path=None):
    """Returns dict with everything the DF loop needs for a golden main case.  profile_path / prf_path: files to
    use where the golden argv says "-i PROFILE" / "-i PRF" (the reference ran with a temporary copy of the same
    text)."""
    o = parse(argv)
    J, K, ns = o["j"], o["k"], o["n"]
    nrm = np.sqrt(o["nx"] ** 2 + o["ny"] ** 2 + o["nz"] ** 2)          # DF:1291-1293
    normal = (o["nx"] / nrm, o["ny"] / nrm, o["nz"] / nrm)
    from_file = o.get("i") in ("PROFILE", "PRF")
    if o.get("i") == "PRF":                                               # DF:1299-1304
        r = dfo.read_prf(prf_path, o["r"], 0.0, 0.0, o["U0"], False, False)
        U, V, W, uu, vv, ww, uv, uw, vw, ln_file, K, J = r[:12]
        lnx, lny, lnz, nfx, nfy, nfz, dt = dfo.derive_filter_sizes(o["l"], o["f"], o["r"], U, o["t"], V, W, ln_file)
        return dict(J=J, K=K, ns=ns, nm=o["m"], et=o["e"], dt=dt, normal=None,
                    lund=("prf", U, V, W, uu, vv, ww, uv, uw, vw), lnx=lnx, lny=lny, lnz=lnz,
                    nfx=nfx, nfy=nfy, nfz=nfz, profile=(U, V, W, uu, vv, ww, uv, uw, vw),
                    mean_profile=o["p"], ring=o["ring"], plane_normal=r[12:15], plane_centre=r[15:18])
    if from_file:
        U, uu, vv, ww, uw = dfo.read_profile(profile_path, K)             # DF:1306-1307
    else:
        U, uu, vv, ww, uw = dfo.build_profile(o["p"], "top-hat", o["U0"], o["u_dash"], K)
    lnx, lny, lnz, nfx, nfy, nfz, dt = dfo.derive_filter_sizes(o["l"], o["f"], o["r"], U, o["t"])
    for a in (uu, vv, ww):                                                # DF:1347-1354
        a[a < 0.0] = 0.0
    if o["p"] in ("double-hyperbolic-tangent", "ring-hyperbolic-tangent", "circular-hyperbolic-tangent"):
        lund = ("2d", U, uu, vv, ww, uw, o["p"], o["ring"])
    else:
        lund = ("1d", U, uu, vv, ww, uw)
    if from_file:
        normal = None                                                     # DF:1476: rotation only without a profile file
    return dict(J=J, K=K, ns=ns, nm=o["m"], et=o["e"], dt=dt, normal=normal, lund=lund,
                lnx=lnx, lny=lny, lnz=lnz, nfx=nfx, nfy=nfy, nfz=nfz, profile=(U, uu, vv, ww, uw),
                mean_profile=o["p"], ring=o["ring"])
