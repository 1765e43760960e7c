"""CPU oracle for the digital-filter (DF) hot path -- TEST INFRASTRUCTURE ONLY.

Python-3 / numpy restatement of the reference's `digitalfilters.py` hot path
(Klein 2003 digital filter + Lund transform).  Nothing in the product package
imports this module: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` do, and only as the
checker / the timed CPU baseline.

Pinning: the reference ships no tests or golden vectors (docs/usage/quickstart.rst:70),
so parity is *unpinned by the reference's own tests*.  It is pinned here instead by
outputs of the reference's OWN code: `oracle/make_golden.py` runs the real
`/root/reference/*.py` through the mechanical py2->py3 shim `oracle/ref_py3.py` and
stores the results in `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks
every function below against those vectors.

Citations `DF:a-b` are `/root/reference/digitalfilters.py` lines a-b.
Conventions: J=jma (span), K=kma (wall-normal, fastest axis), nf* = HALF filter
width (2nf+1 taps), ln* = length scale in cells.
"""
import math

import numpy as np
import scipy.signal as _sig
from scipy import interpolate as _itp

SQRT3 = np.sqrt(3.0)  # DF:1344 pdfr


# --------------------------------------------------------------------------- taps
def calccoeff(n, ln):
    """DF:73-89.  b_i = exp(-pi (i-n)^2 / (2 ln^2)) / sqrt(sum a^2), i = 0..2n.
    The norm is accumulated sequentially exactly as the reference loop does."""
    a = np.zeros(2 * n + 1)
    norm = 0.0
    for i in range(2 * n + 1):
        k = float(i - n)
        a[i] = np.exp(-np.pi * k * k / (2.0 * ln * ln))
        norm = norm + a[i] ** 2
    return a / np.sqrt(norm)


def coeff3d(nfx, nfy, nfz, lnx, lny, lnz):
    """DF:46-60.  Dense outer-product kernel a[0,i,j,k] = bx_i by_j bz_k
    (normalisation block DF:62-70 is commented out in the reference)."""
    bx, by, bz = calccoeff(nfx, lnx), calccoeff(nfy, lny), calccoeff(nfz, lnz)
    a = np.zeros((1, 2 * nfx + 1, 2 * nfy + 1, 2 * nfz + 1))
    # (bx_i*by_j)*bz_k, the reference's left-to-right product order
    a[0] = (bx[:, None, None] * by[None, :, None]) * bz[None, None, :]
    return a


# ------------------------------------------------------------------------- filter
def filter3d_dense(x, a, J, K, nfy, nfz):
    """DF:91-98 (legacy loop filter, kept as the known-answer cross-check of
    DF:1407-1438): y[j,k] = sum(x[:, j:j+2nfy+1, k:k+2nfz+1] * a[0])."""
    y = np.zeros((J, K))
    for k in range(K):
        for j in range(J):
            y[j, k] = np.sum(x[:, j:j + 2 * nfy + 1, k:k + 2 * nfz + 1] * a[0])
    return y


def filter3d_separable(x, J, K, lnx, lny, lnz, nfx, nfy, nfz):
    """DF:100-140.  Three chained scipy 'valid'/'direct' 1-D convolutions,
    x (time) -> y -> z, of the block (2nfx+1, J+2nfy, K+2nfz) -> (J, K)."""
    bx = calccoeff(nfx, lnx)[:, None, None]
    by = calccoeff(nfy, lny)[None, :, None]
    bz = calccoeff(nfz, lnz)[None, None, :]
    t1 = _sig.convolve(x, bx, mode="valid", method="direct")
    t2 = _sig.convolve(t1, by, mode="valid", method="direct")
    t3 = _sig.convolve(t2, bz, mode="valid", method="direct")
    assert t3.shape == (1, J, K)
    return t3[0].copy()


def filter_taps(x, bx, by, bz):
    """Same contraction as filter3d_separable for explicit tap vectors (any
    symmetric or asymmetric taps; convolution flips, as scipy does)."""
    t1 = _sig.convolve(x, bx[:, None, None], mode="valid", method="direct")
    t2 = _sig.convolve(t1, by[None, :, None], mode="valid", method="direct")
    t3 = _sig.convolve(t2, bz[None, None, :], mode="valid", method="direct")
    return t3[0].copy()


# --------------------------------------------------------------------------- Lund
def _chol_rows(R00, R10, R11, R20, R21, R22, mode):
    """Cholesky factor of the Reynolds-stress tensor, vectorised over points.

    mode 'raw'   : DF:157-169 (adapt1d)  -- no clamps, 1e-20 guards only
    mode 'prf'   : DF:199-223 (adapt2prf) -- branchy clamps
    mode 'clamp' : DF:278-299, 365-385, 457-477 (adapt2d) -- radicand clamps; note the
                   reference clamps a temp copy of R00 but still uses sqrt(R00) (DF:279-282).
    Returns a00,a10,a11,a20,a21,a22 (same shape as the inputs)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        a00 = np.sqrt(R00)
        if mode == "raw":
            a10 = R10 / (a00 + 1e-20)
            a11 = np.sqrt(R11 - a10 * a10)
            a20 = R20 / (a00 + 1e-20)
            a21 = (R21 - a10 * a20) / (a11 + 1e-20)
            a22 = np.sqrt(R22 - a20 * a20 - a21 * a21)
        elif mode == "prf":
            pos = a00 > 0.0
            a10 = np.where(pos, R10 / (a00 + 1e-20), 0.0)
            a11 = np.where(a10 ** 2 > R11, 0.0, np.sqrt(R11 - a10 * a10))
            a20 = np.where(pos, R20 / (a00 + 1e-20), 0.0)
            a21 = np.where(a11 > 0.0, (R21 - a10 * a20) / (a11 + 1e-20), 0.0)
            rad = R22 - a20 * a20 - a21 * a21
            a22 = np.where(R22 < a20 * a20 + a21 * a21, 0.0, np.sqrt(rad))
        elif mode == "clamp":
            a10 = R10 / (a00 + 1e-20)
            t = R11 - a10 * a10
            a11 = np.sqrt(np.where(t < 0, 0.0, t))
            a20 = R20 / (a00 + 1e-20)
            a21 = (R21 - a10 * a20) / (a11 + 1e-20)
            t = R22 - a20 * a20 - a21 * a21
            a22 = np.sqrt(np.where(t < 0, 0.0, t))
        else:
            raise ValueError(mode)
    return a00, a10, a11, a20, a21, a22


def _apply(fac, mean, yu, yv, yw):
    """y = A x + mean with A lower-triangular, written like DF:176-178 (the
    explicit zero terms A01, A02, A12 are kept so NaN/inf propagate identically)."""
    a00, a10, a11, a20, a21, a22 = fac
    xu, xv, xw = yu.copy(), yv.copy(), yw.copy()
    with np.errstate(invalid="ignore"):
        yu[...] = a00 * xu + 0.0 * xv + 0.0 * xw + mean[0]
        yv[...] = a10 * xu + a11 * xv + 0.0 * xw + mean[1]
        yw[...] = a20 * xu + a21 * xv + a22 * xw + mean[2]


def adapt1d_coefficients(uin, uuin, vvin, wwin, uwin, J, K):
    """DF:143-169.  Only R00,R11,R22,R20 are set (DF:151-155), so a10=a21=0.
    Returns (fac, mean) broadcast to (J,K)."""
    z = np.zeros(K)
    bc = lambda v: np.broadcast_to(np.asarray(v, dtype=np.float64), (K,)).copy()
    fac = _chol_rows(bc(uuin), z, bc(vvin), bc(uwin), z, bc(wwin), "raw")
    fac = tuple(np.broadcast_to(f[None, :], (J, K)).copy() for f in fac)
    zero = np.zeros((J, K))
    mean = (np.broadcast_to(bc(uin)[None, :], (J, K)).copy(), zero, zero.copy())
    return fac, mean


def adapt1d(yu, yv, yw, uin, uuin, vvin, wwin, uwin, J, K):
    """DF:143-178, in place on yu,yv,yw (J,K)."""
    fac, mean = adapt1d_coefficients(uin, uuin, vvin, wwin, uwin, J, K)
    _apply(fac, mean, yu, yv, yw)


def adapt2prf_coefficients(uin, vin, win, uuin, vvin, wwin, uvin, uwin, vwin):
    """DF:180-223.  Full symmetric R(j,k) and means U,V,W(j,k)."""
    fac = _chol_rows(uuin, uvin, vvin, uwin, vwin, wwin, "prf")
    return fac, (uin, vin, win)


def adapt2prf(yu, yv, yw, uin, vin, win, uuin, vvin, wwin, uvin, uwin, vwin, J, K):
    """DF:180-231, in place."""
    fac, mean = adapt2prf_coefficients(uin, vin, win, uuin, vvin, wwin, uvin, uwin, vwin)
    _apply(fac, mean, yu, yv, yw)


def adapt2d_coefficients(uin, uuin, vvin, wwin, uwin, J, K, mean_profile, inner_d):
    """DF:233-485.  The three shaped variants; everything here is time-invariant
    (the reference recomputes it, splines included, on every step)."""
    uin, uuin, vvin, wwin, uwin = (np.asarray(v, dtype=np.float64) for v in (uin, uuin, vvin, wwin, uwin))
    spl = lambda xs, ys: _itp.splrep(xs, ys, s=0)
    if mean_profile == "double-hyperbolic-tangent":      # DF:238-307 square jet
        zA = np.linspace(-1.0, 1, K)
        zi = np.linspace(-1.0, 1, J)
        prof = []
        for p in (uin, uuin, vvin, wwin, uwin):
            pj = _itp.splev(zi, spl(zA, p), der=0)
            pj[0], pj[-1] = p[0], p[-1]                  # DF:248-258
            prof.append(pj)
        uj, uuj, vvj, wwj, uwj = prof
        for arr in (uuj, vvj, wwj):                      # DF:263-268
            arr[arr < 0.0] = 0.0
        with np.errstate(invalid="ignore"):
            R00 = np.sqrt(uuin[None, :] * uuj[:, None])  # DF:272-276
            R11 = np.sqrt(vvin[None, :] * vvj[:, None])
            R22 = np.sqrt(wwin[None, :] * wwj[:, None])
            R20 = np.sign(uwin[None, :] + uwj[:, None]) * np.sqrt(np.abs(uwin[None, :] * uwj[:, None]))
            mean_u = np.sqrt(uin[None, :] * uj[:, None])  # DF:305
    elif mean_profile in ("circular-hyperbolic-tangent", "ring-hyperbolic-tangent"):
        x = np.linspace(-1.0, 1.0, J)
        y = np.linspace(-1.0, 1.0, K)
        r = np.sqrt(x[:, None] ** 2 + y[None, :] ** 2)   # DF:329, 412
        if mean_profile.startswith("circular"):           # DF:309-393 round jet
            ci = int(np.argmax(uin))
            zA = np.linspace(0, 1, len(uin) - ci)
            src = [p[ci:] for p in (uin, uuin, vvin, wwin, uwin)]
            lo_val = [p[ci] for p in (uin, uuin, vvin, wwin, uwin)]
            lo = 0.0
        else:                                             # DF:395-485 annulus
            zA = np.linspace(inner_d, 1.0, K)
            src = [uin, uuin, vvin, wwin, uwin]
            lo_val = [p[0] for p in src]
            lo = inner_d
        vals = []
        for p, first in zip(src, lo_val):
            v = _itp.splev(r.ravel(), spl(zA, p), der=0).reshape(r.shape)
            v = np.where(r == lo, first, v)               # DF:337-342 / 420-425
            v = np.where(r == 1.0, p[-1], v)              # DF:344-349 / 427-432
            v = np.where(r > 1.0, 0.0, v)                 # DF:352-357 / 435-440
            if not mean_profile.startswith("circular"):
                v = np.where(r < inner_d, 0.0, v)         # DF:443-448
            vals.append(v)
        mean_u, R00, R11, R22, R20 = vals
    else:
        raise ValueError(mean_profile)
    z = np.zeros((J, K))
    fac = _chol_rows(R00, z, R11, R20, z, R22, "clamp")
    return fac, (mean_u, z, z.copy())


def adapt2d(yu, yv, yw, uin, uuin, vvin, wwin, uwin, J, K, mean_profile, inner_d):
    """DF:233-485, in place."""
    fac, mean = adapt2d_coefficients(uin, uuin, vvin, wwin, uwin, J, K, mean_profile, inner_d)
    _apply(fac, mean, yu, yv, yw)


# ----- faithful per-point loop forms (timing baseline: where the reference spends its time)
def adapt1d_loops(yu, yv, yw, uin, uuin, vvin, wwin, uwin, J, K):
    """DF:143-178 with the reference's own per-k / per-j Python loops."""
    R = np.zeros((3, 3))
    A = np.zeros((3, 3))
    for k in range(K):
        R[0, 0], R[1, 1], R[2, 2], R[2, 0] = uuin[k], vvin[k], wwin[k], uwin[k]
        A[:, :] = 1.0
        A[0, 0] = np.sqrt(R[0, 0]); A[0, 1] = 0.0; A[0, 2] = 0.0
        A[1, 0] = R[1, 0] / (A[0, 0] + 1e-20)
        A[1, 1] = np.sqrt(R[1, 1] - A[1, 0] * A[1, 0]); A[1, 2] = 0.0
        A[2, 0] = R[2, 0] / (A[0, 0] + 1e-20)
        A[2, 1] = (R[2, 1] - A[1, 0] * A[2, 0]) / (A[1, 1] + 1e-20)
        A[2, 2] = np.sqrt(R[2, 2] - A[2, 0] * A[2, 0] - A[2, 1] * A[2, 1])
        for j in range(J):
            xu, xv, xw = yu[j, k], yv[j, k], yw[j, k]
            yu[j, k] = A[0, 0] * xu + A[0, 1] * xv + A[0, 2] * xw + uin[k]
            yv[j, k] = A[1, 0] * xu + A[1, 1] * xv + A[1, 2] * xw
            yw[j, k] = A[2, 0] * xu + A[2, 1] * xv + A[2, 2] * xw


def adapt2d_round_loops(yu, yv, yw, uin, uuin, vvin, wwin, uwin, J, K):
    """DF:309-393 with the reference's per-point loop and five scalar `splev`
    calls per point per step (the dominant cost of the round-jet path today)."""
    x = np.linspace(-1.0, 1.0, J)
    y = np.linspace(-1.0, 1.0, K)
    ci = int(np.argmax(uin))
    zA = np.linspace(0, 1, len(uin) - ci)
    tcks = [_itp.splrep(zA, p[ci:], s=0) for p in (uin, uuin, vvin, wwin, uwin)]
    for j in range(J):
        for k in range(K):
            r = np.sqrt(x[j] ** 2 + y[k] ** 2)
            v = [_itp.splev(r, t, der=0) for t in tcks]
            if r == 0.0:
                v = [p[ci] for p in (uin, uuin, vvin, wwin, uwin)]
            if r == 1.0:
                v = [p[-1] for p in (uin, uuin, vvin, wwin, uwin)]
            if r > 1.0:
                v = [0.0] * 5
            um, R00, R11, R22, R20 = v
            a00 = np.sqrt(R00)
            a10 = 0.0 / (a00 + 1e-20)
            t = R11 - a10 * a10
            a11 = np.sqrt(0.0 if t < 0 else t)
            a20 = R20 / (a00 + 1e-20)
            a21 = (0.0 - a10 * a20) / (a11 + 1e-20)
            t = R22 - a20 * a20 - a21 * a21
            a22 = np.sqrt(0.0 if t < 0 else t)
            xu, xv, xw = yu[j, k], yv[j, k], yw[j, k]
            yu[j, k] = a00 * xu + 0.0 * xv + 0.0 * xw + um
            yv[j, k] = a10 * xu + a11 * xv + 0.0 * xw
            yw[j, k] = a20 * xu + a21 * xv + a22 * xw


# ----------------------------------------------------------------------- profiles
def build_profile(mean_profile, turb_profile, bulk_velocity, turbulence_intensity, K):
    """DF:1038-1062."""
    if mean_profile in ("hyperbolic-tangent", "double-hyperbolic-tangent",
                        "circular-hyperbolic-tangent", "ring-hyperbolic-tangent"):
        y = np.linspace(-0.5, 0.5, K)
        U = bulk_velocity / 2 * (1.0 + np.tanh(10.0 * (-np.abs(y) + 0.5)))
    else:
        raise Exception("Invalid mean_profile chosen")
    if turb_profile == "top-hat":
        uu = (turbulence_intensity * U) ** 2
        vv = (turbulence_intensity * U) ** 2
        ww = (turbulence_intensity * U) ** 2
        uw = 0.0 * U
    elif turb_profile == "none":
        uu = vv = ww = uw = 0.0
    else:
        raise Exception("Invalid turb_profile chosen")
    return U, uu, vv, ww, uw


def read_profile(profilefile, K):
    """DF:487-522.  1-D profile file with columns y,U,uu,vv,ww,uv: the half profile is mirrored about its
    last point (uv changes sign), y is normalised to [0,1], cubic splines are evaluated at K points and the
    wall values are forced to zero."""
    d = np.genfromtxt(profilefile, names=True, autostrip=True, comments="#")
    n = d.shape[0]
    for row in reversed(d[0:n - 2]):
        d = np.append(d, row)
    d["y"][n:] = (-(d["y"][n:] - 1.0) + 1)
    d["uv"][n:] = -d["uv"][n:]
    z = d["y"]
    z = (z - np.min(z)) / (np.max(z) - np.min(z))
    zi = np.linspace(np.min(z), np.max(z), K)
    out = [_itp.splev(zi, _itp.splrep(z, d[name], s=0), der=0) for name in ("U", "uu", "vv", "ww", "uv")]
    for v in out:
        v[0] = v[-1] = 0.0
    return tuple(out)


def _axis_rotation(theta, nx, ny, nz):
    """Rodrigues rotation about the (unit) axis n by theta, the 3x3 form written out at DF:648-650, 664-666."""
    C, S = np.cos(theta), np.sin(theta)
    t = 1 - C
    return np.array([[t * nx ** 2 + C, t * nx * ny - S * nz, t * nx * nz + S * ny],
                     [t * nx * ny + S * nz, t * ny ** 2 + C, t * ny * nz - S * nx],
                     [t * nx * nz - S * ny, t * ny * nz + S * nx, t * nz ** 2 + C]])


def read_prf(profilefile, res, mdot, den, bulk_velocity, non_dim, TestGrad):
    """DF:524-1035.  2-D `.prf` ingestion: columns named on the `data,` header line (x,y,z,u,v,w and k with e or
    sdr); plane normal from (last-first) x (second-first); the plane is rotated into y-z, resampled on a
    (kma, jma) raster of spacing `res` with linear `griddata`; optional mass-flow / bulk-velocity rescaling
    at constant turbulence intensity and length scale (python-2 `3/2 == 1`, DF:758-788); velocity gradients
    (2x2 box-smoothed) and the eddy-viscosity Reynolds stresses (DF:1000-1010).  Returns the reference's tuple
    U,V,W,uu,vv,ww,uv,uw,vw (each (jma,kma)), lnx, kma, jma, normal(3), centre(3).  `non_dim` only affects the
    reference's plots."""
    skip = 0
    with open(profilefile) as f:
        while True:
            line = f.readline()
            skip += 1
            if line.startswith("data"):
                names = [c.strip() for c in line.strip().split(",")]
                break
    col = {n: i - 1 for i, n in enumerate(names) if i > 0}
    try:
        d = np.loadtxt(profilefile, skiprows=skip)
    except Exception:
        d = np.loadtxt(profilefile, skiprows=skip, delimiter=",")
    xA, yA, zA = d[:, col["x"]], d[:, col["y"]], d[:, col["z"]]
    v2 = (xA[1] - xA[0], yA[1] - yA[0], zA[1] - zA[0])
    v1 = (xA[-1] - xA[0], yA[-1] - yA[0], zA[-1] - zA[0])
    xn = v1[1] * v2[2] - v1[2] * v2[1]
    yn = v1[2] * v2[0] - v1[0] * v2[2]
    zn = v1[0] * v2[1] - v1[1] * v2[0]
    nnorm = np.sqrt(xn ** 2 + yn ** 2 + zn ** 2)
    xn, yn, zn = xn / nnorm, yn / nnorm, zn / nnorm
    xc = (np.amax(xA) + np.amin(xA)) / 2
    yc = (np.amax(yA) + np.amin(yA)) / 2
    zc = (np.amax(zA) + np.amin(zA)) / 2
    pts = np.array([xA - xc, yA - yc, zA - zc])
    pts = np.dot(_axis_rotation(-np.arccos(xn), 0, -zn, yn), pts)            # normal -> (1,0,0), DF:636-658
    pts = np.dot(_axis_rotation(-np.arctan2(zn, yn), xn, yn, zn), pts)        # twist correction, DF:638, 660-672
    yspan = np.amax(pts[1]) - np.amin(pts[1])
    zspan = np.amax(pts[2]) - np.amin(pts[2])
    kma = int(math.ceil(zspan / res))
    jma = int(math.ceil(yspan / res))
    yi = np.linspace(np.min(pts[1]), np.min(pts[1]) + res * jma, jma)
    zi = np.linspace(np.min(pts[2]), np.min(pts[2]) + res * kma, kma)
    y, z = np.meshgrid(yi, zi)
    yz = pts[1:, :].T

    def grid(name, clip):
        g = _itp.griddata(yz, d[:, col[name]], (y, z), fill_value=0.0, method="linear")
        if clip:
            g[g < 0] = 0
        return g

    U, V, W = grid("u", False), grid("v", False), grid("w", False)
    k = grid("k", True)
    if "e" in col:
        eps = grid("e", True)
    if "sdr" in col:
        eps = 0.09 * k * grid("sdr", True)
        eps[np.where(eps > 100000000)] = 0
    if mdot != 0.0 or bulk_velocity != 1.0:
        udotn = np.mean(U) * xn + np.mean(V) * yn + np.mean(W) * zn
        if mdot != 0.0:
            scale = mdot / (udotn * (res ** 2 * (kma - 1) * (jma - 1)) * den)
        else:
            scale = bulk_velocity / udotn
        flag = eps > 0
        TI = np.sqrt(2. / 3. * k[flag]) / np.sqrt(U[flag] ** 2 + V[flag] ** 2 + W[flag] ** 2)
        L = k[flag] ** 1 / eps[flag]
        U, V, W = U * scale, V * scale, W * scale
        k[flag] = TI ** 2 * (U[flag] ** 2 + W[flag] ** 2 + V[flag] ** 2)
        eps[flag] = k[flag] ** 1 / L
    if TestGrad:
        eps = np.ones(np.shape(U))
        k = np.ones(np.shape(U))
        k[0] = eps[0] = 0.0
    flag = np.where(eps == 0.0)
    flag1 = np.where(U == 0.0)
    U[flag] = 0
    V[flag] = 0
    W[flag] = 0
    k[flag] = 0
    eps[flag1] = 0
    if TestGrad:
        U[:] = 1 * y + 2 * z
        V[:] = 3 * y + 4 * z
        W[:] = 5 * y + 6 * z
    grads = []
    for F in (U, V, W):
        dz, dy = np.gradient(F, res)
        dy[flag] = 0
        dz[flag] = 0
        grads += [dy, dz]
    if not TestGrad:
        for g in grads:                                                      # DF:828-842: backward 2x2 box mean
            g1 = g.copy()
            for i in range(1, kma - 1):
                for j in range(1, jma - 1):
                    g[i, j] = np.mean(g1[i - 1:i + 1, j - 1:j + 1])
    dUdy, dUdz, dVdy, dVdz, dWdy, dWdz = grads
    dUdx = -dVdy - dWdz
    dVdx = np.zeros((kma, jma))
    dWdx = np.zeros((kma, jma))
    B = 2 * np.amax(pts[1])
    Cc = 2 * np.amax(pts[2])
    lnx = math.ceil(0.07 * 2 * B * Cc / (B + Cc) / res)
    nu_t = np.zeros((kma, jma))
    pos = np.where(eps > 0)
    nu_t[pos] = 0.09 * k[pos] ** 2 / eps[pos]
    uu = -2. * nu_t * dUdx + 2. / 3. * k
    vv = -2. * nu_t * dVdy + 2. / 3. * k
    ww = -2. * nu_t * dWdz + 2. / 3. * k
    uv = -nu_t * (dUdy + dVdx)
    uw = -nu_t * (dUdz + dWdx)
    vw = -nu_t * (dVdz + dWdy)
    for a in (uu, vv, ww):
        a[np.where(a < 0.0)] = 0.0
    fields = [np.flip(a, 0).T for a in (U, V, W, uu, vv, ww, uv, uw, vw)]
    return tuple(fields) + (lnx, kma, jma, xn, yn, zn, xc, yc, zc)


def channel_profile(K):
    """synthetic realizable channel profile on z in [0,1]
    (not in the reference; shared definition for CPU and GPU sides of the test)."""
    z = np.linspace(0.0, 1.0, K)
    s = 16 * z ** 2 * (1 - z) ** 2
    U = 1 - (2 * z - 1) ** 8
    uu, vv, ww = 0.01 * s, 0.004 * s, 0.003 * s
    uw = -0.002 * s * (1 - 2 * z)
    for a in (U, uu, vv, ww, uw):   # walls forced to zero like DF:516-520
        a[0] = a[-1] = 0.0
    return U, uu, vv, ww, uw


# ----------------------------------------------------------------------- rotation
def prof_rotation_matrix(nx, ny, nz):
    """DF:1064-1116.  R = Ra(azimuth) . Rp(polar), mapping (1,0,0) -> n."""
    n = np.sqrt(nx ** 2 + ny ** 2 + nz ** 2)
    n_proj = np.sqrt(nx ** 2 + ny ** 2)
    if ny > 0:
        az = np.arccos(nx / n_proj)
    elif ny < 0:
        az = -np.arccos(nx / n_proj)
    elif nx >= 0:
        az = 0.0
    else:
        az = np.pi
    c, s = np.cos(az), np.sin(az)
    Ra = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    if nz > 0:
        pol = np.arccos(n_proj / n)
    elif nz < 0:
        pol = -np.arccos(n_proj / n)
    else:
        pol = 0.0
    c, s = np.cos(pol), np.sin(pol)
    Rp = np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]])
    return Ra.dot(Rp)


def rotate_velocity(A, nx, ny, nz, loops=False):
    """DF:1119-1131.  A is the packed column [u; v; w] (3P,)."""
    pts = len(A) // 3              # py2 integer division (DF:1121)
    R = prof_rotation_matrix(nx, ny, nz)
    out = np.zeros(len(A))
    if loops:
        for i in range(pts):
            V = np.array([A[i], A[i + pts], A[i + 2 * pts]])
            Vr = R.dot(V)
            out[i], out[i + pts], out[i + 2 * pts] = Vr[0], Vr[1], Vr[2]
        return out
    u, v, w = A[:pts], A[pts:2 * pts], A[2 * pts:]
    with np.errstate(invalid="ignore"):
        for c in range(3):
            out[c * pts:(c + 1) * pts] = R[c, 0] * u + R[c, 1] * v + R[c, 2] * w
    return out


# ----------------------------------------------------------------- random streams
def reference_stream(seed, ns, J, K, nfx, nfy, nfz):
    """The reference's random numbers in its own draw order (DF:1361-1366, 1460-1467),
    re-assembled as one stream per component: returns (3, ns+2nfx+1, J+2nfy, K+2nfz).
    Step i sees slabs i .. i+2nfx (the last slab is drawn but never used)."""
    rs = np.random.RandomState(seed)    # legacy MT19937 == seeded global np.random
    Jp, Kp = J + 2 * nfy, K + 2 * nfz
    X = np.empty((3, ns + 2 * nfx + 1, Jp, Kp))
    for c in range(3):
        X[c, :2 * nfx + 1] = rs.uniform(low=-SQRT3, high=SQRT3, size=(2 * nfx + 1, Jp, Kp))
    for i in range(ns):
        for c in range(3):
            X[c, 2 * nfx + 1 + i] = rs.uniform(low=-SQRT3, high=SQRT3, size=(Jp, Kp))
    return X


# -------------------------------------------------------------------- the DF loop
def generate(X, ns, J, K, lnx, lny, lnz, nfx, nfy, nfz, lund, rotate=None, loops=False, timers=None):
    """DF:1403-1477: the time-marching loop on an injected stream X (3, >=ns+2nfx, Jp, Kp).

    lund: ('1d', U,uu,vv,ww,uw) | ('2d', U,uu,vv,ww,uw, mean_profile, inner_d) |
          ('prf', U,V,W,uu,vv,ww,uv,uw,vw) | None
    rotate: (nx,ny,nz) or None (the reference rotates only when no profile file is given, DF:1476).
    loops=True uses the reference's per-point Python loops (timing baseline).
    Returns A (3JK, ns) float64, the snapshot matrix of DF:1397, 1471-1473."""
    import time as _t
    P = J * K
    A = np.zeros((3 * P, ns))
    y = [np.zeros((J, K)) for _ in range(3)]
    tm = timers if timers is not None else {}
    for i in range(ns):
        t0 = _t.perf_counter()
        for c in range(3):   # DF:1440-1442
            y[c][:, :] = filter3d_separable(X[c, i:i + 2 * nfx + 1], J, K, lnx, lny, lnz, nfx, nfy, nfz)
        t1 = _t.perf_counter()
        if lund is not None:  # DF:1445-1451
            kind = lund[0]
            if kind == "1d":
                (adapt1d_loops if loops else adapt1d)(y[0], y[1], y[2], *lund[1:], J, K)
            elif kind == "2d":
                if loops and lund[6] == "circular-hyperbolic-tangent":
                    adapt2d_round_loops(y[0], y[1], y[2], *lund[1:6], J, K)
                else:
                    adapt2d(y[0], y[1], y[2], *lund[1:6], J, K, lund[6], lund[7])
            elif kind == "prf":
                adapt2prf(y[0], y[1], y[2], *lund[1:], J, K)
            else:
                raise ValueError(kind)
        t2 = _t.perf_counter()
        for c in range(3):   # DF:1471-1473
            A[c * P:(c + 1) * P, i] = y[c].reshape(P)
        if rotate is not None:  # DF:1476-1477
            A[:, i] = rotate_velocity(A[:, i], *rotate, loops=loops)
        t3 = _t.perf_counter()
        tm["filter"] = tm.get("filter", 0.0) + (t1 - t0)
        tm["lund"] = tm.get("lund", 0.0) + (t2 - t1)
        tm["pack_rotate"] = tm.get("pack_rotate", 0.0) + (t3 - t2)
    return A


def remove_mean(A):
    """DF:1492-1495.  Returns (mean_field, A - mean) (A is modified in place)."""
    mean_field = np.mean(A, 1)
    for j in range(A.shape[1]):
        A[:, j] = A[:, j] - mean_field[:]
    return mean_field, A


def derive_filter_sizes(lengthscale, fwidth, res, U, dt, V=0, W=0, ln_file=None):
    """DF:1277-1285, 1299-1322.  Returns (lnx,lny,lnz,nfx,nfy,nfz,dt).  ln_file: the length scale a `.prf`
    file yields (DF:1301-1304 overwrite lnx=lny=lnz with it; the filter half-widths stay on the flags)."""
    lnx = lny = lnz = lengthscale
    nf = int(math.ceil(fwidth * lengthscale))
    nfx = nfy = nfz = nf
    if ln_file is not None:
        lnx = lny = lnz = ln_file
    U = np.asarray(U, dtype=np.float64)
    flag = np.where(U ** 2 + V ** 2 + W ** 2 != 0)
    if dt == 0.0:
        dt = res / np.mean(U[flag])
    else:
        factor = (res / np.mean(U[flag])) / dt
        lnx = factor * lnx
        nfx = int(math.ceil(float(fwidth) * lnx))
    return lnx, lny, lnz, nfx, nfy, nfz, dt
