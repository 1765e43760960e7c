"""Host-side set-up of the time-invariant Lund / rotation transform  y = M(j,k) x + m(j,k).

The reference re-derives the Cholesky factor of the Reynolds-stress tensor (and, for the shaped
inlets, five spline evaluations) for every point on EVERY time step inside Python loops
(digitalfilters.py:143-485, "DF" below).  None of it depends on time, so the product evaluates
it once here, vectorised, with the same numpy / FITPACK calls, and ships 12 coefficient planes
to the GPU where the transform is fused into the plane-filter kernel's epilogue.

Returned by every builder:  coef (12, J*K) float64 = M00,M01,M02,M10,..,M22,m0,m1,m2  and the
bit mask of planes that are not identically zero.
"""
import numpy as np
from scipy import interpolate

EPS = 1e-20   # the reference's denominator guard (DF:162-168)


def _factor(R00, R10, R11, R20, R21, R22, clamp):
    """Lower-triangular factor, per point.  clamp: 'none' (DF:157-169), 'branch' (DF:199-223),
    'radicand' (DF:278-299 and its two copies)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        a00 = np.sqrt(R00)   # adapt2d clamps a temp but takes sqrt of the raw value (DF:279-282)
        if clamp == "branch":
            ok = a00 > 0.0
            a10 = np.where(ok, R10 / (a00 + EPS), 0.0)
            a11 = np.where(a10 ** 2 > R11, 0.0, np.sqrt(R11 - a10 * a10))
            a20 = np.where(ok, R20 / (a00 + EPS), 0.0)
            a21 = np.where(a11 > 0.0, (R21 - a10 * a20) / (a11 + EPS), 0.0)
            a22 = np.where(R22 < a20 * a20 + a21 * a21, 0.0, np.sqrt(R22 - a20 * a20 - a21 * a21))
        else:
            a10 = R10 / (a00 + EPS)
            r = R11 - a10 * a10
            a11 = np.sqrt(np.where(r < 0, 0.0, r) if clamp == "radicand" else r)
            a20 = R20 / (a00 + EPS)
            a21 = (R21 - a10 * a20) / (a11 + EPS)
            r = R22 - a20 * a20 - a21 * a21
            a22 = np.sqrt(np.where(r < 0, 0.0, r) if clamp == "radicand" else r)
    return a00, a10, a11, a20, a21, a22


def pack(fac, mean, J, K, rotation=None):
    """(a00..a22, (U,V,W)) -> coef (12, P), mask.  rotation: optional constant 3x3 applied after the
    Lund transform (DF:1119-1131); folded in as M <- R.M, m <- R.m (same 0*NaN = NaN propagation
    as the reference's R.dot(V))."""
    P = J * K
    a00, a10, a11, a20, a21, a22 = (np.broadcast_to(f, (J, K)).reshape(P) for f in fac)
    z = np.zeros(P)
    M = np.array([[a00, z, z], [a10, a11, z], [a20, a21, a22]])          # (3,3,P)
    m = np.array([np.broadcast_to(v, (J, K)).reshape(P) for v in mean])  # (3,P)
    if rotation is not None:
        with np.errstate(invalid="ignore"):
            M = np.einsum("ab,bcp->acp", rotation, M)
            m = rotation.dot(m)
    coef = np.ascontiguousarray(np.concatenate([M.reshape(9, P), m], axis=0))
    mask = 0
    for i in range(12):
        if np.any(coef[i] != 0.0):     # NaN != 0 is True, so NaN planes stay present
            mask |= 1 << i
    return coef, mask


def _kprofile(v, K):
    return np.broadcast_to(np.asarray(v, dtype=np.float64), (K,))


def lund_1d(uin, uuin, vvin, wwin, uwin, J, K):
    """adapt1d (DF:143-178): 1-D wall-normal profiles, only R00,R11,R22,R20 set."""
    z = np.zeros(K)
    fac = _factor(_kprofile(uuin, K), z, _kprofile(vvin, K), _kprofile(uwin, K), z, _kprofile(wwin, K), "none")
    fac = tuple(f[None, :] for f in fac)
    return fac, (_kprofile(uin, K)[None, :], 0.0, 0.0)


def lund_prf(uin, vin, win, uuin, vvin, wwin, uvin, uwin, vwin):
    """adapt2prf (DF:180-231): full symmetric R(j,k) and means."""
    fac = _factor(uuin, uvin, vvin, uwin, vwin, wwin, "branch")
    return fac, (uin, vin, win)


def lund_2d(uin, uuin, vvin, wwin, uwin, J, K, mean_profile, inner_d):
    """adapt2d (DF:233-485): square / round / annular jets built from the 1-D profile."""
    prof = [np.asarray(v, dtype=np.float64) for v in (uin, uuin, vvin, wwin, uwin)]
    if mean_profile == "double-hyperbolic-tangent":
        zk, zj = np.linspace(-1.0, 1, K), np.linspace(-1.0, 1, J)
        pj = []
        for p in prof:
            q = interpolate.splev(zj, interpolate.splrep(zk, p, s=0), der=0)
            q[0], q[-1] = p[0], p[-1]
            pj.append(q)
        for q in pj[1:4]:
            q[q < 0.0] = 0.0
        with np.errstate(invalid="ignore"):
            geo = [np.sqrt(p[None, :] * q[:, None]) for p, q in zip(prof[:4], pj[:4])]
            R20 = np.sign(prof[4][None, :] + pj[4][:, None]) * np.sqrt(np.abs(prof[4][None, :] * pj[4][:, None]))
        mean_u, R00, R11, R22 = geo
    elif mean_profile in ("circular-hyperbolic-tangent", "ring-hyperbolic-tangent"):
        x, y = np.linspace(-1.0, 1.0, J), np.linspace(-1.0, 1.0, K)
        r = np.sqrt(x[:, None] ** 2 + y[None, :] ** 2)
        if mean_profile[0] == "c":
            ci = int(np.argmax(prof[0]))
            knots, src, lo = np.linspace(0, 1, len(prof[0]) - ci), [p[ci:] for p in prof], 0.0
        else:
            knots, src, lo = np.linspace(inner_d, 1.0, K), prof, inner_d
        vals = []
        for p in src:
            v = interpolate.splev(r.ravel(), interpolate.splrep(knots, p, s=0), der=0).reshape(r.shape)
            v = np.where(r == lo, p[0], v)
            v = np.where(r == 1.0, p[-1], v)
            v = np.where(r > 1.0, 0.0, v)
            if mean_profile[0] == "r":
                v = np.where(r < inner_d, 0.0, v)
            vals.append(v)
        mean_u, R00, R11, R22, R20 = vals
    else:
        raise Exception("Invalid mean_profile chosen, type 'python digitalfilters.py -h' for help.")
    z = np.zeros((J, K))
    return _factor(R00, z, R11, R20, z, R22, "radicand"), (mean_u, 0.0, 0.0)
