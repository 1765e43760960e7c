"""LES inflow generator after Klein et al. (2003) -- B200-native drop-in for the reference's
`digitalfilters.py` (same function names, argument order, in-place output arrays and CLI flags).

Python 3 host code over the C ABI in include/podsdf.h (ctypes, `_capi.py`); the arithmetic runs in
hand-written sm_100a CUDA kernels (csrc/df_kernels.cu).  There is no CPU fallback: without a B200
and a built libpodsdf.so every compute entry raises.

What each entry replaces ("DF:a-b" = reference digitalfilters.py lines):
  calccoeff, coeff3D            DF:73-89, 46-70     host (tiny, defines the taps bit for bit)
  filter3DSciPy1D, filter3D     DF:100-140, 91-98   GPU: time filter + plane filter kernels
  adapt1d / adapt2prf / adapt2d DF:143-485          host set-up once (_lund.py) + GPU transform
  rotate_velocity               DF:1119-1131        GPU (constant 3x3 folded into the transform)
  main                          DF:1134-1510        whole time loop fused on the device
"""
import math
import os
import sys
from optparse import OptionParser

import numpy as np
from scipy import interpolate

try:
    from . import _capi, _lund
    from . import PODFS as pod
    from . import HDF5
except ImportError:          # used flat, like the reference: `import digitalfilters`
    import _capi
    import _lund
    import PODFS as pod
    import HDF5

PROG = "DigitalFilters"
VERSION = "1.1.0-b200"
Pi = np.pi
DEVICE = int(os.environ.get("PODSDF_DEVICE", "0"))


class obj(object):     # the reference's attribute bag (DF:31-32)
    a = 0


# ------------------------------------------------------------------------------------ taps
def calccoeff(a, n, ln):
    """DF:73-89.  Gaussian taps normalised so that sum(b^2) = 1; `a` is scratch of 2n+1, filled."""
    norm = 0.0
    for i in range(n * 2 + 1):
        k = float(i - n)
        a[i] = np.exp(-Pi * k * k / (2.0 * ln * ln))
        norm = norm + a[i] ** 2
    return a / np.sqrt(norm)


def _taps(n, ln):
    return calccoeff(np.zeros(2 * n + 1), n, ln)


def coeff3D(a, nfx, nfy, nfz, lnx, lny, lnz):
    """DF:46-60.  Fills a[0,i,j,k] = bx_i by_j bz_k in place."""
    bx, by, bz = _taps(nfx, lnx), _taps(nfy, lny), _taps(nfz, lnz)
    a[0, :, :, :] = (bx[:, None, None] * by[None, :, None]) * bz[None, None, :]


# ----------------------------------------------------------------------------------- plans
_plans = {}


def _plan(J, K, nfx, nfy, nfz, bx, by, bz):
    key = (J, K, nfx, nfy, nfz, bx.tobytes(), by.tobytes(), bz.tobytes())
    p = _plans.get(key)
    if p is None:
        if len(_plans) > 8:
            _plans.clear()
        p = _plans[key] = _capi.DfPlan(J, K, nfx, nfy, nfz, bx, by, bz, device=DEVICE)
    return p


def filter3DSciPy1D(x, y, a, jma, kma, lnx, lny, lnz, nfx, nfy, nfz):
    """DF:100-140.  x (2nfx+1, jma+2nfy, kma+2nfz) -> y (jma, kma) in place; `a` is ignored, as in
    the reference.  One separable contraction on the GPU instead of three scipy convolutions."""
    p = _plan(jma, kma, nfx, nfy, nfz, _taps(nfx, lnx), _taps(nfy, lny), _taps(nfz, lnz))
    p.filter_block(x, y)


def filter3D(x, y, a, jma, kma, nfx, nfy, nfz):
    """DF:91-98 (legacy dense filter).  `a` from coeff3D is an outer product; its three 1-D factors
    are recovered from the centre lines and the separable GPU filter is used.  A non-separable `a`
    is not supported (the reference never builds one)."""
    a0 = np.asarray(a[0], dtype=np.float64)
    c = a0[nfx, nfy, nfz]
    if c == 0.0:
        raise ValueError("filter3D: kernel centre is zero, cannot factor")
    bx, by, bz = a0[:, nfy, nfz].copy(), a0[nfx, :, nfz] / c, a0[nfx, nfy, :] / c
    if np.abs((bx[:, None, None] * by[None, :, None]) * bz[None, None, :] - a0).max() > 1e-13 * abs(c):
        raise NotImplementedError("filter3D: only separable kernels (as built by coeff3D) run on the GPU path")
    # the dense loop correlates (no flip); the separable kernel convolves: flip the factors
    p = _plan(jma, kma, nfx, nfy, nfz, bx[::-1].copy(), by[::-1].copy(), bz[::-1].copy())
    p.filter_block(x, y)


# ------------------------------------------------------------------------------------ Lund
def _apply(yu, yv, yw, fac_mean, jma, kma, rotation=None):
    one = np.array([1.0])
    p = _plan(jma, kma, 0, 0, 0, one, one, one)
    if fac_mean is None:
        p.set_coefficients(None, 0)
    else:
        p.set_coefficients(*_lund.pack(fac_mean[0], fac_mean[1], jma, kma))
    p.set_rotation(rotation)
    p.lund_apply(yu, yv, yw)
    p.set_rotation(None)


def adapt1d(yu, yv, yw, uin, uuin, vvin, wwin, uwin, jma, kma):
    """DF:143-178, in place on yu, yv, yw (jma, kma)."""
    _apply(yu, yv, yw, _lund.lund_1d(uin, uuin, vvin, wwin, uwin, jma, kma), jma, kma)


def adapt2prf(yu, yv, yw, uin, vin, win, uuin, vvin, wwin, uvin, uwin, vwin, jma, kma):
    """DF:180-231, in place."""
    _apply(yu, yv, yw, _lund.lund_prf(uin, vin, win, uuin, vvin, wwin, uvin, uwin, vwin), jma, kma)


def adapt2d(yu, yv, yw, uin, uuin, vvin, wwin, uwin, jma, kma, mean_profile, inner_d):
    """DF:233-485, in place."""
    _apply(yu, yv, yw, _lund.lund_2d(uin, uuin, vvin, wwin, uwin, jma, kma, mean_profile, inner_d), jma, kma)


# -------------------------------------------------------------------------------- profiles
def read_profile(profilefile, kma):
    """DF:487-522.  1-D profile file (columns y,U,uu,vv,ww,uv), mirrored and splined onto kma points."""
    d = np.genfromtxt(profilefile, names=True, autostrip=True, comments="#")
    npoints = d.shape[0]
    for row in reversed(d[0:npoints - 2]):
        d = np.append(d, row)
    d["y"][npoints:] = (-(d["y"][npoints:] - 1.0) + 1)
    d["uv"][npoints:] = -d["uv"][npoints:]
    z = d["y"]
    z = (z - np.min(z)) / (np.max(z) - np.min(z))
    zi = np.linspace(np.min(z), np.max(z), kma)
    out = [interpolate.splev(zi, interpolate.splrep(z, d[name], s=0), der=0) for name in ("U", "uu", "vv", "ww", "uv")]
    for v in out:
        v[0] = v[-1] = 0.0
    return tuple(out)


def _rodrigues(theta, ax, ay, az):
    c, s_ = np.cos(theta), np.sin(theta)
    t = 1 - c
    return np.array([[t * ax ** 2 + c, t * ax * ay - s_ * az, t * ax * az + s_ * ay],
                     [t * ax * ay + s_ * az, t * ay ** 2 + c, t * ay * az - s_ * ax],
                     [t * ax * az - s_ * ay, t * ay * az + s_ * ax, t * az ** 2 + c]])


def _box4(g):
    """Backward 2x2 box mean on the interior, the smoothing of DF:828-842."""
    out = g.copy()
    out[1:-1, 1:-1] = (g[:-2, :-2] + g[:-2, 1:-1] + g[1:-1, :-2] + g[1:-1, 1:-1]) / 4.0
    return out


def read_prf(profilefile, res, mdot, den, bulk_velocity, non_dim, TestGrad):
    """DF:524-1035: 2-D `.prf` profile -> mean velocity and Reynolds-stress planes for adapt2prf.

    Host-side set-up, run once per job.  The columns are named on the `data,...` header line (x,y,z,u,v,w,k and
    e or sdr).  The face list is turned into the y-z plane (normal from the first, second and last point, as
    the reference does), resampled on a (kma, jma) raster of spacing `res` with linear griddata, optionally
    rescaled to a mass flow or bulk velocity at constant turbulence intensity and length scale, and closed
    with an eddy-viscosity model on box-smoothed gradients.  Returns U,V,W,uu,vv,ww,uv,uw,vw (jma,kma), lnx,
    kma, jma, xn,yn,zn, xc,yc,zc.  `non_dim` only rescaled the reference's plot axes."""
    from scipy import interpolate
    with open(profilefile) as f:
        skip = 0
        for line in f:
            skip += 1
            if line.startswith("data"):
                break
        else:
            raise Exception("no 'data' header line in " + profilefile)
    names = [c.strip() for c in line.strip().split(",")][1:]
    col = {n: i for i, n in enumerate(names)}
    try:
        table = np.loadtxt(profilefile, skiprows=skip)
    except Exception:
        table = np.loadtxt(profilefile, skiprows=skip, delimiter=",")
    xyz = np.array([table[:, col[c]] for c in "xyz"])
    e1, e2 = xyz[:, -1] - xyz[:, 0], xyz[:, 1] - xyz[:, 0]
    nrm = np.array([e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]])
    xn, yn, zn = nrm / np.sqrt(nrm[0] ** 2 + nrm[1] ** 2 + nrm[2] ** 2)
    xc, yc, zc = ((np.amax(a) + np.amin(a)) / 2 for a in xyz)
    pts = np.array([xyz[0] - xc, xyz[1] - yc, xyz[2] - zc])
    pts = np.dot(_rodrigues(-np.arccos(xn), 0, -zn, yn), pts)
    pts = np.dot(_rodrigues(-np.arctan2(zn, yn), xn, yn, zn), pts)
    ymin, zmin = np.min(pts[1]), np.min(pts[2])
    jma = int(math.ceil((np.amax(pts[1]) - ymin) / res))
    kma = int(math.ceil((np.amax(pts[2]) - zmin) / res))
    y, z = np.meshgrid(np.linspace(ymin, ymin + res * jma, jma), np.linspace(zmin, zmin + res * kma, kma))

    def raster(name, nonneg=True):
        g = interpolate.griddata(pts[1:].T, table[:, col[name]], (y, z), fill_value=0.0, method="linear")
        return np.maximum(g, 0.0) if nonneg else g

    U, V, W = (raster(c, False) for c in "uvw")
    k = raster("k")
    if "e" in col:
        eps = raster("e")
    if "sdr" in col:
        eps = 0.09 * k * raster("sdr")
        eps[eps > 100000000] = 0
    if mdot != 0.0 or bulk_velocity != 1.0:
        udotn = np.mean(U) * xn + np.mean(V) * yn + np.mean(W) * zn
        area = res ** 2 * (kma - 1) * (jma - 1)
        scale = mdot / (udotn * area * den) if mdot != 0.0 else bulk_velocity / udotn
        live = eps > 0
        TI = np.sqrt(2. / 3. * k[live]) / np.sqrt(U[live] ** 2 + V[live] ** 2 + W[live] ** 2)
        L = k[live] / eps[live]              # the reference's k**(3/2) under python 2 is k**1
        U, V, W = U * scale, V * scale, W * scale
        k[live] = TI ** 2 * (U[live] ** 2 + W[live] ** 2 + V[live] ** 2)
        eps[live] = k[live] / L
    if TestGrad:
        eps = np.ones_like(U)
        k = np.ones_like(U)
        k[0] = eps[0] = 0.0
    dead, still = eps == 0.0, U == 0.0
    for a in (U, V, W, k):
        a[dead] = 0
    eps[still] = 0
    if TestGrad:
        U, V, W = 1 * y + 2 * z, 3 * y + 4 * z, 5 * y + 6 * z
    grads = {}
    for c, F in zip("UVW", (U, V, W)):
        for axis, d in zip("zy", np.gradient(F, res)):
            d[dead] = 0
            grads[c + axis] = d if TestGrad else _box4(d)
    B, C = 2 * np.amax(pts[1]), 2 * np.amax(pts[2])
    lnx = math.ceil(0.07 * 2 * B * C / (B + C) / res)
    nu_t = np.zeros((kma, jma))
    live = eps > 0
    nu_t[live] = 0.09 * k[live] ** 2 / eps[live]
    dUdx = -grads["Vy"] - grads["Wz"]
    uu = np.maximum(-2. * nu_t * dUdx + 2. / 3. * k, 0.0)
    vv = np.maximum(-2. * nu_t * grads["Vy"] + 2. / 3. * k, 0.0)
    ww = np.maximum(-2. * nu_t * grads["Wz"] + 2. / 3. * k, 0.0)
    uv = -nu_t * (grads["Uy"] + 0.0)
    uw = -nu_t * (grads["Uz"] + 0.0)
    vw = -nu_t * (grads["Vz"] + grads["Wy"])
    planes = tuple(np.ascontiguousarray(a[::-1].T) for a in (U, V, W, uu, vv, ww, uv, uw, vw))
    return planes + (lnx, kma, jma, xn, yn, zn, xc, yc, zc)


def build_profile(mean_profile, turb_profile, bulk_velocity, turbulence_intensity, kma):
    """DF:1038-1062."""
    if mean_profile in ["hyperbolic-tangent", "double-hyperbolic-tangent",
                        "circular-hyperbolic-tangent", "ring-hyperbolic-tangent"]:
        y = np.linspace(-0.5, 0.5, kma)
        U = bulk_velocity / 2 * (1. + np.tanh(10. * (-np.abs(y) + 0.5)))
    else:
        raise Exception("Invalid mean_profile chosen, type 'python digitalfilters.py -h' for help.")
    if turb_profile == "top-hat":
        uu = (turbulence_intensity * U) ** 2
        vv = (turbulence_intensity * U) ** 2
        ww = (turbulence_intensity * U) ** 2
        uw = 0.0 * U
    elif turb_profile == "none":
        uu = vv = ww = uw = 0.0
    else:
        raise Exception("Invalid turb_profile chosen, type 'python digitalfilters.py -h' for help.")
    return U, uu, vv, ww, uw


# -------------------------------------------------------------------------------- rotation
def prof_rotation_matrix(nx, ny, nz):
    """DF:1064-1116.  R = Ra(azimuth) . Rp(polar) with R.(1,0,0) = n."""
    n = np.sqrt(nx ** 2 + ny ** 2 + nz ** 2)
    n_proj = np.sqrt(nx ** 2 + ny ** 2)
    if ny > 0:
        azimuth = np.arccos(nx / n_proj)
    elif ny < 0:
        azimuth = -np.arccos(nx / n_proj)
    elif nx >= 0:
        azimuth = 0.
    else:
        azimuth = np.pi
    c, s = np.cos(azimuth), np.sin(azimuth)
    Ra = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    if nz > 0:
        polar = np.arccos(n_proj / n)
    elif nz < 0:
        polar = -np.arccos(n_proj / n)
    else:
        polar = 0.
    c, s = np.cos(polar), np.sin(polar)
    Rp = np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]])
    return Ra.dot(Rp)


def rotate_velocity(A, nx, ny, nz):
    """DF:1119-1131.  A = packed column [u; v; w]; returns the rotated copy."""
    pts = len(A) // 3
    R = prof_rotation_matrix(nx, ny, nz)
    u, v, w = (np.array(A[c * pts:(c + 1) * pts], dtype=np.float64).reshape(1, pts) for c in range(3))
    _apply(u, v, w, None, 1, pts, rotation=R)
    return np.concatenate([u.ravel(), v.ravel(), w.ravel()])


# ------------------------------------------------------------------------------ generation
def reference_stream(nsteps, jma, kma, nfx, nfy, nfz, pdfr=np.sqrt(3.0)):
    """Draws the random field from the global np.random in the reference's own order (DF:1361-1366,
    1460-1467) and returns it as slabs (nsteps+2nfx+1, 3, Jp, Kp): step i uses slabs i..i+2nfx."""
    Jp, Kp = jma + 2 * nfy, kma + 2 * nfz
    X = np.empty((nsteps + 2 * nfx + 1, 3, Jp, Kp))
    for c in range(3):
        X[:2 * nfx + 1, c] = np.random.uniform(low=-pdfr, high=pdfr, size=(2 * nfx + 1, Jp, Kp))
    for i in range(nsteps):
        for c in range(3):
            X[2 * nfx + 1 + i, c] = np.random.uniform(low=-pdfr, high=pdfr, size=(Jp, Kp))
    return X


def generate(nsteps, jma, kma, lnx, lny, lnz, nfx, nfy, nfz, fac_mean=None, rotation=None,
             slabs=None, seed=None, t0=0, out=None, impl=None, chunk=None):
    """The whole time loop DF:1403-1477 on the device.  Returns planes (nsteps, 3, jma, kma), i.e.
    the snapshot matrix transposed (A = planes.reshape(nsteps, -1).T).

    slabs: injected random stream (n, 3, Jp, Kp) with n >= nsteps + 2nfx, or None to generate the
    stream on the device from `seed` (Philox), planes t0 .. t0+nsteps-1."""
    p = _plan(jma, kma, nfx, nfy, nfz, _taps(nfx, lnx), _taps(nfy, lny), _taps(nfz, lnz))
    if impl is not None:
        p.set_impl(impl)
    if chunk is not None:
        p.set_chunk(chunk)
    if fac_mean is None:
        p.set_coefficients(None, 0)
    else:
        p.set_coefficients(*_lund.pack(fac_mean[0], fac_mean[1], jma, kma))
    # the identity is skipped (it is exact, and keeps the three components independent on the device)
    p.set_rotation(None if rotation is None or np.array_equal(rotation, np.eye(3)) else rotation)
    if slabs is not None:
        return p.generate_injected(slabs[:nsteps + 2 * nfx], out)
    return p.generate_seeded(seed, t0, nsteps, out)


def _device_pipeline_available():
    """The device-resident DF -> PODFS path needs PyTorch for its device buffers (plumbing only)."""
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _main_device_pipeline(i_d, options, nsteps, jma, kma, lnx, lny, lnz, nfx, nfy, nfz, fac_mean, rotation, slabs, dt):
    """DF:1403-1500 on the device (pipeline.df_to_podfs); fills the same i_d fields POD() publishes, except that
    i_d.temporal_modes holds the nm kept modes (ns, nm) instead of all ns of them."""
    import torch
    try:
        from . import pipeline
    except ImportError:
        import pipeline
    p = _plan(jma, kma, nfx, nfy, nfz, _taps(nfx, lnx), _taps(nfy, lny), _taps(nfz, lnz))
    p.set_impl(_capi.IMPL_AUTO)
    if fac_mean is None:
        p.set_coefficients(None, 0)
    else:
        p.set_coefficients(*_lund.pack(fac_mean[0], fac_mean[1], jma, kma))
    p.set_rotation(None if rotation is None or np.array_equal(rotation, np.eye(3)) else rotation)
    dev = torch.device("cuda", DEVICE)
    slabs_dev = None
    if slabs is not None:
        slabs_dev = torch.from_numpy(np.ascontiguousarray(slabs[:nsteps + 2 * nfx])).to(dev)
    print("\n==============================================================")
    print("Calculating POD modes on the device ...")
    res = pipeline.df_to_podfs(p, nsteps, options.nm, dt, options.et, seed=options.seed, slabs=slabs_dev,
                               device_index=DEVICE)
    energy, nm = res["energy"], res["nm"]
    nv = pod._count_valid(energy, nsteps, 1.0e-15)
    print("      Number of valid POD modes with positive energies = ", nv)
    i_d.mean_field = res["mean"].cpu().numpy()
    i_d.temporal_modes = res["T"]
    i_d.spatial_modes = res["Phi"].cpu().numpy()
    i_d.nm = nm
    i_d.energy = energy
    i_d.c128 = res["c128"]
    pod.write_eigenvalues(nv, nsteps, energy, "PODFS/" + "POD.eigenvalues.dat")
    print("\nCode completed.")
    print("==============================================================\n")


# ------------------------------------------------------------------------------------- CLI
# (flags, dest, type, default, help) -- the reference's option table DF:1143-1237, names verbatim
_OPTIONS = [
    (("-i", "--inputfile"), "profilefile", "string", "none", "1d turbulent profile file"),
    (("-p", "--mean_profile"), "mean_profile", "string", "hyperbolic-tangent",
     "mean flow profile: hyperbolic-tangent, double-hyperbolic-tangent, ring-hyperbolic-tangent, "
     "circular-hyperbolic-tangent"),
    (("--turb_profile",), "turb_profile", "string", "top-hat", "turbulence profile: top-hat, none"),
    (("--U0", "--bulk_velocity"), "bulk_velocity", "float", 1.0, "bulk velocity magnitude"),
    (("--u_dash",), "turbulence_intensity", "float", 0.02, "u'/U0 with u'=v'=w'"),
    (("-n", "--nsteps"), "nsteps", "int", 20, "number of steps"),
    (("-l", "--lengthscale"), "lengthscale", "float", 3.0, "turbulent lengthscale in grid spacings"),
    (("-f", "--fwidth"), "fwidth", "float", 2.0, "half filter width in lengthscales (>= 2)"),
    (("-k", "--nk"), "kma", "int", 11, "points in k (wall-normal) direction"),
    (("-j", "--nj"), "jma", "int", 10, "points in j (spanwise) direction"),
    (("-t", "--dt"), "dt", "float", 0.0, "time step (s)"),
    (("-m", "--nm"), "nm", "int", 20, "number of POD modes"),
    (("-e", "--et"), "et", "float", 0.9, "target energy for Fourier reconstruction"),
    (("-v", "--verbose"), "verbose", None, False, "save the mean flow, POD spatial and temporal modes"),
    (("--non_dim",), "non_dim", None, False, "non-dimensionalise lengths if using .prf"),
    (("-r", "--resolution"), "res", "float", 0.1, "plane resolution in meters per grid point"),
    (("--nx",), "nx", "float", 1.0, "plane normal, x"),
    (("--ny",), "ny", "float", 0.0, "plane normal, y"),
    (("--nz",), "nz", "float", 0.0, "plane normal, z"),
    (("--ox",), "ox", "float", 0.0, "plane origin, x"),
    (("--oy",), "oy", "float", 0.0, "plane origin, y"),
    (("--oz",), "oz", "float", 0.0, "plane origin, z"),
    (("--rotate",), "rot", "float", 0.0, "rotate plane about its normal (degrees)"),
    (("--ring",), "ring", "float", 0.5, "inner/outer diameter ratio for ring-hyperbolic-tangent"),
    (("--massflow",), "mdot", "float", 0.0, "target mass flow when using a .prf file"),
    (("--density",), "den", "float", 0.0, "density, needed with --massflow"),
    (("-5", "--hdf5"), "hdf5", None, False, "also save a single PODFS.hdf5"),
    (("--test_gradients",), "TestGrad", None, False, "set U,V,W = 1y+2z, 3y+4z, 5y+6z"),
    # additions of this implementation (the reference draws from the unseeded global np.random)
    (("--rng",), "rng", "string", "numpy",
     "numpy: draw the random field from np.random in the reference's order and inject it (bit-compatible "
     "with a seeded reference run); philox: generate it on the GPU (fast path)"),
    (("--seed",), "seed", "int", 0, "seed of the philox stream"),
    (("--pipeline",), "pipeline", "string", "device",
     "device: planes stay in GPU memory from generation to the POD modes (default when not verbose); "
     "host: every step goes through host arrays like the reference"),
]


def _parser():
    parser = OptionParser(usage="usage: %prog [options]", version="%s %s" % (PROG, VERSION),
                          description=" LES Inflow Generator after Klein et.al. ")
    for flags, dest, typ, default, hlp in _OPTIONS:
        if typ is None:
            parser.add_option(*flags, dest=dest, default=default, action="store_true", help=hlp)
        else:
            parser.add_option(*flags, dest=dest, default=default, type=typ, help=hlp)
    return parser


def main(argv=None):
    """DF:1134-1510.  Same flags, same `i_d` attribute bag, same artefacts under ./PODFS/."""
    parser = _parser()
    argv = sys.argv[1:] if argv is None else list(argv)
    if len(argv) == 0:
        parser.parse_args(["--help"])
    options, _ = parser.parse_args(argv)
    profilefile = options.profilefile
    if not os.path.exists("PODFS"):
        os.mkdir("PODFS")
    mean_profile = options.mean_profile
    nsteps, kma, jma = options.nsteps, options.kma, options.jma
    res, dt, inner_d = options.res, options.dt, options.ring
    lnx = lny = lnz = options.lengthscale
    nf = int(math.ceil(options.fwidth * options.lengthscale))          # DF:1282
    nfx = nfy = nfz = nf
    nrm = np.sqrt(options.nx ** 2 + options.ny ** 2 + options.nz ** 2)
    nx, ny, nz = options.nx / nrm, options.ny / nrm, options.nz / nrm   # DF:1291-1293

    ox, oy, oz = options.ox, options.oy, options.oz
    V = W = 0
    from_prf = False
    if options.profilefile and os.path.isfile(profilefile):
        if profilefile.endswith(".prf"):                                 # DF:1299-1304: precise 2-D profile file
            from_prf = True
            (U, V, W, uu, vv, ww, uv, uw, vw, lnx, kma, jma, nx, ny, nz, ox, oy, oz) = read_prf(
                profilefile, res, options.mdot, options.den, options.bulk_velocity, options.non_dim,
                options.TestGrad)
            lny = lnz = lnx
        else:
            U, uu, vv, ww, uw = read_profile(profilefile, kma)
    else:
        U, uu, vv, ww, uw = build_profile(mean_profile, options.turb_profile, options.bulk_velocity,
                                          options.turbulence_intensity, kma)
    flag = np.where(U ** 2 + V ** 2 + W ** 2 != 0)                       # DF:1311-1322
    if dt == 0.:
        dt = res / np.mean(U[flag])
        print("timestep set to: ", dt, " seconds")
    else:
        lnx = (res / np.mean(U[flag])) / dt * lnx
        nfx = int(math.ceil(float(options.fwidth) * lnx))
        print("Lengthscale in x-direction set to: ", lnx, "grid points")
        print("Filter width in x-direction set to: ", nfx, "grid points")
    if max(nfx, nfy, nfz) > 64:
        # the device kernels carry at most 129 taps per axis (include/podsdf.h); fail before any work starts
        raise Exception("half filter width %d exceeds the supported maximum of 64 grid points "
                        "(reduce -l/-f, or choose --dt closer to res/mean(U))" % max(nfx, nfy, nfz))
    if not from_prf:                                                     # DF:1345-1352
        for v in (uu, vv, ww):
            v[v < 0.0] = 0.0

    i_d = obj()
    i_d.kma, i_d.jma, i_d.ns, i_d.dt, i_d.nm, i_d.et = kma, jma, nsteps, dt, options.nm, options.et
    i_d.rot, i_d.t_o, i_d.res, i_d.n = options.rot, [ox, oy, oz], res, [nx, ny, nz]
    num_points = jma * kma
    i_d.num_points = num_points
    i_d.is_POD_var_vec = False
    i_d.grid = pod.make_inflow_plane(i_d)
    i_d.hdf5 = options.hdf5

    if from_prf:
        fac_mean = _lund.lund_prf(U, V, W, uu, vv, ww, uv, uw, vw)                     # DF:1445-1446
    elif mean_profile in ["double-hyperbolic-tangent", "ring-hyperbolic-tangent", "circular-hyperbolic-tangent"]:
        fac_mean = _lund.lund_2d(U, uu, vv, ww, uw, jma, kma, mean_profile, inner_d)   # DF:1447-1449
    else:
        fac_mean = _lund.lund_1d(U, uu, vv, ww, uw, jma, kma)                          # DF:1450-1451
    rotation = prof_rotation_matrix(nx, ny, nz) if profilefile == "none" else None    # DF:1476-1477

    print("Generating", nsteps, "inlets on the GPU")
    if options.rng not in ("numpy", "philox"):
        raise Exception("--rng must be numpy or philox")
    slabs = reference_stream(nsteps, jma, kma, nfx, nfy, nfz) if options.rng == "numpy" else None
    i_d.verbose = options.verbose
    if not options.verbose and options.pipeline != "host" and _device_pipeline_available():
        # DF:1403-1500 without a host round trip: the planes stay in HBM in the generator's layout [t][c][p] (the
        # snapshot matrix transposed), mean removal, Gram matrix, leading eigenpairs and projection run on them in
        # place, and only the mean field, the nm modes and the energies come back (pipeline.df_to_podfs).  Verbose
        # runs write every plane and every temporal mode, so they take the host path below.
        _main_device_pipeline(i_d, options, nsteps, jma, kma, lnx, lny, lnz, nfx, nfy, nfz, fac_mean, rotation, slabs, dt)
    else:
        if slabs is not None:
            planes = generate(nsteps, jma, kma, lnx, lny, lnz, nfx, nfy, nfz, fac_mean, rotation, slabs=slabs)
        else:
            planes = generate(nsteps, jma, kma, lnx, lny, lnz, nfx, nfy, nfz, fac_mean, rotation, seed=options.seed)
        A = planes.reshape(nsteps, 3 * num_points).T       # (3P, ns) view of the device-native layout

        if options.verbose:                                  # DF:1479-1481
            for i in range(nsteps):
                i_d.time = i * dt
                pod.save_plane(A[:, i], i_d)
        nmw = options.nm if options.verbose else 0

        # DF:1492-1495.  The reference reduces a C-contiguous (3P, ns) array, i.e. numpy's pairwise sum along the
        # contiguous time axis; `planes` holds the transpose, so the same reduction order needs a contiguous copy
        # of each row block (bit parity of mean_field with the reference).
        mean_field = np.empty(3 * num_points)
        blk = max(1, (1 << 24) // max(nsteps, 1))
        for r0 in range(0, 3 * num_points, blk):
            mean_field[r0:r0 + blk] = np.mean(np.ascontiguousarray(A[r0:r0 + blk]), 1)
        i_d.mean_field = mean_field
        planes -= mean_field[None, :].reshape(1, 3, jma, kma)

        pod.POD(A, nsteps, num_points, 3, "false", [], "PODFS/", "false", 1.0e-15, options.nm, nmw, "false", "false",
                i_d.grid, mean_field, dt, "velocity", 1, nsteps, 1, 1, i_d)
    pod.fourier_coefficients(i_d)
    pod.pod2prf(i_d)
    if options.hdf5:
        HDF5.write_HDF5(i_d)
    return i_d


if __name__ == "__main__":
    main()
