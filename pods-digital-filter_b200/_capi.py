"""ctypes binding of libpodsdf.so (the C ABI declared in include/podsdf.h).

This is the only place the Python shims touch native code.  There is no CPU fallback: if the
library is missing, or no sm_100 GPU is visible, every compute call raises `PodsdfError`.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PODSDF_LIB") or os.path.join(HERE, "libpodsdf.so")   # PODSDF_LIB: A/B builds (tools/)

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_UNSUPPORTED, ERR_NUMERIC = range(6)
LAYOUT_ROWS_BY_TIME, LAYOUT_TIME_BY_ROWS = 0, 1
IMPL_AUTO, IMPL_GENERIC, IMPL_TUNED = 0, 1, 2
RNG_U53, RNG_U32 = 0, 1


class PodsdfError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "libpodsdf error %d: %s" % (code, msg))
        self.code = code


_c_double_p = ctypes.POINTER(ctypes.c_double)
_vp = ctypes.c_void_p
_i, _l, _d, _u64 = ctypes.c_int, ctypes.c_long, ctypes.c_double, ctypes.c_uint64

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/podsdf.h one to one
SIGNATURES = {
    "podsdf_last_error": [],
    "podsdf_version": [],
    "podsdf_device_count": [ctypes.POINTER(_i)],
    "podsdf_launch_count": [],
    "podsdf_host_alloc": [ctypes.POINTER(_vp), ctypes.c_size_t],
    "podsdf_host_free": [_vp],
    "podsdf_df_plan_create": [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, ctypes.POINTER(_vp)],
    "podsdf_df_plan_destroy": [_vp],
    "podsdf_df_plan_set_impl": [_vp, _i],
    "podsdf_df_plan_set_chunk": [_vp, _i],
    "podsdf_df_plan_set_rng": [_vp, _i],
    "podsdf_df_plan_info": [_vp, ctypes.POINTER(_i)],
    "podsdf_df_set_coefficients": [_vp, _vp, ctypes.c_uint],
    "podsdf_df_set_rotation": [_vp, _vp],
    "podsdf_df_filter_block_host": [_vp, _vp, _vp],
    "podsdf_df_lund_apply_host": [_vp, _vp, _vp, _vp],
    "podsdf_df_generate_injected_host": [_vp, _vp, _l, _vp],
    "podsdf_df_generate_injected_dev": [_vp, _vp, _l, _vp, _vp],
    "podsdf_df_generate_seeded_host": [_vp, _u64, _l, _l, _vp],
    "podsdf_df_generate_seeded_dev": [_vp, _u64, _l, _l, _vp, _vp],
    "podsdf_df_random_slabs_host": [_vp, _u64, _l, _l, _vp],
    "podsdf_df_random_slabs_dev": [_vp, _u64, _l, _l, _vp, _vp],
    "podsdf_df_last_kernel_ms": [_vp, _vp],
    "podsdf_mean_rows_dev": [_i, _vp, _l, _l, _i, _vp, _i, _vp],
    "podsdf_gram_host": [_i, _vp, _l, _l, _i, _vp, _l, _d, _vp],
    "podsdf_gram_dev": [_i, _vp, _l, _l, _i, _vp, _l, _d, _vp, _vp],
    "podsdf_release_scratch": [_i],
    "podsdf_gram_last_ms": [_vp],
    "podsdf_sym_pack_dev": [_i, _vp, _l, _vp, _vp],
    "podsdf_sym_unpack_dev": [_i, _vp, _l, _vp, _vp],
    "podsdf_eig_sym_host": [_i, _vp, _l, _vp, _vp],
    "podsdf_eig_sym_top_dev": [_i, _vp, _l, _i, _vp, ctypes.POINTER(_i), _vp],
    "podsdf_eig_sym_top_host": [_i, _vp, _l, _i, _vp, _vp, ctypes.POINTER(_i)],
    "podsdf_eig_sym_dev": [_i, _vp, _l, _vp, _vp],
    "podsdf_sym_tridiag_host": [_i, _vp, _l, _vp, _vp, _vp, _vp],
    "podsdf_field_moments_dev": [_i, _vp, _l, _l, _i, _vp, _vp, _vp],
    "podsdf_lag_covariance_dev": [_i, _vp, _l, _l, _l, _i, _vp, _i, _vp, _i, _vp, _vp],
    "podsdf_project_host": [_i, _vp, _l, _l, _i, _vp, _l, _vp, _i, _d, _vp],
    "podsdf_project_dev": [_i, _vp, _l, _l, _i, _vp, _l, _vp, _i, _d, _vp, _vp],
    "podsdf_dft_modes_host": [_i, _vp, _l, _l, _i, _vp],
    "podsdf_dft_modes_dev": [_i, _vp, _l, _l, _i, _vp, _vp],
}
_RESTYPES = {"podsdf_last_error": ctypes.c_char_p, "podsdf_launch_count": ctypes.c_longlong}

_lib = None


def lib():
    """The loaded library; raises (loudly) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libpodsdf.so not found at %s -- build it with "
                              "`python pods-digital-filter_b200/build.py` (there is no CPU fallback)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, ctypes.c_int)
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise PodsdfError(rc, lib().podsdf_last_error().decode("utf-8", "replace"))


def last_error():
    return lib().podsdf_last_error().decode("utf-8", "replace")


def device_count():
    n = ctypes.c_int(0)
    rc = lib().podsdf_device_count(ctypes.byref(n))
    return n.value if rc == 0 else 0


def require_device():
    n = ctypes.c_int(0)
    check(lib().podsdf_device_count(ctypes.byref(n)))
    return n.value


def launch_count():
    return int(lib().podsdf_launch_count())


def f64(a, shape=None, name="array", writable=False):
    """Validate a caller array for the C ABI: float64, C-contiguous (no silent copies of outputs)."""
    if not isinstance(a, np.ndarray):
        if writable:
            raise TypeError("%s must be a numpy array (written in place)" % name)
        a = np.ascontiguousarray(a, dtype=np.float64)
    if a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
        if writable:
            raise TypeError("%s must be a C-contiguous float64 array (written in place)" % name)
        a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError("%s has shape %s, expected %s" % (name, a.shape, tuple(shape)))
    return a


def ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def pinned_empty(shape, dtype=np.float64):
    """numpy array backed by pinned host memory (async H2D/D2H); keep the returned owner alive."""
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = ctypes.c_void_p()
    check(lib().podsdf_host_alloc(ctypes.byref(p), max(nbytes, 8)))
    buf = (ctypes.c_char * max(nbytes, 8)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    class _Owner(object):
        def __init__(self, addr):
            self.addr = addr

        def free(self):
            if self.addr:
                lib().podsdf_host_free(ctypes.c_void_p(self.addr))
                self.addr = None

        def __del__(self):
            try:
                self.free()
            except Exception:
                pass

    return arr, _Owner(p.value)


class DfPlan(object):
    """Owns a podsdf_df_plan.  Taps are explicit so the shim controls them bit for bit."""

    def __init__(self, J, K, nfx, nfy, nfz, bx, by, bz, device=0):
        self.J, self.K, self.nfx, self.nfy, self.nfz = int(J), int(K), int(nfx), int(nfy), int(nfz)
        self.Jp, self.Kp = self.J + 2 * self.nfy, self.K + 2 * self.nfz
        self.device = device
        bx, by, bz = f64(bx, (2 * nfx + 1,)), f64(by, (2 * nfy + 1,)), f64(bz, (2 * nfz + 1,))
        h = ctypes.c_void_p()
        check(lib().podsdf_df_plan_create(device, self.J, self.K, self.nfx, self.nfy, self.nfz,
                                          ptr(bx), ptr(by), ptr(bz), ctypes.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            try:
                lib().podsdf_df_plan_destroy(self._h)
            except TypeError:          # interpreter shutdown: module globals are already gone
                pass
            self._h = None

    __del__ = close

    def set_impl(self, impl):
        check(lib().podsdf_df_plan_set_impl(self._h, impl))

    def set_chunk(self, n):
        check(lib().podsdf_df_plan_set_chunk(self._h, int(n)))

    def set_rng(self, kind):
        """Philox stream of the seeded entries: RNG_U32 (default, 4 numbers per call) or RNG_U53 (round-1 stream)."""
        check(lib().podsdf_df_plan_set_rng(self._h, int(kind)))

    def set_coefficients(self, coef, mask):
        """coef (12, J*K) planes M00..M22, m0..m2 (or None for identity); mask: present planes."""
        if coef is None:
            check(lib().podsdf_df_set_coefficients(self._h, None, 0))
            return
        coef = f64(coef, (12, self.J * self.K), "coef")
        check(lib().podsdf_df_set_coefficients(self._h, ptr(coef), int(mask)))

    def set_rotation(self, R):
        """Constant 3x3 rotation applied after the Lund transform (None switches it off)."""
        if R is None:
            check(lib().podsdf_df_set_rotation(self._h, None))
            return
        R = f64(np.asarray(R, dtype=np.float64), (3, 3), "R")
        check(lib().podsdf_df_set_rotation(self._h, ptr(R)))

    def filter_block(self, x, y):
        x = f64(x, (2 * self.nfx + 1, self.Jp, self.Kp), "x")
        y = f64(y, (self.J, self.K), "y", writable=True)
        check(lib().podsdf_df_filter_block_host(self._h, ptr(x), ptr(y)))

    def lund_apply(self, yu, yv, yw):
        yu, yv, yw = (f64(a, (self.J, self.K), n, writable=True) for a, n in ((yu, "yu"), (yv, "yv"), (yw, "yw")))
        check(lib().podsdf_df_lund_apply_host(self._h, ptr(yu), ptr(yv), ptr(yw)))

    def generate_injected(self, slabs, out=None):
        """slabs (n, 3, Jp, Kp) -> out (n - 2nfx, 3, J, K)."""
        n = slabs.shape[0]
        slabs = f64(slabs, (n, 3, self.Jp, self.Kp), "slabs")
        ns = n - 2 * self.nfx
        if out is None:
            out = np.empty((ns, 3, self.J, self.K))
        out = f64(out, (ns, 3, self.J, self.K), "out", writable=True)
        check(lib().podsdf_df_generate_injected_host(self._h, ptr(slabs), n, ptr(out)))
        return out

    def generate_seeded(self, seed, t0, nsteps, out=None):
        if out is None:
            out = np.empty((nsteps, 3, self.J, self.K))
        out = f64(out, (nsteps, 3, self.J, self.K), "out", writable=True)
        check(lib().podsdf_df_generate_seeded_host(self._h, int(seed), int(t0), int(nsteps), ptr(out)))
        return out

    def random_slabs(self, seed, s0, n):
        out = np.empty((n, 3, self.Jp, self.Kp))
        check(lib().podsdf_df_random_slabs_host(self._h, int(seed), int(s0), int(n), ptr(out)))
        return out

    # device-pointer entries (addresses as ints, e.g. torch.Tensor.data_ptr()); asynchronous
    def generate_injected_dev(self, slabs_ptr, n_slabs, out_ptr, stream=0):
        check(lib().podsdf_df_generate_injected_dev(self._h, _vp(slabs_ptr), int(n_slabs), _vp(out_ptr), _vp(stream)))

    def generate_seeded_dev(self, seed, t0, nsteps, out_ptr, stream=0):
        check(lib().podsdf_df_generate_seeded_dev(self._h, int(seed), int(t0), int(nsteps), _vp(out_ptr), _vp(stream)))

    def random_slabs_dev(self, seed, s0, n, out_ptr, stream=0):
        check(lib().podsdf_df_random_slabs_dev(self._h, int(seed), int(s0), int(n), _vp(out_ptr), _vp(stream)))

    def info(self):
        """dict(chunk, tuned_time, tuned_plane, diagonal, passes) -- what the plan runs with its current settings."""
        a = (_i * 5)()
        check(lib().podsdf_df_plan_info(self._h, a))
        return dict(chunk=a[0], tuned_time=bool(a[1]), tuned_plane=bool(a[2]), diagonal=bool(a[3]), passes=a[4])

    def last_kernel_ms(self):
        ms = np.zeros(2)
        check(lib().podsdf_df_last_kernel_ms(self._h, ptr(ms)))
        return float(ms[0]), float(ms[1])


# ---------------------------------------------------------------------------------- PODFS entries
def release_scratch(device=0):
    """Free the grow-only device scratch the PODFS entries keep between calls."""
    check(lib().podsdf_release_scratch(device))


def gram(A, layout=LAYOUT_ROWS_BY_TIME, w=None, scale=1.0, C=None, device=0):
    """C = scale * A^T diag(w) A for the (M x N) snapshot matrix in the given layout."""
    A = f64(A, name="A")
    if layout == LAYOUT_ROWS_BY_TIME:
        M, N = A.shape
    else:
        N, M = A.shape
    P = 0
    if w is not None:
        w = f64(w, name="cell_volume")
        P = w.shape[0]
    if C is None:
        C = np.empty((N, N))
    C = f64(C, (N, N), "C", writable=True)
    check(lib().podsdf_gram_host(device, ptr(A), M, N, layout, ptr(w), P, float(scale), ptr(C)))
    return C


def eig_sym(C, device=0):
    """Ascending eigenvalues and eigenvectors (columns) of the symmetric matrix C."""
    C = f64(C, name="C")
    N = C.shape[0]
    lam = np.empty(N)
    V = np.empty((N, N))
    check(lib().podsdf_eig_sym_host(device, ptr(C), N, ptr(lam), ptr(V)))
    return lam, V


def eig_sym_top(C, nm, device=0):
    """All eigenvalues (ascending) and the unit eigenvectors of the nm largest as ROWS, descending; also returns
    which route ran (1 = tridiagonal route, 0 = full solver)."""
    C = f64(C, name="C")
    N = C.shape[0]
    lam = np.empty(N)
    V = np.empty((int(nm), N))
    path = _i(0)
    check(lib().podsdf_eig_sym_top_host(device, ptr(C), N, int(nm), ptr(lam), ptr(V), ctypes.byref(path)))
    return lam, V, path.value


def sym_tridiag(C, want_band=False, device=0):
    """Two-stage tridiagonalisation of the symmetric C on the device: (d, e, band or None, (ms stage 1, ms stage 2))."""
    C = f64(C, name="C")
    N = C.shape[0]
    d, e = np.empty(N), np.empty(max(N - 1, 0))
    band = np.empty((N, 33)) if want_band else None
    ms = np.zeros(2)
    check(lib().podsdf_sym_tridiag_host(device, ptr(C), N, ptr(d), ptr(e), ptr(band), ptr(ms)))
    return d, e, band, (float(ms[0]), float(ms[1]))


def project(A, T, dinv, nm, scale, layout=LAYOUT_ROWS_BY_TIME, device=0):
    A = f64(A, name="A")
    if layout == LAYOUT_ROWS_BY_TIME:
        M, N = A.shape
    else:
        N, M = A.shape
    T = f64(T, name="T")
    dinv = f64(dinv, (nm,), "dinv")
    Phi = np.empty((M, nm))
    check(lib().podsdf_project_host(device, ptr(A), M, N, layout, ptr(T), T.shape[1], ptr(dinv), int(nm),
                                    float(scale), ptr(Phi)))
    return Phi


def dft_modes(a, nm, device=0):
    """fftshift-ordered DFT/N of the first nm columns of a (N x ld): complex128 (N, nm)."""
    a = f64(a, name="temporal_modes")
    N, ld = a.shape
    c = np.empty((N, nm), dtype=np.complex128)
    check(lib().podsdf_dft_modes_host(device, ptr(a), N, ld, int(nm), ctypes.c_void_p(c.ctypes.data)))
    return c
