"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/podsdf.h
declares, and refuses to compute without a B200 (no silent CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests.util import ROOT, pkg


def header_symbols():
    src = open(os.path.join(ROOT, "include", "podsdf.h")).read()
    return sorted(set(re.findall(r"PODSDF_API\s+[\w\s\*]+?\b(podsdf_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    capi = pkg()._capi
    L = capi.lib()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), "libpodsdf.so lacks %s" % s
    assert sorted(capi.SIGNATURES) == syms, "ctypes table and header disagree"
    assert L.podsdf_version() >= 100


def test_sm100a_code_is_embedded():
    import subprocess
    capi = pkg()._capi
    out = subprocess.run(["cuobjdump", "--list-elf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_device():
    p = pkg()
    capi = p._capi
    if capi.device_count() > 0:
        pytest.skip("a B200 is visible")
    with pytest.raises(capi.PodsdfError) as e:
        capi.require_device()
    assert e.value.code == capi.ERR_NO_DEVICE
    y = np.zeros((6, 7))
    with pytest.raises(capi.PodsdfError):
        p.digitalfilters.filter3DSciPy1D(np.zeros((7, 12, 13)), y, None, 6, 7, 1.5, 1.5, 1.5, 3, 3, 3)
    with pytest.raises(capi.PodsdfError):
        p.PODFS.calculate_correlation_matrix(4, 5, 3, "false", [], np.zeros((15, 4)), np.zeros((4, 4)))
    with pytest.raises(capi.PodsdfError):
        capi.dft_modes(np.zeros((8, 2)), 2)


def test_argument_validation_is_host_side():
    capi = pkg()._capi
    with pytest.raises(TypeError):
        capi.f64(np.zeros((3, 3), dtype=np.float32), writable=True)
    with pytest.raises(ValueError):
        capi.f64(np.zeros((3, 3)), shape=(3, 4))
    a = np.zeros((4, 4))[:, ::2]
    with pytest.raises(TypeError):
        capi.f64(a, writable=True)
    assert capi.f64(a).flags["C_CONTIGUOUS"]


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import it."""
    pdir = os.path.join(ROOT, "pods-digital-filter_b200")
    for fn in os.listdir(pdir):
        if fn.endswith(".py"):
            src = open(os.path.join(pdir, fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), fn


def test_flat_import_verbose_pod_with_stubbed_capi(tmp_path):
    """Script mode (`python pods-digital-filter_b200/digitalfilters.py -v`): the shims are imported flat, with the
    package directory on sys.path.  POD(verbose) reaches nsigproclib_no_mpi -> _capi through the same try/except
    as the module tops.  _capi is stubbed with numpy so the test runs without a GPU."""
    import subprocess
    import sys
    import textwrap
    code = textwrap.dedent("""
        import sys, types, os
        import numpy as np
        sys.path.insert(0, %r)
        stub = types.ModuleType("_capi")
        stub.LAYOUT_ROWS_BY_TIME, stub.LAYOUT_TIME_BY_ROWS = 0, 1
        def gram(a, layout, w, scale, out, device=0):
            A = a if layout == 0 else a.T
            C = scale * A.T.dot(A)
            if out is not None:
                out[:, :] = C
            return C
        def eig_sym(C, device=0):
            return np.linalg.eigh(C)
        def project(a, T, dinv, nm, scale, layout, device=0):
            A = a if layout == 0 else a.T
            return scale * A.dot(T[:, :nm]) * dinv[None, :]
        def dft_modes(a, nm, device=0):
            N = a.shape[0]
            return np.fft.fftshift(np.fft.fft(a[:, :nm], axis=0), axes=0) / N
        stub.gram, stub.eig_sym, stub.project, stub.dft_modes = gram, eig_sym, project, dft_modes
        sys.modules["_capi"] = stub
        import PODFS as pod
        import nsigproclib_no_mpi as sp
        assert pod._capi is stub and sp._capi is stub
        class Bag(object):
            pass
        i_d = Bag()
        i_d.verbose = True
        rng = np.random.RandomState(3)
        ns, P = 16, 10
        A = rng.standard_normal((3 * P, ns))
        A -= A.mean(axis=1)[:, None]
        os.makedirs("PODFS", exist_ok=True)
        pod.POD(A, ns, P, 3, "false", [], "PODFS/", "false", 1.0e-15, 4, 4, "false", "false", None, None, 0.1,
                "velocity", 1, ns, 1, 1, i_d)
        assert i_d.PSD.shape == (ns, 4) and np.isfinite(i_d.PSD).all()
        assert os.path.exists("PODFS/POD_PSD_mode_3.dat") and os.path.exists("PODFS/POD.temporal_mode_0001.dat")
        print("flat-ok")
    """) % os.path.join(ROOT, "pods-digital-filter_b200")
    r = subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0 and "flat-ok" in r.stdout, r.stdout + r.stderr
