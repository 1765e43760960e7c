"""CPU oracle for the PODFS hot path -- TEST INFRASTRUCTURE ONLY.

Python-3 / numpy restatement of `PODFS.py` (snapshot POD + Fourier-series fit,
Treleaven 2019).  Only `tests/`, `__graft_entry__.smoke()` and the CPU-baseline
legs of `bench.py` import it.  Pinned against outputs of the reference's own code
(`oracle/make_golden.py` via `oracle/ref_py3.py`), see df_oracle.py header.

Citations `POD:a-b` are `/root/reference/PODFS.py` lines a-b.

Legacy semantics reproduced on purpose:
  * py2 integer division `num_fcs/2` (POD:1565, 1607, 1636, 1657)
  * Fourier coefficients stored in complex64 (POD:1547, 1579); the pre-rounding
    complex128 values are returned too, because 1e-10 parity is only meaningful there
  * energy-threshold arithmetic: under the numpy-1.x value-based casting the authors
    ran, `energy` accumulates float32 magnitudes in float64 and the threshold is
    float64(float32 sum) * et; under numpy>=2 (NEP 50) both stay float32.  `nep50`
    selects which; the golden vectors (made with numpy 2) use nep50=True.
  * tuple sort with reverse=True: ties broken by the HIGHER index first (POD:1440, 1584)
"""
import math

import numpy as np
from numpy import linalg


def calculate_correlation_matrix(num_snapshots, num_points, num_components,
                                 correct_for_cell_volumes, cell_volume, A, C):
    """POD:1451-1464, in place on C.  'false': C = A^T A / ns;  'true': volume-weighted,
    upper triangle computed then mirrored."""
    C[:, :] = 0.0
    ns = num_snapshots
    if correct_for_cell_volumes == "false":
        C[:, :] = np.dot(A[:, 0:ns].T, A[:, 0:ns]) / ns
    elif correct_for_cell_volumes == "true":
        w = np.asarray(cell_volume, dtype=np.float64)
        for i in range(ns):
            for j in range(i, ns):
                for k in range(num_components):
                    sl = slice(k * num_points, (k + 1) * num_points)
                    C[i, j] = C[i, j] + np.dot(A[sl, i].T * w[:], A[sl, j]) / ns
            for j in range(i):
                C[i, j] = C[j, i]
    # any other string: C stays zero, like the reference


def sort_eigenvalues(num_snapshots, energy, temporal_modes):
    """POD:1430-1447, in place.  NaN eigenvalue -> -1e10 and zeroed vector; descending
    sort of (value, index) tuples."""
    ns = num_snapshots
    vals = np.zeros(ns)
    idx = np.zeros(ns, dtype=int)
    for k in range(ns):
        idx[k] = k
        if math.isnan(energy[k].real) or math.isnan(energy[k].imag):
            vals[k] = -1.0e10
            temporal_modes[:, k] = 0.0
        else:
            vals[k] = energy[k].real
    pairs = sorted(zip(vals.tolist(), idx.tolist()), reverse=True)
    order = np.array([p[1] for p in pairs], dtype=int)
    energy[0:ns] = np.array([p[0] for p in pairs])
    temporal_modes[:, :] = temporal_modes.real.copy()[:, order]


def count_valid_modes(energy, num_snapshots, tol_CN):
    """POD:1312-1317 (the double-increment loop, capped by ns-2)."""
    n = 0
    tol2 = pow(tol_CN, 2.0)
    while (energy[n].real / energy[0].real > tol2) and (n < num_snapshots - 2) and (energy[n].real > 0.0):
        n += 1
        if (energy[n].real / energy[0].real > tol2) and (energy[n].real > 0.0):
            n += 1
    return n


def pod(A, num_snapshots, num_points, num_components, correct_for_cell_volumes="false",
        cell_volume=(), tol_CN=1.0e-15, num_modes_trunc=20, eig=None):
    """POD:1294-1333, 1391-1393 without the writers.  Returns a dict with
    C, energy, temporal_modes (ns,ns, scaled), spatial_modes (ncomp*P, nm), nm, num_valid.
    eig: optional replacement eigen-solver (default numpy.linalg.eig as POD:1309)."""
    ns = num_snapshots
    C = np.zeros((ns, ns))
    calculate_correlation_matrix(ns, num_points, num_components, correct_for_cell_volumes, cell_volume, A, C)
    energy, tm = (eig or linalg.eig)(C)
    energy = np.array(energy)
    tm = np.array(tm)
    sort_eigenvalues(ns, energy, tm)
    nv = count_valid_modes(energy, ns, tol_CN)
    nm = num_modes_trunc
    if nm < 0 or nm > nv:
        nm = nv
    for j in range(nv):                                   # POD:1323-1325
        mag = sum(tm[:, j].real * tm[:, j].real) / ns
        tm[:, j] = tm[:, j] * np.sqrt(energy[j].real / mag)
    einv = np.diag(np.ones(nm) / energy[0:nm].real, 0)    # POD:1330-1333
    phi = np.dot(np.dot(A[:, 0:ns], tm[:, 0:nm].real), einv) / ns
    return dict(C=C, energy=energy, temporal_modes=tm, spatial_modes=phi, nm=nm, num_valid=nv)


def dft_coefficients(temporal_modes, ns, nm, dt, direct=True):
    """POD:1540-1573.  c[n,i] = mean_j a_i(t_j) exp(-2 pi i k t_j / period), k = n - ns//2.
    Returns (c128 (ns,nm) complex128 pre-rounding, period, time)."""
    time = np.linspace(0, (ns - 1) * dt * 1, ns)
    period = time[-1] + (time[1] - time[0])
    c = np.zeros((ns, nm), dtype=np.complex128)
    for i in range(nm):
        y = temporal_modes[:, i]
        if direct:
            for n in range(ns):
                k = n - ns // 2
                ctemp = y * np.exp(-1j * 2 * k * np.pi * time / period)
                c[n, i] = ctemp.sum() / ctemp.size
        else:   # same quantity through the FFT (equal to rounding)
            c[:, i] = np.fft.fftshift(np.fft.fft(np.real(y))) / ns
    return c, period, time


def select_coefficients(c128, et, nep50=False):
    """POD:1547, 1577-1593.  complex64 storage, rank by |c| (ties: higher index first),
    count = smallest set whose magnitude sum reaches et * total.
    Returns (c64 (ns,nm), c_ind (nm,ns) int32, c_count (nm,) int64)."""
    ns, nm = c128.shape
    c = c128.astype(np.complex64)
    c_ind = np.zeros((nm, ns), dtype=np.int32)
    c_count = np.zeros(nm, dtype=np.int64)
    for i in range(nm):
        cmod = np.abs(c[:, i])                             # float32
        pairs = sorted(zip(cmod.tolist(), range(ns)), reverse=True)
        c_ind[i, :] = [p[1] for p in pairs]
        esum = np.sum(np.abs(c[:, i]))                     # float32 pairwise sum
        if nep50:
            energy = np.float32(0.0)
            thresh = np.float32(esum * np.float32(et))
            while energy < thresh:
                energy = np.float32(energy + np.abs(c[c_ind[i, c_count[i]], i]))
                c_count[i] += 1
        else:
            energy = 0.0
            thresh = float(esum) * et
            while energy < thresh:
                energy += float(np.abs(c[c_ind[i, c_count[i]], i]))
                c_count[i] += 1
    return c, c_ind, c_count


def fourier_coefficients(temporal_modes, ns, nm, dt, et, nep50=False, direct=True, reconstruct=False):
    """POD:1523-1659 without plots.  Returns dict(c128, c64, c_ind, c_count, period, FC,
    podfs_dat (text of PODFS.dat), y2 (optional reconstruction, POD:1602-1612))."""
    c128, period, time = dft_coefficients(temporal_modes, ns, nm, dt, direct=direct)
    c, c_ind, c_count = select_coefficients(c128, et, nep50=nep50)
    half = ns // 2
    FC = np.zeros((int(np.sum(c_count)), 3), dtype=np.float64)  # POD:1631-1639
    row = 0
    for i in range(nm):
        for j in range(c_count[i]):
            v = c[c_ind[i, j], i]
            FC[row, :] = [c_ind[i, j] - half, v.real, v.imag]
            row += 1
    lines = [str(nm), str(period)]                               # POD:1650-1657
    for i in range(nm):
        lines.append(str(i + 1) + "\t" + str(c_count[i]))
    for i in range(nm):
        for j in range(c_count[i]):
            v = c[c_ind[i, j], i]
            lines.append(str(c_ind[i, j] - half) + "\t" + str(v.real) + "\t" + str(v.imag))
    out = dict(c128=c128, c64=c, c_ind=c_ind, c_count=c_count, period=period, FC=FC,
               podfs_dat="\n".join(lines))
    if reconstruct:                                              # POD:1602-1612 (plot-only)
        y2 = np.zeros((ns, nm))
        for i in range(nm):
            sel = c_ind[i, :c_count[i]]
            k = sel - half
            ph = np.exp(1j * 2 * np.pi * np.outer(time, k) / period)
            y2[:, i] = (ph @ c[sel, i].astype(np.complex128)).real
        out["y2"] = y2
    return out


def synthetic_snapshots(M, N, seed=0, rank_r=40, noise=1e-3, row0=0, rows=None):
    """A = sum_r sigma_r phi_r(p) g_r(t) + noise, sigma_r = r^-0.8
    (separated spectrum), deterministic per (row block, seed).  Not in the reference;
    shared definition for both sides of the parity tests and the bench.  Rows
    [row0, row0+rows) of the (M, N) matrix, reference layout (row = c*P+p, time contiguous)."""
    rows = M - row0 if rows is None else rows
    t = np.arange(N) / N
    p = (np.arange(row0, row0 + rows) + 0.5) / M
    A = np.zeros((rows, N))
    for r in range(1, rank_r + 1):
        sig = r ** -0.8
        phi = np.sin(np.pi * r * p + 0.3 * r)
        g = np.cos(2 * np.pi * r * t + 0.7 * r) + 0.5 * np.sin(2 * np.pi * (r + 3) * t)
        A += sig * np.outer(phi, g)
    # counter-style noise so any row block is reproducible on its own
    ii = (np.arange(row0, row0 + rows, dtype=np.uint64)[:, None] * np.uint64(N)
          + np.arange(N, dtype=np.uint64)[None, :] + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15))
    ii ^= ii >> np.uint64(33)
    ii *= np.uint64(0xFF51AFD7ED558CCD)
    ii ^= ii >> np.uint64(33)
    ii *= np.uint64(0xC4CEB9FE1A85EC53)
    ii ^= ii >> np.uint64(33)
    A += noise * ((ii >> np.uint64(11)).astype(np.float64) / 9007199254740992.0 - 0.5)
    return A


def fct_welch(x, fs, N, iwindow):
    """nsigproclib_no_mpi.py:10-68 (reached from POD:1373 in verbose mode).  Welch PSD with 50 % overlap; block
    spectra are stored in complex64 before squaring, the running sum is float64 (imaginary part dropped), the
    result is returned as complex64.  `-N/2`, `N/2-1` are python-2 integer divisions."""
    w = {1: np.ones(N), 2: np.hanning(N), 3: np.blackman(N)}[iwindow]
    Cw = N / np.sum(w ** 2)
    noverlap = int(np.floor(N / 2))
    M, end = 1, N
    while end + noverlap <= x.size:
        end += noverlap
        M += 1
    f = np.linspace((-N) // 2, N // 2 - 1, N) / N * fs
    Sxx = np.zeros(N, dtype=np.complex64)
    acc = np.zeros(N)
    for j in range(M):
        Sxx[:] = np.fft.fft(x[j * noverlap:j * noverlap + N] * w)
        Sxx[:] = np.fft.fftshift(Sxx)
        acc[:] = (acc + Cw / N / fs * Sxx * np.conj(Sxx)).real
    Sxx[:] = acc / M
    return f, Sxx, M
