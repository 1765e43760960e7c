"""Welch power spectral density of the temporal modes -- the one diagnostic of the reference's signal library
that `PODFS.POD` reaches (verbose mode, POD:1366-1373 -> nsigproclib_no_mpi.py:10-68).

Same name, arguments and return values as the reference's `fct_welch`; the block transforms run as ONE call of
the device DFT kernel (`podsdf_dft_modes_host`, include/podsdf.h) with the windowed 50 %-overlap blocks as
columns.  Like the reference, the block spectra are held in complex64 before they are squared."""
import sys

import numpy as np

try:
    from . import _capi
except ImportError:                              # flat import (script mode), like the other shims
    import _capi


def _window(N, iwindow):
    if iwindow == 1:
        print("WINDOW: rectangular")
        return np.ones(N, dtype=np.float64)
    if iwindow == 2:
        print("WINDOW: Hanning")
        return np.hanning(N)
    if iwindow == 3:
        print("Window: Blackman")
        return np.blackman(N)
    raise ValueError("iwindow must be 1 (rectangular), 2 (Hanning) or 3 (Blackman)")


def fct_welch(x, fs, N, iwindow, device=0):
    """SIG:10-68.  Returns (f, Sxx, M): frequencies (fftshift order), the averaged PSD as complex64 with zero
    imaginary part, and the number of 50 %-overlap blocks."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    if N > x.size:
        print("*** ERROR: Block size N should not be larger than the signal size.")
        print("*** STOP !!!")
        sys.exit()
    w = _window(N, iwindow)
    Cw = N / np.sum(w ** 2)
    hop = int(np.floor(N / 2))
    M = 1 + (x.size - N) // hop if hop > 0 else 1
    print("Welch: number of blocks: M= " + str(M))
    f = np.linspace((-N) // 2, N // 2 - 1, N) / N * fs          # python-2 integer division of the reference
    blocks = np.empty((N, M))
    for j in range(M):
        blocks[:, j] = x[j * hop:j * hop + N] * w
    spec = (_capi.dft_modes(blocks, M, device) * N).astype(np.complex64)     # fftshift(fft(block)), stored c64
    power = (np.complex64(Cw / N / fs) * spec * np.conj(spec)).real.astype(np.float64)
    Sxx = (power.sum(axis=1) / M).astype(np.complex64)
    return f, Sxx, M
