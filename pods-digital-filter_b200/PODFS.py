"""PODFS (Treleaven 2019): snapshot POD + Fourier-series fit of the temporal coefficients --
B200-native drop-in for the hot path of the reference's `PODFS.py` (same function names, the
22-positional-argument `POD(...)`, results published on the `i_d` attribute bag).

GPU work (csrc/pod_kernels.cu through include/podsdf.h):
  calculate_correlation_matrix  POD:1451-1464  weighted Gram X^T W X, FP64 tensor-core MMA
  eigen-solve                   POD:1309       symmetric solver on the device (C is exactly symmetric)
  spatial modes                 POD:1329-1333  projection GEMM, FP64 tensor-core MMA
  Fourier coefficients          POD:1562-1573  batched direct DFT with exact twiddles
Host work kept verbatim in meaning: eigenvalue sort / NaN guard / validity count / scaling
(POD:1310-1327, 1430-1447), complex64 cast, ranking and energy cut (POD:1577-1593), file writers.
The VTK volume post-processing functions of the reference (POD:33-1240) are not on this path and
are not provided.
"""
import math
import os

import numpy as np

try:
    from . import _capi
except ImportError:
    import _capi

DEVICE = int(os.environ.get("PODSDF_DEVICE", "0"))


def _layout(A):
    """(array to hand over, layout flag).  A (rows, time) C-contiguous is the reference layout; a
    transposed view of the device-native planes (time, rows) is passed through without a copy."""
    if A.flags["C_CONTIGUOUS"]:
        return A, _capi.LAYOUT_ROWS_BY_TIME
    if A.T.flags["C_CONTIGUOUS"]:
        return A.T, _capi.LAYOUT_TIME_BY_ROWS
    return np.ascontiguousarray(A), _capi.LAYOUT_ROWS_BY_TIME


# =====================================================================================
def calculate_correlation_matrix(num_snapshots, num_points, num_components, correct_for_cell_volumes,
                                 cell_volume, A, C):
    """POD:1451-1464, in place on C (num_snapshots x num_snapshots)."""
    C[:, :] = 0.0
    ns = num_snapshots
    if correct_for_cell_volumes not in ("false", "true"):
        return      # the reference leaves C at zero for any other string
    a, layout = _layout(A[:, 0:ns])
    w = None
    if correct_for_cell_volumes == "true":
        w = np.ascontiguousarray(cell_volume, dtype=np.float64)
        if w.shape[0] != num_points or a.size != num_components * num_points * ns:
            raise ValueError("cell_volume / A do not match num_points, num_components")
    out = C if (C.flags["C_CONTIGUOUS"] and C.dtype == np.float64) else None
    res = _capi.gram(a, layout, w, 1.0 / ns, out, device=DEVICE)
    if out is None:
        C[:, :] = res


def sort_eigenvalues(num_snapshots, energy, temporal_modes):
    """POD:1430-1447, in place: NaN eigenvalues -> -1e10 with zeroed vectors, then descending order of
    (value, index) pairs -- ties resolve to the HIGHER original index first, like the tuple sort."""
    ns = num_snapshots
    vals = np.real(energy[:ns]).astype(np.float64)
    bad = np.isnan(np.real(energy[:ns])) | np.isnan(np.imag(energy[:ns]))
    vals[bad] = -1.0e10
    temporal_modes[:, :ns][:, bad] = 0.0
    order = np.lexsort((-np.arange(ns), -vals))
    energy[0:ns] = vals[order]
    temporal_modes[0:ns, 0:ns] = np.real(temporal_modes[0:ns, 0:ns])[:, order]


def _count_valid(energy, num_snapshots, tol_CN):
    """POD:1312-1317, including the double increment and the ns-2 cap."""
    n = 0
    tol2 = pow(tol_CN, 2.0)
    e0 = energy[0].real
    while (energy[n].real / e0 > tol2) and (n < num_snapshots - 2) and (energy[n].real > 0.0):
        n += 1
        if (energy[n].real / e0 > tol2) and (energy[n].real > 0.0):
            n += 1
    return n


def POD(A, num_snapshots, num_points, num_components, correct_for_cell_volumes, cell_volume, restart_dir,
        restart_flag, tol_CN, num_modes_trunc, num_modes_to_write, test_POD_orthogonality, write_matrices,
        grid, mean_field, dt, var_name, ifig, N, iwindow, stride, i_d):
    """POD:1294-1393.  Publishes i_d.temporal_modes (ns,ns), i_d.spatial_modes (ncomp*P, nm), i_d.nm."""
    print("\n==============================================================")
    print("Calculating POD modes ...")
    ns = num_snapshots
    C = np.zeros((ns, ns), dtype=np.float64)
    print("\n   Calculating covariance matrix ...")
    calculate_correlation_matrix(ns, num_points, num_components, correct_for_cell_volumes, cell_volume, A, C)

    print("\n   Solving eigenvalue problem ...")
    energy, temporal_modes = _capi.eig_sym(C, device=DEVICE)     # POD:1309 (C is exactly symmetric)
    sort_eigenvalues(ns, energy, temporal_modes)

    num_valid_modes = _count_valid(energy, ns, tol_CN)
    print("      Number of valid POD modes with positive energies = ", num_valid_modes)
    if (num_modes_trunc < 0) or (num_modes_trunc > num_valid_modes):
        num_modes_trunc = num_valid_modes

    print("\n   Scaling temporal modes ...")                       # POD:1323-1325
    v = temporal_modes[:, :num_valid_modes]
    mag = np.sum(v * v, axis=0) / ns
    with np.errstate(invalid="ignore", divide="ignore"):
        temporal_modes[:, :num_valid_modes] = v * np.sqrt(energy[:num_valid_modes] / mag)[None, :]

    print("\n   Calculating truncated spatial modes ...")          # POD:1329-1333
    nm = num_modes_trunc
    if nm > 0:
        a, layout = _layout(A[:, 0:ns])
        spatial_modes_trunc = _capi.project(a, temporal_modes, 1.0 / energy[0:nm], nm, 1.0 / ns, layout,
                                            device=DEVICE)
    else:
        spatial_modes_trunc = np.zeros((num_components * num_points, 0))

    print("\n==============================================================")
    print("Writing POD output ...")
    write_eigenvalues(num_valid_modes, ns, energy, restart_dir + "POD.eigenvalues.dat")
    if getattr(i_d, "verbose", False):
        write_temporal_modes(num_valid_modes, ns, dt, temporal_modes, restart_dir)
        # POD:1366-1388: Welch PSD of the written temporal modes.  The reference plots them (POD_PSD_mode_j.png,
        # matplotlib); here the same numbers go to a two-column text file next to the temporal modes.
        try:
            from . import nsigproclib_no_mpi as sp
        except ImportError:                      # flat import: `python pods-digital-filter_b200/digitalfilters.py -v`
            import nsigproclib_no_mpi as sp
        nw = num_modes_to_write
        if nw < 0 or nw > num_modes_trunc:
            nw = num_modes_trunc
        i_d.PSD = np.zeros((N, nw))
        for j in range(nw):
            fs, Sxx, _ = sp.fct_welch(temporal_modes[:, j], 1. / dt / stride, N, iwindow)
            i_d.PSD[:, j] = Sxx.real
            np.savetxt(restart_dir + "POD_PSD_mode_" + str(j) + ".dat", np.column_stack([fs, Sxx.real]))
    print("\nCode completed.")
    print("==============================================================\n")
    i_d.temporal_modes = temporal_modes
    i_d.spatial_modes = spatial_modes_trunc
    i_d.nm = num_modes_trunc
    i_d.energy = energy


# =====================================================================================
def write_eigenvalues(num_valid_modes, num_snapshots, energy, filename):
    """POD:1409-1427."""
    if num_valid_modes < 1:
        return
    cum = np.cumsum(np.real(energy[:num_valid_modes]))
    total = cum[num_valid_modes - 1]
    print("\n   Writing energies to")
    print("      ", filename)
    d = os.path.dirname(filename)
    if d and not os.path.exists(d):
        os.makedirs(d)
    with open(filename, "w") as f:
        f.write("#\n")
        f.write("# mode, energy, cumulative, percenterage energy, percentage cumulative, condition number "
                "(absolute value if negative)\n")
        f.write("#		Note: cummulative energies are set to zero after first negative energy")
        f.write("#\n")
        e0 = energy[0].real
        for i in range(0, num_valid_modes):
            e = energy[i].real
            f.write("%4.1d %18.10e %18.10e %18.10e %18.10e %18.10e\n"
                    % (i + 1, e, cum[i], e / total * 100.0, cum[i] / total * 100.0, math.sqrt(e / e0)))
        for i in range(num_valid_modes, num_snapshots):
            e = energy[i].real
            f.write("%4.1d %18.10e %18.10e %18.10e %18.10e %18.10e\n"
                    % (i + 1, e, 0.0, e / total * 100.0, 0.0, math.sqrt(abs(e / e0))))


def write_temporal_modes(num_valid_modes, num_snapshots, dt, temporal_modes, rdir):
    """POD:1468-1482."""
    print("\n   Writing temporal modes to")
    for j in range(0, num_valid_modes):
        name = rdir + "POD.temporal_mode_" + "%04d" % (j + 1) + ".dat"
        print("      ", name)
        with open(name, "w") as f:
            f.write("#\n# time, amplitude\n#\n")
            for i in range(0, num_snapshots):
                f.write("%18.10e %18.10e\n" % (i * dt, temporal_modes[i, j].real))


# =====================================================================================
def select_fourier(c128, energy_target):
    """POD:1547, 1577-1593: complex64 storage, ranking by magnitude (ties: higher index first) and the
    smallest leading set whose magnitude sum reaches energy_target * total.
    Returns (c complex64 (ns,nm), c_ind (nm,ns) int32, c_count (nm,) int64)."""
    num_fcs, num_modes = c128.shape
    c = c128.astype(np.complex64)                                          # POD:1547
    c_count = np.zeros(num_modes, dtype=np.int64)
    c_ind = np.zeros((num_modes, num_fcs), dtype=np.int32)
    cmT = np.ascontiguousarray(np.abs(c).T)                                # float32 magnitudes, one row per mode
    last = num_fcs - 1
    rev_index = (last - np.arange(num_fcs)).astype(np.uint64)

    def one_mode(i):
        cmod = cmT[i]
        # POD:1584 tuple sort, reversed: magnitude descending, ties -> higher index first.  The IEEE bit pattern of
        # a non-negative float32 is monotone in its value, so (~bits(|c|), N-1-index) packed in one uint64 is a
        # unique key whose plain ascending sort IS that order (no stable sort, no index indirection).
        if np.isnan(cmod).any():
            order = np.lexsort((-np.arange(num_fcs), -cmod))
        else:
            key = ((np.uint64(0xFFFFFFFF) - cmod.view(np.uint32).astype(np.uint64)) << np.uint64(32)) | rev_index
            key.sort()
            order = (last - (key & np.uint64(0xFFFFFFFF)).astype(np.int64))
        c_ind[i, :] = order
        # POD:1588-1593 under the numpy-1.x casting the reference ran with: float64 running sum of the
        # float32 magnitudes against float64(float32 total) * et
        thresh = float(np.sum(cmod)) * energy_target
        if thresh > 0.0:
            run = np.cumsum(cmod[order].astype(np.float64))
            c_count[i] = min(int(np.searchsorted(run, thresh, side="left")) + 1, num_fcs)

    # (a thread pool over the modes was tried: 7x SLOWER -- the per-mode numpy calls are too short; one vectorised pass
    # over all modes: 2x slower -- the per-mode rows stay in cache, the (nm, ns) temporaries do not)
    for i in range(num_modes):
        one_mode(i)
    return c, c_ind, c_count


def fourier_coefficients(i_d):
    """POD:1523-1659.  DFT of the first i_d.nm temporal modes on the GPU, then the reference's
    complex64 storage, magnitude ranking and energy cut; writes ./PODFS/PODFS.dat and, when
    i_d.hdf5, publishes i_d.period / i_d.N_FC / i_d.FC.  The plot-only reconstruction y2
    (POD:1602-1612) is not computed."""
    num_snapshots = i_d.ns
    dt = i_d.dt
    num_fcs = num_snapshots
    num_modes = i_d.nm
    energy_target = i_d.et
    rdir = "./PODFS/"
    time = np.linspace(0, (num_snapshots - 1) * dt * 1, num_snapshots)
    period = time[-1] + (time[1] - time[0])

    c128 = getattr(i_d, "c128", None)              # the device pipeline has already transformed the kept modes
    if c128 is None or c128.shape != (num_fcs, num_modes):
        c128 = _capi.dft_modes(i_d.temporal_modes, num_modes, device=DEVICE)   # (ns, nm), k = n - ns//2
    c, c_ind, c_count = select_fourier(c128, energy_target)
    for i in range(num_modes):
        print("The number of fourier coeffients used for mode ", i + 1, " is ", c_count[i])

    i_d.c = c
    i_d.c128 = c128
    i_d.c_ind = c_ind
    i_d.c_count = c_count
    half = num_fcs // 2                                                    # py2 integer division
    if i_d.hdf5:                                                           # POD:1627-1639
        i_d.period = period
        i_d.N_FC = c_count
        FC = np.zeros((int(np.sum(c_count)), 3), dtype=np.float64)
        row = 0
        for i in range(num_modes):
            sel = c_ind[i, :c_count[i]]
            FC[row:row + len(sel), 0] = sel - half
            FC[row:row + len(sel), 1] = c[sel, i].real
            FC[row:row + len(sel), 2] = c[sel, i].imag
            row += len(sel)
        i_d.FC = FC

    print("\n Saving to PODFS.dat ...")                                    # POD:1644-1659
    if not os.path.exists(rdir):
        os.makedirs(rdir)
    parts = [str(num_modes), "\n" + str(period)]
    for i in range(num_modes):
        parts.append("\n" + str(i + 1) + "\t" + str(c_count[i]))
    for i in range(num_modes):
        for j in range(c_count[i]):
            v = c[c_ind[i, j], i]
            parts.append("\n" + str(c_ind[i, j] - half) + "\t" + str(v.real) + "\t" + str(v.imag))
    with open(rdir + "PODFS.dat", "w") as f:
        f.write("".join(parts))


# =====================================================================================
# Inlet-plane geometry and .prf writers without VTK ("next" row 8(f)-1).
class InflowPlane(object):
    """Stand-in for the vtkPolyData the reference carries in i_d.grid: cell centres only."""

    def __init__(self, centers):
        self.centers = centers      # (P, 3) float32, cell id = k + j*K

    def GetNumberOfCells(self):
        return self.centers.shape[0]


def _rot(angle_deg, axis):
    ax = np.asarray(axis, dtype=np.float64)
    n = np.linalg.norm(ax)
    if n == 0.0 or angle_deg == 0.0:
        return np.eye(3)
    ax = ax / n
    a = np.deg2rad(angle_deg)
    Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    return np.eye(3) + np.sin(a) * Kx + (1 - np.cos(a)) * Kx.dot(Kx)


def make_inflow_plane(i_d):
    """POD:1243-1290 without VTK: vtkPlaneSource(K, J) with normal (1,0,0), scaled by
    (0, res J^2/(J-1), res K^2/(K-1)), rotated onto the plane normal, twisted by beta+rot and
    translated to the origin; points are float32 like VTK's."""
    J, K = i_d.jma, i_d.kma
    nx, ny, nz = i_d.n
    alpha = np.arccos(nx) * 180 / np.pi
    beta = np.arctan2(nz, ny) * 180 / np.pi
    s2 = i_d.res * float(J) * float(J) / (float(J) - 1) if J > 1 else 0.0
    s3 = i_d.res * float(K) * float(K) / (float(K) - 1) if K > 1 else 0.0
    y = (-0.5 + (np.arange(J) + 0.5) / J) * s2
    z = (0.5 - (np.arange(K) + 0.5) / K) * s3
    pts = np.zeros((J * K, 3))
    pts[:, 1] = np.repeat(y, K)
    pts[:, 2] = np.tile(z, J)
    pts = pts.astype(np.float32).astype(np.float64)
    pts = pts.dot(_rot(alpha, (0, -nz, ny)).T).astype(np.float32).astype(np.float64)
    pts = pts.dot(_rot(beta + i_d.rot, (nx, ny, nz)).T).astype(np.float32).astype(np.float64)
    pts = (pts + np.asarray(i_d.t_o, dtype=np.float64)[None, :]).astype(np.float32)
    return InflowPlane(pts)


def _fmt(val):       # nsigproclib_no_mpi.str, SIG:880-882
    return "%0.12f" % val


def _write_prf(filename, title, normal, plane_rhs, points, u):
    with open(filename, "w") as t:
        t.write("# " + title + " # name of the profile\n")
        t.write("# turbulence model, none\n")
        t.write("# plane normal and translation " + str(normal[0]) + "\t" + str(normal[1]) + "\t" + str(normal[2])
                + "\t" + str(plane_rhs) + "\n")
        t.write("type, xyz # type of profile (rad or xyz)\n")
        t.write("localcs,origin,0,0,0 # origin of local coordinate system\n")
        t.write("localcs,xaxis,1,0,0 # x axis direction of local coordinate system\n")
        t.write("localcs,yaxis,0,1,0 # y axis direction of local coordinate system\n")
        t.write("localcs,zaxis,0,0,1 # z axis direction of local coordinate system\n")
        t.write("tolerance, 1.00E-08 # tolerance\n")
        t.write("scale,1,1,1,1,1,1 # scaling factors\n")
        t.write("data,x,y,z,u,v,w\n")
        cols = np.concatenate([np.asarray(points, dtype=np.float64), u], axis=1)
        t.write("".join(",".join(_fmt(v) for v in row) + "\n" for row in cols))


def save_plane(u, i_d):
    """POD:854-887: one generated inlet as ./PODFS/<time>.prf."""
    pts = i_d.grid.centers
    npt = pts.shape[0]
    rhs = i_d.t_o[0] * i_d.n[0] + i_d.t_o[1] * i_d.n[1] + i_d.t_o[2] * i_d.n[2]
    uvw = np.stack([u[:npt], u[npt:2 * npt], u[2 * npt:3 * npt]], axis=1)
    _write_prf("./PODFS/" + ("%.5E" % i_d.time) + ".prf", "Generated using the digital filter method",
               i_d.n, rhs, pts, uvw)


def pod2prf(i_d):
    """POD:1662-1800: mean field and spatial modes as .prf profiles (and the arrays HDF5.py needs)."""
    rdir = "./PODFS/"
    num_modes, num_points = i_d.nm, i_d.num_points
    i_d.turbulence_model = "none"
    i_d.t_o = np.array([0, 0, 0])
    pts = i_d.grid.centers
    n1, n2, n3 = i_d.n
    if i_d.hdf5:
        i_d.mean = np.zeros((num_points, 6), dtype=np.float64)
        i_d.modes = np.zeros((num_modes, num_points, 6), dtype=np.float64)
    u = np.asarray(i_d.mean_field).reshape((num_points, 3), order="F")
    if i_d.hdf5:
        i_d.mean[:, 0:3] = pts
        i_d.mean[:, 3:] = u
    rhs = i_d.t_o[0] * n1 + i_d.t_o[1] * n2 + i_d.t_o[2] * n3
    _write_prf(rdir + "PODFS_mean.prf", "PODFS_mean", i_d.n, rhs, pts, u)
    for i in range(num_modes):
        counter = "%4.4i" % (i + 1)
        u = np.asarray(i_d.spatial_modes[:, i]).reshape((num_points, 3), order="F")
        if i_d.hdf5:
            i_d.modes[i, :, 0:3] = pts
            i_d.modes[i, :, 3:] = u
        _write_prf(rdir + "PODFS_mode_" + counter + ".prf", "PODFS_mode_" + counter, i_d.n, 0 * n1 + 0 * n2 + 0 * n3,
                   pts, u)
