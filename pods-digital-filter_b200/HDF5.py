"""Single-file HDF5 output of a PODFS model -- drop-in for the reference's `HDF5.py` (H5:11-64).

Tree (docs/usage/CFDCodeIntegration.rst:61-74):  /main  attrs N_POD, period;  /main/N_FC (nm,) int;
/main/FC (3*sum N_FC,) Fortran-flattened [k, Re, Im];  /main/mean (6*Np,) Fortran-flattened
[x,y,z,u,v,w] with attrs Np, Nvar, Vars, SF;  /main/modes/mode_%04d likewise.
h5py is imported lazily: it is not installed in the build image, and nothing on the GPU hot path
needs it.  `hdf5_tree(i_d)` returns the same content as plain numpy arrays (used by the tests and
by callers that stage results through pinned host memory).
"""
import numpy as np


def hdf5_tree(i_d):
    """The datasets and attributes write_HDF5 stores, as a nested dict of numpy arrays."""
    nfc = int(np.sum(i_d.N_FC))
    attrs = {"Np": i_d.num_points, "Nvar": 6, "Vars": np.bytes_("x,y,z,u,v,w,dummy"),
             "SF": [1., 1., 1., 1., 1., 1.]}
    tree = {
        "main": {
            "attrs": {"N_POD": i_d.nm, "period": i_d.period},
            "N_FC": np.asarray(i_d.N_FC, dtype="i"),
            "FC": np.asarray(i_d.FC, dtype=np.float64).reshape(nfc * 3, order="F"),
            "mean": {"data": np.asarray(i_d.mean).reshape(i_d.num_points * 6, order="F"), "attrs": dict(attrs)},
            "modes": {},
        }
    }
    for i in range(0, i_d.nm):
        tree["main"]["modes"]["mode_%4.4i" % (i + 1)] = {
            "data": np.asarray(i_d.modes[i, :, :]).reshape(i_d.num_points * 6, order="F"), "attrs": dict(attrs)}
    return tree


def write_HDF5(i_d):
    """H5:11-64: writes PODFS/PODFS.hdf5."""
    try:
        import h5py
    except ImportError:
        raise ImportError("write_HDF5 needs h5py, which is not installed here; "
                          "HDF5.hdf5_tree(i_d) returns the same content as numpy arrays")
    tree = hdf5_tree(i_d)["main"]
    f = h5py.File("PODFS/PODFS.hdf5", "w")
    main = f.create_group("main")
    for k, v in tree["attrs"].items():
        main.attrs[k] = v
    main.create_dataset("N_FC", (i_d.nm,), dtype="i")[:] = tree["N_FC"]
    main.create_dataset("FC", tree["FC"].shape, dtype=np.float64)[:] = tree["FC"]

    def put(group, name, node):
        d = group.create_dataset(name, node["data"].shape, dtype=np.float64)
        d[:] = node["data"]
        for k, v in node["attrs"].items():
            d.attrs[k] = v

    put(main, "mean", tree["mean"])
    modes = main.create_group("modes")
    for name, node in tree["modes"].items():
        put(modes, name, node)
    f.close()
