"""Seeded slices of tools/fuzz_df.py and tools/fuzz_pod.py as part of the `-m gpu` suite:
random plane sizes, independent filter half-widths per axis (including every tuned tap count), coefficient masks,
rotations, chunk sizes, plan implementations, injected and generated streams with random time offsets -- each case
against a plain numpy contraction of the same inputs -- and the PODFS entries against numpy/LAPACK."""
import importlib.util
import os

import numpy as np
import pytest

from tests.util import ROOT, pkg

pytestmark = pytest.mark.gpu


def _tool(name):
    pkg()._capi.require_device()
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_fuzz_df_200_cases():
    fz = _tool("fuzz_df")
    rs = np.random.RandomState(20261019)
    ran, worst, worst_cfg = 0, 0.0, None
    for i in range(200):
        r = fz.one_case(rs, i)
        if r is None:
            continue
        ran += 1
        if r[0] > worst:
            worst, worst_cfg = r
    assert ran >= 150
    assert worst < 1e-11, (worst, worst_cfg)


def test_fuzz_pod_60_cases():
    fz = _tool("fuzz_pod")
    rs = np.random.RandomState(20261019)
    for _ in range(60):
        errs, cfg = fz.one_case(rs)
        assert errs["sym"] == 0.0, cfg
        assert max(errs["gram"], errs["eig"], errs["proj"], errs["dft"]) < 1e-11, (errs, cfg)
