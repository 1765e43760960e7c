import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_sessionstart(session):
    """The C-ABI library is built in-tree and git-ignored: build it when a fresh checkout has none (nvcc
    cross-compiles sm_100a without a GPU; on the GPU box the built .so travels with the snapshot)."""
    lib = os.path.join(ROOT, "pods-digital-filter_b200", "libpodsdf.so")
    if not os.path.exists(lib):
        import importlib.util
        spec = importlib.util.spec_from_file_location("podsdf_build", os.path.join(ROOT, "pods-digital-filter_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build(force=False, verbose=False)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
