"""Import helper: the package directory `pods-digital-filter_b200/` has a hyphen in its name, so it
is registered under the importable name `pods_digital_filter_b200`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "pods-digital-filter_b200")
NAME = "pods_digital_filter_b200"


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(NAME, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception:
        sys.modules.pop(NAME, None)
        raise
    return mod
