"""pods-digital-filter_b200: B200-native digital-filter (Klein 2003) + PODFS (Treleaven 2019) hot paths.

The directory name is not a Python identifier; load it with `podsdf_loader.load()` (registers the
package as `pods_digital_filter_b200`), or put this directory on sys.path and `import digitalfilters`,
`import PODFS as pod` exactly as with the reference's flat scripts.
"""
from . import _capi          # noqa: F401  (fails loudly if libpodsdf.so has not been built)
from . import _lund          # noqa: F401
from . import PODFS          # noqa: F401
from . import HDF5           # noqa: F401
from . import nsigproclib_no_mpi  # noqa: F401
from . import digitalfilters  # noqa: F401
from . import pipeline       # noqa: F401  (torch is imported lazily, only by the device drivers)

__all__ = ["_capi", "_lund", "PODFS", "HDF5", "nsigproclib_no_mpi", "digitalfilters", "pipeline"]
