"""pism_b200 -- B200-native drop-in for PISM's SIAFD stress-balance hot path.

The product is the CUDA shared library ``libsiafd_b200.so`` (C ABI in ``include/siafd_b200.h``).
This package is the Python host-side mirror of the reference's ``SSB_Modifier``/``SIAFD``
interface on top of that ABI (ctypes), plus the grid/decomposition helpers and the synthetic
input generators the tests and ``bench.py`` share.  There is no CPU fallback: importing
``pism_b200.capi`` fails loudly when the library is missing, and creating a handle fails
loudly when no CUDA device is present.
"""

__all__ = ["capi", "grid", "sia", "synthetic", "verification", "halo"]
__version__ = "0.1.0"
