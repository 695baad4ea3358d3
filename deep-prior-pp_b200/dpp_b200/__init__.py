"""dpp_b200 - ctypes binding and graph executor for libdpp_b200.so (sm_100a CUDA).

The directory that contains this package (``deep-prior-pp_b200/``) plays the role of the
reference's ``src/``: put it on ``sys.path`` and ``net``, ``trainer``, ``data``, ``util`` import
exactly as they do in moberweger/deep-prior-pp, with all arithmetic running in the CUDA library.
There is no CPU compute path: anything that needs the device raises if CUDA or the library is
missing.
"""
from .lib import lib, DppError, library_path  # noqa: F401
