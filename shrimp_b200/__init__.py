"""shrimp_b200 -- B200-native hot path of SHRiMP2's gmapper (seed scan -> vector SW -> full SW).

Python here is plumbing only: a ctypes mirror of the reference's operator interface
(sw_vector_setup/sw_vector, sw_full_ls, ..., see shrimp_b200/api.py) over the C ABI in
include/shrimp_b200.h.  All compute is in libshrimp_b200.so (hand-written CUDA for sm_100a).
"""
from ._lib import ShrimpGpuError, LIB_PATH  # noqa: F401
from .api import GpuContext, pack_bases, pack_colours, LS_DEFAULT_SCORES, CS_DEFAULT_SCORES  # noqa: F401
