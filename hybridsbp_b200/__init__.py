"""hybridsbp_b200 -- B200-native (sm_100a) hybridized summation-by-parts solve path.

Host-side mirror of the reference's entry points (see host.py) over the C-ABI of
libhsbp.so (include/hsbp.h).  The compute path is hand-written CUDA only; importing
this package never pulls in the CPU oracle.
"""
from ._lib import Context, DeviceArray, HsbpError, lib, declared_symbols, LIB_PATH  # noqa: F401
from .blocks import Blocks, SpdFactor, Trace, LOCAL_PCG, LOCAL_CHOLESKY, LOCAL_BAND, LOCAL_FDM  # noqa: F401

BC_DIRICHLET = 1
BC_NEUMANN = 2
BC_LOCKED_INTERFACE = 0
BC_JUMP_INTERFACE = 7
