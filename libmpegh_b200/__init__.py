"""libmpegh_b200 -- B200-native post-parse signal path of the MPEG-H 3D Audio decoder.

The product is the C-ABI shared library `libmpeghb200.so` (CUDA, sm_100a; sources in `csrc/`, interface in
`include/mpeghb200.h`).  This package is only its ctypes binding, used by the tests and `bench.py`.
There is no CPU implementation: every compute entry point fails with MPEGHB200_ERR_NO_DEVICE when no GPU
is visible, and importing `libmpegh_b200.lib` raises if the shared library has not been built.
"""
from .lib import Lib, MpeghError, load, build, SO_PATH  # noqa: F401
