"""cufinufft_b200 -- B200-native (sm_100a) implementation of cuFINUFFT v1.3's type-1/type-2
NUFFT hot path behind the reference's own C API and Python class.

    from cufinufft_b200 import cufinufft          # same class as `from cufinufft import cufinufft`

Importing this package loads cufinufft_b200/lib/libcufinufft.so and raises if it is missing:
there is no CPU fallback.
"""
from .cufinufft import cufinufft

__all__ = ["cufinufft"]
__version__ = "0.1"
