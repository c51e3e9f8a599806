"""Drop-in alias: put cufinufft_b200/compat on sys.path and `from cufinufft import cufinufft`
resolves to the B200 implementation (same public names as the reference's python/cufinufft)."""
from cufinufft_b200 import cufinufft  # noqa: F401
from cufinufft_b200 import _cufinufft  # noqa: F401

__all__ = ["cufinufft"]
__version__ = "1.3"
