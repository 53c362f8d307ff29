"""zpaqsharp_b200 -- B200-native ZPAQ block codec behind the LibZPAQ API surface.

The product is `libzpaqb200.so` (hand-written CUDA for sm_100a + a host scheduler, C ABI in
include/zpaqb200.h).  `zpaqsharp_b200.libzpaq` binds it for Python; `bindings/csharp/` holds the
P/Invoke facade a ZPAQSharp maintainer would use.
"""
from . import libzpaq  # noqa: F401

__all__ = ["libzpaq"]
