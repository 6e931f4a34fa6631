"""gf2_b200 — host-side Python mirror of the C ABI in include/gf2_abi.h (ctypes over libgf2_b200.so).
The CUDA library is loaded lazily by `lib()`; there is no CPU fallback: without the built extension every
compute entry point raises."""
from . import abi  # noqa: F401
