"""kazan_b200 — B200-native execution back end for the Kazan Vulkan driver's data-parallel hot path.

The product is the C-ABI library `kazan_b200/libkazan_b200.so` (hand-written sm_100a kernels,
include/kazan_b200.h).  This package holds its sources (csrc/), the build recipe and a small ctypes
harness used by tests and bench.  Nothing here falls back to the CPU or to PyTorch.
"""
from . import abi  # noqa: F401
from .host import CommandBuffer, Device, Fence, Image, KazanError, Memory, Queue, Semaphore, library_path, load_library  # noqa: F401

__all__ = ["abi", "Device", "Fence", "Image", "KazanError", "Memory", "Queue", "library_path", "load_library"]
