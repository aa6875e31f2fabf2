"""ctypes binding of the CPU oracle (oracle/libkzn_oracle.so).  TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package (kazan_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from kazan_b200 import abi

_DIR = Path(__file__).resolve().parent
_PATH = _DIR / "libkzn_oracle.so"
_lib = None

P, u32, i32, u64, vp = C.POINTER, C.c_uint32, C.c_int32, C.c_uint64, C.c_void_p

_PROTOS = {
    "kzo_set_threads": (None, [C.c_int]),
    "kzo_get_threads": (C.c_int, []),
    "kzo_hardware_threads": (C.c_int, []),
    "kzo_float_to_unorm8": (C.c_uint8, [C.c_float]),
    "kzo_unorm8_to_float": (C.c_float, [C.c_uint8]),
    "kzo_pack_rgba8": (u32, [P(C.c_float)]),
    "kzo_texel_size": (u32, [u32]),
    "kzo_texel_offset": (u64, [P(abi.ImageDesc), u32, u32, u32, u32]),
    "kzo_image_memory_requirements": (i32, [P(abi.ImageDesc), P(u64), P(u64)]),
    "kzo_clear_color_image": (i32, [P(abi.ImageDesc), P(abi.ClearColorValue), u32, P(abi.ImageSubresourceRange)]),
    "kzo_clear_depth_stencil_image": (i32, [P(abi.ImageDesc), P(abi.ClearDepthStencilValue), u32,
                                            P(abi.ImageSubresourceRange)]),
    "kzo_clear_attachment": (i32, [P(abi.ImageDesc), P(abi.ClearColorValue), u32, P(abi.ClearRect)]),
    "kzo_copy_buffer_to_image": (i32, [vp, u64, P(abi.ImageDesc), u32, P(abi.BufferImageCopy)]),
    "kzo_copy_image_to_buffer": (i32, [P(abi.ImageDesc), vp, u64, u32, P(abi.BufferImageCopy)]),
    "kzo_copy_buffer": (i32, [vp, u64, vp, u64, u32, P(abi.BufferCopy)]),
    "kzo_fill_buffer": (i32, [vp, u64, u64, u64, u32]),
    "kzo_update_buffer": (i32, [vp, u64, u64, u64, vp]),
    "kzo_copy_image": (i32, [P(abi.ImageDesc), P(abi.ImageDesc), u32, P(abi.ImageCopy)]),
    "kzo_blit_image": (i32, [P(abi.ImageDesc), P(abi.ImageDesc), u32, P(abi.ImageBlit), u32]),
    "kzo_convert_tiling": (i32, [P(abi.ImageDesc), P(abi.ImageDesc)]),
    "kzo_draw": (i32, [P(abi.GraphicsPipelineState), P(abi.Framebuffer), P(abi.Viewport), P(abi.Rect2D),
                       P(abi.VertexInput), u32, u32, u32, u32, P(abi.DrawStatistics)]),
    "kzo_draw_indexed": (i32, [P(abi.GraphicsPipelineState), P(abi.Framebuffer), P(abi.Viewport), P(abi.Rect2D),
                               P(abi.VertexInput), vp, u64, u32, u32, u32, u32, i32, u32, P(abi.DrawStatistics)]),
    "kzo_fs_procedural_orbit": (None, [u32, u32, C.c_float, C.c_float, P(C.c_float)]),
    "kzo_splitmix64": (u64, [P(u64)]),
    "kzo_fill_random_bytes": (None, [vp, u64, u64]),
    "kzo_make_grid_mesh": (None, [u32, u32, u64, vp, vp]),
}


def build(force: bool = False) -> Path:
    # make decides whether the sources are newer than the library; on a machine without gcc (never the case in this
    # image) the prebuilt library is used as it is
    import shutil
    if force or not _PATH.exists() or shutil.which("make"):
        subprocess.run(["make", "-C", str(_DIR), "-s"] + (["-B"] if force else []), check=True)
    return _PATH


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(str(_PATH))
        for name, (res, args) in _PROTOS.items():
            if hasattr(lib, name):
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def ptr(a: np.ndarray) -> int:
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def aligned_empty(nbytes: int, align: int = 64) -> np.ndarray:
    """uint8 array whose data pointer is `align`-byte aligned (image.rs IMAGE_ALIGNMENT)."""
    raw = np.empty(nbytes + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + nbytes]


def random_bytes(nbytes: int, seed: int) -> np.ndarray:
    a = aligned_empty(nbytes)
    load().kzo_fill_random_bytes(C.c_void_p(ptr(a)), nbytes, seed)
    return a


def array_of(ctype, items):
    items = list(items)
    return (ctype * max(len(items), 1))(*items), len(items)
