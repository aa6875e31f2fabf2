"""Run one case of tests/cases.py on the oracle (host memory) or on the GPU (through the C ABI).

Both runners return the full destination memory as bytes so that parity is a byte-for-byte comparison.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from kazan_b200 import abi, host
from oracle import binding as ob
from tests import cases, util
from tests.util import check


def _img(spec: dict, address: int) -> abi.ImageDesc:
    f = cases.image_desc_fields(spec)
    return abi.ImageDesc(address, f["format"], f["tiling"], f["width"], f["height"], f["depth"], f["array_layers"], 1, 0)


def _host_img(spec: dict, address: int) -> host.Image:
    f = cases.image_desc_fields(spec)
    return host.Image(address, f["width"], f["height"], f["depth"], f["array_layers"], f["format"], f["tiling"])


def _size(spec: dict) -> int:
    return _host_img(spec, 0x1000).memory_requirements()[0]


def _ranges(rs, aspect=abi.ASPECT_COLOR):
    return [abi.ImageSubresourceRange(aspect, 0, abi.REMAINING, b, n) for b, n in rs]


# ---- oracle --------------------------------------------------------------------------------------------

def oracle_clear(spec, ranges, color=cases.CLEAR_COLOR) -> bytes:
    lib = ob.load()
    mem = ob.random_bytes(_size(spec), 11)
    d = _img(spec, ob.ptr(mem))
    c = abi.ClearColorValue()
    c.float32[:] = list(color)
    a, n = util.arr(abi.ImageSubresourceRange, _ranges(ranges))
    check(lib.kzo_clear_color_image(C.byref(d), C.byref(c), n, a))
    return mem.tobytes()


def oracle_buffer_to_image(spec, regions, buffer_bytes) -> bytes:
    lib = ob.load()
    mem = ob.random_bytes(_size(spec), 12)
    buf = ob.random_bytes(buffer_bytes, 0x4B415A414E)
    d = _img(spec, ob.ptr(mem))
    a, n = util.arr(abi.BufferImageCopy, regions)
    check(lib.kzo_copy_buffer_to_image(C.c_void_p(ob.ptr(buf)), buffer_bytes, C.byref(d), n, a))
    return mem.tobytes()


def oracle_image_to_buffer(spec, regions, buffer_bytes) -> bytes:
    lib = ob.load()
    mem = ob.random_bytes(_size(spec), 13)
    buf = ob.random_bytes(buffer_bytes, 14)
    d = _img(spec, ob.ptr(mem))
    a, n = util.arr(abi.BufferImageCopy, regions)
    check(lib.kzo_copy_image_to_buffer(C.byref(d), C.c_void_p(ob.ptr(buf)), buffer_bytes, n, a))
    return buf.tobytes()


def oracle_blit(src_spec, dst_spec, regions, filt) -> bytes:
    lib = ob.load()
    s = ob.random_bytes(_size(src_spec), 15)
    dmem = ob.random_bytes(_size(dst_spec), 16)
    sd, dd = _img(src_spec, ob.ptr(s)), _img(dst_spec, ob.ptr(dmem))
    a, n = util.arr(abi.ImageBlit, regions)
    check(lib.kzo_blit_image(C.byref(sd), C.byref(dd), n, a, filt))
    return dmem.tobytes()


def oracle_copy_image(src_spec, dst_spec, regions) -> bytes:
    lib = ob.load()
    s = ob.random_bytes(_size(src_spec), 17)
    dmem = ob.random_bytes(_size(dst_spec), 18)
    sd, dd = _img(src_spec, ob.ptr(s)), _img(dst_spec, ob.ptr(dmem))
    a, n = util.arr(abi.ImageCopy, regions)
    check(lib.kzo_copy_image(C.byref(sd), C.byref(dd), n, a))
    return dmem.tobytes()


def oracle_convert_tiling(spec, to_tiled: bool) -> bytes:
    lib = ob.load()
    s_spec = dict(spec, tiling=abi.TILING_LINEAR if to_tiled else abi.TILING_TILED)
    d_spec = dict(spec, tiling=abi.TILING_TILED if to_tiled else abi.TILING_LINEAR)
    s = ob.random_bytes(_size(s_spec), 19)
    dmem = ob.random_bytes(_size(d_spec), 20)
    sd, dd = _img(s_spec, ob.ptr(s)), _img(d_spec, ob.ptr(dmem))
    check(lib.kzo_convert_tiling(C.byref(sd), C.byref(dd)))
    return dmem.tobytes()


# ---- GPU -------------------------------------------------------------------------------------------------

def gpu_clear(ctx, spec, ranges, color=cases.CLEAR_COLOR) -> bytes:
    mem = ctx.buffer(_size(spec), ob.random_bytes(_size(spec), 11))
    ctx.queue.cmd_clear_color_image(_host_img(spec, mem.address), color, _ranges(ranges))
    out = mem.download().tobytes()
    ctx.release()
    return out


def gpu_buffer_to_image(ctx, spec, regions, buffer_bytes) -> bytes:
    mem = ctx.buffer(_size(spec), ob.random_bytes(_size(spec), 12))
    buf = ctx.buffer(buffer_bytes, ob.random_bytes(buffer_bytes, 0x4B415A414E))
    ctx.queue.cmd_copy_buffer_to_image(buf.address, buffer_bytes, _host_img(spec, mem.address), regions)
    out = mem.download().tobytes()
    ctx.release()
    return out


def gpu_image_to_buffer(ctx, spec, regions, buffer_bytes) -> bytes:
    mem = ctx.buffer(_size(spec), ob.random_bytes(_size(spec), 13))
    buf = ctx.buffer(buffer_bytes, ob.random_bytes(buffer_bytes, 14))
    ctx.queue.cmd_copy_image_to_buffer(_host_img(spec, mem.address), buf.address, buffer_bytes, regions)
    out = buf.download().tobytes()
    ctx.release()
    return out


def gpu_blit(ctx, src_spec, dst_spec, regions, filt) -> bytes:
    s = ctx.buffer(_size(src_spec), ob.random_bytes(_size(src_spec), 15))
    d = ctx.buffer(_size(dst_spec), ob.random_bytes(_size(dst_spec), 16))
    ctx.queue.cmd_blit_image(_host_img(src_spec, s.address), _host_img(dst_spec, d.address), regions, filt)
    out = d.download().tobytes()
    ctx.release()
    return out


def gpu_copy_image(ctx, src_spec, dst_spec, regions) -> bytes:
    s = ctx.buffer(_size(src_spec), ob.random_bytes(_size(src_spec), 17))
    d = ctx.buffer(_size(dst_spec), ob.random_bytes(_size(dst_spec), 18))
    ctx.queue.cmd_copy_image(_host_img(src_spec, s.address), _host_img(dst_spec, d.address), regions)
    out = d.download().tobytes()
    ctx.release()
    return out


def gpu_convert_tiling(ctx, spec, to_tiled: bool) -> bytes:
    s_spec = dict(spec, tiling=abi.TILING_LINEAR if to_tiled else abi.TILING_TILED)
    d_spec = dict(spec, tiling=abi.TILING_TILED if to_tiled else abi.TILING_LINEAR)
    s = ctx.buffer(_size(s_spec), ob.random_bytes(_size(s_spec), 19))
    d = ctx.buffer(_size(d_spec), ob.random_bytes(_size(d_spec), 20))
    ctx.queue.cmd_convert_tiling(_host_img(s_spec, s.address), _host_img(d_spec, d.address))
    out = d.download().tobytes()
    ctx.release()
    return out
