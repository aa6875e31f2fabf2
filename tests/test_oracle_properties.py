"""CPU: size-independent properties of the transfer oracle on random legal regions (LINEAR and TILED, layers, 3-D, padded
pitches): what vkCmdCopyImageToBuffer reads back is exactly what vkCmdCopyBufferToImage wrote, texel for texel and at the
same buffer positions, and a LINEAR -> TILED -> LINEAR conversion is the identity."""
import ctypes as C

import numpy as np
import pytest

from kazan_b200 import abi
from oracle import binding as ob
from tests import cases, runners, util
from tests.test_gpu_transfer import _random_region

SHAPES = [dict(width=61, height=37), dict(width=128, height=24, array_layers=3), dict(width=20, height=12, depth=5),
          dict(width=70, height=41, tiling=abi.TILING_TILED, array_layers=2), dict(width=33, height=18, depth=3, tiling=abi.TILING_TILED)]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_image_to_buffer_inverts_buffer_to_image(oracle, seed):
    rng = np.random.default_rng(seed)
    checked = 0
    for k in range(120):
        spec = SHAPES[k % len(SHAPES)]
        region = _random_region(rng, spec)
        nbytes = 64 + 4 * (spec["width"] + 8) * (spec["height"] + 4) * spec.get("depth", 1) * spec.get("array_layers", 1)
        size = runners._size(spec)
        original = ob.random_bytes(size, 10 * k + seed)
        buf_a = ob.random_bytes(nbytes, 7000 + k)
        a, n = util.arr(abi.BufferImageCopy, [region])
        image = original.copy()
        d = runners._img(spec, ob.ptr(image))
        if oracle.kzo_copy_buffer_to_image(C.c_void_p(ob.ptr(buf_a)), nbytes, C.byref(d), n, a) != abi.SUCCESS:
            continue
        buf_b = ob.aligned_empty(nbytes)
        buf_b[:] = 0xA5
        assert oracle.kzo_copy_image_to_buffer(C.byref(d), C.c_void_p(ob.ptr(buf_b)), nbytes, n, a) == abi.SUCCESS
        # every byte the read-back touched equals the source buffer at the same position; everything else is untouched
        touched = buf_b != 0xA5
        assert np.array_equal(buf_b[touched], buf_a[touched])
        texels = region.imageExtent.width * region.imageExtent.height * region.imageExtent.depth * region.imageSubresource.layerCount
        assert touched.sum() <= 4 * texels  # (a source byte may happen to be 0xA5)
        # writing the read-back into the original image reproduces the image after the first copy
        again = original.copy()
        d2 = runners._img(spec, ob.ptr(again))
        assert oracle.kzo_copy_buffer_to_image(C.c_void_p(ob.ptr(buf_b)), nbytes, C.byref(d2), n, a) == abi.SUCCESS
        assert np.array_equal(again, image)
        checked += 1
    assert checked > 70


@pytest.mark.parametrize("shape", [dict(width=50, height=37), dict(width=128, height=64, array_layers=2), dict(width=17, height=5, depth=4)])
def test_tiling_conversion_round_trip_is_the_identity(oracle, shape):
    lin = dict(shape, tiling=abi.TILING_LINEAR)
    til = dict(shape, tiling=abi.TILING_TILED)
    a = ob.random_bytes(runners._size(lin), 5)
    t = ob.aligned_empty(runners._size(til))
    t[:] = 0
    b = ob.aligned_empty(runners._size(lin))
    b[:] = 0
    da, dt, db = runners._img(lin, ob.ptr(a)), runners._img(til, ob.ptr(t)), runners._img(lin, ob.ptr(b))
    assert oracle.kzo_convert_tiling(C.byref(da), C.byref(dt)) == abi.SUCCESS
    assert oracle.kzo_convert_tiling(C.byref(dt), C.byref(db)) == abi.SUCCESS
    f = cases.image_desc_fields(lin)
    used = 4 * f["width"] * f["height"] * f["depth"] * f["array_layers"]
    assert np.array_equal(a[:used], b[:used])
