"""CPU: the transfer oracle against independent numpy restatements, known answers and its golden digests."""
import ctypes as C
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from kazan_b200 import abi
from oracle import binding as ob
from tests import cases, runners

GOLDEN = json.loads((Path(__file__).parent / "golden" / "oracle_golden.json").read_text())


def sha(b):
    return hashlib.sha256(b).hexdigest()


def test_unorm8_known_answers(oracle):
    # SURVEY.md E.1 worked check used by BASELINE config #1
    col = (C.c_float * 4)(0.25, 0.5, 0.75, 1.0)
    assert oracle.kzo_pack_rgba8(col) == 0xFFBF8040
    # RNE ties: 0.5/255 -> 0 (tie to even), 1.5/255 -> 2, 2.5/255 -> 2
    assert [oracle.kzo_float_to_unorm8(C.c_float(x / 255.0)) for x in (0.5, 1.5, 2.5)] == [0, 2, 2]
    assert oracle.kzo_float_to_unorm8(C.c_float(float("nan"))) == 0
    assert oracle.kzo_float_to_unorm8(C.c_float(-3.0)) == 0
    assert oracle.kzo_float_to_unorm8(C.c_float(float("inf"))) == 255
    for u in range(256):  # round trip is the identity
        assert oracle.kzo_float_to_unorm8(oracle.kzo_unorm8_to_float(u)) == u


def _numpy_view(mem: bytes, spec):
    f = cases.image_desc_fields(spec)
    n = f["width"] * f["height"] * f["depth"] * f["array_layers"]
    return np.frombuffer(mem, dtype=np.uint32)[:n].reshape(f["array_layers"], f["depth"], f["height"], f["width"])


@pytest.mark.parametrize("name,spec,regions,nbytes", [c for c in cases.BUFFER_IMAGE_CASES if c[1].get("tiling", 0) == 0],
                         ids=lambda c: c if isinstance(c, str) else "")
def test_buffer_to_image_matches_numpy(name, spec, regions, nbytes):
    """E.3 addressing restated with numpy slicing."""
    f = cases.image_desc_fields(spec)
    out = _numpy_view(runners.oracle_buffer_to_image(spec, regions, nbytes), spec)
    size = runners._size(spec)
    expect = _numpy_view(ob.random_bytes(size, 12).tobytes(), spec).copy()
    buf = np.frombuffer(ob.random_bytes(nbytes, 0x4B415A414E).tobytes(), dtype=np.uint32)
    for r in regions:
        w, h, d = r.imageExtent.width, r.imageExtent.height, r.imageExtent.depth
        rl = r.bufferRowLength or w
        ih = r.bufferImageHeight or h
        for l in range(r.imageSubresource.layerCount):
            for z in range(d):
                s = l * d + z
                for y in range(h):
                    o = r.bufferOffset // 4 + (s * ih + y) * rl
                    expect[r.imageSubresource.baseArrayLayer + l, r.imageOffset.z + z, r.imageOffset.y + y,
                           r.imageOffset.x:r.imageOffset.x + w] = buf[o:o + w]
    assert np.array_equal(out, expect)


def test_tiled_layout_formula(oracle):
    """E.10: 16x16 tiles, row-major texels in a tile, row-major tiles."""
    d = abi.ImageDesc(0, 37, abi.TILING_TILED, 50, 37, 1, 2, 1, 0)
    tiles_x, tiles_y = 4, 3
    for (x, y, l) in [(0, 0, 0), (15, 15, 0), (16, 0, 0), (49, 36, 0), (17, 18, 1), (0, 16, 1)]:
        expect = l * tiles_x * tiles_y * 1024 + (((y // 16) * tiles_x + x // 16) * 256 + (y % 16) * 16 + x % 16) * 4
        assert oracle.kzo_texel_offset(C.byref(d), x, y, 0, l) == expect
    lin = abi.ImageDesc(0, 37, 0, 50, 37, 3, 2, 1, 0)
    assert oracle.kzo_texel_offset(C.byref(lin), 7, 5, 2, 1) == 4 * (7 + 50 * (5 + 37 * (2 + 3 * 1)))


def test_tiling_round_trip():
    for name, spec in cases.CONVERT_TILING_CASES:
        f = cases.image_desc_fields(spec)
        lin_spec = dict(spec, tiling=0)
        src = ob.random_bytes(runners._size(lin_spec), 19)
        tiled = runners.oracle_convert_tiling(spec, True)
        # convert back by hand with the oracle
        t = np.frombuffer(tiled, dtype=np.uint8).copy()
        tmem = ob.aligned_empty(len(t)); tmem[:] = t
        back = ob.aligned_empty(len(src))
        sd = abi.ImageDesc(ob.ptr(tmem), f["format"], 1, f["width"], f["height"], f["depth"], f["array_layers"], 1, 0)
        dd = abi.ImageDesc(ob.ptr(back), f["format"], 0, f["width"], f["height"], f["depth"], f["array_layers"], 1, 0)
        assert ob.load().kzo_convert_tiling(C.byref(sd), C.byref(dd)) == 0
        n = 4 * f["width"] * f["height"] * f["depth"] * f["array_layers"]
        assert bytes(back[:n]) == bytes(src[:n]), name


def test_clear_known_bytes():
    out = runners.oracle_clear(dict(width=1024, height=1024), [(0, abi.REMAINING)])
    assert out == bytes([0x40, 0x80, 0xBF, 0xFF]) * (1024 * 1024)
    # untouched layers keep their content
    spec = dict(width=33, height=17, array_layers=4)
    out = np.frombuffer(runners.oracle_clear(spec, [(1, 2)]), dtype=np.uint32)
    before = np.frombuffer(ob.random_bytes(runners._size(spec), 11).tobytes(), dtype=np.uint32)
    n = 33 * 17
    assert np.array_equal(out[:n], before[:n]) and np.array_equal(out[3 * n:], before[3 * n:])
    assert np.all(out[n:3 * n] == 0xFFBF8040)


def test_blit_nearest_identity_and_integer_upscale():
    s = dict(width=64, height=48)
    name, src, dst, regions, filt = cases.blit_cases()[0]
    out = np.frombuffer(runners.oracle_blit(src, dst, regions, filt), dtype=np.uint32)[:64 * 48]
    srcmem = np.frombuffer(ob.random_bytes(runners._size(s), 15).tobytes(), dtype=np.uint32)[:64 * 48]
    assert np.array_equal(out, srcmem)
    name, src, dst, regions, filt = cases.blit_cases()[1]
    out = np.frombuffer(runners.oracle_blit(src, dst, regions, filt), dtype=np.uint32)[:64 * 48].reshape(48, 64)
    srcmem = np.frombuffer(ob.random_bytes(runners._size(src), 15).tobytes(), dtype=np.uint32)[:32 * 24].reshape(24, 32)
    assert np.array_equal(out, np.repeat(np.repeat(srcmem, 2, axis=0), 2, axis=1))


def test_blit_linear_against_float64_numpy():
    """LINEAR is an fp32 path: the oracle must be within 1 LSB of a float64 evaluation of the same filter."""
    name, src, dst, regions, filt = cases.blit_cases()[3]
    out = np.frombuffer(runners.oracle_blit(src, dst, regions, filt), dtype=np.uint8)[:64 * 48 * 4].reshape(48, 64, 4)
    s = np.frombuffer(ob.random_bytes(runners._size(src), 15).tobytes(), dtype=np.uint8)[:17 * 13 * 4].reshape(13, 17, 4) / 255.0
    j, i = np.mgrid[0:48, 0:64]
    u = (i + 0.5) * 17 / 64 - 0.5
    v = (j + 0.5) * 13 / 48 - 0.5
    x0, y0 = np.floor(u).astype(int), np.floor(v).astype(int)
    a, b = (u - x0)[..., None], (v - y0)[..., None]
    cx = lambda x: np.clip(x, 0, 16)
    cy = lambda y: np.clip(y, 0, 12)
    ref = ((1 - a) * (1 - b) * s[cy(y0), cx(x0)] + a * (1 - b) * s[cy(y0), cx(x0 + 1)] +
           (1 - a) * b * s[cy(y0 + 1), cx(x0)] + a * b * s[cy(y0 + 1), cx(x0 + 1)]) * 255.0
    assert np.max(np.abs(out.astype(np.float64) - ref)) <= 1.0 + 1e-6


def test_invalid_usage_is_rejected(oracle):
    mem = ob.aligned_empty(64 * 64 * 4)
    d = abi.ImageDesc(ob.ptr(mem), 37, 0, 64, 64, 1, 1, 1, 0)
    buf = ob.aligned_empty(64 * 64 * 4)
    V = abi.ERROR_VALIDATION_FAILED
    from kazan_b200.host import buffer_image_copy as bic
    def b2i(region, size=buf.nbytes):
        a, n = ob.array_of(abi.BufferImageCopy, [region])
        return oracle.kzo_copy_buffer_to_image(C.c_void_p(ob.ptr(buf)), size, C.byref(d), n, a)
    assert b2i(bic(64, 64)) == 0
    assert b2i(bic(64, 64, buffer_offset=2)) == V          # VUID 00194
    assert b2i(bic(64, 64, row_length=32)) == V            # VUID 00195
    assert b2i(bic(64, 64, image_height=8)) == V           # VUID 00196
    assert b2i(bic(64, 64, offset=(1, 0, 0))) == V         # VUID 00197
    assert b2i(bic(64, 64), size=64 * 64 * 4 - 4) == V     # VUID 00171
    assert b2i(bic(64, 64, aspect=abi.ASPECT_DEPTH)) == V  # VUID 00212
    assert b2i(bic(64, 64, base_layer=1)) == V
    # bufferRowLength x bufferImageHeight x slices past 2^64: the needed size must not wrap around and look small
    mem2 = ob.aligned_empty(16 * 16 * 4 * 2)
    d2 = abi.ImageDesc(ob.ptr(mem2), 37, 0, 16, 16, 1, 2, 1, 0)
    a, n = ob.array_of(abi.BufferImageCopy, [bic(16, 16, row_length=2**32 - 1, image_height=2**32 - 1, layer_count=2)])
    assert oracle.kzo_copy_buffer_to_image(C.c_void_p(ob.ptr(buf)), buf.nbytes, C.byref(d2), n, a) == V
    c = abi.ClearColorValue()
    a, n = ob.array_of(abi.ImageSubresourceRange, [abi.ImageSubresourceRange(abi.ASPECT_DEPTH, 0, 1, 0, 1)])
    assert oracle.kzo_clear_color_image(C.byref(d), C.byref(c), n, a) == V  # VUID 02498
    dd = abi.ImageDesc(ob.ptr(mem), abi.FORMAT_D32_SFLOAT, 0, 64, 64, 1, 1, 1, 0)
    a, n = ob.array_of(abi.ImageSubresourceRange, [abi.ImageSubresourceRange(abi.ASPECT_COLOR, 0, 1, 0, 1)])
    assert oracle.kzo_clear_color_image(C.byref(dd), C.byref(c), n, a) == abi.ERROR_FORMAT_NOT_SUPPORTED  # VUID 00007


def test_buffer_commands(oracle):
    src = ob.random_bytes(4096, 1)
    dst = ob.random_bytes(4096, 2)
    before = dst.copy()
    a, n = ob.array_of(abi.BufferCopy, [abi.BufferCopy(3, 1001, 777), abi.BufferCopy(2048, 0, 512)])
    assert oracle.kzo_copy_buffer(C.c_void_p(ob.ptr(src)), 4096, C.c_void_p(ob.ptr(dst)), 4096, n, a) == 0
    exp = before.copy(); exp[1001:1778] = src[3:780]; exp[0:512] = src[2048:2560]
    assert np.array_equal(dst, exp)
    assert oracle.kzo_fill_buffer(C.c_void_p(ob.ptr(dst)), 4094, 8, abi.WHOLE_SIZE, 0xDEADBEEF) == 0
    w = dst.view(np.uint32)
    assert np.all(w[2:1023] == 0xDEADBEEF) and w[1] != 0xDEADBEEF and dst[4092:4096].tobytes() == exp[4092:4096].tobytes()
    assert oracle.kzo_fill_buffer(C.c_void_p(ob.ptr(dst)), 4096, 6, 8, 0) == abi.ERROR_VALIDATION_FAILED
    data = bytes(range(64))
    assert oracle.kzo_update_buffer(C.c_void_p(ob.ptr(dst)), 4096, 128, 64, data) == 0
    assert dst[128:192].tobytes() == data
    assert oracle.kzo_update_buffer(C.c_void_p(ob.ptr(dst)), 4096, 128, 65540, data) == abi.ERROR_VALIDATION_FAILED


def test_oracle_matches_golden_digests():
    from tests.golden.make_golden import generate
    now = generate()
    assert now == GOLDEN, [k for k in GOLDEN if now.get(k) != GOLDEN[k]]


def test_threads_do_not_change_results(oracle):
    spec, regions, nbytes = cases.BUFFER_IMAGE_CASES[2][1:]
    oracle.kzo_set_threads(1)
    a = runners.oracle_buffer_to_image(spec, regions, nbytes)
    oracle.kzo_set_threads(4)
    b = runners.oracle_buffer_to_image(spec, regions, nbytes)
    oracle.kzo_set_threads(1)
    assert a == b


def test_image_sizes_that_overflow_are_rejected(oracle):
    """The reference computes an image's byte size with checked_mul().unwrap() (image.rs:74-82): a size that does not fit is a
    panic there, an error here — in the oracle and (tests/test_abi.py) in the library alike, LINEAR and TILED."""
    for w, h, d, layers, tiling in [(2**32 - 1, 2**32 - 1, 1, 1, 0), (2**31, 2**31, 4, 4, 0), (2**32 - 1, 2**32 - 1, 1, 1, 1),
                                    (2**32 - 1, 2**28, 1, 1, 1)]:
        desc = abi.ImageDesc(0x1000, abi.FORMAT_R8G8B8A8_UNORM, tiling, w, h, d, layers, 1, 0)
        size, align = C.c_uint64(), C.c_uint64()
        assert oracle.kzo_image_memory_requirements(C.byref(desc), C.byref(size), C.byref(align)) == abi.ERROR_VALIDATION_FAILED
    desc = abi.ImageDesc(0x1000, abi.FORMAT_R8G8B8A8_UNORM, 0, 2**31, 1, 1, 1, 1, 0)  # small, but wider than any VkOffset3D
    size, align = C.c_uint64(), C.c_uint64()
    assert oracle.kzo_image_memory_requirements(C.byref(desc), C.byref(size), C.byref(align)) == abi.ERROR_VALIDATION_FAILED
    desc = abi.ImageDesc(0x1000, abi.FORMAT_R8G8B8A8_UNORM, 0, 65536, 65536, 256, 256, 1, 0)  # 2^50 bytes: large but representable
    size, align = C.c_uint64(), C.c_uint64()
    assert oracle.kzo_image_memory_requirements(C.byref(desc), C.byref(size), C.byref(align)) == 0 and size.value == 1 << 50


def test_division_free_unorm8_to_float_is_exact():
    """transfer.cu `unorm8_exact`: q0 = u * fl(1/255); q = fma(fma(-q0, 255, u), fl(1/255), q0) equals the correctly rounded
    u / 255.0f (E.1) for every byte — checked here in exact rational arithmetic with explicit roundings to fp32."""
    from fractions import Fraction
    f = np.float32

    def rnd(fr):
        x = f(float(fr))
        cands = [np.nextafter(x, f(-np.inf)), x, np.nextafter(x, f(np.inf))]
        best = min(cands, key=lambda t: abs(Fraction(float(t)) - fr))
        ties = [t for t in cands if abs(Fraction(float(t)) - fr) == abs(Fraction(float(best)) - fr)]
        if len(ties) > 1:  # halfway: to the even significand
            best = next(t for t in ties if (int(np.float32(t).view(np.uint32)) & 1) == 0)
        return best

    rcp = f(1.0) / f(255.0)
    assert float(rcp) == 0.003921568859368563
    for u in range(256):
        q0 = rnd(Fraction(u) * Fraction(float(rcp)))
        rem = rnd(Fraction(u) - Fraction(float(q0)) * 255)
        assert Fraction(float(rem)) == Fraction(u) - Fraction(float(q0)) * 255, "the remainder is exact"
        q = rnd(Fraction(float(q0)) + Fraction(float(rem)) * Fraction(float(rcp)))
        assert q == f(u) / f(255.0), u
