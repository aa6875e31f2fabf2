"""GPU parity: the CUDA transfer kernels, called through the C ABI, against the CPU oracle — bit exact."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from kazan_b200 import abi, host
from kazan_b200.host import buffer_image_copy as bic
from oracle import binding as ob
from tests import cases, runners, util

pytestmark = pytest.mark.gpu
GOLDEN = json.loads((Path(__file__).parent / "golden" / "oracle_golden.json").read_text())


def sha(b):
    return hashlib.sha256(b).hexdigest()


def same(gpu_bytes, oracle_bytes, key):
    assert gpu_bytes == oracle_bytes, f"{key}: CUDA result differs from the oracle"
    assert sha(gpu_bytes) == GOLDEN[key], f"{key}: result differs from the committed golden digest"


@pytest.mark.parametrize("case", cases.CLEAR_CASES, ids=lambda c: c[0])
def test_clear_color_image(gpu, case):
    name, spec, ranges = case
    same(runners.gpu_clear(gpu, spec, ranges), runners.oracle_clear(spec, ranges), f"clear/{name}")


@pytest.mark.parametrize("case", cases.BUFFER_IMAGE_CASES, ids=lambda c: c[0])
def test_copy_buffer_to_image(gpu, case):
    name, spec, regions, nbytes = case
    same(runners.gpu_buffer_to_image(gpu, spec, regions, nbytes), runners.oracle_buffer_to_image(spec, regions, nbytes), f"b2i/{name}")


@pytest.mark.parametrize("case", cases.BUFFER_IMAGE_CASES, ids=lambda c: c[0])
def test_copy_image_to_buffer(gpu, case):
    name, spec, regions, nbytes = case
    same(runners.gpu_image_to_buffer(gpu, spec, regions, nbytes), runners.oracle_image_to_buffer(spec, regions, nbytes), f"i2b/{name}")


@pytest.mark.parametrize("case", cases.blit_cases(), ids=lambda c: c[0])
def test_blit_image(gpu, case):
    name, s, d, regions, filt = case
    # NEAREST is an integer path; LINEAR is fp32 with the oracle's operation order and is bit exact as well
    same(runners.gpu_blit(gpu, s, d, regions, filt), runners.oracle_blit(s, d, regions, filt), f"blit/{name}")


@pytest.mark.parametrize("case", cases.copy_image_cases(), ids=lambda c: c[0])
def test_copy_image(gpu, case):
    name, s, d, regions = case
    same(runners.gpu_copy_image(gpu, s, d, regions), runners.oracle_copy_image(s, d, regions), f"copy_image/{name}")


@pytest.mark.parametrize("case", cases.CONVERT_TILING_CASES, ids=lambda c: c[0])
def test_convert_tiling(gpu, case):
    name, spec = case
    same(runners.gpu_convert_tiling(gpu, spec, True), runners.oracle_convert_tiling(spec, True), f"to_tiled/{name}")
    same(runners.gpu_convert_tiling(gpu, spec, False), runners.oracle_convert_tiling(spec, False), f"to_linear/{name}")


def test_depth_clear_and_attachment_clear(gpu, oracle):
    import ctypes as C
    spec = dict(width=70, height=33, array_layers=3, format=abi.FORMAT_D32_SFLOAT)
    size = runners._size(spec)
    init = ob.random_bytes(size, 31)
    mem = gpu.buffer(size, init)
    img = runners._host_img(spec, mem.address)
    gpu.queue.cmd_clear_depth_stencil_image(img, 0.625, [abi.ImageSubresourceRange(abi.ASPECT_DEPTH, 0, 1, 1, 2)])
    rects = [abi.ClearRect(abi.Rect2D(abi.Offset2D(3, 4), abi.Extent2D(50, 20)), 0, 1),
             abi.ClearRect(abi.Rect2D(abi.Offset2D(64, 0), abi.Extent2D(6, 33)), 2, 1)]
    gpu.queue.cmd_clear_attachment(img, [0.25], rects)
    out = mem.download().tobytes()
    gpu.release()
    ref = init.copy()
    d = runners._img(spec, ob.ptr(ref))
    v = abi.ClearDepthStencilValue(0.625, 0)
    a, n = ob.array_of(abi.ImageSubresourceRange, [abi.ImageSubresourceRange(abi.ASPECT_DEPTH, 0, 1, 1, 2)])
    assert oracle.kzo_clear_depth_stencil_image(C.byref(d), C.byref(v), n, a) == 0
    c = abi.ClearColorValue(); c.float32[0] = 0.25
    a, n = ob.array_of(abi.ClearRect, rects)
    assert oracle.kzo_clear_attachment(C.byref(d), C.byref(c), n, a) == 0
    assert out == ref.tobytes()


def test_buffer_commands(gpu, oracle):
    import ctypes as C
    src_h, dst_h = ob.random_bytes(1 << 20, 1), ob.random_bytes(1 << 20, 2)
    src, dst = gpu.buffer(1 << 20, src_h), gpu.buffer(1 << 20, dst_h)
    regions = [abi.BufferCopy(3, 1001, 777), abi.BufferCopy(4096, 65536, 262144), abi.BufferCopy(8, 400000, 40),
               abi.BufferCopy(500004, 600004, 4000)]
    gpu.queue.cmd_copy_buffer(src.address, 1 << 20, dst.address, 1 << 20, regions)
    gpu.queue.cmd_fill_buffer(dst.address, (1 << 20) - 2, 900000, abi.WHOLE_SIZE, 0xDEADBEEF)
    gpu.queue.cmd_fill_buffer(dst.address, 1 << 20, 12, 20, 0x01020304)
    gpu.queue.cmd_update_buffer(dst.address, 1 << 20, 128, bytes(range(200)))
    out = dst.download().tobytes()
    gpu.release()
    ref = dst_h.copy()
    a, n = ob.array_of(abi.BufferCopy, regions)
    assert oracle.kzo_copy_buffer(C.c_void_p(ob.ptr(src_h)), 1 << 20, C.c_void_p(ob.ptr(ref)), 1 << 20, n, a) == 0
    assert oracle.kzo_fill_buffer(C.c_void_p(ob.ptr(ref)), (1 << 20) - 2, 900000, abi.WHOLE_SIZE, 0xDEADBEEF) == 0
    assert oracle.kzo_fill_buffer(C.c_void_p(ob.ptr(ref)), 1 << 20, 12, 20, 0x01020304) == 0
    assert oracle.kzo_update_buffer(C.c_void_p(ob.ptr(ref)), 1 << 20, 128, 200, bytes(range(200))) == 0
    assert out == ref.tobytes()


def test_invalid_usage_returns_the_oracles_error_codes(gpu):
    mem = gpu.buffer(64 * 64 * 4)
    buf = gpu.buffer(64 * 64 * 4)
    img = host.Image(mem.address, 64, 64)
    V = abi.ERROR_VALIDATION_FAILED

    def code(fn):
        with pytest.raises(host.KazanError) as e:
            fn()
        return e.value.code

    q = gpu.queue
    assert code(lambda: q.cmd_copy_buffer_to_image(buf.address, buf.nbytes, img, [bic(64, 64, buffer_offset=2)])) == V
    assert code(lambda: q.cmd_copy_buffer_to_image(buf.address, buf.nbytes, img, [bic(64, 64, row_length=32)])) == V
    assert code(lambda: q.cmd_copy_buffer_to_image(buf.address, buf.nbytes, img, [bic(64, 64, image_height=8)])) == V
    assert code(lambda: q.cmd_copy_buffer_to_image(buf.address, buf.nbytes, img, [bic(64, 64, offset=(1, 0, 0))])) == V
    assert code(lambda: q.cmd_copy_buffer_to_image(buf.address, buf.nbytes - 4, img, [bic(64, 64)])) == V
    assert code(lambda: q.cmd_copy_image_to_buffer(img, buf.address, buf.nbytes, [bic(64, 64, aspect=abi.ASPECT_DEPTH)])) == V
    assert code(lambda: q.cmd_clear_color_image(img, (0, 0, 0, 0), [abi.ImageSubresourceRange(abi.ASPECT_DEPTH, 0, 1, 0, 1)])) == V
    dimg = host.Image(mem.address, 64, 64, format=abi.FORMAT_D32_SFLOAT)
    assert code(lambda: q.cmd_clear_color_image(dimg, (0, 0, 0, 0))) == abi.ERROR_FORMAT_NOT_SUPPORTED
    assert code(lambda: q.cmd_fill_buffer(buf.address, buf.nbytes, 6, 8, 0)) == V
    assert code(lambda: q.cmd_copy_buffer_to_image(buf.address, buf.nbytes, img, [])) == V
    # bufferRowLength x bufferImageHeight x slices past 2^64 must not wrap around and look small
    img2 = host.Image(mem.address, 16, 16, array_layers=2)
    huge = bic(16, 16, row_length=2**32 - 1, image_height=2**32 - 1, layer_count=2)
    assert code(lambda: q.cmd_copy_buffer_to_image(buf.address, buf.nbytes, img2, [huge])) == V
    assert code(lambda: q.cmd_copy_image_to_buffer(img2, buf.address, buf.nbytes, [huge])) == V
    gpu.release()


def test_full_size_properties_4k_and_8k(gpu):
    """BASELINE sizes, through size-independent properties (no oracle pass over 132 MB needed):
    clear -> every word equals the packed colour; copy b2i -> i2b round trip is the identity; checksum of
    checksums of a padded-pitch copy equals numpy's."""
    for (w, h) in [(3840, 2160), (7680, 4320)]:
        n = w * h * 4
        img_mem, buf_a, buf_b = gpu.buffer(n), gpu.buffer(n), gpu.buffer(n)
        img = host.Image(img_mem.address, w, h)
        gpu.queue.cmd_clear_color_image(img, cases.CLEAR_COLOR)
        words = img_mem.download().view(np.uint32)
        assert words.min() == words.max() == 0xFFBF8040
        data = ob.random_bytes(n, w)
        buf_a.upload(data)
        gpu.queue.cmd_copy_buffer_to_image(buf_a.address, n, img, [bic(w, h)])
        gpu.queue.cmd_copy_image_to_buffer(img, buf_b.address, n, [bic(w, h)])
        assert np.array_equal(buf_b.download(), data)
        # sub-rectangle with padded pitch into a cleared image: digest against numpy slicing
        gpu.queue.cmd_clear_color_image(img, (0, 0, 0, 0))
        rw, rh, pitch = w - 48, h - 40, w - 16
        gpu.queue.cmd_copy_buffer_to_image(buf_a.address, n, img, [bic(rw, rh, row_length=pitch, offset=(16, 8, 0))])
        got = img_mem.download().view(np.uint32).reshape(h, w)
        exp = np.zeros((h, w), dtype=np.uint32)
        src = data.view(np.uint32)[:pitch * rh].reshape(rh, pitch)[:, :rw]
        exp[8:8 + rh, 16:16 + rw] = src
        assert sha(got.tobytes()) == sha(exp.tobytes())
        gpu.release()


def test_tma_and_ldst_paths_agree(gpu):
    """KZN_TRANSFER_PATH picks the engine for contiguous runs; results must be identical."""
    import os
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, '.');"
        "from tests.util import GpuContext; from tests import runners, cases; import hashlib;"
        "g = GpuContext();"
        "n, spec, regions, nb = cases.BUFFER_IMAGE_CASES[0];"
        "print(hashlib.sha256(runners.gpu_buffer_to_image(g, spec, regions, nb)).hexdigest(),"
        "      hashlib.sha256(runners.gpu_clear(g, *cases.CLEAR_CASES[0][1:])).hexdigest())")
    outs = []
    for path in ("ldst", "tma"):
        env = dict(os.environ, KZN_TRANSFER_PATH=path)
        outs.append(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True,
                                   cwd=str(Path(__file__).parent.parent)).stdout.split())
    assert outs[0] == outs[1] == [GOLDEN["b2i/cfg1_1024"], GOLDEN["clear/cfg1_1024"]]


def _random_region(rng, spec):
    """A vkCmdCopyBufferToImage region that is usually legal and sometimes breaks exactly one rule."""
    f = cases.image_desc_fields(spec)
    w, h, d, layers = f["width"], f["height"], f["depth"], f["array_layers"]
    ew, eh = int(rng.integers(1, w + 1)), int(rng.integers(1, h + 1))
    ox, oy = int(rng.integers(0, w - ew + 1)), int(rng.integers(0, h - eh + 1))
    ed = int(rng.integers(1, d + 1))
    oz = int(rng.integers(0, d - ed + 1))
    base = int(rng.integers(0, layers))
    count = int(rng.integers(1, layers - base + 1))
    row_length = 0 if rng.random() < 0.5 else ew + int(rng.integers(0, 9))
    image_height = 0 if rng.random() < 0.5 else eh + int(rng.integers(0, 5))
    offset = 4 * int(rng.integers(0, 9))
    flaw = rng.random()
    if flaw < 0.06:
        ox = w - ew + 1 + int(rng.integers(0, 3))          # extent sticks out of the image
    elif flaw < 0.12:
        base = layers                                      # layer out of range
    elif flaw < 0.18 and row_length:
        row_length = max(ew - 1, 0)                        # bufferRowLength smaller than the extent (0 means tight)
    elif flaw < 0.24:
        offset += 2                                        # bufferOffset not a multiple of the texel size
    elif flaw < 0.30:
        ew = 0                                             # empty extent: legal, copies nothing
    return bic(ew, eh, depth=ed, buffer_offset=offset, row_length=row_length, image_height=image_height,
                             offset=(ox, oy, oz), base_layer=base, layer_count=count)


import os  # noqa: E402


@pytest.mark.parametrize("seed", [int(x) for x in os.environ.get("KZN_FUZZ_SEEDS", "20261018").split(",")])
def test_random_regions_agree_with_the_oracle_in_result_code_and_bytes(gpu, seed):
    """Differential test on 150 random regions over four image shapes: the GPU path and the oracle must return the same
    VkResult, and when they accept the region, the same bytes (buffer -> image and back)."""
    import ctypes as C
    rng = np.random.default_rng(seed)
    lib = ob.load()
    shapes = [dict(width=61, height=37), dict(width=128, height=24, array_layers=3), dict(width=20, height=12, depth=5),
              dict(width=70, height=41, tiling=abi.TILING_TILED, array_layers=2)]
    accepted = rejected = 0
    n_regions = int(os.environ.get("KZN_FUZZ_REGIONS", "150"))
    for k in range(n_regions):
        spec = shapes[k % len(shapes)]
        region = _random_region(rng, spec)
        buffer_bytes = 64 + 4 * (spec["width"] + 8) * (spec["height"] + 4) * spec.get("depth", 1) * spec.get("array_layers", 1)
        size = runners._size(spec)
        img0, buf0 = ob.random_bytes(size, 100 + k), ob.random_bytes(buffer_bytes, 500 + k)
        # oracle
        o_img, o_buf = img0.copy(), buf0.copy()
        d = runners._img(spec, ob.ptr(o_img))
        a, n = util.arr(abi.BufferImageCopy, [region])
        code_in = lib.kzo_copy_buffer_to_image(C.c_void_p(ob.ptr(o_buf)), buffer_bytes, C.byref(d), n, a)
        code_out = lib.kzo_copy_image_to_buffer(C.byref(d), C.c_void_p(ob.ptr(o_buf)), buffer_bytes, n, a)
        # GPU
        g_img, g_buf = gpu.buffer(size, img0), gpu.buffer(buffer_bytes, buf0)
        himg = runners._host_img(spec, g_img.address)
        got_in = got_out = abi.SUCCESS
        try:
            gpu.queue.cmd_copy_buffer_to_image(g_buf.address, buffer_bytes, himg, [region])
        except host.KazanError as e:
            got_in = e.code
        try:
            gpu.queue.cmd_copy_image_to_buffer(himg, g_buf.address, buffer_bytes, [region])
        except host.KazanError as e:
            got_out = e.code
        assert (got_in, got_out) == (code_in, code_out), (k, spec, region.bufferOffset, region.bufferRowLength)
        if code_in == abi.SUCCESS:
            accepted += 1
            assert g_img.download().tobytes() == o_img.tobytes(), (k, "image")
            assert g_buf.download().tobytes() == o_buf.tobytes(), (k, "buffer")
        else:
            rejected += 1
        gpu.release()
    assert accepted > 0.6 * n_regions and rejected > 0.1 * n_regions


@pytest.mark.parametrize("seed", [int(x) for x in os.environ.get("KZN_FUZZ_SEEDS", "20261018").split(",")])
def test_random_image_copies_blits_and_rect_clears_agree_with_the_oracle(gpu, seed):
    """Differential test of vkCmdCopyImage, vkCmdBlitImage (NEAREST and LINEAR, flips and scales included) and
    vkCmdClearAttachments on random boxes between LINEAR / TILED images of different shapes: result code and bytes."""
    import ctypes as C
    rng = np.random.default_rng(seed)
    lib = ob.load()
    shapes = [dict(width=45, height=33), dict(width=64, height=20, array_layers=2), dict(width=18, height=10, depth=4),
              dict(width=50, height=37, tiling=abi.TILING_TILED), dict(width=96, height=16, tiling=abi.TILING_TILED, array_layers=2)]
    n_cases = int(os.environ.get("KZN_FUZZ_REGIONS", "150"))
    done = {"copy": 0, "blit": 0, "clear": 0, "rejected": 0}

    def box(spec_f, rng):
        w, h, d = spec_f["width"], spec_f["height"], spec_f["depth"]
        ew, eh, ed = int(rng.integers(1, w + 1)), int(rng.integers(1, h + 1)), int(rng.integers(1, d + 1))
        return (int(rng.integers(0, w - ew + 1)), int(rng.integers(0, h - eh + 1)), int(rng.integers(0, d - ed + 1))), (ew, eh, ed)

    for k in range(n_cases):
        s_spec, d_spec = shapes[int(rng.integers(0, len(shapes)))], shapes[int(rng.integers(0, len(shapes)))]
        sf, df = cases.image_desc_fields(s_spec), cases.image_desc_fields(d_spec)
        s0, d0 = ob.random_bytes(runners._size(s_spec), 900 + k), ob.random_bytes(runners._size(d_spec), 1900 + k)
        o_s, o_d = s0.copy(), d0.copy()
        g_s, g_d = gpu.buffer(s0.nbytes, s0), gpu.buffer(d0.nbytes, d0)
        osd, odd = runners._img(s_spec, ob.ptr(o_s)), runners._img(d_spec, ob.ptr(o_d))
        hs, hd = runners._host_img(s_spec, g_s.address), runners._host_img(d_spec, g_d.address)
        kind = ("copy", "blit", "clear")[k % 3]
        sl = abi.ImageSubresourceLayers(abi.ASPECT_COLOR, 0, int(rng.integers(0, sf["array_layers"])), 1)
        dl = abi.ImageSubresourceLayers(abi.ASPECT_COLOR, 0, int(rng.integers(0, df["array_layers"] + (rng.random() < 0.1))), 1)
        if kind == "copy":
            (sx, sy, sz), (ew, eh, ed) = box(sf, rng)
            dx, dy, dz = int(rng.integers(0, df["width"])), int(rng.integers(0, df["height"])), int(rng.integers(0, df["depth"]))
            if rng.random() < 0.75:  # usually make the destination box fit
                ew, eh, ed = min(ew, df["width"] - dx), min(eh, df["height"] - dy), min(ed, df["depth"] - dz)
            reg = abi.ImageCopy(sl, abi.Offset3D(sx, sy, sz), dl, abi.Offset3D(dx, dy, dz), abi.Extent3D(ew, eh, ed))
            a, n = util.arr(abi.ImageCopy, [reg])
            want = lib.kzo_copy_image(C.byref(osd), C.byref(odd), n, a)
            run = lambda: gpu.queue.cmd_copy_image(hs, hd, [reg])
        elif kind == "blit":
            (sx, sy, sz), (sw, sh, sd_) = box(sf, rng)
            (dx, dy, dz), (dw, dh, dd_) = box(df, rng)
            so = [abi.Offset3D(sx, sy, sz), abi.Offset3D(sx + sw, sy + sh, sz + sd_)]
            do = [abi.Offset3D(dx, dy, dz), abi.Offset3D(dx + dw, dy + dh, dz + dd_)]
            if rng.random() < 0.3:
                so[0].x, so[1].x = so[1].x, so[0].x  # mirrored source
            if rng.random() < 0.3:
                do[0].y, do[1].y = do[1].y, do[0].y  # mirrored destination
            reg = abi.ImageBlit(sl, (abi.Offset3D * 2)(*so), dl, (abi.Offset3D * 2)(*do))
            filt = int(rng.integers(0, 2))
            a, n = util.arr(abi.ImageBlit, [reg])
            want = lib.kzo_blit_image(C.byref(osd), C.byref(odd), n, a, filt)
            run = lambda: gpu.queue.cmd_blit_image(hs, hd, [reg], filt)
        else:
            if df["depth"] != 1:
                d_spec = shapes[0]
                df = cases.image_desc_fields(d_spec)
                d0 = ob.random_bytes(runners._size(d_spec), 1900 + k)
                o_d = d0.copy()
                g_d = gpu.buffer(d0.nbytes, d0)
                odd, hd = runners._img(d_spec, ob.ptr(o_d)), runners._host_img(d_spec, g_d.address)
            (dx, dy, _), (dw, dh, _) = box(df, rng)
            if rng.random() < 0.1:
                dw = df["width"] - dx + 1  # sticks out
            rect = abi.ClearRect(abi.Rect2D(abi.Offset2D(dx, dy), abi.Extent2D(dw, dh)), int(rng.integers(0, df["array_layers"])), 1)
            value = tuple(float(v) for v in rng.uniform(-0.2, 1.2, 4))
            cv = abi.ClearColorValue()
            cv.float32[:] = list(value)
            a, n = util.arr(abi.ClearRect, [rect])
            want = lib.kzo_clear_attachment(C.byref(odd), C.byref(cv), n, a)
            run = lambda: gpu.queue.cmd_clear_attachment(hd, value, [rect])
        got = abi.SUCCESS
        try:
            run()
        except host.KazanError as e:
            got = e.code
        assert got == want, (k, kind, s_spec, d_spec)
        if want == abi.SUCCESS:
            done[kind] += 1
            assert g_d.download().tobytes() == o_d.tobytes(), (k, kind, s_spec, d_spec)
        else:
            done["rejected"] += 1
        gpu.release()
    assert min(done["copy"], done["blit"], done["clear"]) > 0.15 * n_cases and done["rejected"] > 0


@pytest.mark.parametrize("seed", [int(x) for x in os.environ.get("KZN_FUZZ_SEEDS", "20261018").split(",")])
def test_random_buffer_commands_agree_with_the_oracle(gpu, oracle, seed):
    """Differential test of vkCmdCopyBuffer (byte-granular and aligned regions, every alignment of source and destination
    modulo 16), vkCmdFillBuffer and vkCmdUpdateBuffer in random sequences on one pair of buffers, immediately and
    recorded into a command buffer (where disjoint aligned commands merge): result codes and final bytes."""
    import ctypes as C
    rng = np.random.default_rng(seed)
    size = 1 << 18
    for recorded in (False, True):
        src_h, dst_h = ob.random_bytes(size, 10 + seed), ob.random_bytes(size, 20 + seed)
        ref = dst_h.copy()
        src, dst = gpu.buffer(size, src_h), gpu.buffer(size, dst_h)
        q = gpu.queue
        ops = []
        for k in range(60):
            kind = rng.random()
            if kind < 0.5:
                n = int(rng.choice([1, 3, 16, 100, 4096, 65536, 70001]))
                if rng.random() < 0.6:
                    n = max(16, n // 16 * 16)
                so, do = int(rng.integers(0, size - n + 1)), int(rng.integers(0, size - n + 1))
                if rng.random() < 0.6:
                    so, do = so // 4 * 4, do // 16 * 16
                if rng.random() < 0.05:
                    n = size  # does not fit at this offset
                region = abi.BufferCopy(so, do, n)
                a, cnt = ob.array_of(abi.BufferCopy, [region])
                want = oracle.kzo_copy_buffer(C.c_void_p(ob.ptr(src_h)), size, C.c_void_p(ob.ptr(ref)), size, cnt, a)
                ops.append((want, lambda region=region: q.cmd_copy_buffer(src.address, size, dst.address, size, [region])))
            elif kind < 0.8:
                off = 4 * int(rng.integers(0, size // 4))
                n = abi.WHOLE_SIZE if rng.random() < 0.1 else 4 * int(rng.integers(1, (size - off) // 4 + 1))
                if rng.random() < 0.05:
                    off += 2  # misaligned offset: rejected
                word = int(rng.integers(0, 2**32))
                want = oracle.kzo_fill_buffer(C.c_void_p(ob.ptr(ref)), size, off, n, word)
                ops.append((want, lambda off=off, n=n, word=word: q.cmd_fill_buffer(dst.address, size, off, n, word)))
            else:
                n = 4 * int(rng.integers(1, 200))
                off = 4 * int(rng.integers(0, (size - n) // 4))
                data = bytes(rng.integers(0, 256, n, dtype=np.uint8))
                want = oracle.kzo_update_buffer(C.c_void_p(ob.ptr(ref)), size, off, n, data)
                ops.append((want, lambda off=off, data=data: q.cmd_update_buffer(dst.address, size, off, data)))

        def issue():
            for want, fn in ops:
                got = abi.SUCCESS
                try:
                    fn()
                except host.KazanError as e:
                    got = e.code
                assert got == want

        if recorded:
            with q.record() as cb:
                issue()
            assert cb.launch_count < sum(1 for w, _ in ops if w == abi.SUCCESS)  # some of them merged
            q.submit([cb])
            q.wait_idle()
            cb.destroy()
        else:
            issue()
        assert dst.download().tobytes() == ref.tobytes(), ("recorded" if recorded else "immediate")
        gpu.release()
