"""GPU: command buffers (vkBeginCommandBuffer .. vkEndCommandBuffer .. vkQueueSubmit).

BASELINE configs[0] — vkCmdClearColorImage + vkCmdCopyBufferToImage on a 1024x1024 R8G8B8A8_UNORM image — recorded once
and replayed; the result is compared byte for byte with the oracle's.
"""
import ctypes as C

import numpy as np
import pytest

from kazan_b200 import abi, host
from kazan_b200.host import buffer_image_copy
from oracle import binding as ob
from tests import cases, util
from tests.util import check

pytestmark = pytest.mark.gpu


def _oracle_clear_then_copy(w, h, color, buf, region):
    lib = ob.load()
    mem = ob.aligned_empty(w * h * 4)
    mem[:] = 0x5A
    d = abi.ImageDesc(ob.ptr(mem), abi.FORMAT_R8G8B8A8_UNORM, 0, w, h, 1, 1, 1, 0)
    c = abi.ClearColorValue()
    c.float32[:] = list(color)
    rng, n = util.arr(abi.ImageSubresourceRange, [abi.ImageSubresourceRange(abi.ASPECT_COLOR, 0, abi.REMAINING, 0, abi.REMAINING)])
    check(lib.kzo_clear_color_image(C.byref(d), C.byref(c), n, rng))
    reg, n = util.arr(abi.BufferImageCopy, [region])
    check(lib.kzo_copy_buffer_to_image(C.c_void_p(ob.ptr(buf)), buf.nbytes, C.byref(d), n, reg))
    return mem


def test_config0_clear_and_copy_recorded_once_replayed_twice(gpu):
    w = h = 1024
    color = (0.25, 0.5, 0.75, 1.0)  # -> 40 80 BF FF
    region = buffer_image_copy(992, 1000, offset=(16, 8, 0))  # sub-rect so that the clear stays visible around it
    buf = ob.random_bytes(992 * 1000 * 4, 0x4B415A414E)
    want = _oracle_clear_then_copy(w, h, color, buf, region)
    image = gpu.buffer(w * h * 4, np.full(w * h * 4, 0x5A, dtype=np.uint8))
    staging = gpu.buffer(buf.nbytes, buf)
    img = host.Image(image.address, w, h)
    q = gpu.queue
    before = q.launch_count
    with q.record() as cb:
        q.cmd_clear_color_image(img, color)
        q.cmd_copy_buffer_to_image(staging.address, buf.nbytes, img, [region])
    assert q.launch_count == before and cb.launch_count == 2  # recording launches nothing
    assert np.array_equal(image.download(), np.full(w * h * 4, 0x5A, dtype=np.uint8))
    fence = gpu.device.create_fence()
    q.submit([cb], fence)
    assert fence.wait() == abi.SUCCESS
    assert q.launch_count == before + 2
    assert np.array_equal(image.download(), want)
    assert want[:4].tolist() == [0x40, 0x80, 0xBF, 0xFF]
    image.upload(np.zeros(w * h * 4, dtype=np.uint8))
    q.submit([cb, cb])  # replays are idempotent here
    assert np.array_equal(image.download(), want)
    cb.destroy(); fence.destroy()
    gpu.release()


def test_recorded_draw_needs_reserved_scratch_and_matches_oracle(gpu):
    scene = next(s for s in cases.draw_scenes() if s.name == "mesh_40x20_depth_blend")
    want_c, want_d, _ = scene.run_oracle()
    color0, depth0 = scene._initial()
    dev = host.Device(0)
    q = dev.create_queue()  # a fresh queue: no raster scratch yet
    mem = [dev.allocate(a.nbytes) for a in (color0, depth0, scene.vertices, scene.indices)]
    for m, a in zip(mem, (color0, depth0, scene.vertices, scene.indices)):
        a = np.ascontiguousarray(a)
        q.upload(m.address, a.ctypes.data, a.nbytes)
    q.wait_idle()
    fb = scene._framebuffer(mem[0].address, mem[1].address)
    vi = abi.VertexInput(mem[2].address, scene.vertices.nbytes, scene.stride, 0, scene.pos_comps, scene.color_offset)

    def draw():
        q.cmd_draw_indexed(scene.state, fb, scene.viewport, scene.scissor, vi, mem[3].address, scene.indices.nbytes,
                           scene.index_type, scene.count, scene.instances, scene.first, scene.vertex_offset, 0)

    with pytest.raises(host.KazanError) as e:
        with q.record():
            draw()
    assert e.value.code == abi.ERROR_VALIDATION_FAILED and "kzn_queue_reserve_draw" in str(e.value)
    q.reserve_draw(scene.count // 3, scene.width, scene.height)
    with q.record() as cb:
        draw()
    q.submit([cb])
    q.wait_idle()
    got_c, got_d = np.empty_like(color0), np.empty_like(depth0)
    q.download(got_c.ctypes.data, mem[0].address, got_c.nbytes)
    q.download(got_d.ctypes.data, mem[1].address, got_d.nbytes)
    q.wait_idle()
    assert got_d.tobytes() == want_d.tobytes() and got_c.tobytes() == want_c.tobytes()
    cb.destroy()
    for m in mem:
        m.free()
    q.destroy(); dev.destroy()


def test_update_buffer_data_is_captured_at_record_time(gpu):
    dst = gpu.buffer(64, np.zeros(64, dtype=np.uint8))
    data = np.arange(16, dtype=np.uint32)
    q = gpu.queue
    with q.record() as cb:
        q.cmd_update_buffer(dst.address, 64, 0, data.tobytes())
    data[:] = 0xDEAD  # the application may reuse its memory as soon as the command is recorded
    q.submit([cb])
    assert np.array_equal(dst.download().view(np.uint32), np.arange(16, dtype=np.uint32))
    cb.destroy()
    gpu.release()


def test_recording_rules(gpu):
    q = gpu.queue
    f = gpu.device.create_fence()
    with pytest.raises(host.KazanError):
        with q.record():
            q.wait_idle()  # host waits cannot be recorded
    with pytest.raises(host.KazanError):
        with q.record():
            q.signal_fence(f)
    with q.record() as cb:
        with pytest.raises(host.KazanError):
            q.record().__enter__()  # one recording per queue
        with pytest.raises(host.KazanError):
            q.submit([cb])  # not executable yet
    q.submit([cb])  # an empty command buffer is valid
    q.wait_idle()
    cb.destroy(); f.destroy()
