"""GPU: the data path of WSI presentation (include/kazan_b200.h "WSI presentation").

The reference stops before swapchain creation (xcb_swapchain.rs:342-349 `unimplemented!()`), so there is nothing upstream to
compare with: what a present must deliver is fixed by X11 itself — the window's byte order, its scanline pitch — and is
checked here by reading the SysV segment the way the X server would (a second, independent attach by id) and comparing with
the oracle's frame: rendered frames arrive bit for bit, R and B exchanged for B8G8R8A8 windows, rows at bytes_per_line."""
import ctypes as C
import ctypes.util

import numpy as np
import pytest

from kazan_b200 import abi, host
from tests import util

pytestmark = pytest.mark.gpu

libc = C.CDLL(ctypes.util.find_library("c") or "libc.so.6", use_errno=True)
libc.shmat.restype, libc.shmat.argtypes = C.c_void_p, [C.c_int, C.c_void_p, C.c_int]
libc.shmdt.argtypes = [C.c_void_p]


def read_segment(seg: abi.SwapchainSegment) -> np.ndarray:
    """attach the segment by its id, as the X server does after xcb_shm_attach, and copy it out"""
    at = libc.shmat(seg.shm_id, None, 0o10000)  # SHM_RDONLY
    assert at not in (None, C.c_void_p(-1).value), C.get_errno()
    data = np.ctypeslib.as_array((C.c_uint8 * seg.size).from_address(at)).copy()
    libc.shmdt(at)
    return data


def expected_rows(frame_rgba: np.ndarray, width, height, byte_order, bytes_per_line) -> np.ndarray:
    px = frame_rgba[:width * height * 4].reshape(height, width, 4)
    if byte_order == abi.WINDOW_B8G8R8A8:
        px = px[:, :, [2, 1, 0, 3]]
    out = np.zeros((height, bytes_per_line), dtype=np.uint8)
    out[:, :width * 4] = px.reshape(height, width * 4)
    return out


@pytest.mark.parametrize("byte_order", [abi.WINDOW_B8G8R8A8, abi.WINDOW_R8G8B8A8])
@pytest.mark.parametrize("size", [(256, 144), (131, 77)])
def test_rendered_frames_reach_the_window_segments(gpu, byte_order, size):
    W, H = size
    sc = host.Swapchain(gpu.device, W, H, image_count=3, window_byte_order=byte_order)
    scene = util.Scene("fill", W, H, util.FULLSCREEN_TRIANGLE, state=util.pipeline_state(fs=abi.FS_PROCEDURAL_ORBIT))
    want_frame, _, _ = scene.run_oracle()
    vbuf = gpu.buffer(scene.vertices.nbytes, scene.vertices)
    vi = abi.VertexInput(vbuf.address, scene.vertices.nbytes, scene.stride, 0, scene.pos_comps, scene.color_offset)
    q = gpu.queue
    seen = []
    for frame in range(5):  # more presents than images: acquire has to wait for a segment to be free again
        idx = sc.acquire_next_image()
        assert idx == frame % 3
        seen.append(idx)
        image = sc.image(idx)
        assert (image.width, image.height, image.format, image.tiling) == (W, H, abi.FORMAT_R8G8B8A8_UNORM, 0)
        fb = abi.Framebuffer()
        fb.color = image.desc()
        q.cmd_clear_color_image(image, (0.1 * frame, 0.2, 0.3, 1.0))
        if frame != 3:  # frame 3 presents the clear alone
            q.cmd_draw(scene.state, fb, scene.viewport, scene.scissor, vi, 3, 1, 0, 0)
        q.present(sc, idx)
        sc.wait_presented(idx)
        seg = sc.segment(idx)
        assert seg.bytes_per_line == W * 4 and seg.size == W * 4 * H and seg.shm_id >= 0
        got = read_segment(seg).reshape(H, seg.bytes_per_line)
        if frame != 3:
            want = expected_rows(want_frame, W, H, byte_order, seg.bytes_per_line)
        else:
            from oracle import binding as ob
            packed = np.uint32(ob.load().kzo_pack_rgba8((C.c_float * 4)(0.1 * frame, 0.2, 0.3, 1.0)))
            want = expected_rows(np.full(W * H, packed, dtype=np.uint32).view(np.uint8), W, H, byte_order, seg.bytes_per_line)
        assert np.array_equal(got, want), f"frame {frame}"
    sc.destroy()
    gpu.release()


def test_acquire_and_present_rules(gpu):
    sc = host.Swapchain(gpu.device, 64, 32, image_count=2)
    a, b = sc.acquire_next_image(), sc.acquire_next_image()
    assert (a, b) == (0, 1)
    assert sc.acquire_next_image(0) == -1  # VK_NOT_READY: both images are with the application
    with pytest.raises(host.KazanError) as e:
        gpu.queue.present(sc, 5)
    assert e.value.code == abi.ERROR_VALIDATION_FAILED
    gpu.queue.present(sc, 0)
    with pytest.raises(host.KazanError):  # presented already, not acquired again
        gpu.queue.present(sc, 0)
    assert sc.acquire_next_image() == 0
    with gpu.queue.record():
        with pytest.raises(host.KazanError):  # a queue operation, not a command
            gpu.queue.present(sc, 1)
    sc.destroy()
    for bad in (dict(image_count=1), dict(image_count=17), dict(window_byte_order=7), dict(scanline_alignment=3), dict(width=0)):
        args = dict(width=64, height=32)
        args.update(bad)
        with pytest.raises(host.KazanError):
            host.Swapchain(gpu.device, **args)
    gpu.release()


def test_present_4k_rate(gpu):
    """a 4K frame into a B8G8R8A8 window segment: the byte-swapping kernel stores straight into the mapped segment; report
    the rate (PCIe-bound) and check the frame"""
    import time
    W, H = 3840, 2160
    sc = host.Swapchain(gpu.device, W, H, image_count=2)
    q = gpu.queue
    rng = np.random.default_rng(7)
    frame = rng.integers(0, 2**32, W * H, dtype=np.uint32)
    staging = gpu.buffer(frame.nbytes, frame)
    region = host.buffer_image_copy(W, H)
    for _ in range(2):
        idx = sc.acquire_next_image()
        q.cmd_copy_buffer_to_image(staging.address, frame.nbytes, sc.image(idx), [region])
        q.wait_idle()
        t0 = time.perf_counter()
        q.present(sc, idx)
        sc.wait_presented(idx)
        dt = time.perf_counter() - t0
    got = read_segment(sc.segment(idx)).view(np.uint32)
    want = (frame & 0xFF00FF00) | ((frame & 0xFF) << 16) | ((frame >> 16) & 0xFF)
    assert np.array_equal(got, want)
    print(f"present 4K B8G8R8A8: {dt * 1e3:.2f} ms, {frame.nbytes / dt / 1e9:.1f} GB/s into the segment")
    sc.destroy()
    gpu.release()
