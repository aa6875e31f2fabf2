"""GPU: split-frame rendering into the presenter's image over peer memory.

Two processes share cuda:0 (CUDA IPC works between processes on one device just as between the GPUs of a box), each
renders its band of the frame into the PRESENTER's image through an imported allocation, flags carry the composite.
The assembled frame must equal the oracle's single-pass frame bit for bit.
"""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from kazan_b200 import host

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, result):
    import ctypes as C
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kazan_b200 import abi, host
    from kazan_b200.split_frame import PeerSplitFrame
    from tests import cases
    from tests.util import FULLSCREEN_TRIANGLE, pipeline_state
    try:
        W, H, frames = 256, 144, 5
        dev = host.Device(0)
        q = dev.create_queue()
        pf = PeerSplitFrame(dev, q, W, H, world, rank, dist, n_targets=2, mode=mode)
        verts = dev.allocate(FULLSCREEN_TRIANGLE.nbytes)
        q.upload(verts.address, FULLSCREEN_TRIANGLE.ctypes.data, FULLSCREEN_TRIANGLE.nbytes)
        st = pipeline_state(fs=abi.FS_PROCEDURAL_ORBIT)
        vp = abi.Viewport(0, 0, W, H, 0, 1)
        sc = abi.Rect2D(abi.Offset2D(0, 0), abi.Extent2D(W, H))
        vi = abi.VertexInput(verts.address, FULLSCREEN_TRIANGLE.nbytes, 28, 0, 3, 12)
        for f in range(frames):  # more frames than targets: exercises the credit back-pressure
            fb = abi.Framebuffer()
            fb.color = abi.ImageDesc(pf.begin_frame(), abi.FORMAT_R8G8B8A8_UNORM, 0, W, H, 1, 1, 1, 0)
            fb.band_y0, fb.band_y1 = pf.split.band
            q.cmd_draw(st, fb, vp, sc, vi, 3)
            pf.end_frame()
        pf.check()
        if rank == 0:
            ok = True
            for k in range(2):
                out = np.empty(W * H * 4, dtype=np.uint8)
                q.download(out.ctypes.data, pf.target(k), out.nbytes)
                q.wait_idle()
                scene = next(s for s in cases.draw_scenes() if s.name == "fill_orbit_256x144")
                want, _, _ = scene.run_oracle()
                ok = ok and bool(np.array_equal(out, want))
            result.put(ok)
        pf.close()
        verts.free(); q.destroy(); dev.destroy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["push", "store"])
def test_two_ranks_render_bands_into_the_presenters_image(mode):
    ctx = mp.get_context("spawn")
    result = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, result)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert result.get(timeout=10) is True


def test_semaphore_orders_two_queues(gpu):
    """vkQueueSubmit semaphores: the second queue's copy starts only after the first queue's fill has finished."""
    from kazan_b200.host import Semaphore
    n = 64 << 20
    a, b = gpu.device.allocate(n), gpu.device.allocate(n)
    q2 = gpu.device.create_queue()
    sem = Semaphore(gpu.device)
    with pytest.raises(host.KazanError):
        q2.wait_semaphore(sem)  # no signal pending
    gpu.queue.cmd_fill_buffer(a.address, n, 0, n, 0x01020304)
    gpu.queue.signal_semaphore(sem)
    q2.wait_semaphore(sem)
    q2.copy_dma(b.address, a.address, n)
    q2.wait_idle()
    out = np.empty(n // 4, dtype=np.uint32)
    q2.download(out.ctypes.data, b.address, n)
    q2.wait_idle()
    assert (out == 0x01020304).all()
    sem.destroy(); q2.destroy(); a.free(); b.free()


def test_flags_order_work_on_one_queue(gpu):
    """signal then wait on the same queue; a wait that can never be satisfied gives up and reports through status."""
    mem = gpu.device.allocate(256)
    gpu.queue.cmd_fill_buffer(mem.address, 256, 0, 256, 0)
    gpu.queue.cmd_signal_flag(mem.address, 3)
    gpu.queue.cmd_signal_flag(mem.address, 4)
    gpu.queue.cmd_wait_flag(mem.address, 7, mem.address + 64)
    gpu.queue.wait_idle()
    out = np.zeros(32, dtype=np.uint64)
    gpu.queue.download(out.ctypes.data, mem.address, 256)
    gpu.queue.wait_idle()
    assert out[0] == 7 and out[8] == 0
    mem.free()
