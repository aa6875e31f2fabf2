"""GPU: memory, queue and fence objects of the C ABI (vkAllocateMemory / vkMapMemory / vkQueueSubmit / fences)."""
import ctypes as C

import numpy as np
import pytest

from kazan_b200 import abi, host

pytestmark = pytest.mark.gpu


def test_host_coherent_memory_is_mapped_and_device_visible(gpu):
    m = gpu.device.allocate(1000, abi.MEMORY_TYPE_HOST_COHERENT)
    assert m.size == 1024  # rounded up to MIN_MEMORY_MAP_ALIGNMENT (device_memory.rs:166-181)
    p = m.map(128)
    view = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(64,))
    view[:] = 7
    gpu.queue.cmd_fill_buffer(m.address, 1024, 128, 32, 0x0A0B0C0D)
    gpu.queue.wait_idle()
    assert view[:4].tolist() == [0x0D, 0x0C, 0x0B, 0x0A] and view[32] == 7
    dl = gpu.device.allocate(4096)
    with pytest.raises(host.KazanError) as e:
        dl.map()
    assert e.value.code == abi.ERROR_VALIDATION_FAILED
    m.free(); dl.free()


def test_fences(gpu):
    f = gpu.device.create_fence()
    assert f.status() == abi.NOT_READY
    assert f.wait(1000) == abi.TIMEOUT
    mem = gpu.device.allocate(1 << 26)
    gpu.queue.cmd_fill_buffer(mem.address, 1 << 26, 0, abi.WHOLE_SIZE, 1)
    gpu.queue.signal_fence(f)
    assert f.wait() == abi.SUCCESS and f.status() == abi.SUCCESS
    with pytest.raises(host.KazanError):
        gpu.queue.signal_fence(f)  # must be reset first
    f.reset()
    assert f.status() == abi.NOT_READY
    s = gpu.device.create_fence(signaled=True)
    assert s.status() == abi.SUCCESS
    f.destroy(); s.destroy(); mem.free()


def test_launch_counter_counts_kernels(gpu):
    mem = gpu.device.allocate(1 << 20)
    before = gpu.queue.launch_count
    gpu.queue.cmd_fill_buffer(mem.address, 1 << 20, 0, abi.WHOLE_SIZE, 1)
    gpu.queue.wait_idle()
    assert gpu.queue.launch_count == before + 1
    mem.free()


def test_out_of_device_memory_is_reported(gpu):
    with pytest.raises(host.KazanError) as e:
        gpu.device.allocate(1 << 45)
    assert e.value.code == abi.ERROR_OUT_OF_DEVICE_MEMORY


def test_stream_order_between_copies_and_dependently_launched_kernels(gpu):
    """The library launches its kernels with programmatic dependent launch.  That relaxes only the kernel -> kernel launch
    gap: a kernel must still see everything an earlier asynchronous upload wrote, an earlier kernel wrote (read after
    write through a chain of kernels), and a later download must see what the kernel wrote — with no host wait in
    between.  Buffers are large so that an early start would be caught (16 MiB upload = a few hundred microseconds)."""
    lib = host.load_library()
    n = 16 << 20
    rounds = 12
    pinned = []
    for _ in range(2):
        p = C.c_void_p()
        assert lib.kzn_host_alloc_pinned(n, C.byref(p)) == abi.SUCCESS
        pinned.append(p)
    src = [np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(n // 4,)) for p in pinned]
    back = C.c_void_p()
    assert lib.kzn_host_alloc_pinned(rounds * 64, C.byref(back)) == abi.SUCCESS
    got = np.ctypeslib.as_array(C.cast(back, C.POINTER(C.c_uint32)), shape=(rounds, 16))
    got[:] = 0
    a, b, c = (gpu.device.allocate(n) for _ in range(3))
    q = gpu.queue
    region = [abi.BufferCopy(0, 0, n)]
    tail = [abi.BufferCopy(n - 64, 0, 64)]
    small = gpu.device.allocate(64 * rounds)
    for k in range(rounds):
        src[k % 2][:] = 0xA000 + k  # the host may refill this pinned buffer: its previous upload was two rounds ago ...
        q.upload(a.address, pinned[k % 2].value, n)          # async H2D into A
        q.cmd_copy_buffer(a.address, n, b.address, n, region)  # kernel 1 reads A (must follow the upload)
        q.cmd_copy_buffer(b.address, n, c.address, n, region)  # kernel 2 reads what kernel 1 wrote
        q.cmd_fill_buffer(a.address, n, 0, n, 0xDEAD0000 + k)  # kernel 3 overwrites A (must not overtake kernel 1)
        q.cmd_copy_buffer(c.address, n, small.address + 64 * k, 64, tail)  # last 64 bytes of C
        if k % 2 == 1:
            q.wait_idle()  # ... so wait before the pinned buffers are rewritten
    q.download(back.value, small.address, 64 * rounds)         # async D2H after the last kernel
    q.wait_idle()
    for k in range(rounds):
        assert np.all(got[k] == 0xA000 + k), (k, [hex(int(v)) for v in got[k][:4]])
    for m in (a, b, c, small):
        m.free()
    for p in pinned + [back]:
        lib.kzn_host_free_pinned(p)
