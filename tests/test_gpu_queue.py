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
