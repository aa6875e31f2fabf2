"""Thin host-side harness over the C ABI (include/kazan_b200.h), used by tests and bench.

The product's host layer is C++ (the ICD, kazan_b200/csrc/icd) because the reference's host is
compiled Rust; this module only lets Python drive the same entry points with the reference's
names: Queue.cmd_clear_color_image <-> vkCmdClearColorImage (vulkan-driver/src/api_impl.rs:4780) etc.

There is no fallback: if libkazan_b200.so is missing or no CUDA device exists, this raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from pathlib import Path
from typing import Iterable, Sequence

from . import abi

_LIB_PATH = Path(__file__).resolve().parent / "libkazan_b200.so"
_lib = None


class KazanError(RuntimeError):
    def __init__(self, code: int, message: str):
        self.code = code
        super().__init__(f"{abi.RESULT_NAMES.get(code, code)}: {message}")


def library_path() -> Path:
    return _LIB_PATH


def load_library() -> C.CDLL:
    """Load the CUDA extension.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise ImportError(
                f"{_LIB_PATH} is missing: build it with `python -m kazan_b200.build` "
                "(kazan_b200 has no CPU or PyTorch fallback)")
        _lib = abi.bind(C.CDLL(str(_LIB_PATH)))
    return _lib


def _check(code: int) -> int:
    if code < 0:
        raise KazanError(code, load_library().kzn_last_error().decode())
    return code


def _array(ctype, items: Sequence):
    items = list(items)
    return (ctype * max(len(items), 1))(*items), len(items)


@dataclass
class Image:
    """vulkan-driver/src/image.rs `ImageProperties` + bound memory address."""
    address: int
    width: int
    height: int
    depth: int = 1
    array_layers: int = 1
    format: int = abi.FORMAT_R8G8B8A8_UNORM
    tiling: int = abi.TILING_LINEAR
    mip_levels: int = 1

    def desc(self) -> abi.ImageDesc:
        return abi.ImageDesc(self.address, self.format, self.tiling, self.width, self.height, self.depth,
                             self.array_layers, self.mip_levels, 0)

    def memory_requirements(self) -> tuple[int, int]:
        size, align = abi.u64(), abi.u64()
        d = self.desc()
        _check(load_library().kzn_image_memory_requirements(C.byref(d), C.byref(size), C.byref(align)))
        return size.value, align.value


def color_range(base_layer=0, layer_count=abi.REMAINING, aspect=abi.ASPECT_COLOR, base_mip=0, level_count=abi.REMAINING):
    return abi.ImageSubresourceRange(aspect, base_mip, level_count, base_layer, layer_count)


def buffer_image_copy(width, height, depth=1, buffer_offset=0, row_length=0, image_height=0, offset=(0, 0, 0),
                      base_layer=0, layer_count=1, aspect=abi.ASPECT_COLOR, mip=0) -> abi.BufferImageCopy:
    return abi.BufferImageCopy(buffer_offset, row_length, image_height,
                               abi.ImageSubresourceLayers(aspect, mip, base_layer, layer_count),
                               abi.Offset3D(*offset), abi.Extent3D(width, height, depth))


class Device:
    def __init__(self, ordinal: int = 0):
        self._lib = load_library()
        h = C.c_void_p()
        _check(self._lib.kzn_device_create(ordinal, C.byref(h)))
        self.handle = h
        self.ordinal = ordinal

    @staticmethod
    def count() -> int:
        n = abi.i32()
        code = load_library().kzn_device_count(C.byref(n))
        return n.value if code == 0 else 0

    @property
    def sm_count(self) -> int:
        return self._lib.kzn_device_sm_count(self.handle)

    def wait_idle(self):
        _check(self._lib.kzn_device_wait_idle(self.handle))

    def create_queue(self) -> "Queue":
        h = C.c_void_p()
        _check(self._lib.kzn_queue_create(self.handle, C.byref(h)))
        return Queue(self, h)

    def wrap_stream(self, cuda_stream: int) -> "Queue":
        h = C.c_void_p()
        _check(self._lib.kzn_queue_wrap_stream(self.handle, C.c_void_p(cuda_stream), C.byref(h)))
        return Queue(self, h)

    def allocate(self, size: int, memory_type: int = abi.MEMORY_TYPE_DEVICE_LOCAL) -> "Memory":
        h = C.c_void_p()
        _check(self._lib.kzn_memory_allocate(self.handle, size, memory_type, C.byref(h)))
        return Memory(self, h)

    def import_memory(self, handle: bytes, size: int) -> "Memory":
        """Open another process's exported DEVICE_LOCAL allocation as peer memory (kzn_memory_import)."""
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        h = C.c_void_p()
        _check(self._lib.kzn_memory_import(self.handle, buf, size, C.byref(h)))
        return Memory(self, h)

    def create_fence(self, signaled: bool = False) -> "Fence":
        h = C.c_void_p()
        _check(self._lib.kzn_fence_create(self.handle, int(signaled), C.byref(h)))
        return Fence(self, h)

    def destroy(self):
        if self.handle:
            self._lib.kzn_device_destroy(self.handle)
            self.handle = None


class Memory:
    def __init__(self, device: Device, handle):
        self.device, self.handle, self._lib = device, handle, device._lib

    @property
    def address(self) -> int:
        return self._lib.kzn_memory_device_address(self.handle)

    @property
    def size(self) -> int:
        return self._lib.kzn_memory_size(self.handle)

    def export(self) -> bytes:
        buf = (C.c_uint8 * 64)()
        _check(self._lib.kzn_memory_export(self.handle, buf))
        return bytes(buf)

    def map(self, offset: int = 0) -> int:
        p = C.c_void_p()
        _check(self._lib.kzn_memory_map(self.handle, offset, C.byref(p)))
        return p.value

    def free(self):
        if self.handle:
            self._lib.kzn_memory_free(self.handle)
            self.handle = None


class Semaphore:
    """VkSemaphore (binary): orders work between two queues of one device."""

    def __init__(self, device: Device):
        self.device, self._lib = device, device._lib
        h = C.c_void_p()
        _check(self._lib.kzn_semaphore_create(device.handle, C.byref(h)))
        self.handle = h

    def destroy(self):
        if self.handle:
            self._lib.kzn_semaphore_destroy(self.handle)
            self.handle = None


class Fence:
    def __init__(self, device: Device, handle):
        self.device, self.handle, self._lib = device, handle, device._lib

    def status(self) -> int:
        return _check(self._lib.kzn_fence_status(self.handle))

    def wait(self, timeout_ns: int = 2**64 - 1) -> int:
        return _check(self._lib.kzn_fence_wait(self.handle, timeout_ns))

    def reset(self):
        _check(self._lib.kzn_fence_reset(self.handle))

    def destroy(self):
        if self.handle:
            self._lib.kzn_fence_destroy(self.handle)
            self.handle = None


class CommandBuffer:
    """vkBeginCommandBuffer .. vkEndCommandBuffer: commands issued on the queue inside the `with` block are recorded."""

    def __init__(self, queue: "Queue"):
        self.queue, self._lib, self.handle = queue, queue._lib, C.c_void_p()

    def __enter__(self):
        _check(self._lib.kzn_command_buffer_begin(self.queue.handle, C.byref(self.handle)))
        return self

    def __exit__(self, exc_type, exc, tb):
        if exc_type is not None:
            self.destroy()
            return False
        _check(self._lib.kzn_command_buffer_end(self.handle))
        return False

    @property
    def launch_count(self) -> int:
        return self._lib.kzn_command_buffer_launch_count(self.handle)

    def destroy(self):
        if self.handle:
            self._lib.kzn_command_buffer_destroy(self.handle)
            self.handle = C.c_void_p()


class Swapchain:
    """VkSwapchainKHR, GPU side (include/kazan_b200.h "WSI presentation"): presentable images in HBM, one SysV shared memory
    segment per image for xcb_shm_put_image."""

    def __init__(self, device: "Device", width: int, height: int, image_count: int = 2, window_byte_order: int = abi.WINDOW_B8G8R8A8,
                 scanline_alignment: int = 4):
        self.device, self._lib = device, device._lib
        info = abi.SwapchainCreateInfo(width, height, image_count, window_byte_order, scanline_alignment, 0)
        self.handle = C.c_void_p()
        _check(self._lib.kzn_swapchain_create(device.handle, C.byref(info), C.byref(self.handle)))
        self.width, self.height, self.image_count = width, height, image_count

    def image(self, index: int) -> "Image":
        d = abi.ImageDesc()
        _check(self._lib.kzn_swapchain_get_image(self.handle, index, C.byref(d)))
        return Image(d.address, d.width, d.height, d.depth, d.array_layers, d.format, d.tiling, d.mip_levels)

    def segment(self, index: int) -> abi.SwapchainSegment:
        s = abi.SwapchainSegment()
        _check(self._lib.kzn_swapchain_get_segment(self.handle, index, C.byref(s)))
        return s

    def acquire_next_image(self, timeout_ns: int = 2**64 - 1) -> int:
        """index of the acquired image; raises KazanError on errors, returns -1 for VK_NOT_READY / VK_TIMEOUT"""
        idx = C.c_uint32()
        code = _check(self._lib.kzn_swapchain_acquire_next_image(self.handle, timeout_ns, C.byref(idx)))
        return idx.value if code == 0 else -1

    def wait_presented(self, index: int):
        _check(self._lib.kzn_swapchain_wait_presented(self.handle, index))

    def destroy(self):
        if self.handle:
            self._lib.kzn_swapchain_destroy(self.handle)
            self.handle = C.c_void_p()


class Queue:
    """One VkQueue == one CUDA stream.  Method names follow the Vulkan commands they implement."""

    def __init__(self, device: Device, handle):
        self.device, self.handle, self._lib = device, handle, device._lib

    # -- queue -------------------------------------------------------------------------------
    def wait_idle(self):
        _check(self._lib.kzn_queue_wait_idle(self.handle))

    @property
    def launch_count(self) -> int:
        return self._lib.kzn_queue_launch_count(self.handle)

    @property
    def stream(self) -> int:
        return self._lib.kzn_queue_stream(self.handle) or 0

    def record(self) -> CommandBuffer:
        """with queue.record() as cb: queue.cmd_...()   then   queue.submit([cb])"""
        return CommandBuffer(self)

    def submit(self, command_buffers, fence: "Fence | None" = None):
        arr = (C.c_void_p * max(len(command_buffers), 1))(*[cb.handle.value for cb in command_buffers])
        _check(self._lib.kzn_queue_submit(self.handle, len(command_buffers), arr, fence.handle if fence else None))

    def present(self, swapchain: Swapchain, index: int):
        """vkQueuePresentKHR: after the queue's earlier work, the image goes into its shared memory segment"""
        _check(self._lib.kzn_queue_present(self.handle, swapchain.handle, index))

    def reserve_draw(self, max_triangles: int, width: int, height: int):
        _check(self._lib.kzn_queue_reserve_draw(self.handle, max_triangles, width, height))

    def set_kernel_timing(self, enable: bool = True):
        _check(self._lib.kzn_queue_set_kernel_timing(self.handle, int(enable)))

    def last_kernel_ms(self) -> float:
        """Device duration of the dominant kernel of the last command (waits for it)."""
        ms = abi.f32()
        code = _check(self._lib.kzn_queue_last_kernel_ms(self.handle, C.byref(ms)))
        if code != 0:
            raise KazanError(code, self._lib.kzn_last_error().decode())
        return ms.value

    def signal_semaphore(self, semaphore: "Semaphore"):
        _check(self._lib.kzn_queue_signal_semaphore(self.handle, semaphore.handle))

    def wait_semaphore(self, semaphore: "Semaphore"):
        _check(self._lib.kzn_queue_wait_semaphore(self.handle, semaphore.handle))

    def calibrate_fp32_issue(self) -> float:
        """Measured fp32 issue rate, tera lane-operations per second."""
        v = abi.f32()
        _check(self._lib.kzn_calibrate_fp32_issue(self.handle, C.byref(v)))
        return v.value

    def cmd_pipeline_barrier(self):
        _check(self._lib.kzn_cmd_pipeline_barrier(self.handle))

    def copy_dma(self, dst_address: int, src_address: int, size: int):
        _check(self._lib.kzn_queue_copy_dma(self.handle, dst_address, src_address, size))

    def cmd_signal_flag(self, flag_address: int, add: int = 1):
        _check(self._lib.kzn_cmd_signal_flag(self.handle, flag_address, add))

    def cmd_wait_flag(self, flag_address: int, expected: int, status_address: int = 0):
        _check(self._lib.kzn_cmd_wait_flag(self.handle, flag_address, expected, status_address))

    def cmd_wait_flags(self, first_flag_address: int, flag_count: int, expected: int, status_address: int = 0):
        _check(self._lib.kzn_cmd_wait_flags(self.handle, first_flag_address, flag_count, expected, status_address))

    def signal_fence(self, fence: Fence):
        _check(self._lib.kzn_queue_signal_fence(self.handle, fence.handle))

    def upload(self, dst_address: int, src_ptr: int, size: int):
        _check(self._lib.kzn_queue_upload(self.handle, dst_address, C.c_void_p(src_ptr), size))

    def download(self, dst_ptr: int, src_address: int, size: int):
        _check(self._lib.kzn_queue_download(self.handle, C.c_void_p(dst_ptr), src_address, size))

    def destroy(self):
        if self.handle:
            self._lib.kzn_queue_destroy(self.handle)
            self.handle = None

    # -- transfer ------------------------------------------------------------------------------
    def cmd_clear_color_image(self, image: Image, color: Sequence[float], ranges: Iterable | None = None):
        d = image.desc()
        c = abi.ClearColorValue()
        c.float32[:] = list(color)
        arr, n = _array(abi.ImageSubresourceRange, ranges if ranges is not None else [color_range()])
        _check(self._lib.kzn_cmd_clear_color_image(self.handle, C.byref(d), C.byref(c), n, arr))

    def cmd_clear_depth_stencil_image(self, image: Image, depth: float, ranges: Iterable | None = None):
        d = image.desc()
        v = abi.ClearDepthStencilValue(depth, 0)
        arr, n = _array(abi.ImageSubresourceRange,
                        ranges if ranges is not None else [color_range(aspect=abi.ASPECT_DEPTH)])
        _check(self._lib.kzn_cmd_clear_depth_stencil_image(self.handle, C.byref(d), C.byref(v), n, arr))

    def cmd_clear_attachment(self, image: Image, value: Sequence[float], rects: Iterable[abi.ClearRect]):
        d = image.desc()
        c = abi.ClearColorValue()
        c.float32[:] = (list(value) + [0.0] * 4)[:4]
        arr, n = _array(abi.ClearRect, rects)
        _check(self._lib.kzn_cmd_clear_attachment(self.handle, C.byref(d), C.byref(c), n, arr))

    def cmd_begin_render_pass(self, framebuffer: abi.Framebuffer, render_area: abi.Rect2D | None = None,
                              clear_color: Sequence[float] | None = None, clear_depth: float | None = None):
        """vkCmdBeginRenderPass: loadOp CLEAR for the attachments a clear value is given for, LOAD for the others."""
        b = abi.RenderPassBegin()
        b.framebuffer = framebuffer
        w = framebuffer.color.width if framebuffer.color.format else framebuffer.depth.width
        h = framebuffer.color.height if framebuffer.color.format else framebuffer.depth.height
        b.render_area = render_area if render_area is not None else abi.Rect2D(abi.Offset2D(0, 0), abi.Extent2D(w, h))
        if clear_color is not None:
            b.color_load_op = abi.LOAD_OP_CLEAR
            b.clear_color.float32[:] = (list(clear_color) + [0.0] * 4)[:4]
        if clear_depth is not None:
            b.depth_load_op = abi.LOAD_OP_CLEAR
            b.clear_depth_stencil = abi.ClearDepthStencilValue(clear_depth, 0)
        _check(self._lib.kzn_cmd_begin_render_pass(self.handle, C.byref(b)))

    def cmd_end_render_pass(self):
        _check(self._lib.kzn_cmd_end_render_pass(self.handle))

    def cmd_copy_buffer_to_image(self, buffer_address: int, buffer_size: int, image: Image,
                                 regions: Iterable[abi.BufferImageCopy]):
        d = image.desc()
        arr, n = _array(abi.BufferImageCopy, regions)
        _check(self._lib.kzn_cmd_copy_buffer_to_image(self.handle, buffer_address, buffer_size, C.byref(d), n, arr))

    def cmd_copy_image_to_buffer(self, image: Image, buffer_address: int, buffer_size: int,
                                 regions: Iterable[abi.BufferImageCopy]):
        d = image.desc()
        arr, n = _array(abi.BufferImageCopy, regions)
        _check(self._lib.kzn_cmd_copy_image_to_buffer(self.handle, C.byref(d), buffer_address, buffer_size, n, arr))

    def cmd_copy_buffer(self, src_address: int, src_size: int, dst_address: int, dst_size: int,
                        regions: Iterable[abi.BufferCopy]):
        arr, n = _array(abi.BufferCopy, regions)
        _check(self._lib.kzn_cmd_copy_buffer(self.handle, src_address, src_size, dst_address, dst_size, n, arr))

    def cmd_fill_buffer(self, dst_address: int, dst_size: int, offset: int, size: int, data: int):
        _check(self._lib.kzn_cmd_fill_buffer(self.handle, dst_address, dst_size, offset, size, data))

    def cmd_update_buffer(self, dst_address: int, dst_size: int, offset: int, data: bytes):
        buf = C.create_string_buffer(data, len(data))
        _check(self._lib.kzn_cmd_update_buffer(self.handle, dst_address, dst_size, offset, len(data),
                                               C.cast(buf, C.c_void_p)))

    def cmd_copy_image(self, src: Image, dst: Image, regions: Iterable[abi.ImageCopy]):
        s, d = src.desc(), dst.desc()
        arr, n = _array(abi.ImageCopy, regions)
        _check(self._lib.kzn_cmd_copy_image(self.handle, C.byref(s), C.byref(d), n, arr))

    def cmd_blit_image(self, src: Image, dst: Image, regions: Iterable[abi.ImageBlit], filter: int):
        s, d = src.desc(), dst.desc()
        arr, n = _array(abi.ImageBlit, regions)
        _check(self._lib.kzn_cmd_blit_image(self.handle, C.byref(s), C.byref(d), n, arr, filter))

    def cmd_convert_tiling(self, src: Image, dst: Image):
        s, d = src.desc(), dst.desc()
        _check(self._lib.kzn_cmd_convert_tiling(self.handle, C.byref(s), C.byref(d)))

    # -- draw -----------------------------------------------------------------------------------
    def set_draw_statistics(self, address: int):
        _check(self._lib.kzn_queue_set_draw_statistics(self.handle, address))

    def cmd_draw(self, state: abi.GraphicsPipelineState, framebuffer: abi.Framebuffer, viewport: abi.Viewport,
                 scissor: abi.Rect2D, vertices: abi.VertexInput, vertex_count: int, instance_count: int = 1,
                 first_vertex: int = 0, first_instance: int = 0):
        _check(self._lib.kzn_cmd_draw(self.handle, C.byref(state), C.byref(framebuffer), C.byref(viewport),
                                      C.byref(scissor), C.byref(vertices), vertex_count, instance_count,
                                      first_vertex, first_instance))

    def cmd_draw_indexed(self, state: abi.GraphicsPipelineState, framebuffer: abi.Framebuffer,
                         viewport: abi.Viewport, scissor: abi.Rect2D, vertices: abi.VertexInput,
                         index_buffer_address: int, index_buffer_size: int, index_type: int, index_count: int,
                         instance_count: int = 1, first_index: int = 0, vertex_offset: int = 0,
                         first_instance: int = 0):
        _check(self._lib.kzn_cmd_draw_indexed(self.handle, C.byref(state), C.byref(framebuffer), C.byref(viewport),
                                              C.byref(scissor), C.byref(vertices), index_buffer_address,
                                              index_buffer_size, index_type, index_count, instance_count,
                                              first_index, vertex_offset, first_instance))
