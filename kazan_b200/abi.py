"""ctypes mirror of include/kazan_b200.h (struct layouts and prototypes).

The structs are layout-identical to the Vulkan 1.1.130 structs they are named after
(reference: external/Vulkan-Headers/include/vulkan/vulkan_core.h:2899-2919); tests/test_abi.py
checks this file, the C header and the Vulkan header against each other.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "kazan_b200.h"
BACKEND_HEADER = ROOT / "include" / "kazan_b200_backend.h"
HEADERS = (HEADER, BACKEND_HEADER)

# results -------------------------------------------------------------------------------------
SUCCESS = 0
NOT_READY = 1
TIMEOUT = 2
ERROR_OUT_OF_HOST_MEMORY = -1
ERROR_OUT_OF_DEVICE_MEMORY = -2
ERROR_INITIALIZATION_FAILED = -3
ERROR_DEVICE_LOST = -4
ERROR_FEATURE_NOT_PRESENT = -8
ERROR_FORMAT_NOT_SUPPORTED = -11
ERROR_VALIDATION_FAILED = -1000011001

RESULT_NAMES = {
    SUCCESS: "VK_SUCCESS",
    NOT_READY: "VK_NOT_READY",
    TIMEOUT: "VK_TIMEOUT",
    ERROR_OUT_OF_HOST_MEMORY: "VK_ERROR_OUT_OF_HOST_MEMORY",
    ERROR_OUT_OF_DEVICE_MEMORY: "VK_ERROR_OUT_OF_DEVICE_MEMORY",
    ERROR_INITIALIZATION_FAILED: "VK_ERROR_INITIALIZATION_FAILED",
    ERROR_DEVICE_LOST: "VK_ERROR_DEVICE_LOST",
    ERROR_FEATURE_NOT_PRESENT: "VK_ERROR_FEATURE_NOT_PRESENT",
    ERROR_FORMAT_NOT_SUPPORTED: "VK_ERROR_FORMAT_NOT_SUPPORTED",
    ERROR_VALIDATION_FAILED: "VK_ERROR_VALIDATION_FAILED_EXT",
}

# constants -----------------------------------------------------------------------------------
FORMAT_R8G8B8A8_UNORM = 37
FORMAT_D32_SFLOAT = 126
TILING_LINEAR = 0
TILING_TILED = 1
TILE_DIM = 16
ASPECT_COLOR = 1
ASPECT_DEPTH = 2
FILTER_NEAREST = 0
FILTER_LINEAR = 1
INDEX_TYPE_UINT16 = 0
INDEX_TYPE_UINT32 = 1
WHOLE_SIZE = 2**64 - 1
REMAINING = 2**32 - 1
MEMORY_TYPE_HOST_COHERENT = 0
MEMORY_TYPE_DEVICE_LOCAL = 1
VS_PASSTHROUGH = 0
FS_VERTEX_COLOR = 0
FS_PROCEDURAL_ORBIT = 1
FS_COMPILED = 2
VS_PASSTHROUGH, VS_COMPILED = 0, 1
# VkCompareOp
COMPARE_NEVER, COMPARE_LESS, COMPARE_EQUAL, COMPARE_LESS_OR_EQUAL, COMPARE_GREATER, COMPARE_NOT_EQUAL, \
    COMPARE_GREATER_OR_EQUAL, COMPARE_ALWAYS = range(8)
# VkBlendFactor (subset)
BLEND_ZERO, BLEND_ONE, BLEND_SRC_COLOR, BLEND_ONE_MINUS_SRC_COLOR, BLEND_DST_COLOR, BLEND_ONE_MINUS_DST_COLOR, \
    BLEND_SRC_ALPHA, BLEND_ONE_MINUS_SRC_ALPHA, BLEND_DST_ALPHA, BLEND_ONE_MINUS_DST_ALPHA = range(10)
# VkBlendOp
BLEND_OP_ADD, BLEND_OP_SUBTRACT, BLEND_OP_REVERSE_SUBTRACT, BLEND_OP_MIN, BLEND_OP_MAX = range(5)
CULL_NONE, CULL_FRONT, CULL_BACK, CULL_BOTH = range(4)
FRONT_FACE_CCW, FRONT_FACE_CW = 0, 1

u32, i32, u64, f32 = C.c_uint32, C.c_int32, C.c_uint64, C.c_float


class Offset3D(C.Structure):
    _fields_ = [("x", i32), ("y", i32), ("z", i32)]


class Extent3D(C.Structure):
    _fields_ = [("width", u32), ("height", u32), ("depth", u32)]


class Offset2D(C.Structure):
    _fields_ = [("x", i32), ("y", i32)]


class Extent2D(C.Structure):
    _fields_ = [("width", u32), ("height", u32)]


class Rect2D(C.Structure):
    _fields_ = [("offset", Offset2D), ("extent", Extent2D)]


class Viewport(C.Structure):
    _fields_ = [("x", f32), ("y", f32), ("width", f32), ("height", f32), ("minDepth", f32), ("maxDepth", f32)]


class ClearColorValue(C.Union):
    _fields_ = [("float32", f32 * 4), ("int32", i32 * 4), ("uint32", u32 * 4)]


class ClearDepthStencilValue(C.Structure):
    _fields_ = [("depth", f32), ("stencil", u32)]


class ImageSubresourceRange(C.Structure):
    _fields_ = [("aspectMask", u32), ("baseMipLevel", u32), ("levelCount", u32), ("baseArrayLayer", u32),
                ("layerCount", u32)]


class ImageSubresourceLayers(C.Structure):
    _fields_ = [("aspectMask", u32), ("mipLevel", u32), ("baseArrayLayer", u32), ("layerCount", u32)]


class BufferCopy(C.Structure):
    _fields_ = [("srcOffset", u64), ("dstOffset", u64), ("size", u64)]


class BufferImageCopy(C.Structure):
    _fields_ = [("bufferOffset", u64), ("bufferRowLength", u32), ("bufferImageHeight", u32),
                ("imageSubresource", ImageSubresourceLayers), ("imageOffset", Offset3D), ("imageExtent", Extent3D)]


class ImageCopy(C.Structure):
    _fields_ = [("srcSubresource", ImageSubresourceLayers), ("srcOffset", Offset3D),
                ("dstSubresource", ImageSubresourceLayers), ("dstOffset", Offset3D), ("extent", Extent3D)]


class ImageBlit(C.Structure):
    _fields_ = [("srcSubresource", ImageSubresourceLayers), ("srcOffsets", Offset3D * 2),
                ("dstSubresource", ImageSubresourceLayers), ("dstOffsets", Offset3D * 2)]


class ClearRect(C.Structure):
    _fields_ = [("rect", Rect2D), ("baseArrayLayer", u32), ("layerCount", u32)]


class ImageDesc(C.Structure):
    _fields_ = [("address", u64), ("format", u32), ("tiling", u32), ("width", u32), ("height", u32), ("depth", u32),
                ("array_layers", u32), ("mip_levels", u32), ("reserved", u32)]


class GraphicsPipelineState(C.Structure):
    _fields_ = [(n, u32) for n in (
        "vertex_program", "fragment_program", "cull_mode", "front_face", "depth_test_enable", "depth_write_enable",
        "depth_compare_op", "blend_enable", "src_color_blend_factor", "dst_color_blend_factor", "color_blend_op",
        "src_alpha_blend_factor", "dst_alpha_blend_factor", "alpha_blend_op", "color_write_mask", "logic_op", "topology", "reserved")]


class VertexInput(C.Structure):
    _fields_ = [("buffer_address", u64), ("buffer_size", u64), ("stride", u32), ("position_offset", u32),
                ("position_components", u32), ("color_offset", u32)]


class Framebuffer(C.Structure):
    _fields_ = [("color", ImageDesc), ("depth", ImageDesc), ("color_layer", u32), ("depth_layer", u32),
                ("band_y0", u32), ("band_y1", u32)]


TOPOLOGY_LINE_LIST, TOPOLOGY_LINE_STRIP, TOPOLOGY_TRIANGLE_LIST, TOPOLOGY_TRIANGLE_STRIP, TOPOLOGY_TRIANGLE_FAN = 1, 2, 3, 4, 5
TOPOLOGY_POINT_LIST = 0x10  # VK_PRIMITIVE_TOPOLOGY_POINT_LIST is 0, which KznGraphicsPipelineState reads as "default"
LOAD_OP_LOAD, LOAD_OP_CLEAR = 0, 1


class RenderPassBegin(C.Structure):  # KznRenderPassBegin
    _fields_ = [("framebuffer", Framebuffer), ("render_area", Rect2D), ("color_load_op", u32), ("depth_load_op", u32),
                ("clear_color", ClearColorValue), ("clear_depth_stencil", ClearDepthStencilValue)]


class DrawStatistics(C.Structure):
    _fields_ = [(n, u64) for n in (
        "triangles_in", "triangles_culled", "triangles_binned", "bin_entries", "fragments_covered",
        "fragments_depth_passed", "tiles_launched", "bin_overflow_tiles", "triangles_clipped", "clip_triangles", "triangles_discarded",
        "clip_overflow")]


class SwapchainCreateInfo(C.Structure):  # KznSwapchainCreateInfo
    _fields_ = [("width", u32), ("height", u32), ("image_count", u32), ("window_byte_order", u32), ("scanline_alignment", u32),
                ("reserved", u32)]


class SwapchainSegment(C.Structure):  # KznSwapchainSegment
    _fields_ = [("shm_id", i32), ("bytes_per_line", u32), ("size", u64), ("mapped", C.c_void_p)]


WINDOW_R8G8B8A8, WINDOW_B8G8R8A8 = 0, 1
MAX_SWAPCHAIN_IMAGES = 16

P = C.POINTER
vp = C.c_void_p

# name -> (restype, argtypes).  Opaque handles cross as void*.
PROTOTYPES = {
    "kzn_abi_version": (u32, []),
    "kzn_last_error": (C.c_char_p, []),
    "kzn_image_memory_requirements": (i32, [P(ImageDesc), P(u64), P(u64)]),
    "kzn_device_count": (i32, [P(i32)]),
    "kzn_device_create": (i32, [i32, P(vp)]),
    "kzn_device_destroy": (None, [vp]),
    "kzn_device_wait_idle": (i32, [vp]),
    "kzn_device_sm_count": (i32, [vp]),
    "kzn_device_ordinal": (i32, [vp]),
    "kzn_memory_allocate": (i32, [vp, u64, u32, P(vp)]),
    "kzn_memory_free": (None, [vp]),
    "kzn_memory_device_address": (u64, [vp]),
    "kzn_memory_size": (u64, [vp]),
    "kzn_memory_map": (i32, [vp, u64, P(vp)]),
    "kzn_memory_export": (i32, [vp, P(C.c_uint8)]),
    "kzn_memory_import": (i32, [vp, P(C.c_uint8), u64, P(vp)]),
    "kzn_cmd_signal_flag": (i32, [vp, u64, u64]),
    "kzn_cmd_wait_flag": (i32, [vp, u64, u64, u64]),
    "kzn_queue_create": (i32, [vp, P(vp)]),
    "kzn_queue_wrap_stream": (i32, [vp, vp, P(vp)]),
    "kzn_queue_destroy": (None, [vp]),
    "kzn_queue_wait_idle": (i32, [vp]),
    "kzn_queue_stream": (vp, [vp]),
    "kzn_queue_launch_count": (u64, [vp]),
    "kzn_queue_set_kernel_timing": (i32, [vp, i32]),
    "kzn_queue_last_kernel_ms": (i32, [vp, P(f32)]),
    "kzn_queue_upload": (i32, [vp, u64, vp, u64]),
    "kzn_queue_download": (i32, [vp, vp, u64, u64]),
    "kzn_host_alloc_pinned": (i32, [u64, P(vp)]),
    "kzn_host_free_pinned": (None, [vp]),
    "kzn_command_buffer_begin": (i32, [vp, P(vp)]),
    "kzn_command_buffer_end": (i32, [vp]),
    "kzn_command_buffer_destroy": (None, [vp]),
    "kzn_command_buffer_launch_count": (u64, [vp]),
    "kzn_queue_submit": (i32, [vp, u32, P(vp), vp]),
    "kzn_queue_reserve_draw": (i32, [vp, u64, u32, u32]),
    "kzn_semaphore_create": (i32, [vp, P(vp)]),
    "kzn_semaphore_destroy": (None, [vp]),
    "kzn_queue_signal_semaphore": (i32, [vp, vp]),
    "kzn_queue_wait_semaphore": (i32, [vp, vp]),
    "kzn_queue_copy_dma": (i32, [vp, u64, u64, u64]),
    "kzn_cmd_wait_flags": (i32, [vp, u64, u32, u64, u64]),
    "kzn_calibrate_fp32_issue": (i32, [vp, P(f32)]),
    "kzn_cmd_pipeline_barrier": (i32, [vp]),
    "kzn_fence_create": (i32, [vp, i32, P(vp)]),
    "kzn_fence_destroy": (None, [vp]),
    "kzn_fence_reset": (i32, [vp]),
    "kzn_queue_signal_fence": (i32, [vp, vp]),
    "kzn_fence_status": (i32, [vp]),
    "kzn_fence_wait": (i32, [vp, u64]),
    "kzn_cmd_clear_color_image": (i32, [vp, P(ImageDesc), P(ClearColorValue), u32, P(ImageSubresourceRange)]),
    "kzn_cmd_clear_depth_stencil_image": (i32, [vp, P(ImageDesc), P(ClearDepthStencilValue), u32,
                                                P(ImageSubresourceRange)]),
    "kzn_cmd_clear_attachment": (i32, [vp, P(ImageDesc), P(ClearColorValue), u32, P(ClearRect)]),
    "kzn_cmd_begin_render_pass": (i32, [vp, P(RenderPassBegin)]),
    "kzn_cmd_end_render_pass": (i32, [vp]),
    "kzn_cmd_copy_buffer_to_image": (i32, [vp, u64, u64, P(ImageDesc), u32, P(BufferImageCopy)]),
    "kzn_cmd_copy_image_to_buffer": (i32, [vp, P(ImageDesc), u64, u64, u32, P(BufferImageCopy)]),
    "kzn_cmd_copy_buffer": (i32, [vp, u64, u64, u64, u64, u32, P(BufferCopy)]),
    "kzn_cmd_fill_buffer": (i32, [vp, u64, u64, u64, u64, u32]),
    "kzn_cmd_update_buffer": (i32, [vp, u64, u64, u64, u64, vp]),
    "kzn_cmd_copy_image": (i32, [vp, P(ImageDesc), P(ImageDesc), u32, P(ImageCopy)]),
    "kzn_cmd_blit_image": (i32, [vp, P(ImageDesc), P(ImageDesc), u32, P(ImageBlit), u32]),
    "kzn_cmd_convert_tiling": (i32, [vp, P(ImageDesc), P(ImageDesc)]),
    "kzn_cmd_draw": (i32, [vp, P(GraphicsPipelineState), P(Framebuffer), P(Viewport), P(Rect2D), P(VertexInput),
                           u32, u32, u32, u32]),
    "kzn_cmd_draw_indexed": (i32, [vp, P(GraphicsPipelineState), P(Framebuffer), P(Viewport), P(Rect2D),
                                   P(VertexInput), u64, u64, u32, u32, u32, u32, i32, u32]),
    "kzn_queue_set_draw_statistics": (i32, [vp, u64]),
    "kzn_split_frame_create": (i32, [u32, P(i32), u32, u32, u32, P(vp)]),
    "kzn_split_frame_destroy": (None, [vp]),
    "kzn_split_frame_device_count": (u32, [vp]),
    "kzn_split_frame_device": (vp, [vp, u32]),
    "kzn_split_frame_queue": (vp, [vp, u32]),
    "kzn_split_frame_framebuffer": (i32, [vp, u32, P(Framebuffer)]),
    "kzn_split_frame_composite": (i32, [vp]),
    "kzn_split_frame_present_address": (u64, [vp]),
    "kzn_split_frame_present_queue": (vp, [vp]),
    "kzn_split_frame_wait_idle": (i32, [vp]),
    "kzn_swapchain_create": (i32, [vp, P(SwapchainCreateInfo), P(vp)]),
    "kzn_swapchain_destroy": (None, [vp]),
    "kzn_swapchain_get_image": (i32, [vp, u32, P(ImageDesc)]),
    "kzn_swapchain_get_segment": (i32, [vp, u32, P(SwapchainSegment)]),
    "kzn_swapchain_acquire_next_image": (i32, [vp, u64, P(u32)]),
    "kzn_queue_present": (i32, [vp, vp, u32]),
    "kzn_swapchain_wait_presented": (i32, [vp, u32]),
}


class LaunchDims(C.Structure):  # kazan_b200_backend.h KznLaunchDims
    _fields_ = [("grid", u32 * 3), ("block", u32 * 3)]


CALLABLE_FN = C.CFUNCTYPE(i32, vp, vp, P(LaunchDims), P(vp))


class Callable(C.Structure):  # kazan_b200_backend.h KznCallable
    _fields_ = [("call", CALLABLE_FN), ("self", vp)]


sz = C.c_size_t
cp = C.c_char_p
i64 = C.c_int64


class DescriptorSetLayoutBinding(C.Structure):  # kazan_b200_backend.h KznDescriptorSetLayoutBinding
    _fields_ = [("binding", u32), ("descriptor_type", u32), ("descriptor_count", u32)]


class DescriptorBufferInfo(C.Structure):  # kazan_b200_backend.h KznDescriptorBufferInfo
    _fields_ = [("address", u64), ("offset", u64), ("range", u64)]


DESCRIPTOR_TYPE_UNIFORM_BUFFER, DESCRIPTOR_TYPE_STORAGE_BUFFER = 6, 7
PIPELINE_BIND_POINT_GRAPHICS, PIPELINE_BIND_POINT_COMPUTE = 0, 1
IR_VERSION_1, IR_VERSION_2 = 1, 2
# include/kazan_b200_backend.h
BACKEND_PROTOTYPES = {
    "kzn_ir_parse": (i32, [cp, sz, P(vp)]),
    "kzn_ir_parse_ex": (i32, [cp, sz, u32, P(vp)]),
    "kzn_ir_module_destroy": (None, [vp]),
    "kzn_descriptor_set_create": (i32, [vp, u32, P(DescriptorSetLayoutBinding), P(vp)]),
    "kzn_descriptor_set_destroy": (None, [vp]),
    "kzn_descriptor_set_write_buffers": (i32, [vp, u32, u32, u32, P(DescriptorBufferInfo)]),
    "kzn_descriptor_set_table_address": (u64, [vp]),
    "kzn_descriptor_set_slot": (i64, [vp, u32]),
    "kzn_cmd_bind_descriptor_set": (i32, [vp, u32, vp]),
    "kzn_cmd_push_constants": (i32, [vp, u32, u32, vp]),
    "kzn_fragment_program_create": (i32, [vp, vp, P(vp)]),
    "kzn_fragment_program_destroy": (None, [vp]),
    "kzn_cmd_bind_fragment_program": (i32, [vp, vp]),
    "kzn_vertex_program_create": (i32, [vp, vp, P(vp)]),
    "kzn_vertex_program_destroy": (None, [vp]),
    "kzn_cmd_bind_vertex_program": (i32, [vp, vp]),
    "kzn_vertex_program_inputs_size": (u64, [vp]),
    "kzn_vertex_program_outputs_size": (u64, [vp]),
    "kzn_queue_reserve_vertex_outputs": (i32, [vp, u64, vp]),
    "kzn_fragment_program_ptx": (i32, [vp, P(vp), P(sz)]),
    "kzn_fragment_program_module_ptx": (i32, [vp, P(vp), P(sz)]),
    "kzn_ir_to_text": (i32, [vp, P(vp), P(sz)]),
    "kzn_ir_round_trip": (i32, [cp, cp, sz, P(vp), P(sz)]),
    "kzn_free": (None, [vp]),
    "kzn_translate_spirv_to_ir": (i32, [P(u32), sz, cp, u32, P(vp), P(u32)]),
    "kzn_translate_spirv_to_ir_ex": (i32, [P(u32), sz, cp, u32, u32, P(vp), P(u32)]),
    "kzn_spirv_disassemble": (i32, [P(u32), sz, P(vp), P(sz)]),
    "kzn_ptx_from_ir": (i32, [vp, P(vp), P(sz)]),
    "kzn_compute_pipeline_create": (i32, [vp, vp, P(u32), P(vp)]),
    "kzn_pipeline_destroy": (None, [vp]),
    "kzn_cmd_dispatch": (i32, [vp, vp, P(u32), P(u32), u64, u64, u64]),
    "kzn_pipeline_inputs_size": (u64, [vp]),
    "kzn_pipeline_outputs_size": (u64, [vp]),
    "kzn_backend_compiler_name": (cp, []),
    "kzn_backend_context_create": (i32, [P(vp)]),
    "kzn_backend_context_destroy": (None, [vp]),
    "kzn_backend_type_bool": (vp, [vp]),
    "kzn_backend_type_int": (vp, [vp, u32, i32]),
    "kzn_backend_type_f32": (vp, [vp]),
    "kzn_backend_type_f64": (vp, [vp]),
    "kzn_backend_type_pointer": (vp, [vp, vp]),
    "kzn_backend_type_array": (vp, [vp, vp, u64]),
    "kzn_backend_type_vector": (vp, [vp, vp, u64]),
    "kzn_backend_type_struct": (vp, [vp, P(vp), u64]),
    "kzn_backend_type_function": (vp, [vp, vp, P(vp), u64]),
    "kzn_backend_type_size": (u64, [vp]),
    "kzn_backend_create_module": (vp, [vp, cp]),
    "kzn_backend_module_set_source_file_name": (i32, [vp, cp]),
    "kzn_backend_add_function": (vp, [vp, cp, vp]),
    "kzn_backend_append_new_basic_block": (vp, [vp, cp]),
    "kzn_backend_function_parameter_count": (u64, [vp]),
    "kzn_backend_function_parameter": (vp, [vp, u64]),
    "kzn_backend_function_as_value": (vp, [vp]),
    "kzn_backend_block_as_value": (vp, [vp]),
    "kzn_backend_create_builder": (vp, [vp]),
    "kzn_backend_builder_attach": (i32, [vp, vp]),
    "kzn_backend_builder_current_basic_block": (vp, [vp]),
    "kzn_backend_build_alloca": (vp, [vp, vp]),
    "kzn_backend_build_return": (i32, [vp, vp]),
    "kzn_backend_module_verify": (i32, [vp]),
    "kzn_backend_module_to_ptx": (i32, [vp, P(vp), P(sz)]),
    "kzn_backend_compile": (i32, [vp, vp, P(cp), P(vp), u64, P(vp)]),
    "kzn_compiled_code_get": (i32, [vp, cp, P(Callable)]),
    "kzn_compiled_code_destroy": (None, [vp]),
}
ALL_PROTOTYPES = {**PROTOTYPES, **BACKEND_PROTOTYPES}


def declared_symbols(header: Path | None = None) -> list[str]:
    """Every function the C headers declare with KZN_API (what the library must export)."""
    out = []
    for h in ((header,) if header else HEADERS):
        text = h.read_text()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        out += re.findall(r"KZN_API\s+[^;(]*?\b(kzn_\w+)\s*\(", text)
    return out


def bind(lib: C.CDLL, prototypes: dict | None = None) -> C.CDLL:
    for name, (res, args) in (prototypes or ALL_PROTOTYPES).items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib
