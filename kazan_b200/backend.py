"""Host-side mirror of `shader-compiler-backend-ptx` (include/kazan_b200_backend.h) for tests and bench.

Names follow the reference: IrModule <-> shader_compiler_ir::Module (module.rs:128-150, text.rs FromText/ToText),
translate_spirv_to_ir <-> TranslatedSPIRVShader::new (shader-compiler-translate-spirv-to-ir/src/lib.rs:190-215),
ComputePipeline <-> vulkan-driver/src/pipeline.rs:40-43, Queue.cmd_dispatch <-> vkCmdDispatch (api_impl.rs:4677),
Context / Module / Function / Builder / TypeBuilder <-> shader-compiler-backend/src/lib.rs:15-257.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

from . import abi
from .host import Device, KazanError, Queue, _check, load_library

EXECUTION_MODEL_VERTEX, EXECUTION_MODEL_FRAGMENT, EXECUTION_MODEL_GLCOMPUTE = 0, 4, 5


def _take_text(ptr: C.c_void_p, length: C.c_size_t) -> str:
    lib = load_library()
    try:
        return C.string_at(ptr.value, length.value).decode()
    finally:
        lib.kzn_free(ptr)


def _fail(code: int):
    raise KazanError(code, load_library().kzn_last_error().decode())


def round_trip(kind: str, text: str) -> str:
    """Parse one syntactic category and print it canonically (what the reference's text unit tests check)."""
    out, n = C.c_void_p(), C.c_size_t()
    b = text.encode()
    _check(load_library().kzn_ir_round_trip(kind.encode(), b, len(b), C.byref(out), C.byref(n)))
    return _take_text(out, n)


class IrModule:
    def __init__(self, handle):
        self.handle, self._lib = handle, load_library()

    @classmethod
    def parse(cls, text: str, version: int = 1) -> "IrModule":
        """version 1: the reference's seven instructions; version 2: the C++ mirror's additions (kazan_b200_backend.h)."""
        h = C.c_void_p()
        b = text.encode()
        if version == 1:
            _check(load_library().kzn_ir_parse(b, len(b), C.byref(h)))
        else:
            _check(load_library().kzn_ir_parse_ex(b, len(b), version, C.byref(h)))
        return cls(h)

    def to_text(self) -> str:
        out, n = C.c_void_p(), C.c_size_t()
        _check(self._lib.kzn_ir_to_text(self.handle, C.byref(out), C.byref(n)))
        return _take_text(out, n)

    def to_ptx(self) -> str:
        out, n = C.c_void_p(), C.c_size_t()
        _check(self._lib.kzn_ptx_from_ir(self.handle, C.byref(out), C.byref(n)))
        return _take_text(out, n)

    def destroy(self):
        if self.handle:
            self._lib.kzn_ir_module_destroy(self.handle)
            self.handle = None


def translate_spirv_to_ir(words: Sequence[int], entry_point: str, execution_model: int, version: int = 1) -> tuple[IrModule, tuple[int, int, int]]:
    """version 1: what the reference translates (control-flow-free entry points); version 2: shaders with a body, into the
    version 2 instructions of the IR (kazan_b200/csrc/spirv_to_ir_v2.cpp)."""
    arr = (abi.u32 * len(words))(*words)
    h = C.c_void_p()
    ls = (abi.u32 * 3)()
    if version == 1:
        _check(load_library().kzn_translate_spirv_to_ir(arr, len(words), entry_point.encode(), execution_model, C.byref(h), ls))
    else:
        _check(load_library().kzn_translate_spirv_to_ir_ex(arr, len(words), entry_point.encode(), execution_model, version, C.byref(h), ls))
    return IrModule(h), tuple(ls)


def disassemble_spirv(words: Sequence[int]) -> str:
    """spirv-parser's parse-and-dump (spirv-parser/src/lib.rs:26-44): header block + one line per instruction."""
    arr = (abi.u32 * len(words))(*words)
    out, n = C.c_void_p(), C.c_size_t()
    _check(load_library().kzn_spirv_disassemble(arr, len(words), C.byref(out), C.byref(n)))
    return _take_text(out, n)


class ComputePipeline:
    def __init__(self, device: Device, module: IrModule, local_size=(1, 1, 1)):
        self._lib, self.device = load_library(), device
        h = C.c_void_p()
        ls = (abi.u32 * 3)(*local_size)
        _check(self._lib.kzn_compute_pipeline_create(device.handle, module.handle, ls, C.byref(h)))
        self.handle = h
        self.local_size = tuple(local_size)

    @property
    def inputs_size(self) -> int:
        return self._lib.kzn_pipeline_inputs_size(self.handle)

    @property
    def outputs_size(self) -> int:
        return self._lib.kzn_pipeline_outputs_size(self.handle)

    def destroy(self):
        if self.handle:
            self._lib.kzn_pipeline_destroy(self.handle)
            self.handle = None


def cmd_dispatch(queue: Queue, pipeline: ComputePipeline, group_count, user_inputs=0, user_outputs=0, base_group=(0, 0, 0),
                 first_invocation=0):
    """vkCmdDispatchBase (api_impl.rs:5079-5089); vkCmdDispatch when base_group is zero."""
    base = (abi.u32 * 3)(*base_group)
    cnt = (abi.u32 * 3)(*group_count)
    _check(load_library().kzn_cmd_dispatch(queue.handle, pipeline.handle, base, cnt, user_inputs, user_outputs, first_invocation))


Queue.cmd_dispatch = cmd_dispatch  # same spelling as the other vkCmd* mirrors


class DescriptorSet:
    """vulkan-driver/src/descriptor_set.rs DescriptorSet + its layout: bindings = [(binding, descriptor_type, count), ...]."""

    def __init__(self, device: Device, bindings):
        self._lib = load_library()
        arr = (abi.DescriptorSetLayoutBinding * max(len(bindings), 1))(*[abi.DescriptorSetLayoutBinding(*b) for b in bindings])
        h = C.c_void_p()
        _check(self._lib.kzn_descriptor_set_create(device.handle, len(bindings), arr, C.byref(h)))
        self.handle = h

    def write_buffers(self, binding: int, buffers, first_element: int = 0):
        """vkUpdateDescriptorSets with pBufferInfo; buffers = [(address, offset, range), ...]"""
        arr = (abi.DescriptorBufferInfo * max(len(buffers), 1))(*[abi.DescriptorBufferInfo(*b) for b in buffers])
        _check(self._lib.kzn_descriptor_set_write_buffers(self.handle, binding, first_element, len(buffers), arr))

    @property
    def table_address(self) -> int:
        return self._lib.kzn_descriptor_set_table_address(self.handle)

    def slot(self, binding: int) -> int:
        return self._lib.kzn_descriptor_set_slot(self.handle, binding)

    def destroy(self):
        if self.handle:
            self._lib.kzn_descriptor_set_destroy(self.handle)
            self.handle = None


def cmd_bind_descriptor_set(queue: Queue, bind_point: int, descriptor_set: "DescriptorSet | None"):
    """vkCmdBindDescriptorSets (api_impl.rs:4597)"""
    _check(load_library().kzn_cmd_bind_descriptor_set(queue.handle, bind_point, descriptor_set.handle if descriptor_set else None))


def cmd_push_constants(queue: Queue, data: bytes, offset: int = 0):
    """vkCmdPushConstants (api_impl.rs:4932)"""
    buf = C.create_string_buffer(bytes(data), len(data))
    _check(load_library().kzn_cmd_push_constants(queue.handle, offset, len(data), buf))


class FragmentProgram:
    """The fragment stage of a graphics pipeline (GraphicsPipeline::create, pipeline.rs:167-183): a shader-compiler-ir module
    compiled into the raster kernels."""

    def __init__(self, device: Device, module: IrModule):
        self._lib = load_library()
        h = C.c_void_p()
        _check(self._lib.kzn_fragment_program_create(device.handle, module.handle, C.byref(h)))
        self.handle = h

    @staticmethod
    def call_site_ptx(module: IrModule) -> str:
        out, n = C.c_void_p(), C.c_size_t()
        _check(load_library().kzn_fragment_program_ptx(module.handle, C.byref(out), C.byref(n)))
        return _take_text(out, n)

    @staticmethod
    def module_ptx(module: IrModule) -> str:
        out, n = C.c_void_p(), C.c_size_t()
        _check(load_library().kzn_fragment_program_module_ptx(module.handle, C.byref(out), C.byref(n)))
        return _take_text(out, n)

    def destroy(self):
        if self.handle:
            self._lib.kzn_fragment_program_destroy(self.handle)
            self.handle = None


class VertexProgram:
    """The vertex stage of a graphics pipeline: a vertex shader in version 2 of the IR, run by the PTX back end over the
    vertices of a draw before setup (state.vertex_program = VS_COMPILED)."""

    def __init__(self, device: Device, module: IrModule):
        self._lib = load_library()
        h = C.c_void_p()
        _check(self._lib.kzn_vertex_program_create(device.handle, module.handle, C.byref(h)))
        self.handle = h

    @property
    def inputs_size(self) -> int:
        return self._lib.kzn_vertex_program_inputs_size(self.handle)

    @property
    def outputs_size(self) -> int:
        return self._lib.kzn_vertex_program_outputs_size(self.handle)

    def destroy(self):
        if self.handle:
            self._lib.kzn_vertex_program_destroy(self.handle)
            self.handle = None


def cmd_bind_vertex_program(queue: Queue, program: "VertexProgram | None"):
    """vkCmdBindPipeline (api_impl.rs:4506) for the vertex stage"""
    _check(load_library().kzn_cmd_bind_vertex_program(queue.handle, program.handle if program else None))


def reserve_vertex_outputs(queue: Queue, max_vertices: int, program: VertexProgram):
    _check(load_library().kzn_queue_reserve_vertex_outputs(queue.handle, max_vertices, program.handle))


def cmd_bind_fragment_program(queue: Queue, program: "FragmentProgram | None"):
    """vkCmdBindPipeline (api_impl.rs:4506) for the fragment stage"""
    _check(load_library().kzn_cmd_bind_fragment_program(queue.handle, program.handle if program else None))


Queue.cmd_bind_fragment_program = cmd_bind_fragment_program
Queue.cmd_bind_vertex_program = cmd_bind_vertex_program
Queue.reserve_vertex_outputs = reserve_vertex_outputs
Queue.cmd_bind_descriptor_set = cmd_bind_descriptor_set
Queue.cmd_push_constants = cmd_push_constants


# ---- the backend trait surface -------------------------------------------------------------------------------------------

class Context:
    """trait Context (lib.rs:162-189) + TypeBuilder (types.rs:113-153)."""

    def __init__(self):
        self._lib = load_library()
        h = C.c_void_p()
        _check(self._lib.kzn_backend_context_create(C.byref(h)))
        self.handle = h

    def _t(self, p):
        if not p:
            _fail(abi.ERROR_VALIDATION_FAILED)
        return C.c_void_p(p)

    # TypeBuilder
    def build_bool(self): return self._t(self._lib.kzn_backend_type_bool(self.handle))
    def build_int(self, bits, signed=True): return self._t(self._lib.kzn_backend_type_int(self.handle, bits, int(signed)))
    def build_i8(self): return self.build_int(8)
    def build_i16(self): return self.build_int(16)
    def build_i32(self): return self.build_int(32)
    def build_i64(self): return self.build_int(64)
    def build_u8(self): return self.build_int(8, False)
    def build_u16(self): return self.build_int(16, False)
    def build_u32(self): return self.build_int(32, False)
    def build_u64(self): return self.build_int(64, False)
    def build_f32(self): return self._t(self._lib.kzn_backend_type_f32(self.handle))
    def build_f64(self): return self._t(self._lib.kzn_backend_type_f64(self.handle))
    def build_pointer(self, target): return self._t(self._lib.kzn_backend_type_pointer(self.handle, target))
    def build_array(self, element, count): return self._t(self._lib.kzn_backend_type_array(self.handle, element, count))
    def build_vector(self, element, length): return self._t(self._lib.kzn_backend_type_vector(self.handle, element, length))

    def build_struct(self, members):
        arr = (C.c_void_p * max(len(members), 1))(*[m.value for m in members])
        return self._t(self._lib.kzn_backend_type_struct(self.handle, arr, len(members)))

    def build_function(self, arguments, return_type=None):
        arr = (C.c_void_p * max(len(arguments), 1))(*[a.value for a in arguments])
        return self._t(self._lib.kzn_backend_type_function(self.handle, return_type, arr, len(arguments)))

    def type_size(self, t) -> int:
        return self._lib.kzn_backend_type_size(t)

    def create_module(self, name: str) -> "Module":
        return Module(self, self._t(self._lib.kzn_backend_create_module(self.handle, name.encode())))

    def create_builder(self) -> "Builder":
        return Builder(self, self._t(self._lib.kzn_backend_create_builder(self.handle)))

    def destroy(self):
        if self.handle:
            self._lib.kzn_backend_context_destroy(self.handle)
            self.handle = None


class Module:
    def __init__(self, context: Context, handle):
        self.context, self.handle, self._lib = context, handle, context._lib

    def set_source_file_name(self, name: str):
        _check(self._lib.kzn_backend_module_set_source_file_name(self.handle, name.encode()))

    def add_function(self, name: str, function_type) -> "Function":
        p = self._lib.kzn_backend_add_function(self.handle, name.encode(), function_type)
        if not p:
            _fail(abi.ERROR_VALIDATION_FAILED)
        return Function(self, C.c_void_p(p), name)

    def verify(self) -> "Module":
        _check(self._lib.kzn_backend_module_verify(self.handle))
        return self

    def to_ptx(self) -> str:
        out, n = C.c_void_p(), C.c_size_t()
        _check(self._lib.kzn_backend_module_to_ptx(self.handle, C.byref(out), C.byref(n)))
        return _take_text(out, n)


class Function:
    def __init__(self, module: Module, handle, name: str):
        self.module, self.handle, self.name, self._lib = module, handle, name, module._lib

    def append_new_basic_block(self, name: str | None = None):
        p = self._lib.kzn_backend_append_new_basic_block(self.handle, name.encode() if name is not None else None)
        return C.c_void_p(p)

    def parameters(self):
        n = self._lib.kzn_backend_function_parameter_count(self.handle)
        return [C.c_void_p(self._lib.kzn_backend_function_parameter(self.handle, i)) for i in range(n)]


class Builder:
    """DetachedBuilder until attach(); build_return detaches again (lib.rs:15-41)."""

    def __init__(self, context: Context, handle):
        self.context, self.handle, self._lib = context, handle, context._lib

    def attach(self, basic_block) -> "Builder":
        _check(self._lib.kzn_backend_builder_attach(self.handle, basic_block))
        return self

    def build_alloca(self, variable_type):
        p = self._lib.kzn_backend_build_alloca(self.handle, variable_type)
        if not p:
            _fail(abi.ERROR_VALIDATION_FAILED)
        return C.c_void_p(p)

    def build_return(self, value=None) -> "Builder":
        _check(self._lib.kzn_backend_build_return(self.handle, value))
        return self


class CompiledCode:
    """Compiler::run + CompiledCode::get (lib.rs:200-205,252-256)."""

    def __init__(self, device: Device, module: Module, callable_functions: dict[str, Function]):
        self._lib = load_library()
        keys = list(callable_functions)
        karr = (C.c_char_p * max(len(keys), 1))(*[k.encode() for k in keys])
        farr = (C.c_void_p * max(len(keys), 1))(*[callable_functions[k].handle.value for k in keys])
        h = C.c_void_p()
        _check(self._lib.kzn_backend_compile(device.handle, module.handle, karr, farr, len(keys), C.byref(h)))
        self.handle = h

    def get(self, key: str):
        """Returns launch(queue, grid, block, *ctypes_arguments) or None when the key is unknown."""
        c = abi.Callable()
        code = self._lib.kzn_compiled_code_get(self.handle, key.encode(), C.byref(c))
        if code != 0:
            return None

        def launch(queue: Queue, grid=(1, 1, 1), block=(1, 1, 1), *arguments):
            dims = abi.LaunchDims((abi.u32 * 3)(*grid), (abi.u32 * 3)(*block))
            args = (C.c_void_p * max(len(arguments), 1))(*[C.cast(C.pointer(a), C.c_void_p).value for a in arguments])
            _check(c.call(c.self, queue.handle, C.byref(dims), args))
        launch._keepalive = c
        return launch

    def destroy(self):
        if self.handle:
            self._lib.kzn_compiled_code_destroy(self.handle)
            self.handle = None
