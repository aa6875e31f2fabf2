"""Shared helpers for the parity tests: device buffers through the C ABI and oracle-side twins."""
from __future__ import annotations

import ctypes as C

import numpy as np

from kazan_b200 import abi, host
from oracle import binding as ob


class GpuBuffer:
    def __init__(self, ctx: "GpuContext", nbytes: int, memory_type=abi.MEMORY_TYPE_DEVICE_LOCAL):
        self.ctx = ctx
        self.nbytes = nbytes
        self.memory = ctx.device.allocate(max(nbytes, 4), memory_type)
        self.address = self.memory.address

    def upload(self, array: np.ndarray):
        a = np.ascontiguousarray(array).view(np.uint8).reshape(-1)
        assert a.nbytes <= self.nbytes
        self.ctx.queue.upload(self.address, a.ctypes.data, a.nbytes)
        self.ctx.queue.wait_idle()
        return self

    def download(self) -> np.ndarray:
        out = np.empty(self.nbytes, dtype=np.uint8)
        self.ctx.queue.wait_idle()
        self.ctx.queue.download(out.ctypes.data, self.address, self.nbytes)
        self.ctx.queue.wait_idle()
        return out

    def free(self):
        self.memory.free()


class GpuContext:
    def __init__(self, ordinal: int = 0):
        self.device = host.Device(ordinal)
        self.queue = self.device.create_queue()
        self._buffers: list[GpuBuffer] = []

    def buffer(self, nbytes: int, init: np.ndarray | None = None, **kw) -> GpuBuffer:
        b = GpuBuffer(self, nbytes, **kw)
        self._buffers.append(b)
        if init is not None:
            b.upload(init)
        return b

    def release(self):
        for b in self._buffers:
            b.free()
        self._buffers.clear()

    def close(self):
        self.release()
        self.queue.destroy()
        self.device.destroy()


def image_size(width, height, depth=1, layers=1, tiling=abi.TILING_LINEAR, fmt=abi.FORMAT_R8G8B8A8_UNORM) -> int:
    size, _ = host.Image(0x1000, width, height, depth, layers, fmt, tiling).memory_requirements()
    return size


def host_image(array: np.ndarray, width, height, depth=1, layers=1, tiling=abi.TILING_LINEAR,
               fmt=abi.FORMAT_R8G8B8A8_UNORM) -> abi.ImageDesc:
    """ImageDesc whose address is the numpy buffer (oracle side)."""
    return abi.ImageDesc(ob.ptr(array), fmt, tiling, width, height, depth, layers, 1, 0)


def check(code: int):
    assert code == 0, abi.RESULT_NAMES.get(code, code)


def arr(ctype, items):
    return ob.array_of(ctype, items)


def blit_region(src0, src1, dst0, dst1, src_layer=0, dst_layer=0, layers=1) -> abi.ImageBlit:
    b = abi.ImageBlit()
    b.srcSubresource = abi.ImageSubresourceLayers(abi.ASPECT_COLOR, 0, src_layer, layers)
    b.dstSubresource = abi.ImageSubresourceLayers(abi.ASPECT_COLOR, 0, dst_layer, layers)
    b.srcOffsets[0] = abi.Offset3D(*src0)
    b.srcOffsets[1] = abi.Offset3D(*src1)
    b.dstOffsets[0] = abi.Offset3D(*dst0)
    b.dstOffsets[1] = abi.Offset3D(*dst1)
    return b


def image_copy(extent, src_off=(0, 0, 0), dst_off=(0, 0, 0), src_layer=0, dst_layer=0, layers=1,
               aspect=abi.ASPECT_COLOR) -> abi.ImageCopy:
    return abi.ImageCopy(abi.ImageSubresourceLayers(aspect, 0, src_layer, layers), abi.Offset3D(*src_off),
                         abi.ImageSubresourceLayers(aspect, 0, dst_layer, layers), abi.Offset3D(*dst_off),
                         abi.Extent3D(*extent))


# ---- draw scenes (shared by oracle tests, GPU parity tests, golden generation, bench) ----------------------

def pipeline_state(fs=abi.FS_VERTEX_COLOR, cull=abi.CULL_NONE, front=abi.FRONT_FACE_CCW, depth=None, blend=None,
                   write_mask=15, logic_op=None, topology=0) -> abi.GraphicsPipelineState:
    """logic_op: a VkLogicOp 0..15 (logicOpEnable = TRUE) or None; topology: 0 / 3 list, 4 strip, 5 fan, 1 / 2 line list / strip,
    abi.TOPOLOGY_POINT_LIST."""
    st = abi.GraphicsPipelineState()
    st.topology = topology
    st.logic_op = 0 if logic_op is None else 1 + logic_op
    st.fragment_program = fs
    st.cull_mode = cull
    st.front_face = front
    st.color_write_mask = write_mask
    if depth is not None:
        st.depth_test_enable = 1
        st.depth_compare_op, st.depth_write_enable = depth
    if blend is not None:
        st.blend_enable = 1
        (st.src_color_blend_factor, st.dst_color_blend_factor, st.color_blend_op, st.src_alpha_blend_factor,
         st.dst_alpha_blend_factor, st.alpha_blend_op) = blend
    return st


ALPHA_BLEND = (abi.BLEND_SRC_ALPHA, abi.BLEND_ONE_MINUS_SRC_ALPHA, abi.BLEND_OP_ADD,
               abi.BLEND_SRC_ALPHA, abi.BLEND_ONE_MINUS_SRC_ALPHA, abi.BLEND_OP_ADD)
ADDITIVE_BLEND = (abi.BLEND_ONE, abi.BLEND_ONE, abi.BLEND_OP_ADD, abi.BLEND_ONE, abi.BLEND_ONE, abi.BLEND_OP_ADD)


def grid_mesh(quads_x: int, quads_y: int, seed: int):
    """(vertices float32 [n,7], indices uint32) from the oracle's deterministic generator."""
    nv = (quads_x + 1) * (quads_y + 1)
    verts = np.zeros((nv, 7), dtype=np.float32)
    idx = np.zeros(6 * quads_x * quads_y, dtype=np.uint32)
    ob.load().kzo_make_grid_mesh(quads_x, quads_y, seed, C.c_void_p(ob.ptr(verts)), C.c_void_p(ob.ptr(idx)))
    return verts, idx


FULLSCREEN_TRIANGLE = np.array([[-1, -1, 0, 1, 1, 1, 1], [3, -1, 0, 1, 1, 1, 1], [-1, 3, 0, 1, 1, 1, 1]],
                               dtype=np.float32)


class Scene:
    """One draw: geometry + state + framebuffer size.  Can run on the oracle or on the GPU."""

    def __init__(self, name, width, height, vertices, indices=None, state=None, viewport=None, scissor=None,
                 use_depth=False, clear_color=(0, 0, 0, 0), clear_depth=1.0, color_tiling=abi.TILING_LINEAR,
                 index_type=abi.INDEX_TYPE_UINT32, pos_comps=3, color_offset=12, instances=1, band=(0, 0),
                 first=0, vertex_offset=0):
        self.name, self.width, self.height = name, width, height
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float32)
        self.indices = None if indices is None else np.ascontiguousarray(indices)
        self.state = state or pipeline_state()
        self.viewport = viewport or abi.Viewport(0, 0, width, height, 0, 1)
        self.scissor = scissor or abi.Rect2D(abi.Offset2D(0, 0), abi.Extent2D(width, height))
        self.use_depth, self.clear_color, self.clear_depth = use_depth, clear_color, clear_depth
        self.color_tiling, self.index_type = color_tiling, index_type
        self.pos_comps, self.color_offset, self.instances, self.band = pos_comps, color_offset, instances, band
        self.first, self.vertex_offset = first, vertex_offset
        self.position_offset = 0
        self.stride = self.vertices.shape[1] * 4 if self.vertices.ndim == 2 else 28

    @property
    def count(self):
        return (len(self.indices) if self.indices is not None else len(self.vertices)) - self.first

    @property
    def primitives(self):
        """primitives of one instance (Vulkan 1.1 "Primitive Topologies")"""
        n, topo = self.count, self.state.topology
        return {abi.TOPOLOGY_POINT_LIST: n, abi.TOPOLOGY_LINE_LIST: n // 2, abi.TOPOLOGY_LINE_STRIP: max(0, n - 1),
                0: n // 3, abi.TOPOLOGY_TRIANGLE_LIST: n // 3}.get(topo, max(0, n - 2))

    def _framebuffer(self, color_addr, depth_addr) -> abi.Framebuffer:
        fb = abi.Framebuffer()
        fb.color = abi.ImageDesc(color_addr, abi.FORMAT_R8G8B8A8_UNORM, self.color_tiling, self.width, self.height,
                                 1, 1, 1, 0)
        if self.use_depth:
            fb.depth = abi.ImageDesc(depth_addr, abi.FORMAT_D32_SFLOAT, 0, self.width, self.height, 1, 1, 1, 0)
        fb.band_y0, fb.band_y1 = self.band
        return fb

    def _initial(self):
        csize = image_size(self.width, self.height, tiling=self.color_tiling)
        color = ob.aligned_empty(csize)
        packed = np.uint32(ob.load().kzo_pack_rgba8((C.c_float * 4)(*self.clear_color)))
        color.view(np.uint32)[:] = packed
        depth = ob.aligned_empty(self.width * self.height * 4)
        depth.view(np.float32)[:] = self.clear_depth
        return color, depth

    def render_area(self):
        """The draw's clip rectangle without the band (scissor x viewport x framebuffer), as a render area: what a render pass
        around this one draw would name so that the draw can take its LOAD_OP_CLEAR over."""
        f = np.float32
        vx0, vx1 = f(self.viewport.x), f(self.viewport.x) + f(self.viewport.width)
        vy0, vy1 = sorted((f(self.viewport.y), f(self.viewport.y) + f(self.viewport.height)))
        x0 = max(0, self.scissor.offset.x, int(np.ceil(vx0 - f(0.5))))
        y0 = max(0, self.scissor.offset.y, int(np.ceil(vy0 - f(0.5))))
        x1 = min(self.width - 1, self.scissor.offset.x + self.scissor.extent.width - 1, int(np.floor(vx1 - f(0.5))))
        y1 = min(self.height - 1, self.scissor.offset.y + self.scissor.extent.height - 1, int(np.floor(vy1 - f(0.5))))
        return abi.Rect2D(abi.Offset2D(x0, y0), abi.Extent2D(max(0, x1 - x0 + 1), max(0, y1 - y0 + 1)))

    def _initial_render_pass(self, area: abi.Rect2D):
        """What the attachments hold after vkCmdBeginRenderPass with LOAD_OP_CLEAR on `area` (cut to the band) when they held
        a byte pattern before: the oracle's vkCmdClearAttachments on that rectangle."""
        lib = ob.load()
        color, depth = self._initial()
        color[:] = 0xA5
        depth.view(np.uint32)[:] = 0x3E99999A  # 0.3
        y0, y1 = area.offset.y, area.offset.y + area.extent.height
        if self.band[1]:
            y0, y1 = max(y0, self.band[0]), min(y1, self.band[1])
        if area.extent.width and y1 > y0:
            rect, n = arr(abi.ClearRect, [abi.ClearRect(abi.Rect2D(abi.Offset2D(area.offset.x, y0), abi.Extent2D(area.extent.width, y1 - y0)), 0, 1)])
            fb = self._framebuffer(ob.ptr(color), ob.ptr(depth))
            c = abi.ClearColorValue()
            c.float32[:] = list(self.clear_color)
            check(lib.kzo_clear_attachment(C.byref(fb.color), C.byref(c), n, rect))
            if self.use_depth:
                d = abi.ClearColorValue()
                d.float32[0] = self.clear_depth
                check(lib.kzo_clear_attachment(C.byref(fb.depth), C.byref(d), n, rect))
        return color, depth

    def run_oracle(self, threads: int = 0, render_pass: abi.Rect2D | None = None):
        lib = ob.load()
        if threads:
            lib.kzo_set_threads(threads)
        color, depth = self._initial() if render_pass is None else self._initial_render_pass(render_pass)
        fb = self._framebuffer(ob.ptr(color), ob.ptr(depth))
        vi = abi.VertexInput(ob.ptr(self.vertices), self.vertices.nbytes, self.stride, self.position_offset, self.pos_comps,
                             self.color_offset)
        stats = abi.DrawStatistics()
        if self.indices is not None:
            code = lib.kzo_draw_indexed(C.byref(self.state), C.byref(fb), C.byref(self.viewport),
                                        C.byref(self.scissor), C.byref(vi), C.c_void_p(ob.ptr(self.indices)),
                                        self.indices.nbytes, self.index_type, self.count, self.instances,
                                        self.first, self.vertex_offset, 0, C.byref(stats))
        else:
            code = lib.kzo_draw(C.byref(self.state), C.byref(fb), C.byref(self.viewport), C.byref(self.scissor),
                                C.byref(vi), self.count, self.instances, self.first, 0, C.byref(stats))
        check(code)
        return color, depth, stats

    def run_gpu(self, ctx: GpuContext, compiled=None, vertex_program=None, push_constants: bytes = b"", render_pass: abi.Rect2D | None = None):
        """compiled: {fragment program enum: backend.FragmentProgram} — the draw then runs the compiled form of the scene's
        built-in fragment program (state.fragment_program = KZN_FS_COMPILED, the program bound to the queue, the frame's
        1/width and 1/height as push constants, which is what fs_orbit.kazan-ir reads).
        vertex_program: a backend.VertexProgram — the draw runs it as its vertex stage (state.vertex_program = KZN_VS_COMPILED)
        with `push_constants`; the scene's vertices are then the shader's input block, one per stride.
        render_pass: the draw runs inside vkCmdBeginRenderPass(renderArea = render_pass, LOAD_OP_CLEAR) .. vkCmdEndRenderPass
        on attachments that hold a byte pattern (compare with run_oracle(render_pass=...))."""
        color0, depth0 = self._initial()
        if render_pass is not None:
            color0[:] = 0xA5
            depth0.view(np.uint32)[:] = 0x3E99999A
        cbuf = ctx.buffer(color0.nbytes, color0)
        dbuf = ctx.buffer(depth0.nbytes, depth0)
        vbuf = ctx.buffer(self.vertices.nbytes, self.vertices)
        sbuf = ctx.buffer(C.sizeof(abi.DrawStatistics), np.zeros(C.sizeof(abi.DrawStatistics) // 8, dtype=np.uint64))
        fb = self._framebuffer(cbuf.address, dbuf.address)
        vi = abi.VertexInput(vbuf.address, self.vertices.nbytes, self.stride, self.position_offset, self.pos_comps, self.color_offset)
        q = ctx.queue
        state = self.state
        if compiled is not None or vertex_program is not None:
            state = abi.GraphicsPipelineState.from_buffer_copy(bytes(self.state))
        if vertex_program is not None:
            state.vertex_program = abi.VS_COMPILED
            q.cmd_bind_vertex_program(vertex_program)
            if push_constants:
                q.cmd_push_constants(push_constants)
        if compiled is not None:
            state.fragment_program = abi.FS_COMPILED
            q.cmd_bind_fragment_program(compiled[self.state.fragment_program])
            q.cmd_push_constants(np.array([np.float32(1.0) / np.float32(self.width), np.float32(1.0) / np.float32(self.height)],
                                          dtype=np.float32).tobytes())
        q.set_draw_statistics(sbuf.address)
        ibuf = ctx.buffer(self.indices.nbytes, self.indices) if self.indices is not None else None
        if render_pass is not None:  # (after the uploads: any other command on the queue performs the pending clear)
            q.cmd_begin_render_pass(fb, render_pass, self.clear_color, self.clear_depth if self.use_depth else None)
        if self.indices is not None:
            q.cmd_draw_indexed(state, fb, self.viewport, self.scissor, vi, ibuf.address, self.indices.nbytes,
                               self.index_type, self.count, self.instances, self.first, self.vertex_offset, 0)
        else:
            q.cmd_draw(state, fb, self.viewport, self.scissor, vi, self.count, self.instances, self.first, 0)
        if render_pass is not None:
            q.cmd_end_render_pass()
        q.set_draw_statistics(0)
        if compiled is not None:
            q.cmd_bind_fragment_program(None)
        if vertex_program is not None:
            q.cmd_bind_vertex_program(None)
        color, depth = cbuf.download(), dbuf.download()
        stats = abi.DrawStatistics.from_buffer_copy(sbuf.download().tobytes())
        ctx.release()
        return color, depth, stats
