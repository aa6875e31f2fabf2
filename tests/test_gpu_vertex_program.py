"""GPU: a compiled vertex stage on the draw path (include/kazan_b200_backend.h "vertex programs").

The vertex shader is a version-2 shader-compiler-ir module — hand-written (tests/ir_programs_v2/vs_transform.kazan-ir: a 2x2
matrix, an offset, a perspective term and a tint from the push constants) or the translation of the reference's
simple_test.spv — run by the PTX back end, one invocation per vertex, before setup.  The checker runs the same module in the
SIMT interpreter (oracle/ir_simt.py) over the same vertices and hands its output to the raster oracle: the frames must agree
bit for bit, on the direct path, the binned path, with indices, with firstVertex, and together with a compiled fragment
stage."""
import numpy as np
import pytest

from kazan_b200 import abi, backend, host
from tests import ir_v2_cases as v2
from tests import util
from tests.util import Scene, pipeline_state

pytestmark = pytest.mark.gpu

PC = np.array([0.9, -0.15, 0.1, 0.95, 0.03, -0.02, 0.45, 0.75], dtype=np.float32).tobytes()


def shade_in_interpreter(text_or_module, vertices: np.ndarray, push_constants: bytes) -> np.ndarray:
    """what the vertex stage must write: one output slot per vertex, as float32 rows"""
    from oracle import ir_oracle as iro
    from oracle import ir_simt as sim
    m = iro.parse_module(text_or_module, 2)
    n = len(vertices)
    slot = sim.output_slot_size(m)
    out = bytearray(n * slot)
    sim.run_dispatch(m, n, sim.Memory({0x10000000: bytearray(vertices.tobytes()), 0x20000000: out}), 0x10000000, 0x20000000,
                     push_constants=push_constants)
    return np.frombuffer(bytes(out), dtype=np.float32).reshape(n, slot // 4).copy()


def reference_scene(scene: Scene, shaded: np.ndarray, position_offset: int, color_offset) -> Scene:
    """the same draw over the interpreter's vertices, for the raster oracle"""
    ref = Scene(scene.name + "_ref", scene.width, scene.height, shaded, scene.indices, state=scene.state, viewport=scene.viewport,
                scissor=scene.scissor, use_depth=scene.use_depth, clear_color=scene.clear_color, clear_depth=scene.clear_depth,
                color_tiling=scene.color_tiling, index_type=scene.index_type, pos_comps=4, color_offset=color_offset,
                instances=scene.instances, band=scene.band, first=scene.first, vertex_offset=scene.vertex_offset)
    ref.position_offset = position_offset
    return ref


def mesh_vertices(quads_x, quads_y, seed):
    v, i = util.grid_mesh(quads_x, quads_y, seed)  # x y z r g b a
    out = np.zeros((len(v), 8), dtype=np.float32)
    out[:, :3], out[:, 4:] = v[:, :3], v[:, 3:]
    return out, i


@pytest.fixture(scope="module")
def transform(gpu):
    module = backend.IrModule.parse(v2.program_text("vs_transform"), 2)
    vp = backend.VertexProgram(gpu.device, module)
    assert (vp.inputs_size, vp.outputs_size) == (32, 32)
    yield vp
    vp.destroy()
    module.destroy()


def check(gpu, scene, program, text, pc, position_offset, color_offset, compiled=None):
    shaded = shade_in_interpreter(text, scene.vertices, pc)
    oc, od, os_ = reference_scene(scene, shaded, position_offset, color_offset).run_oracle()
    gc, gd, gs = scene.run_gpu(gpu, vertex_program=program, push_constants=pc, compiled=compiled)
    assert (gs.triangles_in, gs.triangles_culled, gs.fragments_covered, gs.fragments_depth_passed, gs.triangles_clipped) == \
        (os_.triangles_in, os_.triangles_culled, os_.fragments_covered, os_.fragments_depth_passed, os_.triangles_clipped)
    assert gd.tobytes() == od.tobytes() and gc.tobytes() == oc.tobytes()
    assert os_.fragments_covered > 1000
    return os_


def test_binned_indexed_mesh_through_the_compiled_vertex_stage(gpu, transform):
    v, i = mesh_vertices(40, 24, 12)  # 1920 triangles: binned path
    sc = Scene("vs_mesh", 320, 200, v, i, state=pipeline_state(depth=(abi.COMPARE_LESS, 1), blend=util.ALPHA_BLEND), use_depth=True)
    check(gpu, sc, transform, v2.program_text("vs_transform"), PC, 16, 0)


def test_direct_draw_with_first_vertex_and_tiled_target(gpu, transform):
    v, i = mesh_vertices(6, 5, 13)
    flat = v[i.astype(np.int64)]  # 180 vertices, non-indexed: direct path behind setup_kernel
    flat = np.concatenate([np.full((7, 8), 9.0e9, dtype=np.float32), flat])  # firstVertex = 7: the vertices before are never shaded
    sc = Scene("vs_direct", 200, 120, flat, first=7, state=pipeline_state(blend=util.ADDITIVE_BLEND), color_tiling=abi.TILING_TILED)
    sc.vertices[:, 4:] *= 0.25
    check(gpu, sc, transform, v2.program_text("vs_transform"), PC, 16, 0)


def test_vertices_the_shader_sends_behind_the_eye_are_clipped(gpu, transform):
    """a perspective term strong enough to push part of the mesh to w = 0.25 (the shader's f_max) and the guard band"""
    v, i = mesh_vertices(30, 20, 14)
    pc = np.array([40.0, 0.0, 0.0, 40.0, 0.0, 0.0, -3.0, 1.0], dtype=np.float32).tobytes()
    sc = Scene("vs_clip", 256, 160, v, i, state=pipeline_state(depth=(abi.COMPARE_LESS_OR_EQUAL, 1)), use_depth=True)
    os_ = check(gpu, sc, transform, v2.program_text("vs_transform"), pc, 16, 0)
    assert os_.triangles_in == 1200


def test_compiled_vertex_and_fragment_stages_together(gpu, transform):
    fmod = backend.IrModule.parse(v2.program_text("fs_vertex_color"), 2)
    fp = backend.FragmentProgram(gpu.device, fmod)
    v, i = mesh_vertices(24, 16, 15)
    sc = Scene("vs_fs", 240, 150, v, i, state=pipeline_state(blend=util.ALPHA_BLEND))
    shaded = shade_in_interpreter(v2.program_text("vs_transform"), sc.vertices, PC)
    oc, od, os_ = reference_scene(sc, shaded, 16, 0).run_oracle()
    # Scene.run_gpu pushes the compiled fragment program's constants after the vertex program's: bind by hand here
    state = abi.GraphicsPipelineState.from_buffer_copy(bytes(sc.state))
    state.vertex_program, state.fragment_program = abi.VS_COMPILED, abi.FS_COMPILED
    sc.state = state
    q = gpu.queue
    q.cmd_bind_vertex_program(transform)
    q.cmd_bind_fragment_program(fp)
    q.cmd_push_constants(PC)
    gc, gd, gs = sc.run_gpu(gpu)
    q.cmd_bind_vertex_program(None)
    q.cmd_bind_fragment_program(None)
    assert gc.tobytes() == oc.tobytes() and gs.fragments_covered == os_.fragments_covered
    fp.destroy()
    fmod.destroy()


def test_simple_test_spv_as_the_vertex_stage_of_a_draw(gpu):
    """SPIR-V -> IR -> PTX -> vertex stage -> raster: the reference's simple_test.spv (position = (x, y, z, 1), x forced to 0
    unless trunc(z) is 1, 2 or 8) drawing triangles at z = 1 (kept) and z = 0.5 (collapsed onto x = 0)"""
    from tests.test_ir_v2 import _simple_test_words
    module, _ = backend.translate_spirv_to_ir(_simple_test_words(), "main", backend.EXECUTION_MODEL_VERTEX, 2)
    vp = backend.VertexProgram(gpu.device, module)
    assert (vp.inputs_size, vp.outputs_size) == (16, 16)
    rng = np.random.default_rng(16)
    tris = []
    for k in range(40):
        c = rng.uniform(-0.8, 0.8, 2)
        z = 1.0 if k % 3 else 0.5
        for _ in range(3):
            o = rng.uniform(-0.25, 0.25, 2)
            tris.append([c[0] + o[0], c[1] + o[1], z, 0.0])
    sc = Scene("vs_spv", 160, 120, np.array(tris, dtype=np.float32), state=pipeline_state(blend=util.ADDITIVE_BLEND), color_offset=0xFFFFFFFF,
               clear_color=(0, 0, 0, 0))
    sc.state.src_color_blend_factor = abi.BLEND_ONE
    shaded = shade_in_interpreter(module.to_text(), sc.vertices, b"")
    oc, od, os_ = reference_scene(sc, shaded, 0, 0xFFFFFFFF).run_oracle()
    gc, gd, gs = sc.run_gpu(gpu, vertex_program=vp)
    assert gc.tobytes() == oc.tobytes() and gs.fragments_covered == os_.fragments_covered > 500
    assert gs.triangles_culled == os_.triangles_culled >= 13  # the z = 0.5 triangles have zero area
    vp.destroy()
    module.destroy()


def test_vertex_program_validation(gpu, transform):
    v, i = mesh_vertices(4, 4, 17)
    sc = Scene("bad", 64, 64, v, i)
    state = abi.GraphicsPipelineState.from_buffer_copy(bytes(sc.state))
    state.vertex_program = abi.VS_COMPILED
    sc.state = state
    with pytest.raises(host.KazanError) as e:  # nothing bound
        sc.run_gpu(gpu)
    assert e.value.code == abi.ERROR_VALIDATION_FAILED
    gpu.queue.cmd_bind_vertex_program(transform)
    bad_stride = Scene("bad", 64, 64, v[:, :7].copy(), i)  # 28-byte vertices against a 32-byte input block
    bad_stride.state = state
    with pytest.raises(host.KazanError) as e:
        bad_stride.run_gpu(gpu)
    assert "stride" in str(e.value)
    gpu.queue.cmd_bind_vertex_program(None)
    with pytest.raises(host.KazanError):  # a fragment shader is not a vertex program: no vertex_position
        backend.VertexProgram(gpu.device, backend.IrModule.parse(v2.program_text("fs_vertex_color"), 2))
    gpu.release()


def test_recorded_draw_with_a_compiled_vertex_stage(gpu, transform):
    """the vertex stage's dispatch and the draw behind it replay from a command buffer (the vertex scratch has to be reserved
    before recording, like the raster scratch); the push constants are those of record time"""
    v, i = mesh_vertices(40, 24, 18)
    sc = Scene("vs_recorded", 320, 200, v, i, state=pipeline_state(depth=(abi.COMPARE_LESS, 1), blend=util.ALPHA_BLEND), use_depth=True)
    shaded = shade_in_interpreter(v2.program_text("vs_transform"), sc.vertices, PC)
    want_c, want_d, _ = reference_scene(sc, shaded, 16, 0).run_oracle()
    color0, depth0 = sc._initial()
    dev = host.Device(0)
    q = dev.create_queue()  # a fresh queue: no scratch yet
    arrays = (color0, depth0, sc.vertices, sc.indices)
    mem = [dev.allocate(a.nbytes) for a in arrays]

    def reset():
        for m, a in zip(mem, arrays):
            a = np.ascontiguousarray(a)
            q.upload(m.address, a.ctypes.data, a.nbytes)
        q.wait_idle()

    reset()
    fb = sc._framebuffer(mem[0].address, mem[1].address)
    vi = abi.VertexInput(mem[2].address, sc.vertices.nbytes, sc.stride, 0, sc.pos_comps, sc.color_offset)
    state = abi.GraphicsPipelineState.from_buffer_copy(bytes(sc.state))
    state.vertex_program = abi.VS_COMPILED
    vp = backend.VertexProgram(dev, backend.IrModule.parse(v2.program_text("vs_transform"), 2))
    q.cmd_bind_vertex_program(vp)
    q.cmd_push_constants(PC)

    def draw():
        q.cmd_draw_indexed(state, fb, sc.viewport, sc.scissor, vi, mem[3].address, sc.indices.nbytes, sc.index_type, sc.count, 1, 0, 0, 0)

    q.reserve_draw(sc.count // 3, sc.width, sc.height)
    with pytest.raises(host.KazanError) as e:
        with q.record():
            draw()
    assert e.value.code == abi.ERROR_VALIDATION_FAILED
    q.reserve_vertex_outputs(len(v), vp)
    with q.record() as cb:
        draw()
    q.cmd_push_constants(np.zeros(8, dtype=np.float32).tobytes())  # later state does not reach the recorded commands
    for _ in range(2):
        reset()
        q.submit([cb])
        q.wait_idle()
        got_c, got_d = np.empty_like(color0), np.empty_like(depth0)
        q.download(got_c.ctypes.data, mem[0].address, got_c.nbytes)
        q.download(got_d.ctypes.data, mem[1].address, got_d.nbytes)
        q.wait_idle()
        assert got_d.tobytes() == want_d.tobytes() and got_c.tobytes() == want_c.tobytes()
    cb.destroy()
    vp.destroy()
    for m in mem:
        m.free()
    q.destroy()
    dev.destroy()
