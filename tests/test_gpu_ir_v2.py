"""GPU: version 2 of the IR through the PTX back end (north_star (1) and (2)).

Compute: programs using float arithmetic, compares, select, conversions, store, ptr_add, the built-in inputs, a bound
descriptor set and push constants, and the subgroup operations (warp votes / shuffles / redux) are dispatched through the C
ABI and compared bit for bit with the SIMT interpreter (oracle/ir_simt.py); the buffer-to-buffer shader also at 2^24
invocations against its closed form.
Draw: every parity scene of tests/cases.py is drawn again with its fragment program COMPILED from shader-compiler-ir
(tests/ir_programs_v2/fs_*.kazan-ir) into the raster kernels; the frames must be the oracle's, bit for bit, like the frames
of the built-in programs."""
import numpy as np
import pytest

from kazan_b200 import abi, backend
from tests import cases
from tests import ir_v2_cases as v2
from tests.test_gpu_raster import GOLDEN, SCENES, sha

pytestmark = pytest.mark.gpu
# the loop shaders that have run on a B200 (tests/test_spirv_loops.py runs all five through the interpreter and ptxas on the CPU)
LOOP_KINDS_ON_GPU = ["variables", "phis", "do_while", "nested", "header_values_after_loop", "calls"]


def _dispatch(gpu, name, inputs, group_count, local_size=(1, 1, 1)):
    dt_in, dt_out = v2.CASES[name]
    module = backend.IrModule.parse(v2.program_text(name), 2)
    pipe = backend.ComputePipeline(gpu.device, module, local_size)
    assert pipe.inputs_size == dt_in.itemsize and pipe.outputs_size == dt_out.itemsize
    n = len(inputs)
    ibuf = gpu.buffer(inputs.nbytes, inputs)
    obuf = gpu.buffer(n * dt_out.itemsize, np.full(n * dt_out.itemsize, 0xEE, dtype=np.uint8))
    gpu.queue.cmd_dispatch(pipe, group_count, ibuf.address, obuf.address)
    gpu.queue.wait_idle()
    raw = obuf.download()[:n * dt_out.itemsize]
    pipe.destroy()
    module.destroy()
    return raw.view(dt_out)


def test_integer_program_matches_interpreter(gpu):
    inputs = v2.make_inputs("int_ops", 333, seed=21)
    got, want = _dispatch(gpu, "int_ops", inputs, (333, 1, 1)), v2.interpret("int_ops", inputs)
    for f in want.dtype.names:
        if f != "r14":  # r14 is a pointer difference: layout, checked below
            assert np.array_equal(got[f], want[f]), f
    assert np.array_equal(got["r14"], want["r14"]), "select on pointers / pointer difference"
    gpu.release()


def test_float_program_matches_interpreter(gpu):
    inputs = v2.make_inputs("float_ops", 333, seed=22)
    got, want = _dispatch(gpu, "float_ops", inputs, (333, 1, 1)), v2.interpret("float_ops", inputs)
    for f in want.dtype.names:
        if f in ("r0", "r1", "r2", "r3", "r4", "r5", "r9", "r10"):
            assert v2.same_floats(got[f], want[f]), f  # IEEE results: bit patterns, NaN == NaN
        else:
            assert np.array_equal(got[f], want[f]), f
    gpu.release()


@pytest.mark.parametrize("n", [96, 256 + 32 + 7])
def test_subgroup_operations_match_interpreter(gpu, n):
    """warp votes, shuffles and redux: full subgroups, an all-odd and a none-odd one, and (second size) a last subgroup of 7
    invocations whose member mask is cut from the launch size"""
    inputs = v2.make_inputs("subgroup_ops", n, seed=23)
    got, want = _dispatch(gpu, "subgroup_ops", inputs, (n, 1, 1)), v2.interpret("subgroup_ops", inputs)
    for f in want.dtype.names:
        assert np.array_equal(got[f], want[f]), f
    v = inputs["v"]
    assert got["r0"][0] == sum(int(v[i] & 1) << i for i in range(32))  # ballot, closed form
    assert got["r1"][32] == 3 and got["r1"][64] == 0  # all | any
    assert np.array_equal(got["r6"][:32], np.full(32, np.uint32(v[:32].astype(np.uint64).sum() & 0xFFFFFFFF)))
    gpu.release()


def test_translated_simple_test_spv_runs_and_matches_interpreter(gpu):
    """SPIR-V -> IR (version 2) -> PTX -> sm_100a for the reference's simple_test.spv (the shader its own translator cannot
    translate): one vertex per invocation, gl_Position read back from the built-in part of the output slot"""
    from oracle import ir_oracle as iro
    from oracle import ir_simt as sim
    from tests.test_ir_v2 import _simple_test_words
    module, _ = backend.translate_spirv_to_ir(_simple_test_words(), "main", backend.EXECUTION_MODEL_VERTEX, 2)
    rng = np.random.default_rng(31)
    n = 200
    inp = np.zeros((n, 4), dtype=np.float32)
    inp[:, :2] = rng.uniform(-5, 5, (n, 2))
    inp[:, 2] = rng.choice(np.array([0.5, 1.0, 1.5, 2.0, 2.99, 3.0, 8.0, 8.7, 9.0, -1.0, 1e20, np.inf, np.nan], dtype=np.float32), n)
    pipe = backend.ComputePipeline(gpu.device, module, (1, 1, 1))
    assert pipe.inputs_size == 16 and pipe.outputs_size == 16
    ibuf = gpu.buffer(inp.nbytes, inp)
    obuf = gpu.buffer(16 * n, np.full(16 * n, 0xEE, dtype=np.uint8))
    gpu.queue.cmd_dispatch(pipe, (n, 1, 1), ibuf.address, obuf.address)
    gpu.queue.wait_idle()
    got = obuf.download()[:16 * n].view(np.uint32).reshape(n, 4)
    m = iro.parse_module(module.to_text(), 2)
    out = bytearray(16 * n)
    sim.run_dispatch(m, n, sim.Memory({0x1000: bytearray(inp.tobytes()), 0x9000: out}), 0x1000, 0x9000)
    want = np.frombuffer(bytes(out), dtype=np.uint32).reshape(n, 4)
    assert np.array_equal(got, want)
    assert (got[:, 0] == 0).any() and (got[:, 0] != 0).any()  # both arms of the switch were taken
    pipe.destroy()
    module.destroy()
    gpu.release()


def _saxpy(gpu, W, H, local, check_interpreter):
    rng = np.random.default_rng(24)
    x = rng.integers(-1000, 1000, W * H).astype(np.float32)
    y = rng.integers(-1000, 1000, W * H).astype(np.float32)
    alpha = np.float32(1.5)
    xb, yb = gpu.buffer(x.nbytes, x), gpu.buffer(y.nbytes, y)
    ds = backend.DescriptorSet(gpu.device, [(0, abi.DESCRIPTOR_TYPE_STORAGE_BUFFER, 1), (3, abi.DESCRIPTOR_TYPE_STORAGE_BUFFER, 1)])
    ds.write_buffers(0, [(xb.address, 0, x.nbytes)])
    ds.write_buffers(3, [(yb.address, 0, y.nbytes)])
    assert (ds.slot(0), ds.slot(3), ds.slot(1)) == (0, 1, -1)
    module = backend.IrModule.parse(v2.program_text("saxpy_buffers"), 2)
    pipe = backend.ComputePipeline(gpu.device, module, local)
    q = gpu.queue
    q.cmd_bind_descriptor_set(abi.PIPELINE_BIND_POINT_COMPUTE, ds)
    q.cmd_push_constants(alpha.tobytes() + np.uint32(W).tobytes())
    q.cmd_dispatch(pipe, (W // local[0], H // local[1], 1))
    q.wait_idle()
    got = yb.download().view(np.float32)[:W * H]
    assert np.array_equal(got, alpha * x + y), "closed form (every value is exactly representable)"
    if check_interpreter:
        from oracle import ir_oracle as iro
        from oracle import ir_simt as sim
        table = np.array([0x3000_0000, x.nbytes, 0x4000_0000, y.nbytes], dtype=np.uint64)
        ybuf = bytearray(y.tobytes())
        mem = sim.Memory({0x3000_0000: bytearray(x.tobytes()), 0x4000_0000: ybuf, 0x5000_0000: bytearray(table.tobytes())})
        sim.run_dispatch(iro.parse_module(v2.program_text("saxpy_buffers"), 2), W * H, mem, global_size=(W, H, 1),
                         push_constants=alpha.tobytes() + np.uint32(W).tobytes(), descriptor_table=0x5000_0000)
        assert np.array_equal(got, np.frombuffer(bytes(ybuf), dtype=np.float32))
    q.cmd_bind_descriptor_set(abi.PIPELINE_BIND_POINT_COMPUTE, None)
    pipe.destroy()
    module.destroy()
    ds.destroy()
    gpu.release()


@pytest.mark.parametrize("kind", LOOP_KINDS_ON_GPU)
def test_translated_spirv_loops_run_and_match_the_source_loops(gpu, kind):
    """SPIR-V loop -> the IR's `loop` -> PTX -> sm_100a (tests/spirv_loop_shaders.py: loop state in Function variables with a
    `continue` and a `break` inside selections; loop state in OpPhis): 4096 invocations whose trip counts differ inside every
    warp (divergent loops), against what the source loops compute in fp32"""
    from tests import spirv_loop_shaders as shaders
    module, _ = backend.translate_spirv_to_ir(getattr(shaders, kind)(), "main", backend.EXECUTION_MODEL_VERTEX, 2)
    a = shaders.inputs(4096, seed=9)
    pipe = backend.ComputePipeline(gpu.device, module, (64, 1, 1))
    assert pipe.inputs_size == shaders.SLOT.itemsize and pipe.outputs_size == shaders.OUT.itemsize
    ibuf = gpu.buffer(a.nbytes, a)
    obuf = gpu.buffer(shaders.OUT.itemsize * len(a), np.full(shaders.OUT.itemsize * len(a), 0xEE, dtype=np.uint8))
    gpu.queue.cmd_dispatch(pipe, (len(a) // 64, 1, 1), ibuf.address, obuf.address)
    gpu.queue.wait_idle()
    got = obuf.download()[:shaders.OUT.itemsize * len(a)].view(shaders.OUT)
    want = shaders.expected(kind, a)
    assert np.array_equal(got["i"], want["i"])
    assert np.array_equal(got["acc"].view(np.uint32), want["acc"].view(np.uint32))
    pipe.destroy()
    module.destroy()
    gpu.release()


def test_buffer_to_buffer_shader_through_a_descriptor_set_matches_interpreter(gpu):
    """f1: y = alpha * x + y over two STORAGE_BUFFER descriptors, alpha and the row length in push constants, the element
    index from global_invocation_id (16 x 12 invocations, LocalSize 8 x 4)"""
    _saxpy(gpu, 16, 12, (8, 4, 1), True)


def test_buffer_to_buffer_shader_over_2_pow_24_invocations(gpu):
    """the configs[1] size: 4096 x 4096 invocations, closed form"""
    _saxpy(gpu, 4096, 4096, (16, 16, 1), False)


def test_descriptor_set_validation(gpu):
    from kazan_b200.host import KazanError
    with pytest.raises(KazanError) as e:
        backend.DescriptorSet(gpu.device, [(0, 1, 1)])  # VK_DESCRIPTOR_TYPE_COMBINED_IMAGE_SAMPLER
    assert e.value.code == -8
    ds = backend.DescriptorSet(gpu.device, [(0, abi.DESCRIPTOR_TYPE_UNIFORM_BUFFER, 2)])
    with pytest.raises(KazanError):
        ds.write_buffers(0, [(0x1000, 0, 16)] * 3)
    with pytest.raises(KazanError):
        ds.write_buffers(7, [(0x1000, 0, 16)])
    with pytest.raises(KazanError):
        gpu.queue.cmd_push_constants(b"\0" * 132)
    ds.destroy()


@pytest.fixture(scope="module")
def compiled(gpu):
    """the two built-in fragment programs, compiled from their shader-compiler-ir form into the raster kernels"""
    progs = {}
    for enum, name in ((abi.FS_VERTEX_COLOR, "fs_vertex_color"), (abi.FS_PROCEDURAL_ORBIT, "fs_orbit")):
        m = backend.IrModule.parse(v2.program_text(name), 2)
        progs[enum] = backend.FragmentProgram(gpu.device, m)
        m.destroy()
    yield progs
    for p in progs.values():
        p.destroy()


@pytest.mark.parametrize("scene", SCENES, ids=lambda s: s.name)
def test_draw_with_compiled_fragment_program_matches_oracle(gpu, compiled, scene):
    oc, od, os_ = scene.run_oracle()
    gc, gd, gs = scene.run_gpu(gpu, compiled=compiled)
    assert (gs.triangles_in, gs.triangles_culled) == (os_.triangles_in, os_.triangles_culled)
    assert (gs.fragments_covered, gs.fragments_depth_passed) == (os_.fragments_covered, os_.fragments_depth_passed)
    assert gd.tobytes() == od.tobytes(), "depth buffer differs (must be bit exact)"
    if gc.tobytes() != oc.tobytes():
        diff = np.abs(gc.astype(np.int16) - oc.astype(np.int16))
        pytest.fail(f"colour differs: max {diff.max()} LSB on {(diff > 0).sum()} bytes")
    assert sha(gc.tobytes()) == GOLDEN[f"draw/{scene.name}/color"]


def test_draw_without_a_bound_program_is_refused(gpu):
    from kazan_b200.host import KazanError
    sc = next(s for s in SCENES if s.name == "fill_orbit_256x144")
    with pytest.raises(KazanError):
        sc.run_gpu(gpu, compiled={abi.FS_PROCEDURAL_ORBIT: None})
    gpu.release()


def test_full_size_fill_4k_with_the_compiled_orbit_shader(gpu, compiled):
    """BASELINE configs[2] at full size with the fragment shader coming from .kazan-ir: the frame of the built-in program,
    bit for bit"""
    from tests import util
    W, H = 3840, 2160
    sc = util.Scene("fill4k", W, H, util.FULLSCREEN_TRIANGLE, state=util.pipeline_state(fs=abi.FS_PROCEDURAL_ORBIT))
    a, _, st_a = sc.run_gpu(gpu)
    b, _, st_b = sc.run_gpu(gpu, compiled=compiled)
    assert st_b.fragments_covered == W * H == st_a.fragments_covered
    assert a.tobytes() == b.tobytes()


def test_compiled_shader_keeps_products_and_sums_apart(gpu):
    """No contraction: f_mul followed by f_add / f_sub are two roundings (ir.h), also when two fragments share one copy of the
    code with packed f32x2 arithmetic — ptxas contracts add.rn.f32x2 of a mul.rn.f32x2 result into FFMA2, which the emitter
    avoids.  fs_mul_add.kazan-ir turns the difference between (a * b + c) and fma(a, b, c) into the colour, so a contracted
    product would show as a frame of constant grey.  Checked against the SIMT interpreter on every pixel."""
    from oracle import ir_oracle as iro
    from oracle import ir_simt as sim
    from tests import util
    W, H = 64, 36
    text = v2.program_text("fs_mul_add")
    m = backend.IrModule.parse(text, 2)
    prog = backend.FragmentProgram(gpu.device, m)
    sc = util.Scene("mul_add", W, H, util.FULLSCREEN_TRIANGLE, state=util.pipeline_state(fs=abi.FS_PROCEDURAL_ORBIT))
    got, _, st = sc.run_gpu(gpu, compiled={abi.FS_PROCEDURAL_ORBIT: prog})
    assert st.fragments_covered == W * H
    coords = [(x + 0.5, y + 0.5, 0.0, 1.0) for y in range(H) for x in range(W)]
    out = np.array(sim.run_fragments(iro.parse_module(text, 2), coords, [(0, 0, 0, 0)] * len(coords)), dtype=np.uint32).view(np.float32)
    want = np.rint(np.clip(out, np.float32(0), np.float32(1)) * np.float32(255.0)).astype(np.uint8).reshape(H, W, 4)
    frame = got[:W * H * 4].reshape(H, W, 4)
    assert len(np.unique(want[..., :3])) > 2, "the shader must expose the two roundings"
    assert np.array_equal(frame, want)
    prog.destroy(); m.destroy()
