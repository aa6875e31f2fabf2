"""GPU: vkCmdDispatch through the PTX back end (BASELINE configs[1]).

Every IR program is compiled IR -> PTX -> cuModuleLoadDataEx, dispatched one invocation per thread through the C ABI,
and its user_outputs_block compared bit for bit with the oracle interpreter (oracle/ir_oracle.py) on the same seeded
inputs; at 2^24 invocations (the configs[1] size) against the programs' closed forms.
"""
import ctypes as C

import numpy as np
import pytest

from kazan_b200 import abi, backend, host
from tests import ir_cases as ic
from tests.test_ir import SPIRV, _build_test_module, _words

pytestmark = pytest.mark.gpu


def _dispatch(gpu, name, inputs, group_count, local_size=(1, 1, 1), first=0, table=None, slots=None):
    dt_in, dt_out, _, _ = ic.CASES[name]
    module = backend.IrModule.parse(ic.program_text(name))
    pipe = backend.ComputePipeline(gpu.device, module, local_size)
    assert pipe.inputs_size == dt_in.itemsize and pipe.outputs_size == dt_out.itemsize
    n = slots if slots is not None else len(inputs)
    ibuf = gpu.buffer(max(inputs.nbytes, 4), inputs if inputs.nbytes else None)
    obuf = gpu.buffer(max(n * dt_out.itemsize, 4), np.full(max(n * dt_out.itemsize, 4), 0xEE, dtype=np.uint8))
    gpu.queue.cmd_dispatch(pipe, group_count, ibuf.address, obuf.address, first_invocation=first)
    gpu.queue.wait_idle()
    raw = obuf.download()[:n * dt_out.itemsize]
    pipe.destroy()
    module.destroy()
    return raw.view(dt_out) if dt_out.itemsize else raw


@pytest.mark.parametrize("name", sorted(n for n in ic.CASES if n not in ("pointer_chase", "trivial_test")))
def test_dispatch_matches_interpreter(gpu, name):
    n = 777  # ragged: not a multiple of the 256-invocation CTA
    inputs = ic.make_inputs(name, n, seed=11)
    got = _dispatch(gpu, name, inputs, (n, 1, 1))
    want = ic.interpret(name, inputs)
    for field in want.dtype.names:
        if field != "pad":
            assert np.array_equal(got[field], want[field]), field
    gpu.release()


def test_pointer_chase_matches_interpreter(gpu):
    table = np.arange(1000, 1512, dtype=np.uint32)
    tbuf = gpu.buffer(table.nbytes, table)
    inputs = ic.make_inputs("pointer_chase", 300, seed=12, table_address=tbuf.address, table_len=len(table))
    got = _dispatch(gpu, "pointer_chase", inputs, (3, 100, 1))
    want = ic.interpret("pointer_chase", inputs, {tbuf.address: table.tobytes()})
    assert np.array_equal(got["w"], want["w"])
    gpu.release()


def test_trivial_shader_from_spirv_over_2_pow_24_invocations(gpu):
    """configs[1]: the reference's trivial_test.spv -> IR -> PTX, dispatch (65536, 256, 1) x LocalSize (1,1,1)."""
    g = SPIRV[0]
    module, local_size = backend.translate_spirv_to_ir(_words(g["spv_hex"]), "main", backend.EXECUTION_MODEL_GLCOMPUTE)
    pipe = backend.ComputePipeline(gpu.device, module, local_size)
    assert pipe.inputs_size == 0 and pipe.outputs_size == 0
    before = gpu.queue.launch_count
    gpu.queue.cmd_dispatch(pipe, (65536, 256, 1))
    gpu.queue.cmd_dispatch(pipe, (0, 5, 5))  # empty dispatch: valid, no launch
    gpu.queue.wait_idle()
    assert gpu.queue.launch_count == before + 1
    pipe.destroy()


def test_sum_loop_closed_form_at_2_pow_24(gpu):
    n = 1 << 24
    inputs = ic.make_inputs("sum_loop", n, seed=13)
    got = _dispatch(gpu, "sum_loop", inputs, (4096, 16, 1), local_size=(64, 2, 2))
    assert np.array_equal(got["y"], ic.CASES["sum_loop"][3](inputs))
    gpu.release()


def test_first_invocation_shards_a_dispatch(gpu):
    """§8e: a dispatch split over GPUs gives each rank a slice of the interface-block slots."""
    n = 1000
    inputs = ic.make_inputs("branchy", n, seed=14)
    dt_out = ic.CASES["branchy"][1]
    module = backend.IrModule.parse(ic.program_text("branchy"))
    pipe = backend.ComputePipeline(gpu.device, module)
    ibuf = gpu.buffer(inputs.nbytes, inputs)
    obuf = gpu.buffer(n * dt_out.itemsize, np.zeros(n * dt_out.itemsize, dtype=np.uint8))
    for first, count in ((0, 300), (300, 450), (750, 250)):
        gpu.queue.cmd_dispatch(pipe, (count, 1, 1), ibuf.address, obuf.address, base_group=(first, 0, 0), first_invocation=first)
    got = obuf.download().view(dt_out)
    t, r = ic.CASES["branchy"][3](inputs)
    assert np.array_equal(got["t"], t) and np.array_equal(got["r"], r)
    pipe.destroy()
    gpu.release()


def test_dispatch_validation(gpu):
    module = backend.IrModule.parse(ic.program_text("load_i32"))
    pipe = backend.ComputePipeline(gpu.device, module)
    with pytest.raises(host.KazanError) as e:
        gpu.queue.cmd_dispatch(pipe, (1, 1, 1), 0, 0)  # reads inputs but none bound
    assert e.value.code == abi.ERROR_VALIDATION_FAILED
    with pytest.raises(host.KazanError):
        backend.ComputePipeline(gpu.device, module, (0, 1, 1))
    pipe.destroy()


def test_backend_trait_compile_and_call(gpu):
    """shader-compiler-backend-llvm-7/src/tests.rs test_basic + test_names: build, compile, get by key, call with (0u32)."""
    ctx = backend.Context()
    module, functions = _build_test_module(ctx, ["main", "abc123-$._"])
    code = backend.CompiledCode(gpu.device, module, functions)
    before = gpu.queue.launch_count
    for key in functions:
        fn = code.get(key)
        assert fn is not None
        fn(gpu.queue, (1, 1, 1), (1, 1, 1), C.c_uint32(0))
    gpu.queue.wait_idle()
    assert gpu.queue.launch_count == before + 2
    assert code.get("missing") is None
    code.destroy()
    ctx.destroy()


def test_backend_trait_return_slot(gpu):
    """A non-void function returns through the trailing return-slot parameter."""
    ctx = backend.Context()
    u32 = ctx.build_u32()
    m = ctx.create_module("ret")
    f = m.add_function("echo", ctx.build_function([u32, u32], u32))
    b = ctx.create_builder().attach(f.append_new_basic_block("entry"))
    b.build_return(f.parameters()[1])
    code = backend.CompiledCode(gpu.device, m.verify(), {"echo": f})
    out = gpu.buffer(4, np.zeros(1, dtype=np.uint32))
    code.get("echo")(gpu.queue, (1, 1, 1), (1, 1, 1), C.c_uint32(5), C.c_uint32(0xCAFE), C.c_uint64(out.address))
    assert out.download().view(np.uint32)[0] == 0xCAFE
    code.destroy()
    ctx.destroy()
    gpu.release()


@pytest.mark.parametrize("seed_base", [0, 1000])
def test_random_programs_agree_with_the_interpreter(gpu, seed_base):
    """Differential test of the PTX back end: 40 random integer programs (tests/ir_random.py: nested blocks, multi-way
    branches that break outer blocks, bounded loops, all integer widths, under-aligned loads) x 61 invocations each, every
    output compared with the oracle interpreter."""
    from oracle import ir_oracle as iro
    from tests import ir_random
    rng = np.random.default_rng(seed_base + 99)
    n = 61
    for seed in range(seed_base, seed_base + 40):
        text, dt_in, dt_out = ir_random.make_program(seed)
        inputs = np.zeros(n, dtype=dt_in)
        inputs["x"] = rng.integers(0, 2**63, n, dtype=np.uint64) * 2 + rng.integers(0, 2, n, dtype=np.uint64)
        inputs["x"][:8] = [0, 1, 2, 3, 0xFF, 0xFFFF, 0xFFFFFFFF, 0xFFFFFFFFFFFFFFFF]  # small selectors hit the branch arms
        module = backend.IrModule.parse(text)
        pipe = backend.ComputePipeline(gpu.device, module, (1, 1, 1))
        assert pipe.inputs_size == 8 and pipe.outputs_size == dt_out.itemsize
        ibuf = gpu.buffer(inputs.nbytes, inputs)
        obuf = gpu.buffer(n * dt_out.itemsize, np.full(n * dt_out.itemsize, 0xEE, dtype=np.uint8))
        gpu.queue.cmd_dispatch(pipe, (n, 1, 1), ibuf.address, obuf.address)
        gpu.queue.wait_idle()
        got = obuf.download()[:n * dt_out.itemsize].view(dt_out)
        pipe.destroy()
        module.destroy()
        m = iro.parse_module(text)
        members = m.user_outputs.members
        for g in range(n):
            vals = iro.Interpreter(m, {0x10000000: inputs[g].tobytes()}).run_entry(0x10000000, 0x20000000)
            for i, ((off, ty), v) in enumerate(zip(members, vals)):
                assert int(got[f"o{i}"][g]) == int(v), (seed, g, i, ty, hex(int(inputs["x"][g])))
        gpu.release()
