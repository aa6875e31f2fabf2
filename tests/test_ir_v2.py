"""CPU: version 2 of the IR (kazan_b200/csrc/ir.h) — the instructions a shader with a body needs, which the reference does not
have (instructions_impl.rs:395-433): reader / printer agreement between the C++ mirror and the Python oracle, the version
gate, PTX acceptance by ptxas, the fragment-program module, and the SIMT interpreter against the version 1 interpreter and
against the C oracle's procedural shader."""
import ctypes as C
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from kazan_b200 import backend, host
from tests import ir_cases as ic
from tests import ir_v2_cases as v2

PTXAS = shutil.which("ptxas") or "/usr/local/cuda/bin/ptxas"
ALL_V2 = sorted(p.stem for p in v2.PROGRAMS.glob("*.kazan-ir"))


@pytest.mark.parametrize("name", ALL_V2)
def test_reader_and_printer_agree_on_version_2_programs(name):
    from oracle import ir_oracle as iro
    text = v2.program_text(name)
    m = backend.IrModule.parse(text, 2)
    assert m.to_text() == text
    assert iro.module_to_text(iro.parse_module(text, 2)) == text
    m.destroy()


@pytest.mark.parametrize("name", ALL_V2)
def test_version_1_readers_reject_version_2_programs(name):
    """the gate: without the flag both readers are the reference's grammar, where these instructions do not exist"""
    from oracle import ir_oracle as iro
    text = v2.program_text(name)
    with pytest.raises(host.KazanError) as e:
        backend.IrModule.parse(text)
    assert "unknown" in str(e.value)
    with pytest.raises(iro.IrError):
        iro.parse_module(text)


def test_every_version_2_instruction_is_exercised_by_a_program():
    from oracle import ir_oracle as iro
    used = set()
    for name in ALL_V2:
        text = v2.program_text(name)
        used |= {op for op in iro.V2_OPS if f" {op} [" in text or f"\n{op} [" in text.replace("            ", "\n")}
    assert used == set(iro.V2_OPS), sorted(set(iro.V2_OPS) - used)


@pytest.mark.parametrize("name", [n for n in ALL_V2 if not n.startswith("fs_")])
def test_version_2_kernels_assemble_for_sm_100a(name, tmp_path):
    if not os.path.exists(PTXAS):
        pytest.skip("ptxas not available")
    ptx = backend.IrModule.parse(v2.program_text(name), 2).to_ptx()
    f = tmp_path / f"{name}.ptx"
    f.write_text(ptx)
    r = subprocess.run([PTXAS, "-arch=sm_100a", "-o", "/dev/null", str(f)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if name == "subgroup_ops":
        for mnemonic in ("vote.sync.ballot.b32", "vote.sync.all.pred", "vote.sync.any.pred", "shfl.sync.idx.b32", "shfl.sync.bfly.b32",
                         "redux.sync.add.s32"):
            assert mnemonic in ptx, mnemonic
    if name == "saxpy_buffers":
        assert "ld.param.u32" in ptx and "kzn_push_constants" in ptx and "kzn_descriptor_table" in ptx
        assert ".local" not in ptx, "built-in variables that are only loaded from live in registers"


def test_fragment_module_assembles_with_the_shader_at_every_call_site(tmp_path):
    """kzn_fragment_program_create's module: the raster kernels' PTX with the compiled shader in place of every call"""
    if not os.path.exists(PTXAS):
        pytest.skip("ptxas not available")
    m = backend.IrModule.parse(v2.program_text("fs_orbit"), 2)
    ptx = backend.FragmentProgram.module_ptx(m)
    assert "kzn_fs_main" not in ptx and ptx.count("// fragment program, call site") > 10
    for k in ("kzn_fs_raster_direct_0", "kzn_fs_raster_direct_1", "kzn_fs_raster_direct_2", "kzn_fs_raster_0", "kzn_fs_raster_1", "kzn_fs_raster_2"):
        assert f".visible .entry {k}(" in ptx
    f = tmp_path / "fs.ptx"
    f.write_text(ptx)
    cubin = tmp_path / "fs.cubin"
    r = subprocess.run([PTXAS, "-arch=sm_100a", "-o", str(cubin), str(f)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "kzn_fs_raster_direct_0", str(cubin)], capture_output=True, text=True).stdout
    assert sass.count("FFMA") > 1000, "the orbit is inlined and unrolled into the fill kernel"
    assert "kzn_fs_main" not in sass  # (the CALL.REL.NOINC left in the kernel are ptxas' own division slow paths)


def test_fragment_shaders_that_do_not_fit_the_stage_are_rejected():
    text = v2.program_text("int_ops")  # a compute shader: 15 outputs
    with pytest.raises(host.KazanError) as e:
        backend.FragmentProgram.call_site_ptx(backend.IrModule.parse(text, 2))
    assert e.value.code == -8  # VK_ERROR_FEATURE_NOT_PRESENT


@pytest.mark.parametrize("name", ["load_i32", "sum_loop", "branchy", "widths", "fibonacci", "bool_float"])
def test_simt_interpreter_runs_version_1_programs_like_the_scalar_interpreter(name):
    from oracle import ir_oracle as iro
    from oracle import ir_simt as sim
    dt_in, dt_out, _, _ = ic.CASES[name]
    inputs = ic.make_inputs(name, 70, seed=5)
    want = ic.interpret(name, inputs)
    m = iro.parse_module(ic.program_text(name))
    out = bytearray(len(inputs) * dt_out.itemsize)
    mem = sim.Memory({0x1000: bytearray(inputs.tobytes()), 0x100000: out})
    sim.run_dispatch(m, len(inputs), mem, 0x1000, 0x100000)
    got = np.frombuffer(bytes(out), dtype=dt_out)
    for f in dt_out.names:
        if f != "pad":
            assert np.array_equal(got[f], want[f]), f


def test_orbit_shader_in_ir_is_the_c_oracles_shader_bit_for_bit(oracle):
    """tests/ir_programs_v2/fs_orbit.kazan-ir against kzo_fs_procedural_orbit (oracle/raster.c) on a 64 x 36 frame"""
    from oracle import ir_oracle as iro
    from oracle import ir_simt as sim
    W, H = 64, 36
    m = iro.parse_module(v2.program_text("fs_orbit"), 2)
    pc = np.array([1.0 / W, 1.0 / H], dtype=np.float32).tobytes()
    coords = [(x + 0.5, y + 0.5, 0.0, 1.0) for y in range(0, H, 5) for x in range(0, W, 3)]
    got = sim.run_fragments(m, coords, [(0, 0, 0, 0)] * len(coords), pc)
    oracle.kzo_fs_procedural_orbit.argtypes = [C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.POINTER(C.c_float)]
    for (x, y, _, _), g in zip(coords, got):
        rgba = (C.c_float * 4)()
        oracle.kzo_fs_procedural_orbit(int(x), int(y), C.c_float(W), C.c_float(H), rgba)
        assert [struct.unpack("<I", struct.pack("<f", v))[0] for v in rgba] == g, (x, y)


def test_interpreter_corner_cases():
    """the implementation-defined corners ir.h fixes: over-wide shifts, saturating conversions, min / max of NaN and zeros"""
    from oracle import ir_simt as sim
    f = sim.Simt.float_op
    nan, one = 0x7FC00000, 0x3F800000
    assert f("f_min", "f32", [nan, one]) == one and f("f_max", "f32", [one, nan]) == one
    assert f("f_min", "f32", [0x00000000, 0x80000000]) == 0x80000000 and f("f_max", "f32", [0x80000000, 0x00000000]) == 0
    assert f("fma", "f32", [0x3F800001, 0x3F800001, 0xBF800002]) == 0x28800000  # (1+e)^2 - (1+2e) = e^2: only a fused multiply-add sees it
    got = v2.interpret("int_ops", v2.make_inputs("int_ops", 64, seed=3))
    inp = v2.make_inputs("int_ops", 64, seed=3)
    big = inp["sh"] >= 32
    assert np.all(got["r3"][big] == 0), "shift_left_logical by the width or more is 0"
    fo = v2.interpret("float_ops", v2.make_inputs("float_ops", 64, seed=3))
    fi = v2.make_inputs("float_ops", 64, seed=3)
    t7 = fi["a"] * np.float32(1000.0)
    assert np.all(fo["r7"][t7 < 0] == 0) and np.all(fo["r13"][t7 > 255] == 255) and np.all(fo["r14"][t7 > 127] == 127)


def _simple_test_words():
    """the reference's second translator test shader: test_data/simple_test.spv == spirv-parser/test_inputs/test.spv
    (a vertex shader: gl_Position = vec4(in, 1), float -> u64 -> u32 -> i32 conversions into a Function variable, an OpSwitch
    on the 64-bit value whose default stores 0 to gl_Position.x)"""
    import base64
    import json
    from pathlib import Path
    g = json.loads((Path(__file__).parent / "golden" / "spirv_goldens.json").read_text())
    case = next(c for c in g["cases"] if c["file"].endswith("test.spv"))
    return np.frombuffer(base64.b64decode(case["spv_base64"]), dtype="<u4").tolist()


def test_simple_test_spv_translates_to_version_2_ir():
    """the shader the reference's translator cannot translate (src/lib.rs:260-287 is #[ignore], its golden
    "not yet entered"): OpLoad / OpStore / OpAccessChain / OpCompositeExtract / OpCompositeConstruct / OpConvertFToU / OpUConvert
    / OpBitcast / OpSelectionMerge + OpSwitch.  Version 1 of the translator still refuses it, like the reference."""
    from oracle import ir_oracle as iro
    words = _simple_test_words()
    with pytest.raises(host.KazanError) as e:
        backend.translate_spirv_to_ir(words, "main", backend.EXECUTION_MODEL_VERTEX)
    assert e.value.code == -8
    module, _ = backend.translate_spirv_to_ir(words, "main", backend.EXECUTION_MODEL_VERTEX, 2)
    text = module.to_text()
    for needle in ("kind: vertex_position", "convert_f_to_u [id_30] -> [id_32: i64]", "u_convert [id_32] -> [id_33: i32]", "bitcast [id_33] -> [id_34: i32]",
                   "block id_38 -> [] {", "0x8i64 -> break id_36[];", "store [id_10, c3: f32 0x0], align: 0x4 -> [];"):
        assert needle in text, needle
    assert backend.IrModule.parse(text, 2).to_text() == text and iro.module_to_text(iro.parse_module(text, 2)) == text
    if os.path.exists(PTXAS):
        import tempfile
        with tempfile.TemporaryDirectory() as d:
            (open(f"{d}/k.ptx", "w")).write(module.to_ptx())
            r = subprocess.run([PTXAS, "-arch=sm_100a", "-o", "/dev/null", f"{d}/k.ptx"], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr


def test_translated_simple_test_runs_in_the_interpreter():
    """what the shader means, per vertex: position = (x, y, z, 1), and x = 0 unless trunc(z) is 1, 2 or 8"""
    from oracle import ir_oracle as iro
    from oracle import ir_simt as sim
    module, _ = backend.translate_spirv_to_ir(_simple_test_words(), "main", backend.EXECUTION_MODEL_VERTEX, 2)
    m = iro.parse_module(module.to_text(), 2)
    zs = [0.5, 1.0, 1.9, 2.2, 3.0, 8.5, 9.0, -4.0, 1e30, float("nan")]
    inp = np.zeros((len(zs), 4), dtype=np.float32)
    inp[:, 0], inp[:, 1], inp[:, 2] = 7.0, -3.0, zs
    out = bytearray(16 * len(zs))
    assert sim.output_slot_size(m) == 16
    mem = sim.Memory({0x1000: bytearray(inp.tobytes()), 0x9000: out})
    sim.run_dispatch(m, len(zs), mem, 0x1000, 0x9000)
    pos = np.frombuffer(bytes(out), dtype=np.float32).reshape(-1, 4)
    for z, p in zip(zs, pos):
        keeps_x = z == z and int(min(max(z, 0.0), 2.0**64)) in (1, 2, 8)
        assert p[0] == (7.0 if keeps_x else 0.0) and p[1] == -3.0 and p[3] == 1.0, (z, p)
        assert (p[2] == np.float32(z)) or (z != z and p[2] != p[2])
