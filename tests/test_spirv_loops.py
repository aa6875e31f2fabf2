"""CPU: SPIR-V loops -> the IR's `loop` (kazan_b200/csrc/spirv_to_ir_v2.cpp loop_construct).  The reference's translator has
no loop support (parse/unimplemented_instructions.rs: OpLoopMerge is a todo!()) and its fixtures no loop, so the shaders are
assembled here (tests/spirv_loop_shaders.py); the IR text round-trips through both readers, assembles for sm_100a, and the SIMT
interpreter computes what the source loops compute."""
import os
import subprocess

import numpy as np
import pytest

from kazan_b200 import backend
from tests import spirv_loop_shaders as shaders

PTXAS = "/usr/local/cuda/bin/ptxas"


def _translate(kind):
    words = getattr(shaders, kind)()
    module, _ = backend.translate_spirv_to_ir(words, "main", backend.EXECUTION_MODEL_VERTEX, 2)
    return module


@pytest.mark.parametrize("kind", ["variables", "phis", "do_while", "nested", "header_values_after_loop", "calls"])
def test_loop_shader_translates_round_trips_and_assembles(kind, tmp_path):
    from oracle import ir_oracle as iro
    module = _translate(kind)
    text = module.to_text()
    assert "loop id_" in text and "continue id_" in text
    if kind == "calls":  # both functions inlined at their call sites: blocks named after the callee, no second function in the module
        assert text.count("_call -> [") == 2 and text.count("fn id_") == 1
    if kind == "phis":  # header OpPhis are the loop's parameters, the merge block's OpPhis its results
        assert "[c0: 0x0i32, c2: f32 0x0] -> [" in text.replace("\n", " ") or "loop" in text
        header = next(line for line in text.splitlines() if line.strip().startswith("loop "))
        assert header.count(":") >= 2 and "-> [" in header, header
    elif kind == "do_while":  # the back edge under a condition: a block whose results `continue` passes on
        assert "_continue -> [" in text and text.count("continue id_") == 1
    elif kind == "nested":
        assert text.count("loop id_") == 2 and text.count("continue id_") == 2
    else:
        assert "_branch -> !" in text, "the test block's conditional branch has no selection merge: a block nothing follows"
    assert backend.IrModule.parse(text, 2).to_text() == text
    assert iro.module_to_text(iro.parse_module(text, 2)) == text
    if os.path.exists(PTXAS):
        (tmp_path / "k.ptx").write_text(module.to_ptx())
        r = subprocess.run([PTXAS, "-arch=sm_100a", "-o", "/dev/null", str(tmp_path / "k.ptx")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


@pytest.mark.parametrize("kind", ["variables", "phis", "do_while", "nested", "header_values_after_loop", "calls"])
def test_loop_shader_runs_in_the_interpreter(kind):
    from oracle import ir_oracle as iro
    from oracle import ir_simt as sim
    m = iro.parse_module(_translate(kind).to_text(), 2)
    a = shaders.inputs()
    out = bytearray(shaders.OUT.itemsize * len(a))
    assert sim.output_slot_size(m) == shaders.OUT.itemsize
    mem = sim.Memory({0x1000: bytearray(a.tobytes()), 0x9000: out})
    sim.run_dispatch(m, len(a), mem, 0x1000, 0x9000)
    got, want = np.frombuffer(bytes(out), dtype=shaders.OUT), shaders.expected(kind, a)
    assert np.array_equal(got["i"], want["i"])
    assert np.array_equal(got["acc"].view(np.uint32), want["acc"].view(np.uint32))


def test_version_1_still_refuses_loops():
    from kazan_b200 import host
    with pytest.raises(host.KazanError):
        backend.translate_spirv_to_ir(shaders.variables(), "main", backend.EXECUTION_MODEL_VERTEX)


def test_recursion_is_refused():
    from kazan_b200 import host
    with pytest.raises(host.KazanError) as e:
        backend.translate_spirv_to_ir(shaders.recursive(), "main", backend.EXECUTION_MODEL_VERTEX, 2)
    assert e.value.code == -8 and "recursion" in str(e.value)
