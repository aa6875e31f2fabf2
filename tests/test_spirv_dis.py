"""SPIR-V reader / disassembler against the reference's own known-answer tests.

The three goldens of spirv-parser/src/lib.rs:46-215 (test.spv, test2.spv, test3.spv) must be reproduced byte for byte by
kzn_spirv_disassemble; no GPU is involved.  Error behaviour follows spirv-parser-generator/src/generate.rs:1600-1642.
"""
import base64
import json
import struct
from pathlib import Path

import pytest

from kazan_b200 import abi, backend, host

GOLDENS = json.loads((Path(__file__).parent / "golden" / "spirv_goldens.json").read_text())["cases"]


def _words(data: bytes):
    assert len(data) % 4 == 0
    return list(struct.unpack(f"<{len(data) // 4}I", data))


@pytest.mark.parametrize("case", GOLDENS, ids=[c["test"] for c in GOLDENS])
def test_reference_disassembly_goldens(case):
    got = backend.disassemble_spirv(_words(base64.b64decode(case["spv_base64"])))
    want = case["expected"]
    for a, b in zip(got.splitlines(), want.splitlines()):
        assert a == b
    assert got == want


def _module(*instructions, bound=16, version=0x00010000, generator=0x00080001):
    words = [0x07230203, version, generator, bound, 0]
    for opcode, *operands in instructions:
        words.append(((len(operands) + 1) << 16) | opcode)
        words.extend(operands)
    return words


def _string(s: str):
    data = s.encode() + b"\0"
    data += b"\0" * (-len(data) % 4)
    return list(struct.unpack(f"<{len(data) // 4}I", data))


def test_header_and_columns():
    text = backend.disassemble_spirv(_module((17, 1), (19, 2), bound=3, version=0x00010300, generator=0xABC))
    assert text == ("; SPIR-V\n; Version: 1.3\n; Generator: 0xABC\n; Bound: 3\n; Schema: 0\n"
                    "               OpCapability Shader ; 0x00000014\n"
                    "          %2 = OpTypeVoid ; 0x0000001c\n")


def test_64_bit_constants_and_switch_literals():
    # OpTypeInt 64 0; OpConstant (two words, 16 hex digits); OpSwitch on a 64-bit selector takes two-word literals
    m = _module((21, 1, 64, 0), (43, 1, 2, 0x89ABCDEF, 0x01234567), (251, 2, 3, 5, 1, 4), bound=6)
    lines = backend.disassemble_spirv(m).splitlines()[5:]
    assert lines[0].startswith("          %1 = OpTypeInt 64 0")
    assert lines[1].startswith("          %2 = OpConstant %1 0x0123456789ABCDEF")
    assert lines[2].startswith("               OpSwitch %2 %3 4294967301 %4")


def test_strings_are_rust_debug_formatted():
    m = _module((5, 1, *_string('a"b\\c\n')), bound=2)
    assert 'OpName %1 "a\\"b\\\\c\\n"' in backend.disassemble_spirv(m)


def test_bit_enums_and_their_parameters():
    # OpLoad with Volatile|Aligned 16; OpLoopMerge None; OpFunction Inline|Pure
    m = _module((61, 1, 2, 3, 0x3, 16), (246, 4, 5, 0), (54, 1, 6, 0x5, 7), bound=8)
    text = backend.disassemble_spirv(m)
    assert "%2 = OpLoad %1 %3 Volatile|Aligned 16 ;" in text
    assert "OpLoopMerge %4 %5 None ;" in text
    assert "%6 = OpFunction %1 Inline|Pure %7 ;" in text


def test_decorations_with_parameters():
    m = _module((71, 1, 11, 28), (71, 1, 41, *_string("sym"), 1), (72, 2, 3, 35, 64), bound=4)
    text = backend.disassemble_spirv(m)
    assert "OpDecorate %1 BuiltIn GlobalInvocationId ;" in text
    assert 'OpDecorate %1 LinkageAttributes "sym" Import ;' in text
    assert "OpMemberDecorate %2 3 Offset 64 ;" in text


@pytest.mark.parametrize("words, message", [
    ([0x07230203, 0x10000, 0, 4], "missing the file header"),
    ([0x03022307, 0x10000, 0, 4, 0], "invalid file header"),
    (_module((0xFFF, ), bound=4), "unknown opcode: 4095"),
    (_module((19, 9), bound=4), "id is out of bounds: 9"),
    (_module((19, 1), (19, 1), bound=4), "id is already defined: 1"),
    ([0x07230203, 0x10000, 0, 4, 0, 19], "length of zero"),
    ([0x07230203, 0x10000, 0, 4, 0, (3 << 16) | 19, 1], "prematurely ended"),
    (_module((19, 1, 2), bound=4), "instruction is too long"),
    (_module((17, 9999), bound=4), "enum has invalid value"),
    (_module((43, 1, 2, 5), bound=4), "undefined type 1"),
    (_module((21, 1, 128, 0), bound=4), "unsupported int size"),
    (_module((5, 1, 0x00FF41), bound=4), "invalid utf-8 sequence of 1 bytes from index 1"),
    (_module((5, 1, 0x00C341), bound=4), "incomplete utf-8 byte sequence from index 1"),
    (_module((19, 1), bound=0x400000), "exceeds the universal limit"),
])
def test_malformed_modules_report_the_reference_errors(words, message):
    with pytest.raises(host.KazanError) as e:
        backend.disassemble_spirv(words)
    assert e.value.code == abi.ERROR_VALIDATION_FAILED and message in str(e.value)


def test_subgroup_instructions_of_vulkan_1_1():
    """OpGroupNonUniform*: the scope is an id, the group operation an enumerant, ClusteredReduce carries a cluster size."""
    m = _module((21, 1, 32, 0), (43, 1, 2, 3), (339, 1, 4, 2, 5), (349, 1, 6, 2, 3, 7, 8), (366, 1, 9, 2, 7, 10),
                (342, 1, 11, 2, 0, 12), bound=16, version=0x00010300)
    text = backend.disassemble_spirv(m)
    assert "%4 = OpGroupNonUniformBallot %1 %2 %5 ;" in text
    assert "%6 = OpGroupNonUniformIAdd %1 %2 ClusteredReduce %7 %8 ;" in text
    assert "%9 = OpGroupNonUniformQuadSwap %1 %2 %7 %10 ;" in text
    assert "%11 = OpGroupNonUniformBallotBitCount %1 %2 Reduce %12 ;" in text


def test_the_translators_fixture_disassembles_consistently_with_its_translation():
    """trivial_test.spv (the SPIR-V -> IR golden of shader-compiler-translate-spirv-to-ir): what the reader prints and what
    the translator produces describe the same module — entry point, LocalSize, the OpLine location, the source string."""
    ir = json.loads((Path(__file__).parent / "golden" / "ir_goldens.json").read_text())["spirv_to_ir"][0]
    words = _words(bytes.fromhex(ir["spv_hex"]))
    text = backend.disassemble_spirv(words)
    assert 'OpEntryPoint GLCompute %5 "main" ;' in text and "OpExecutionMode %5 LocalSize 1 1 1 ;" in text
    assert '%1 = OpString "stdin" ;' in text and "OpLine %1 4 0 ;" in text
    assert '\\n#version 450\\n\\nvoid main() {\\n    return;\\n}"' in text  # the source literal, Rust-debug escaped
    module, local_size = backend.translate_spirv_to_ir(words, "main", backend.EXECUTION_MODEL_GLCOMPUTE)
    assert local_size == (1, 1, 1)
    translated = module.to_text()
    module.destroy()
    assert translated == ir["expected_ir"] and '"stdin":4:0' in translated  # the OpLine the reader printed, attached in the IR
