"""A few lines of SPIR-V assembler for tests: the image has no glslang / spirv-as, and the reference's fixtures
(trivial_test.spv, simple_test.spv, spirv-parser/test_inputs) contain no loop.  Instructions are (opcode, operands...) tuples
with integer ids; strings are encoded as SPIR-V literal strings."""
from __future__ import annotations

OP = dict(Capability=17, ExtInstImport=11, MemoryModel=14, EntryPoint=15, Name=5, Decorate=71, TypeVoid=19, TypeBool=20, TypeInt=21,
          TypeFloat=22, TypePointer=32, TypeFunction=33, Constant=43, Variable=59, Function=54, FunctionEnd=56, Label=248, Load=61,
          Store=62, Branch=249, BranchConditional=250, LoopMerge=246, SelectionMerge=247, Return=253, Phi=245, IAdd=128, FAdd=129,
          FMul=133, SLessThan=177, BitwiseAnd=199, IEqual=170, ConvertSToF=111, FOrdGreaterThan=186, FunctionParameter=55, FunctionCall=57, ReturnValue=254)
INPUT, OUTPUT, FUNCTION = 1, 3, 7
LOCATION = 30


def lit(s: str) -> list[int]:
    b = s.encode() + b"\0"
    b += b"\0" * (-len(b) % 4)
    return [int.from_bytes(b[i:i + 4], "little") for i in range(0, len(b), 4)]


def assemble(instructions, bound: int) -> list[int]:
    words = [0x07230203, 0x00010000, 0, bound, 0]
    for op, *operands in instructions:
        flat = []
        for o in operands:
            flat += o if isinstance(o, list) else [o]
        words.append((len(flat) + 1) << 16 | OP[op])
        words += flat
    return words
