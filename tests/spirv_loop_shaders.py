"""Two hand-assembled SPIR-V vertex shaders with a loop (tests/spirv_asm.py), and what they compute.

    in  location 0: float x      in  location 1: int n
    out location 0: float acc    out location 1: int i

`variables` — what glslang emits for
    float acc = 0; int i;
    for (i = 0; i < n; ++i) { if ((i & 1) == 1) continue; acc = acc * 0.5 + x + float(i); if (acc > 100.0) break; }
  loop state in Function variables; the header branches to a test block whose conditional branch has no selection merge; a
  `continue` and a `break` from inside selection constructs.
`phis` — the optimised shape of  for (i = 0; i < n; ++i) acc = acc * 0.5 + x;  loop state in OpPhis of the header, the
  conditional branch in the header itself, OpPhis in the continue block and in the merge block.
`do_while`, `nested` — see the functions.
"""
import struct

import numpy as np

from tests.spirv_asm import FUNCTION, INPUT, LOCATION, OUTPUT, assemble, lit


def _f32(x):
    return struct.unpack("<I", struct.pack("<f", x))[0]


class _Ids(dict):
    def __call__(self, name):
        return self.setdefault(name, len(self) + 1)


def _prologue(i):
    return [("Capability", 1), ("MemoryModel", 0, 1),
            ("EntryPoint", 0, i("main"), lit("main"), i("xin"), i("nin"), i("acc_out"), i("i_out")),
            ("Decorate", i("xin"), LOCATION, 0), ("Decorate", i("nin"), LOCATION, 1), ("Decorate", i("acc_out"), LOCATION, 0),
            ("Decorate", i("i_out"), LOCATION, 1),
            ("TypeVoid", i("void")), ("TypeFunction", i("fn"), i("void")), ("TypeFloat", i("float"), 32), ("TypeInt", i("int"), 32, 1),
            ("TypeBool", i("bool")),
            ("TypePointer", i("p_in_f"), INPUT, i("float")), ("TypePointer", i("p_in_i"), INPUT, i("int")),
            ("TypePointer", i("p_out_f"), OUTPUT, i("float")), ("TypePointer", i("p_out_i"), OUTPUT, i("int")),
            ("TypePointer", i("p_fn_f"), FUNCTION, i("float")), ("TypePointer", i("p_fn_i"), FUNCTION, i("int")),
            ("Constant", i("int"), i("i0"), 0), ("Constant", i("int"), i("i1"), 1), ("Constant", i("float"), i("f0"), _f32(0.0)),
            ("Constant", i("float"), i("f05"), _f32(0.5)), ("Constant", i("float"), i("f100"), _f32(100.0)),
            ("Variable", i("p_in_f"), i("xin"), INPUT), ("Variable", i("p_in_i"), i("nin"), INPUT),
            ("Variable", i("p_out_f"), i("acc_out"), OUTPUT), ("Variable", i("p_out_i"), i("i_out"), OUTPUT),
            ("Function", i("void"), i("main"), 0, i("fn"))]


def variables():
    i = _Ids()
    body = [("Label", i("entry")),
            ("Variable", i("p_fn_f"), i("acc"), FUNCTION), ("Variable", i("p_fn_i"), i("i"), FUNCTION),
            ("Store", i("acc"), i("f0")), ("Store", i("i"), i("i0")),
            ("Load", i("int"), i("n"), i("nin")), ("Load", i("float"), i("x"), i("xin")),
            ("Branch", i("H")),
            ("Label", i("H")), ("LoopMerge", i("M"), i("C"), 0), ("Branch", i("T")),
            ("Label", i("T")), ("Load", i("int"), i("iv"), i("i")), ("SLessThan", i("bool"), i("c"), i("iv"), i("n")),
            ("BranchConditional", i("c"), i("B"), i("M")),
            ("Label", i("B")), ("Load", i("int"), i("iv2"), i("i")), ("BitwiseAnd", i("int"), i("odd"), i("iv2"), i("i1")),
            ("IEqual", i("bool"), i("isodd"), i("odd"), i("i1")), ("SelectionMerge", i("S1"), 0),
            ("BranchConditional", i("isodd"), i("Cont1"), i("S1")),
            ("Label", i("Cont1")), ("Branch", i("C")),
            ("Label", i("S1")), ("Load", i("float"), i("a0"), i("acc")), ("FMul", i("float"), i("a1"), i("a0"), i("f05")),
            ("FAdd", i("float"), i("a2"), i("a1"), i("x")), ("ConvertSToF", i("float"), i("fi"), i("iv2")),
            ("FAdd", i("float"), i("a3"), i("a2"), i("fi")), ("Store", i("acc"), i("a3")),
            ("FOrdGreaterThan", i("bool"), i("big"), i("a3"), i("f100")), ("SelectionMerge", i("S2"), 0),
            ("BranchConditional", i("big"), i("Brk"), i("S2")),
            ("Label", i("Brk")), ("Branch", i("M")),
            ("Label", i("S2")), ("Branch", i("C")),
            ("Label", i("C")), ("Load", i("int"), i("iv3"), i("i")), ("IAdd", i("int"), i("inc"), i("iv3"), i("i1")),
            ("Store", i("i"), i("inc")), ("Branch", i("H")),
            ("Label", i("M")), ("Load", i("float"), i("af"), i("acc")), ("Store", i("acc_out"), i("af")),
            ("Load", i("int"), i("if"), i("i")), ("Store", i("i_out"), i("if")),
            ("Return",), ("FunctionEnd",)]
    pro = _prologue(i)
    return assemble(pro + body, 200)


def phis():
    i = _Ids()
    pro = _prologue(i)
    body = [("Label", i("entry")), ("Load", i("int"), i("n"), i("nin")), ("Load", i("float"), i("x"), i("xin")), ("Branch", i("H")),
            ("Label", i("H")),
            ("Phi", i("int"), i("iv"), i("i0"), i("entry"), i("inc"), i("C")),
            ("Phi", i("float"), i("acc"), i("f0"), i("entry"), i("accn"), i("C")),
            ("LoopMerge", i("M"), i("C"), 0), ("SLessThan", i("bool"), i("c"), i("iv"), i("n")),
            ("BranchConditional", i("c"), i("B"), i("M")),
            ("Label", i("B")), ("FMul", i("float"), i("a1"), i("acc"), i("f05")), ("FAdd", i("float"), i("a2"), i("a1"), i("x")),
            ("Branch", i("C")),
            ("Label", i("C")), ("Phi", i("float"), i("accn"), i("a2"), i("B")), ("IAdd", i("int"), i("inc"), i("iv"), i("i1")),
            ("Branch", i("H")),
            ("Label", i("M")), ("Phi", i("float"), i("rf"), i("acc"), i("H")), ("Phi", i("int"), i("ri"), i("iv"), i("H")),
            ("Store", i("acc_out"), i("rf")), ("Store", i("i_out"), i("ri")),
            ("Return",), ("FunctionEnd",)]
    return assemble(pro + body, 200)


def header_values_after_loop():
    """`phis` without OpPhis in the continue and merge blocks: the continue block uses a value of the body, the code after the
    loop values of the header — legal SPIR-V (the definitions dominate the uses), outside the defining block in the IR: the
    translator hands them out through block results"""
    i = _Ids()
    pro = _prologue(i)
    body = [("Label", i("entry")), ("Load", i("int"), i("n"), i("nin")), ("Load", i("float"), i("x"), i("xin")), ("Branch", i("H")),
            ("Label", i("H")),
            ("Phi", i("int"), i("iv"), i("i0"), i("entry"), i("inc"), i("C")),
            ("Phi", i("float"), i("acc"), i("f0"), i("entry"), i("a2"), i("C")),
            ("LoopMerge", i("M"), i("C"), 0), ("SLessThan", i("bool"), i("c"), i("iv"), i("n")),
            ("FAdd", i("float"), i("hsum"), i("acc"), i("f0")),
            ("BranchConditional", i("c"), i("B"), i("M")),
            ("Label", i("B")), ("FMul", i("float"), i("a1"), i("acc"), i("f05")), ("FAdd", i("float"), i("a2"), i("a1"), i("x")),
            ("Branch", i("C")),
            ("Label", i("C")), ("IAdd", i("int"), i("inc"), i("iv"), i("i1")), ("Branch", i("H")),
            ("Label", i("M")), ("Store", i("acc_out"), i("hsum")), ("Store", i("i_out"), i("iv")),
            ("Return",), ("FunctionEnd",)]
    return assemble(pro + body, 200)


def do_while():
    """do { acc = acc * 0.5 + x; ++i; } while (i < n);  — the conditional back edge sits in the continue block (no selection
    merge: `true` continues, `false` leaves), the header only branches into the body; loop state in OpPhis"""
    i = _Ids()
    pro = _prologue(i)
    body = [("Label", i("entry")), ("Load", i("int"), i("n"), i("nin")), ("Load", i("float"), i("x"), i("xin")), ("Branch", i("H")),
            ("Label", i("H")),
            ("Phi", i("int"), i("iv"), i("i0"), i("entry"), i("inc"), i("C")),
            ("Phi", i("float"), i("acc"), i("f0"), i("entry"), i("a2"), i("C")),
            ("LoopMerge", i("M"), i("C"), 0), ("Branch", i("B")),
            ("Label", i("B")), ("FMul", i("float"), i("a1"), i("acc"), i("f05")), ("FAdd", i("float"), i("a2"), i("a1"), i("x")),
            ("Branch", i("C")),
            ("Label", i("C")), ("IAdd", i("int"), i("inc"), i("iv"), i("i1")), ("SLessThan", i("bool"), i("c"), i("inc"), i("n")),
            ("BranchConditional", i("c"), i("H"), i("M")),
            ("Label", i("M")), ("Phi", i("float"), i("rf"), i("a2"), i("C")), ("Phi", i("int"), i("ri"), i("inc"), i("C")),
            ("Store", i("acc_out"), i("rf")), ("Store", i("i_out"), i("ri")),
            ("Return",), ("FunctionEnd",)]
    return assemble(pro + body, 200)


def nested():
    """for (i = 0; i < n; ++i) for (j = 0; j < i; ++j) acc = acc * 0.5 + x + float(j);  — a loop inside a loop's body, both
    with their state in Function variables; out.i = i"""
    i = _Ids()
    pro = _prologue(i)
    body = [("Label", i("entry")),
            ("Variable", i("p_fn_f"), i("acc"), FUNCTION), ("Variable", i("p_fn_i"), i("i"), FUNCTION), ("Variable", i("p_fn_i"), i("j"), FUNCTION),
            ("Store", i("acc"), i("f0")), ("Store", i("i"), i("i0")),
            ("Load", i("int"), i("n"), i("nin")), ("Load", i("float"), i("x"), i("xin")), ("Branch", i("H")),
            ("Label", i("H")), ("LoopMerge", i("M"), i("C"), 0), ("Branch", i("T")),
            ("Label", i("T")), ("Load", i("int"), i("iv"), i("i")), ("SLessThan", i("bool"), i("c"), i("iv"), i("n")),
            ("BranchConditional", i("c"), i("B"), i("M")),
            ("Label", i("B")), ("Store", i("j"), i("i0")), ("Branch", i("H2")),
            ("Label", i("H2")), ("LoopMerge", i("M2"), i("C2"), 0), ("Branch", i("T2")),
            ("Label", i("T2")), ("Load", i("int"), i("jv"), i("j")), ("Load", i("int"), i("iv4"), i("i")),
            ("SLessThan", i("bool"), i("c2"), i("jv"), i("iv4")), ("BranchConditional", i("c2"), i("B2"), i("M2")),
            ("Label", i("B2")), ("Load", i("float"), i("a0"), i("acc")), ("FMul", i("float"), i("a1"), i("a0"), i("f05")),
            ("FAdd", i("float"), i("a2"), i("a1"), i("x")), ("Load", i("int"), i("jv2"), i("j")), ("ConvertSToF", i("float"), i("fj"), i("jv2")),
            ("FAdd", i("float"), i("a3"), i("a2"), i("fj")), ("Store", i("acc"), i("a3")), ("Branch", i("C2")),
            ("Label", i("C2")), ("Load", i("int"), i("jv3"), i("j")), ("IAdd", i("int"), i("jinc"), i("jv3"), i("i1")), ("Store", i("j"), i("jinc")),
            ("Branch", i("H2")),
            ("Label", i("M2")), ("Branch", i("C")),
            ("Label", i("C")), ("Load", i("int"), i("iv3"), i("i")), ("IAdd", i("int"), i("inc"), i("iv3"), i("i1")), ("Store", i("i"), i("inc")),
            ("Branch", i("H")),
            ("Label", i("M")), ("Load", i("float"), i("af"), i("acc")), ("Store", i("acc_out"), i("af")),
            ("Load", i("int"), i("if"), i("i")), ("Store", i("i_out"), i("if")),
            ("Return",), ("FunctionEnd",)]
    return assemble(pro + body, 200)


def calls():
    """Functions, the way glslang emits them (arguments passed as pointers to Function variables) and the way an optimiser leaves
    them (by value):
        float halve_add(float a, float b) { return a * 0.5 + b; }               // by value
        void  step(inout float acc, float x, inout int i) { acc = halve_add(acc, x); i = i + 1; }   // pointer parameters, a nested call
        main: acc = 0; i = 0; while (i < n) step(acc, x, i); out = acc, i      // a call inside a loop: inlined once per call site
    """
    i = _Ids()
    pro = _prologue(i)
    pro = pro[:-1] + [("TypeFunction", i("fn_ff_f"), i("float"), i("float"), i("float")),
                      ("TypeFunction", i("fn_step"), i("void"), i("p_fn_f"), i("float"), i("p_fn_i")), pro[-1]]
    main = [("Label", i("entry")),
            ("Variable", i("p_fn_f"), i("acc"), FUNCTION), ("Variable", i("p_fn_i"), i("i"), FUNCTION),
            ("Store", i("acc"), i("f0")), ("Store", i("i"), i("i0")),
            ("Load", i("int"), i("n"), i("nin")), ("Load", i("float"), i("x"), i("xin")), ("Branch", i("H")),
            ("Label", i("H")), ("LoopMerge", i("M"), i("C"), 0), ("Branch", i("T")),
            ("Label", i("T")), ("Load", i("int"), i("iv"), i("i")), ("SLessThan", i("bool"), i("c"), i("iv"), i("n")),
            ("BranchConditional", i("c"), i("B"), i("M")),
            ("Label", i("B")), ("FunctionCall", i("void"), i("call1"), i("step"), i("acc"), i("x"), i("i")), ("Branch", i("C")),
            ("Label", i("C")), ("Branch", i("H")),
            ("Label", i("M")), ("Load", i("float"), i("af"), i("acc")), ("Store", i("acc_out"), i("af")),
            ("Load", i("int"), i("if"), i("i")), ("Store", i("i_out"), i("if")),
            ("Return",), ("FunctionEnd",)]
    halve_add = [("Function", i("float"), i("halve_add"), 0, i("fn_ff_f")),
                 ("FunctionParameter", i("float"), i("pa")), ("FunctionParameter", i("float"), i("pb")),
                 ("Label", i("ha_entry")), ("FMul", i("float"), i("h1"), i("pa"), i("f05")), ("FAdd", i("float"), i("h2"), i("h1"), i("pb")),
                 ("ReturnValue", i("h2")), ("FunctionEnd",)]
    step = [("Function", i("void"), i("step"), 0, i("fn_step")),
            ("FunctionParameter", i("p_fn_f"), i("s_acc")), ("FunctionParameter", i("float"), i("s_x")), ("FunctionParameter", i("p_fn_i"), i("s_i")),
            ("Label", i("s_entry")), ("Load", i("float"), i("s_a"), i("s_acc")),
            ("FunctionCall", i("float"), i("s_r"), i("halve_add"), i("s_a"), i("s_x")), ("Store", i("s_acc"), i("s_r")),
            ("Load", i("int"), i("s_iv"), i("s_i")), ("IAdd", i("int"), i("s_inc"), i("s_iv"), i("i1")), ("Store", i("s_i"), i("s_inc")),
            ("Return",), ("FunctionEnd",)]
    return assemble(pro + main + halve_add + step, 200)


def recursive():
    """a function that calls itself: refused (the IR has no call instruction and recursion cannot be inlined)"""
    i = _Ids()
    pro = _prologue(i)
    pro = pro[:-1] + [("TypeFunction", i("fn_f_f"), i("float"), i("float")), pro[-1]]
    main = [("Label", i("entry")), ("Load", i("float"), i("x"), i("xin")), ("FunctionCall", i("float"), i("r"), i("rec"), i("x")),
            ("Store", i("acc_out"), i("r")), ("Return",), ("FunctionEnd",)]
    rec = [("Function", i("float"), i("rec"), 0, i("fn_f_f")), ("FunctionParameter", i("float"), i("p")), ("Label", i("r_entry")),
           ("FunctionCall", i("float"), i("r2"), i("rec"), i("p")), ("ReturnValue", i("r2")), ("FunctionEnd",)]
    return assemble(pro + main + rec, 200)


SLOT = np.dtype([("x", "<f4"), ("pad0", "V12"), ("n", "<i4"), ("pad1", "V12")])       # one Location per 16 bytes
OUT = np.dtype([("acc", "<f4"), ("pad0", "V12"), ("i", "<i4"), ("pad1", "V12")])


def inputs(count=64, seed=5):
    rng = np.random.default_rng(seed)
    a = np.zeros(count, dtype=SLOT)
    a["x"] = rng.uniform(-3, 12, count).astype(np.float32)
    a["n"] = rng.integers(-2, 24, count)
    a["n"][:4] = [0, 1, 2, 39]
    return a


def expected(kind, a):
    out = np.zeros(len(a), dtype=OUT)
    f = np.float32
    for k, (x, n) in enumerate(zip(a["x"], a["n"])):
        acc, i = f(0.0), 0
        if kind == "variables":
            while i < n:
                if (i & 1) != 1:
                    acc = f(f(f(acc * f(0.5)) + x) + f(i))
                    if acc > f(100.0):
                        break
                i += 1
        elif kind in ("phis", "header_values_after_loop", "calls"):
            while i < n:
                acc = f(f(acc * f(0.5)) + x)
                i += 1
        elif kind == "do_while":
            while True:
                acc = f(f(acc * f(0.5)) + x)
                i += 1
                if not i < n:
                    break
        else:  # nested
            while i < n:
                for j in range(i):
                    acc = f(f(f(acc * f(0.5)) + x) + f(j))
                i += 1
        out["acc"][k], out["i"][k] = acc, i
    return out
