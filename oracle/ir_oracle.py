"""oracle/ir_oracle.py — reader, printer and interpreter for the `.kazan-ir` text format.  TEST INFRASTRUCTURE.

Restates (kazan-3d/kazan @ f2820899):
    lexer / names / printer rules      shader-compiler-ir/src/text.rs:409-416,493-540,613-645,1133-1326,1417-1613
    module / interface blocks          shader-compiler-ir/src/module.rs:47-150, interface.rs:61-67,199-252,490-544
    functions, variables               shader-compiler-ir/src/function.rs:155-193,217-231,357-434
    blocks, loops, break / continue    shader-compiler-ir/src/block.rs:16-124,251-300,455-503,623-767
    add / load / branch                shader-compiler-ir/src/instructions_impl.rs:326-411,369-386,502-526
    types, constants                   shader-compiler-ir/src/types.rs:42-126,1049-1204, consts.rs:277-447,533-567
PINNED: the printer/parser pair is checked against every text vector the reference's own unit tests hold
(tests/golden/ir_goldens.json, extracted by tests/golden/extract_ir_goldens.py) in tests/test_ir.py.
UNPINNED: execution semantics — the reference never executes IR (its LLVM back end builds only `alloca`/`ret`,
shader-compiler-backend/src/lib.rs:21-29); the interpreter below follows docs/Shader Compiler IR Example.md and the
structural rules in block.rs (a block is left by `break`, a loop re-entered by `continue`).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

KEYWORDS = set("align bool const data_ptr dont_inline entry_point f16 f32 f64 false fixed fn i16 i32 i64 i8 inline kind "
               "module none normal null pure rf32 ri32 size struct true undef variable vscale x".split())
INT_TYPES = {"i8": 8, "i16": 16, "i32": 32, "ri32": 32, "i64": 64}
FLOAT_TYPES = {"f16": 16, "f32": 32, "rf32": 32, "f64": 64}


class IrError(ValueError):
    pass


# Version 2 of the instruction set (kazan_b200/csrc/ir.h `kExtOps`): name -> (argument count, result count).  The reference
# has none of these (instructions_impl.rs:395-433); the C++ mirror reads them only when asked for version 2, and so does
# this parser.  Execution: oracle/ir_simt.py.
V2_OPS = {
    **{n: (2, 1) for n in ("i_sub i_mul bitwise_and bitwise_or bitwise_xor shift_left_logical shift_right_logical "
                           "shift_right_arithmetic f_add f_sub f_mul f_div f_min f_max i_equal i_not_equal u_less_than s_less_than "
                           "u_less_than_equal s_less_than_equal u_greater_than s_greater_than u_greater_than_equal "
                           "s_greater_than_equal f_ord_equal f_ord_less_than f_ord_less_than_equal f_ord_greater_than "
                           "f_ord_greater_than_equal f_unord_not_equal logical_and logical_or ptr_add group_non_uniform_shuffle "
                           "group_non_uniform_shuffle_xor").split()},
    **{n: (1, 1) for n in ("s_negate not f_negate f_abs logical_not convert_f_to_u convert_f_to_s convert_s_to_f convert_u_to_f "
                           "u_convert s_convert f_convert bitcast group_non_uniform_all group_non_uniform_any "
                           "group_non_uniform_ballot group_non_uniform_broadcast_first group_non_uniform_i_add").split()},
    "fma": (3, 1), "select": (3, 1), "store": (2, 0), "group_non_uniform_elect": (0, 1),
}
V2_BUILT_INS = ("global_invocation_id", "subgroup_local_invocation_id", "subgroup_size", "frag_coord", "push_constants",
                "descriptor_table")


# ---- lexer (text.rs:1133-1326) ----------------------------------------------------------------------------------

_TOKEN = re.compile(r"""
    (?P<ws>[ \t\r\n\f\v]+|\#[^\n]*)
  | (?P<id>[A-Za-z_][A-Za-z0-9_]*)
  | (?P<int>0[xX][0-9A-Fa-f]+|0[oO][0-7]+|0[bB][01]+|[0-9]+)(?P<suffix>ri32|i8|i16|i32|i64)?
  | (?P<str>"(?:[^"\\]|\\.)*")
  | (?P<punct>->|[{}\[\]()<>,;:!@*=])
""", re.X)


@dataclass
class Tok:
    kind: str
    text: str
    value: object = None
    suffix: str | None = None
    pos: int = 0


def _unescape(s: str) -> str:
    out, i = [], 0
    while i < len(s):
        c = s[i]
        if c != "\\":
            out.append(c); i += 1; continue
        n = s[i + 1]
        if n == "u":
            j = s.index("}", i)
            out.append(chr(int(s[i + 3:j], 16))); i = j + 1
        else:
            out.append({"0": "\0", "n": "\n", "r": "\r", "t": "\t", "'": "'", '"': '"', "\\": "\\"}[n]); i += 2
    return "".join(out)


def _escape(s: str) -> str:
    out = []
    for c in s:
        if c in '"\\':
            out.append("\\" + c)
        elif c == "\n":
            out.append("\\n")
        elif c == "\r":
            out.append("\\r")
        elif c == "\t":
            out.append("\\t")
        elif c == "\0":
            out.append("\\0")
        elif ord(c) < 0x20 or ord(c) == 0x7f:
            out.append("\\u{%X}" % ord(c))
        else:
            out.append(c)
    return "".join(out)


def lex(text: str) -> list[Tok]:
    toks, i = [], 0
    while i < len(text):
        m = _TOKEN.match(text, i)
        if not m:
            raise IrError(f"unexpected character {text[i]!r} at offset {i}")
        i = m.end()
        if m.group("ws") is not None:
            continue
        if m.group("id") is not None:
            toks.append(Tok("id", m.group("id"), pos=m.start()))
        elif m.group("int") is not None:
            t = m.group("int")
            if len(t) > 1 and t[0] == "0" and t[1].isdigit():
                raise IrError("octal numbers must start with 0o")  # text.rs:1210-1213
            if re.match(r"[A-Za-z0-9_]", text[i:i + 1] or " "):
                raise IrError(f"invalid number suffix at offset {i}")
            base = {"x": 16, "o": 8, "b": 2}.get(t[1:2].lower(), 10) if len(t) > 1 else 10
            toks.append(Tok("int", t, int(t[2:], base) if base != 10 else int(t), m.group("suffix"), m.start()))
        elif m.group("str") is not None:
            toks.append(Tok("str", m.group("str"), _unescape(m.group("str")[1:-1]), pos=m.start()))
        else:
            toks.append(Tok("punct", m.group("punct"), pos=m.start()))
    toks.append(Tok("eof", ""))
    return toks


# ---- IR data model -------------------------------------------------------------------------------------------------

@dataclass(frozen=True)
class Type:
    kind: str                       # int float bool data_ptr fn vector struct
    name: str = ""                  # int/float keyword
    args: tuple = ()                # fn
    results: tuple | None = None    # fn; None = diverging
    scalable: bool = False          # vector
    length: int = 0
    element: "Type | None" = None
    size_kind: str = ""             # struct: fixed | variable
    size: int = 0
    align: int = 1
    members: tuple = ()             # ((offset, Type), ...)


I32, I64, BOOL, DATA_PTR = Type("int", "i32"), Type("int", "i64"), Type("bool"), Type("data_ptr")


@dataclass(frozen=True)
class Const:
    kind: str        # int float bool vector undef null fn
    type: Type | None = None
    value: object = None


@dataclass(eq=False)
class Value:
    name: str
    type: Type
    const: Const | None = None
    suffix_hint: int = 0


@dataclass(eq=False)
class Variable:
    type: Type
    align: int
    pointer: Value
    kind: str | None = None  # interface built-in kind


@dataclass(eq=False)
class Block:
    name: str
    results: list | None  # list[Value] or None (diverging)
    body: list = field(default_factory=list)


@dataclass(eq=False)
class Loop:
    name: str
    arguments: list
    params: list
    results: list | None
    body: Block = None


@dataclass(eq=False)
class Instr:
    op: str                       # add load break continue branch block loop
    args: list = field(default_factory=list)
    results: list = field(default_factory=list)
    align: int = 0
    target: object = None         # Block / Loop for break / continue; Block/Loop object for block / loop
    cases: list = field(default_factory=list)   # branch: [(Const, Block, [uses])]
    location: tuple | None = None


@dataclass(eq=False)
class Function:
    name: str
    params: list
    returns: list | None
    inlining_hint: str = "none"
    side_effects: str = "normal"
    locals: list = field(default_factory=list)
    body: Block = None

    @property
    def type(self) -> Type:
        return Type("fn", args=tuple(p.type for p in self.params),
                    results=None if self.returns is None else tuple(self.returns))


@dataclass(eq=False)
class InterfaceBlock:
    pointer: Value
    size_kind: str
    size: int
    align: int
    members: list  # [(offset, Type)]


@dataclass(eq=False)
class Module:
    data_ptr_type: str = "i64"
    fn_ptr_type: str = "i64"
    built_in_inputs: list = field(default_factory=list)
    user_inputs: InterfaceBlock = None
    built_in_outputs: list = field(default_factory=list)
    user_outputs: InterfaceBlock = None
    globals: list = field(default_factory=list)
    functions: list = field(default_factory=list)
    entry_point: Function = None


# ---- parser ---------------------------------------------------------------------------------------------------------

class Parser:
    def __init__(self, text: str, version: int = 1):
        self.version = version
        self.toks, self.i = lex(text), 0
        self.structs: dict[int, Type] = {}
        self.values: list[dict] = [{}]     # scope stack: (name, suffix) -> Value
        self.blocks: list[dict] = [{}]
        self.loops: list[dict] = [{}]
        self.functions: dict = {}

    # token helpers
    @property
    def t(self) -> Tok:
        return self.toks[self.i]

    def next(self) -> Tok:
        t = self.toks[self.i]; self.i += 1; return t

    def at(self, text: str) -> bool:
        return self.t.kind in ("punct", "id") and self.t.text == text

    def accept(self, text: str) -> bool:
        if self.at(text):
            self.i += 1; return True
        return False

    def expect(self, text: str):
        if not self.accept(text):
            raise IrError(f"expected {text!r}, found {self.t.text!r} at offset {self.t.pos}")

    def finish(self):
        if self.t.kind != "eof":
            raise IrError(f"unexpected token {self.t.text!r} at offset {self.t.pos}")

    def integer(self) -> int:
        t = self.next()
        if t.kind != "int" or t.suffix:
            raise IrError(f"expected an unsuffixed integer at offset {t.pos}")
        return t.value

    def name(self) -> tuple[str, int]:
        t = self.next()
        if t.kind == "id" and t.text not in KEYWORDS:
            return t.text, 0
        if t.kind == "str":
            return t.value, self.integer()
        raise IrError(f"expected a name, found {t.text!r} at offset {t.pos}")

    # scopes
    def push(self):
        self.values.append({}); self.blocks.append({}); self.loops.append({})

    def pop(self):
        self.values.pop(); self.blocks.pop(); self.loops.pop()

    def lookup(self, stack, key, what):
        for scope in reversed(stack):
            if key in scope:
                return scope[key]
        raise IrError(f"{what} {key[0]!r} is not defined")

    def define_value(self, key, value, root=False):
        scope = self.values[0] if root else self.values[-1]
        if key in scope:
            raise IrError(f"value {key[0]!r} defined twice")
        scope[key] = value

    # types (types.rs:1049-1204)
    def type(self) -> Type:
        t = self.t
        if t.kind == "id" and t.text in INT_TYPES:
            self.i += 1; return Type("int", t.text)
        if t.kind == "id" and t.text in FLOAT_TYPES:
            self.i += 1; return Type("float", t.text)
        if self.accept("bool"):
            return BOOL
        if self.accept("data_ptr"):
            return DATA_PTR
        if self.accept("("):
            ty = self.type(); self.expect(")"); return ty
        if self.accept("<"):
            scalable = False
            if self.accept("vscale"):
                self.expect("x"); scalable = True
            n = self.integer(); self.expect("x"); el = self.type(); self.expect(">")
            return Type("vector", scalable=scalable, length=n, element=el)
        if self.accept("fn"):
            return self.fn_type_tail()
        if self.accept("struct"):
            sid = self.integer(); self.expect("{")
            if self.accept("}"):
                if sid not in self.structs:
                    raise IrError(f"struct {sid} referenced before its definition")
                return self.structs[sid]
            self.expect("size"); self.expect(":")
            kind = self.next().text
            if kind not in ("fixed", "variable"):
                raise IrError("struct size must be fixed or variable")
            size = self.integer(); self.expect(",")
            align = self.alignment(); self.expect(",")
            members = []
            while not self.at("}"):
                off = self.integer(); self.expect(":"); mt = self.type(); self.expect(",")
                members.append((off, mt))
            self.expect("}")
            ty = Type("struct", size_kind=kind, size=size, align=align, members=tuple(members))
            if sid in self.structs and self.structs[sid] != ty:
                raise IrError(f"struct {sid} redefined")
            self.structs[sid] = ty
            return ty
        raise IrError(f"expected a type, found {t.text!r} at offset {t.pos}")

    def fn_type_tail(self) -> Type:
        self.expect("[")
        args = []
        while not self.at("]"):
            args.append(self.type())
            if not self.accept(","):
                break
        self.expect("]"); self.expect("->")
        if self.accept("!"):
            return Type("fn", args=tuple(args), results=None)
        self.expect("[")
        res = []
        while not self.at("]"):
            res.append(self.type())
            if not self.accept(","):
                break
        self.expect("]")
        return Type("fn", args=tuple(args), results=tuple(res))

    def alignment(self) -> int:
        self.expect("align"); self.expect(":")
        a = self.integer()
        if a == 0 or a & (a - 1):
            raise IrError("alignment must be a power of two")
        return a

    # constants (consts.rs:533-567)
    def const(self) -> Const:
        t = self.t
        if t.kind == "int":
            self.i += 1
            if not t.suffix:
                raise IrError(f"integer constant needs a type suffix at offset {t.pos}")
            bits = INT_TYPES[t.suffix]
            if t.value >> bits:
                raise IrError(f"constant {t.text} does not fit {t.suffix}")
            return Const("int", Type("int", t.suffix), t.value)
        if t.kind == "id" and t.text in FLOAT_TYPES:
            self.i += 1
            bits = self.integer()
            if bits >> FLOAT_TYPES[t.text]:
                raise IrError("float bit pattern too wide")
            return Const("float", Type("float", t.text), bits)
        if self.accept("true"):
            return Const("bool", BOOL, True)
        if self.accept("false"):
            return Const("bool", BOOL, False)
        if self.accept("<"):
            elems = [self.const()]
            while self.accept(","):
                elems.append(self.const())
            self.expect(">")
            if any(e.type != elems[0].type for e in elems):
                raise IrError("vector constant elements differ in type")
            return Const("vector", Type("vector", length=len(elems), element=elems[0].type), tuple(elems))
        if self.accept("undef"):
            return Const("undef", self.type())
        if self.accept("null"):
            ty = self.type()
            if ty.kind not in ("data_ptr", "fn"):
                raise IrError("null needs a pointer type")
            return Const("null", ty)
        if self.accept("fn"):
            key = self.name()
            if key not in self.functions:
                raise IrError(f"function {key[0]!r} is not defined")
            f = self.functions[key]
            return Const("fn", f.type, f)
        raise IrError(f"expected a constant, found {t.text!r} at offset {t.pos}")

    # values (values.rs:141-200)
    def value_def(self) -> Value:
        key = self.name(); self.expect(":")
        v = Value(key[0], self.type(), suffix_hint=key[1])
        self.define_value(key, v)
        return v

    def value_use(self) -> Value:
        key = self.name()
        if self.accept(":"):
            c = self.const()
            v = Value(key[0], c.type, const=c, suffix_hint=key[1])
            self.define_value(key, v, root=True)
            return v
        return self.lookup(self.values, key, "value")

    def uses(self, close="]") -> list:
        out = []
        while not self.at(close):
            out.append(self.value_use())
            if not self.accept(","):
                break
        self.expect(close)
        return out

    def defs(self) -> list:
        self.expect("[")
        out = []
        while not self.at("]"):
            out.append(self.value_def())
            if not self.accept(","):
                break
        self.expect("]")
        return out

    def results(self):
        if self.accept("!"):
            return None
        return self.defs()

    def variable(self) -> Variable:
        ty = self.type(); self.expect(","); al = self.alignment(); self.expect("->")
        p = self.value_def()
        if p.type != DATA_PTR:
            raise IrError("a variable's pointer must have type data_ptr")
        return Variable(ty, al, p)

    # instructions
    def location(self):
        if self.accept("@"):
            f = self.next()
            if f.kind != "str":
                raise IrError("expected a file name string")
            self.expect(":"); line = self.integer(); self.expect(":"); col = self.integer()
            return (f.value, line, col)
        return None

    def break_tail(self) -> Instr:
        key = self.name()
        blk = self.lookup(self.blocks, key, "block")
        self.expect("[")
        vals = self.uses()
        if blk.results is not None and len(vals) != len(blk.results):
            raise IrError(f"break {blk.name}: wrong number of results")
        return Instr("break", args=vals, target=blk)

    def instruction(self) -> Instr:
        t = self.next()
        if t.kind != "id":
            raise IrError(f"expected an instruction, found {t.text!r} at offset {t.pos}")
        if t.text == "add":
            self.expect("["); a = self.uses(); self.expect("->"); r = self.defs()
            if len(a) != 2 or len(r) != 1:
                raise IrError("add takes two arguments and one result")
            ins = Instr("add", args=a, results=r)
        elif t.text == "load":
            self.expect("["); a = self.uses(); self.expect(","); al = self.alignment(); self.expect("->"); r = self.defs()
            if len(a) != 1 or len(r) != 1:
                raise IrError("load takes one argument and one result")
            ins = Instr("load", args=a, results=r, align=al)
        elif t.text == "break":
            ins = self.break_tail()
        elif t.text == "continue":
            key = self.name()
            lp = self.lookup(self.loops, key, "loop")
            self.expect("[")
            ins = Instr("continue", args=self.uses(), target=lp)
        elif t.text == "branch":
            self.expect("["); a = self.uses(); self.expect(","); self.expect("{")
            cases = []
            while not self.at("}"):
                c = self.const(); self.expect("->"); self.expect("break")
                b = self.break_tail(); self.expect(";")
                cases.append((c, b.target, b.args))
            self.expect("}"); self.expect("->"); self.expect("["); self.expect("]")
            ins = Instr("branch", args=a, cases=cases)
        elif t.text == "block":
            ins = Instr("block", target=self.block_instruction())
        elif t.text == "loop":
            ins = Instr("loop", target=self.loop_instruction())
        elif self.version >= 2 and t.text in V2_OPS:
            self.expect("["); a = self.uses()
            al = 0
            if t.text == "store":
                self.expect(","); al = self.alignment()
            self.expect("->"); r = self.defs()
            if (len(a), len(r)) != V2_OPS[t.text]:
                raise IrError(f"{t.text} takes {V2_OPS[t.text][0]} argument(s) and {V2_OPS[t.text][1]} result(s)")
            ins = Instr(t.text, args=a, results=r, align=al)
        else:
            raise IrError(f"unknown instruction {t.text!r} at offset {t.pos}")
        ins.location = self.location()
        return ins

    def body(self, blk: Block):
        self.expect("{")
        while not self.at("}"):
            if blk.body and _diverges(blk.body[-1]):
                raise IrError("instructions after a diverging instruction")  # block.rs:291-300
            blk.body.append(self.instruction()); self.expect(";")
        self.expect("}")
        if not blk.body or not _diverges(blk.body[-1]):
            raise IrError(f"block {blk.name!r} must end in break/continue or a diverging block")

    def block_instruction(self) -> Block:
        key = self.name(); self.expect("->")
        blk = Block(key[0], None)
        blk.results = self.results()      # result definitions live in the enclosing scope
        self.push(); self.blocks[-1][key] = blk
        self.body(blk)
        self.pop()
        return blk

    def named_body(self, results) -> Block:
        key = self.name()
        blk = Block(key[0], results)
        self.push(); self.blocks[-1][key] = blk
        self.body(blk)
        self.pop()
        return blk

    def loop_instruction(self) -> Loop:
        key = self.name(); self.expect("[")
        args = self.uses(); self.expect("->")
        lp = Loop(key[0], args, [], None)
        lp.results = self.results()
        self.expect("{")
        self.push(); self.loops[-1][key] = lp
        self.expect("->"); lp.params = self.defs(); self.expect(";")
        if len(lp.params) != len(args):
            raise IrError("loop argument count differs from its parameter count")
        lp.body = self.named_body(lp.results)
        self.pop()
        self.expect("}")
        return lp

    # functions, module
    def function(self) -> Function:
        self.expect("fn")
        key = self.name()
        self.push()
        params = self.defs(); self.expect("->")
        if self.accept("!"):
            rets = None
        else:
            self.expect("["); rets = []
            while not self.at("]"):
                rets.append(self.type())
                if not self.accept(","):
                    break
            self.expect("]")
        f = Function(key[0], params, rets)
        self.functions[key] = f  # visible to its own body (function.rs golden: `fn function1` inside function1)
        self.expect("{"); self.expect("hints"); self.expect("{")
        while not self.at("}"):
            h = self.next().text; self.expect(":"); v = self.next().text; self.accept(",")
            if h == "inlining_hint" and v in ("none", "inline", "dont_inline"):
                f.inlining_hint = v
            elif h == "side_effects" and v in ("normal", "pure", "const"):
                f.side_effects = v
            else:
                raise IrError(f"bad hint {h}: {v}")
        self.expect("}"); self.expect("{")
        while not self.at("}"):
            f.locals.append(self.variable()); self.expect(";")
        self.expect("}")
        body_results = None if rets is None else [Value("", t) for t in rets]
        f.body = self.named_body(body_results)
        self.expect("}")
        self.pop()
        return f

    def target_properties(self) -> tuple[str, str]:
        self.expect("target_properties"); self.expect("{")
        dp = fp = "i64"
        while not self.at("}"):
            k = self.next().text; self.expect(":"); v = self.next().text
            if v not in INT_TYPES:
                raise IrError("pointer underlying type must be an integer type")
            if k == "data_pointer_underlying_type":
                dp = v
            elif k == "function_pointer_underlying_type":
                fp = v
            else:
                raise IrError(f"unknown target property {k}")
            if not self.accept(","):
                break
        self.expect("}")
        return dp, fp

    def interface_block(self) -> InterfaceBlock:
        self.expect("{"); self.expect("->"); p = self.value_def(); self.expect(";")
        self.expect("size"); self.expect(":"); kind = self.next().text; size = self.integer(); self.expect(";")
        al = self.alignment(); self.expect(";")
        members = []
        while not self.at("}"):
            off = self.integer(); self.expect(":"); ty = self.type(); self.expect(":"); self.expect("{"); self.expect("}")
            self.expect(";")
            members.append((off, ty))
        self.expect("}")
        return InterfaceBlock(p, kind, size, al, members)

    def interface_vars(self) -> list:
        self.expect("{")
        out = []
        while not self.at("}"):
            v = self.variable(); self.expect(":"); self.expect("{"); self.expect("kind"); self.expect(":")
            v.kind = self.next().text
            if v.kind != "vertex_position" and not (self.version >= 2 and v.kind in V2_BUILT_INS):  # interface.rs:61-67: the only built-in kind
                raise IrError(f"unknown built-in kind {v.kind}")
            self.expect("}"); self.expect(";")
            out.append(v)
        self.expect("}")
        return out

    def module(self) -> Module:
        m = Module()
        self.expect("module"); self.expect("{")
        m.data_ptr_type, m.fn_ptr_type = self.target_properties()
        self.expect("built_in_inputs"); m.built_in_inputs = self.interface_vars()
        self.expect("user_inputs_block"); m.user_inputs = self.interface_block()
        self.expect("built_in_outputs"); m.built_in_outputs = self.interface_vars()
        self.expect("user_outputs_block"); m.user_outputs = self.interface_block()
        self.expect("invocation_global_variables"); self.expect("{")
        while not self.at("}"):
            m.globals.append(self.variable()); self.expect(";")
        self.expect("}")
        while self.at("fn"):
            m.functions.append(self.function())
        self.expect("entry_point"); self.expect(":")
        key = self.name()
        if key not in self.functions:
            raise IrError(f"entry point {key[0]!r} is not defined")
        m.entry_point = self.functions[key]
        self.expect(";"); self.expect("}")
        return m


def _diverges(ins: Instr) -> bool:
    if ins.op in ("break", "continue"):
        return True
    if ins.op in ("block", "loop"):
        return ins.target.results is None
    return False


# ---- printer (text.rs:1544-1613) --------------------------------------------------------------------------------------

class Printer:
    def __init__(self):
        self.out: list[str] = []
        self.indent = 0
        self.names: dict[str, dict[str, list]] = {"value": {}, "block": {}, "loop": {}, "fn": {}}
        self.struct_ids: dict[Type, int] = {}
        self.printed_consts: set[int] = set()

    def w(self, s: str):
        self.out.append(s)

    def nl(self):
        self.out.append("\n" + "    " * self.indent)

    def name(self, kind: str, obj) -> str:
        lst = self.names[kind].setdefault(obj.name, [])
        for i, o in enumerate(lst):
            if o is obj:
                idx = i; break
        else:
            lst.append(obj); idx = len(lst) - 1
        n = obj.name
        if idx == 0 and re.fullmatch(r"[A-Za-z_][A-Za-z0-9_]*", n) and n not in KEYWORDS:
            return n
        return f'"{_escape(n)}"{idx}'

    def type(self, t: Type):
        if t.kind in ("int", "float"):
            self.w(t.name)
        elif t.kind == "bool":
            self.w("bool")
        elif t.kind == "data_ptr":
            self.w("data_ptr")
        elif t.kind == "vector":
            self.w("<vscale x " if t.scalable else "<"); self.w(f"{t.length} x "); self.type(t.element); self.w(">")
        elif t.kind == "fn":
            self.w("fn["); self.comma(t.args, self.type); self.w("] -> ")
            if t.results is None:
                self.w("!")
            else:
                self.w("["); self.comma(t.results, self.type); self.w("]")
        elif t.kind == "struct":
            if t in self.struct_ids:
                self.w(f"struct {self.struct_ids[t]} {{}}"); return
            sid = len(self.struct_ids); self.struct_ids[t] = sid
            self.w(f"struct {sid} {{"); self.indent += 1
            self.nl(); self.w(f"size: {t.size_kind} {t.size:#X},".replace("0X", "0x"))
            self.nl(); self.w(f"align: {t.align:#X},".replace("0X", "0x"))
            for off, mt in t.members:
                self.nl(); self.w(f"{off:#X}: ".replace("0X", "0x")); self.type(mt); self.w(",")
            self.indent -= 1; self.nl(); self.w("}")

    def comma(self, items, fn):
        for i, it in enumerate(items):
            if i:
                self.w(", ")
            fn(it)

    def const(self, c: Const):
        if c.kind == "int":
            self.w(f"{c.value:#X}".replace("0X", "0x") + c.type.name)
        elif c.kind == "float":
            self.w(c.type.name + " " + f"{c.value:#X}".replace("0X", "0x"))
        elif c.kind == "bool":
            self.w("true" if c.value else "false")
        elif c.kind == "vector":
            self.w("<"); self.comma(c.value, self.const); self.w(">")
        elif c.kind in ("undef", "null"):
            self.w(c.kind + " "); self.type(c.type)
        elif c.kind == "fn":
            self.w("fn " + self.name("fn", c.value))

    def value_def(self, v: Value):
        self.w(self.name("value", v) + ": "); self.type(v.type)

    def value_use(self, v: Value):
        self.w(self.name("value", v))
        if v.const is not None and id(v) not in self.printed_consts:
            self.printed_consts.add(id(v))
            self.w(": "); self.const(v.const)

    def results(self, r):
        if r is None:
            self.w("!")
        else:
            self.w("["); self.comma(r, self.value_def); self.w("]")

    def variable(self, v: Variable):
        self.type(v.type); self.w(f", align: {v.align:#X} -> ".replace("0X", "0x")); self.value_def(v.pointer)

    def break_(self, blk: Block, vals):
        self.w("break " + self.name("block", blk) + "["); self.comma(vals, self.value_use); self.w("]")

    def instr(self, ins: Instr):
        if ins.op == "add":
            self.w("add ["); self.comma(ins.args, self.value_use); self.w("] -> "); self.results(ins.results)
        elif ins.op == "load":
            self.w("load ["); self.comma(ins.args, self.value_use)
            self.w(f"], align: {ins.align:#X} -> ".replace("0X", "0x")); self.results(ins.results)
        elif ins.op == "break":
            self.break_(ins.target, ins.args)
        elif ins.op == "continue":
            self.w("continue " + self.name("loop", ins.target) + "["); self.comma(ins.args, self.value_use); self.w("]")
        elif ins.op == "branch":
            self.w("branch ["); self.comma(ins.args, self.value_use); self.w("], {"); self.indent += 1
            for c, blk, vals in ins.cases:
                self.nl(); self.const(c); self.w(" -> "); self.break_(blk, vals); self.w(";")
            self.indent -= 1; self.nl(); self.w("} -> []")
        elif ins.op == "block":
            b = ins.target
            self.w("block " + self.name("block", b) + " -> "); self.results(b.results); self.w(" "); self.body(b)
        elif ins.op in V2_OPS:
            self.w(ins.op + " ["); self.comma(ins.args, self.value_use); self.w("]")
            if ins.op == "store":
                self.w(f", align: {ins.align:#X}".replace("0X", "0x"))
            self.w(" -> "); self.results(ins.results)
        elif ins.op == "loop":
            lp = ins.target
            self.w("loop " + self.name("loop", lp) + "["); self.comma(lp.arguments, self.value_use); self.w("] -> ")
            self.results(lp.results); self.w(" {"); self.indent += 1
            self.nl(); self.w("-> ["); self.comma(lp.params, self.value_def); self.w("];")
            self.nl(); self.w(self.name("block", lp.body) + " "); self.body(lp.body)
            self.indent -= 1; self.nl(); self.w("}")
        if ins.location:
            f, line, col = ins.location
            self.w(f' @ "{_escape(f)}":{line}:{col}')

    def body(self, b: Block):
        self.w("{"); self.indent += 1
        for ins in b.body:
            self.nl(); self.instr(ins); self.w(";")
        self.indent -= 1; self.nl(); self.w("}")

    def function(self, f: Function):
        self.w("fn " + self.name("fn", f) + "["); self.comma(f.params, self.value_def); self.w("] -> ")
        if f.returns is None:
            self.w("!")
        else:
            self.w("["); self.comma(f.returns, self.type); self.w("]")
        self.w(" {"); self.indent += 1
        self.nl(); self.w("hints {"); self.indent += 1
        self.nl(); self.w(f"inlining_hint: {f.inlining_hint},")
        self.nl(); self.w(f"side_effects: {f.side_effects},")
        self.indent -= 1; self.nl(); self.w("}")
        self.nl(); self.w("{"); self.indent += 1
        for v in f.locals:
            self.nl(); self.variable(v); self.w(";")
        self.indent -= 1; self.nl(); self.w("}")
        self.nl(); self.w(self.name("block", f.body) + " "); self.body(f.body)
        self.indent -= 1; self.nl(); self.w("}")

    def target_properties(self, dp: str, fp: str):
        self.w("target_properties {"); self.indent += 1
        self.nl(); self.w(f"data_pointer_underlying_type: {dp},")
        self.nl(); self.w(f"function_pointer_underlying_type: {fp},")
        self.indent -= 1; self.nl(); self.w("}")

    def interface_block(self, title: str, b: InterfaceBlock):
        self.nl(); self.w(title + " {"); self.indent += 1
        self.nl(); self.w("-> "); self.value_def(b.pointer); self.w(";")
        self.nl(); self.w(f"size: {b.size_kind} {b.size:#X};".replace("0X", "0x"))
        self.nl(); self.w(f"align: {b.align:#X};".replace("0X", "0x"))
        for off, ty in b.members:
            self.nl(); self.w(f"{off:#X}: ".replace("0X", "0x")); self.type(ty); self.w(": {};")
        self.indent -= 1; self.nl(); self.w("}")

    def interface_vars(self, title: str, vs: list):
        self.nl(); self.w(title + " {"); self.indent += 1
        for v in vs:
            self.nl(); self.variable(v); self.w(f": {{ kind: {v.kind} }};")
        self.indent -= 1; self.nl(); self.w("}")

    def module(self, m: Module):
        self.w("module {"); self.indent += 1
        self.nl(); self.target_properties(m.data_ptr_type, m.fn_ptr_type)
        self.interface_vars("built_in_inputs", m.built_in_inputs)
        self.interface_block("user_inputs_block", m.user_inputs)
        self.interface_vars("built_in_outputs", m.built_in_outputs)
        self.interface_block("user_outputs_block", m.user_outputs)
        self.nl(); self.w("invocation_global_variables {"); self.indent += 1
        for v in m.globals:
            self.nl(); self.variable(v); self.w(";")
        self.indent -= 1; self.nl(); self.w("}")
        for f in m.functions:
            self.nl(); self.function(f)
        self.nl(); self.w("entry_point: " + self.name("fn", m.entry_point) + ";")
        self.indent -= 1; self.nl(); self.w("}")

    def text(self) -> str:
        return "".join(self.out)


# ---- entry points used by the tests --------------------------------------------------------------------------------------

def parse_module(text: str, version: int = 1) -> Module:
    p = Parser(text, version); m = p.module(); p.finish(); return m


def module_to_text(m: Module) -> str:
    pr = Printer(); pr.module(m); return pr.text() + "\n"


def round_trip(kind: str, text: str) -> str:
    """parse `text` as `kind` and print it canonically (what the reference's test_from_to_text macros do)."""
    p, pr = Parser(text), Printer()
    if kind == "type":
        v = p.type(); p.finish(); pr.type(v)
    elif kind == "const":
        v = p.const(); p.finish(); pr.const(v)
    elif kind == "target_properties":
        v = p.target_properties(); p.finish(); pr.target_properties(*v)
    elif kind == "block":
        p.expect("block"); v = p.block_instruction(); p.finish(); pr.instr(Instr("block", target=v))
    elif kind == "function":
        v = p.function(); p.finish(); pr.function(v)
    elif kind == "definitions_then_instruction":
        defs = p.defs(); ins = p.instruction(); p.finish()
        pr.w("["); pr.comma(defs, pr.value_def); pr.w("]\n"); pr.instr(ins)
    elif kind == "module":
        return module_to_text(parse_module(text))
    else:
        raise IrError(f"unknown kind {kind}")
    return pr.text()


# ---- interpreter ----------------------------------------------------------------------------------------------------------------

class _Break(Exception):
    def __init__(self, block, values):
        self.block, self.values = block, values


class _Continue(Exception):
    def __init__(self, loop, values):
        self.loop, self.values = loop, values


def type_size(t: Type) -> int:
    if t.kind == "int":
        return INT_TYPES[t.name] // 8
    if t.kind == "float":
        return FLOAT_TYPES[t.name] // 8
    if t.kind in ("data_ptr", "fn"):
        return 8
    if t.kind == "bool":
        return 1
    raise IrError(f"type {t.kind} has no scalar size")


class Interpreter:
    """Executes the entry point for ONE invocation.  Memory is a dict address -> bytes-like; `data_ptr` values are ints."""

    def __init__(self, module: Module, memory: dict[int, bytes], step_limit: int = 10_000_000):
        self.m, self.memory, self.steps, self.limit = module, memory, 0, step_limit

    def read(self, addr: int, size: int) -> int:
        for base, buf in self.memory.items():
            if base <= addr and addr + size <= base + len(buf):
                return int.from_bytes(bytes(buf[addr - base:addr - base + size]), "little")
        raise IrError(f"load from unmapped address {addr:#x}")

    def val(self, env: dict, v: Value):
        if v.const is not None:
            c = v.const
            if c.kind in ("int", "float"):
                return c.value
            if c.kind == "bool":
                return int(c.value)
            if c.kind == "null":
                return 0
            raise IrError(f"constant kind {c.kind} cannot be evaluated")
        return env[id(v)]

    def run_body(self, env: dict, blk: Block):
        for ins in blk.body:
            self.steps += 1
            if self.steps > self.limit:
                raise IrError("step limit exceeded")
            if ins.op == "add":
                a, b = (self.val(env, x) for x in ins.args)
                t = ins.results[0].type
                if t.kind != "int":
                    raise IrError("add is defined on integer types only")
                env[id(ins.results[0])] = (a + b) & ((1 << INT_TYPES[t.name]) - 1)
            elif ins.op == "load":
                t = ins.results[0].type
                env[id(ins.results[0])] = self.read(self.val(env, ins.args[0]), type_size(t))
            elif ins.op == "break":
                raise _Break(ins.target, [self.val(env, x) for x in ins.args])
            elif ins.op == "continue":
                raise _Continue(ins.target, [self.val(env, x) for x in ins.args])
            elif ins.op == "branch":
                x = self.val(env, ins.args[0])
                for c, blk2, vals in ins.cases:
                    if c.value == x:
                        raise _Break(blk2, [self.val(env, u) for u in vals])
            elif ins.op == "block":
                self.run_block(env, ins.target)
            elif ins.op == "loop":
                self.run_loop(env, ins.target)

    def run_block(self, env: dict, blk: Block):
        try:
            self.run_body(env, blk)
        except _Break as b:
            if b.block is not blk:
                raise
            for d, v in zip(blk.results or [], b.values):
                env[id(d)] = v
            return b.values
        raise IrError(f"block {blk.name!r} fell off its end")

    def run_loop(self, env: dict, lp: Loop):
        vals = [self.val(env, a) for a in lp.arguments]
        while True:
            for d, v in zip(lp.params, vals):
                env[id(d)] = v
            try:
                self.run_body(env, lp.body)
            except _Continue as c:
                if c.loop is not lp:
                    raise
                vals = c.values
                continue
            except _Break as b:
                if b.block is not lp.body:
                    raise
                for d, v in zip(lp.results or [], b.values):
                    env[id(d)] = v
                return
            raise IrError(f"loop body {lp.body.name!r} fell off its end")

    def run_entry(self, user_inputs: int, user_outputs: int) -> list[int]:
        env = {id(self.m.user_inputs.pointer): user_inputs, id(self.m.user_outputs.pointer): user_outputs}
        f = self.m.entry_point
        return self.run_block(env, f.body)
