"""Random .kazan-ir programs for the differential dispatch test: integer programs built from the instructions the IR has
(add, load, block, loop, branch, break, continue — instructions_impl.rs:395-433), with nested blocks, multi-way branches that
break out of different enclosing blocks, bounded loops, and every integer width.  The oracle interpreter
(oracle/ir_oracle.py) and the PTX back end must agree on every output of every invocation."""
from __future__ import annotations

import numpy as np

TYPES = ["i8", "i16", "i32", "i64"]
BITS = {"i8": 8, "i16": 16, "i32": 32, "i64": 64}
NP = {"i8": "u1", "i16": "<u2", "i32": "<u4", "i64": "<u8"}


class Gen:
    def __init__(self, seed: int):
        self.rng = np.random.default_rng(seed)
        self.n_const = 0
        self.n_val = 0
        self.lines: list[str] = []
        self.indent = 3

    def const(self, ty: str, value: int | None = None) -> str:
        if value is None:
            value = int(self.rng.integers(0, 1 << min(BITS[ty], 62))) if self.rng.random() < 0.7 else (1 << BITS[ty]) - 1
        k = self.n_const
        self.n_const += 1
        return f'""{k}: 0x{value & ((1 << BITS[ty]) - 1):X}{ty}'

    def name(self, stem="v") -> str:
        self.n_val += 1
        return f"{stem}{self.n_val}"

    def emit(self, text: str):
        self.lines.append("    " * self.indent + text)

    def pick(self, pool: dict, ty: str) -> str:
        have = pool.get(ty, [])
        if have and self.rng.random() < 0.8:
            return have[int(self.rng.integers(0, len(have)))]
        return self.const(ty)

    def add(self, pool: dict, ty: str) -> str:
        out = self.name()
        self.emit(f"add [{self.pick(pool, ty)}, {self.pick(pool, ty)}] -> [{out}: {ty}];")
        pool.setdefault(ty, []).append(out)
        return out

    def straight(self, pool: dict, n: int):
        for _ in range(n):
            self.add(pool, TYPES[int(self.rng.integers(0, 4))])

    def block(self, pool: dict, depth: int, outer: list):
        """block with a multi-way branch: arms break this block or an enclosing one (with that block's result types)."""
        ty = TYPES[int(self.rng.integers(0, 4))]
        label = self.name("blk")
        res = self.name("r")
        sel_ty = TYPES[int(self.rng.integers(0, 4))]
        sel = self.pick(pool, sel_ty)
        if sel.startswith('""'):  # branch on a value, not on a constant definition
            sel = self.add(pool, sel_ty)
        self.emit(f"block {label} -> [{res}: {ty}] {{")
        self.indent += 1
        inner = dict((k, list(v)) for k, v in pool.items())
        self.straight(inner, int(self.rng.integers(0, 3)))
        self.emit(f"branch [{sel}], {{")
        self.indent += 1
        used = set()
        for _ in range(int(self.rng.integers(1, 4))):
            case = int(self.rng.integers(0, 4))
            if case in used:
                continue
            used.add(case)
            targets = [(label, [ty])] + outer
            tgt, tys = targets[int(self.rng.integers(0, len(targets)))] if self.rng.random() < 0.3 else targets[0]
            args = ", ".join(self.pick(inner, t) for t in tys)
            self.emit(f"0x{case:X}{sel_ty} -> break {tgt}[{args}];")
        self.indent -= 1
        self.emit("} -> [];")
        if depth > 0 and self.rng.random() < 0.5:
            self.block(inner, depth - 1, [(label, [ty])] + outer)
        self.straight(inner, int(self.rng.integers(0, 3)))
        self.emit(f"break {label}[{self.pick(inner, ty)}];")
        self.indent -= 1
        self.emit("};")
        pool.setdefault(ty, []).append(res)

    def loop(self, pool: dict):
        """bounded loop: an accumulator of a random type and an i32 counter that breaks the body block at a constant"""
        ty = TYPES[int(self.rng.integers(0, 4))]
        label, body, res = self.name("lp"), self.name("step"), self.name("total")
        acc, cnt = self.name("acc"), self.name("n")
        trips = int(self.rng.integers(1, 40))
        self.emit(f"loop {label}[{self.pick(pool, ty)}, {self.const('i32', 0)}] -> [{res}: {ty}] {{")
        self.indent += 1
        self.emit(f"-> [{acc}: {ty}, {cnt}: i32];")
        self.emit(f"{body} {{")
        self.indent += 1
        inner = dict((k, list(v)) for k, v in pool.items())
        inner.setdefault(ty, []).append(acc)
        a2 = self.name("acc")
        self.emit(f"add [{acc}, {self.pick(inner, ty)}] -> [{a2}: {ty}];")
        n2 = self.name("n")
        self.emit(f"add [{cnt}, {self.const('i32', 1)}] -> [{n2}: i32];")
        self.emit(f"branch [{n2}], {{")
        self.emit(f"    0x{trips:X}i32 -> break {body}[{a2}];")
        self.emit("} -> [];")
        self.emit(f"continue {label}[{a2}, {n2}];")
        self.indent -= 1
        self.emit("}")
        self.indent -= 1
        self.emit("};")
        pool.setdefault(ty, []).append(res)


def make_program(seed: int):
    """-> (module text, numpy dtype of an input slot, numpy dtype of an output slot)"""
    g = Gen(seed)
    pool: dict = {}
    for ty, align in (("i8", 1), ("i16", 2), ("i32", 4), ("i64", 8), ("i32", 1), ("i16", 1)):
        v = g.name("in")
        g.emit(f"load [user_inputs_block], align: 0x{align:X} -> [{v}: {ty}];")
        pool.setdefault(ty, []).append(v)
    for _ in range(int(g.rng.integers(2, 6))):
        what = g.rng.random()
        if what < 0.4:
            g.straight(pool, int(g.rng.integers(1, 5)))
        elif what < 0.8:
            g.block(pool, 2, [])
        else:
            g.loop(pool)
    out_tys = [TYPES[int(g.rng.integers(0, 4))] for _ in range(int(g.rng.integers(1, 5)))]
    out_tys.sort(key=lambda t: -BITS[t])  # natural alignment without padding
    outs = [g.pick(pool, t) for t in out_tys]
    members, off = [], 0
    for t in out_tys:
        members.append((off, t))
        off += BITS[t] // 8
    size = max((off + 7) // 8 * 8, 8)
    fields = [(f"o{i}", NP[t]) for i, (_, t) in enumerate(members)]
    if size > off:
        fields.append(("pad", f"V{size - off}"))
    body = "\n".join(g.lines)
    mem_lines = "\n".join(f"        0x{o:X}: {t}: {{}};" for o, t in members)
    text = f"""module {{
    target_properties {{
        data_pointer_underlying_type: i64,
        function_pointer_underlying_type: i64,
    }}
    built_in_inputs {{
    }}
    user_inputs_block {{
        -> user_inputs_block: data_ptr;
        size: fixed 0x8;
        align: 0x8;
        0x0: i64: {{}};
    }}
    built_in_outputs {{
    }}
    user_outputs_block {{
        -> user_outputs_block: data_ptr;
        size: fixed 0x{size:X};
        align: 0x8;
{mem_lines}
    }}
    invocation_global_variables {{
    }}
    fn main[] -> [{", ".join(out_tys)}] {{
        hints {{
            inlining_hint: none,
            side_effects: normal,
        }}
        {{
        }}
        body {{
{body}
            break body[{", ".join(outs)}];
        }}
    }}
    entry_point: main;
}}
"""
    return text, np.dtype([("x", "<u8")]), np.dtype(fields)
