"""oracle/ir_simt.py — SIMT interpreter for `.kazan-ir` modules of version 1 and 2.  TEST INFRASTRUCTURE.

Executes the entry point for the invocations of a dispatch (or the fragments of a draw) 32 at a time, in lock step, the way
the PTX back end lays them out: a subgroup is 32 consecutive invocations, lanes that leave a block by different `break`s
meet again where their targets end (structured control flow: block.rs:16-124, 492-503).  This is the checker for

    kazan_b200/csrc/ptx_emit.cpp   the instructions of version 2 (kazan_b200/csrc/ir.h `kExtOps`), the built-in kinds
                                   (global_invocation_id, subgroup_local_invocation_id, subgroup_size, frag_coord,
                                   push_constants, descriptor_table) and the subgroup operations

PARITY UNPINNED: the reference has none of these instructions (shader-compiler-ir/src/instructions_impl.rs:395-433 stops at
add / load / break / branch / continue / block / loop) and never executes IR; they are spelled like the SPIR-V instructions
they translate and mean what the SPIR-V specification says for scalars, with the implementation-defined corners fixed as
written in ir.h.  Version 1 programs run here exactly as in ir_oracle.Interpreter (tests/test_ir_v2.py checks that).

Values are bit patterns (Python ints); floating point goes through numpy scalars (IEEE, round to nearest even, denormals
kept) and libm's correctly rounded fmaf / fma.  Memory is a dict base address -> bytearray.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import struct

import numpy as np

from . import ir_oracle as iro
from .ir_oracle import FLOAT_TYPES, INT_TYPES, IrError, V2_OPS

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.fmaf.restype, _libm.fmaf.argtypes = ctypes.c_float, [ctypes.c_float] * 3
_libm.fma.restype, _libm.fma.argtypes = ctypes.c_double, [ctypes.c_double] * 3

SUBGROUP = 32
LOCAL_BASE = 0x7000_0000_0000  # private address space of the invocations' variables
LANE_STRIDE = 1 << 20


def _f32(bits: int) -> np.float32:
    return np.frombuffer(struct.pack("<I", bits & 0xFFFFFFFF), dtype=np.float32)[0]


def _f64(bits: int) -> np.float64:
    return np.frombuffer(struct.pack("<Q", bits & 0xFFFFFFFFFFFFFFFF), dtype=np.float64)[0]


def _bits32(f) -> int:
    return struct.unpack("<I", np.float32(f).tobytes())[0]


def _bits64(f) -> int:
    return struct.unpack("<Q", np.float64(f).tobytes())[0]


def _signed(v: int, w: int) -> int:
    return v - (1 << w) if v >> (w - 1) & 1 else v


def _round_int_to_float(n: int, mant: int, f_of, bits_of) -> int:
    """n (any size) -> nearest float with `mant` significand bits, ties to even (cvt.rn.f32.s64 etc.)."""
    if n == 0:
        return 0
    neg, a = n < 0, abs(n)
    length = a.bit_length()
    if length > mant:
        shift = length - mant
        q, r = a >> shift, a & ((1 << shift) - 1)
        half = 1 << (shift - 1)
        if r > half or (r == half and q & 1):
            q += 1
        a = q << shift
    return bits_of(f_of(-a if neg else a))  # exact now: at most `mant` significant bits


def _is_nan32(b):
    return (b & 0x7F800000) == 0x7F800000 and (b & 0x7FFFFF) != 0


def _is_nan64(b):
    return (b & 0x7FF0000000000000) == 0x7FF0000000000000 and (b & 0xFFFFFFFFFFFFF) != 0


class Memory:
    """address -> bytes; regions are bytearrays (or numpy uint8 arrays) keyed by their base address"""

    def __init__(self, regions: dict | None = None):
        self.regions = dict(regions or {})

    def _find(self, addr: int, size: int):
        for base, buf in self.regions.items():
            if base <= addr and addr + size <= base + len(buf):
                return base, buf
        raise IrError(f"access to unmapped address {addr:#x} ({size} bytes)")

    def read(self, addr: int, size: int) -> int:
        base, buf = self._find(addr, size)
        return int.from_bytes(bytes(buf[addr - base:addr - base + size]), "little")

    def write(self, addr: int, size: int, value: int):
        base, buf = self._find(addr, size)
        buf[addr - base:addr - base + size] = (value & ((1 << (8 * size)) - 1)).to_bytes(size, "little")


def _type_bytes(t) -> int:
    if t.kind == "vector":
        return _type_bytes(t.element) * t.length
    if t.kind == "struct":
        return t.size
    return iro.type_size(t)


class Simt:
    def __init__(self, module, memory: Memory, step_limit: int = 50_000_000):
        self.m, self.mem, self.steps, self.limit = module, memory, 0, step_limit
        self.env: dict[int, dict] = {}
        self.uniform_lanes: list[int] = []

    # ---- values ----
    def val(self, lane: int, v) -> int:
        if v.const is not None:
            c = v.const
            if c.kind in ("int", "float"):
                return c.value
            if c.kind == "bool":
                return int(c.value)
            if c.kind in ("null", "undef"):
                return 0
            raise IrError(f"constant kind {c.kind} cannot be evaluated")
        return self.env[lane][id(v)]

    def put(self, lane: int, v, x: int):
        self.env[lane][id(v)] = x

    # ---- one non-control instruction on one lane ----
    def simple(self, lane: int, ins):
        op = ins.op
        a = [self.val(lane, x) for x in ins.args]
        r = ins.results[0] if ins.results else None
        if op == "add":
            if r.type.kind != "int":
                raise IrError("add is defined on integer types only")
            return self.put(lane, r, (a[0] + a[1]) & ((1 << INT_TYPES[r.type.name]) - 1))
        if op == "load":
            return self.put(lane, r, self.mem.read(a[0], iro.type_size(r.type)))
        if op == "store":
            return self.mem.write(a[0], iro.type_size(ins.args[1].type), a[1])
        if op == "ptr_add":
            return self.put(lane, r, (a[0] + _signed(a[1], 64)) & ((1 << 64) - 1))
        t0 = ins.args[0].type if ins.args else None
        if op in ("i_sub", "i_mul", "bitwise_and", "bitwise_or", "bitwise_xor", "shift_left_logical", "shift_right_logical",
                  "shift_right_arithmetic", "s_negate", "not"):
            w = INT_TYPES[r.type.name]
            mask = (1 << w) - 1
            x = a[0]
            y = a[1] if len(a) > 1 else 0
            if op == "i_sub":
                v = x - y
            elif op == "i_mul":
                v = x * y
            elif op == "bitwise_and":
                v = x & y
            elif op == "bitwise_or":
                v = x | y
            elif op == "bitwise_xor":
                v = x ^ y
            elif op == "shift_left_logical":
                v = 0 if y >= w else x << y
            elif op == "shift_right_logical":
                v = 0 if y >= w else x >> y
            elif op == "shift_right_arithmetic":
                v = _signed(x, w) >> min(y, w - 1)
            elif op == "s_negate":
                v = -x
            else:
                v = ~x
            return self.put(lane, r, v & mask)
        if op in ("f_add", "f_sub", "f_mul", "f_div", "f_min", "f_max", "f_negate", "f_abs", "fma"):
            return self.put(lane, r, self.float_op(op, r.type.name, a))
        if op in ("i_equal", "i_not_equal", "u_less_than", "s_less_than", "u_less_than_equal", "s_less_than_equal", "u_greater_than",
                  "s_greater_than", "u_greater_than_equal", "s_greater_than_equal"):
            w = 64 if t0.kind == "data_ptr" else INT_TYPES[t0.name]
            x, y = (a[0], a[1]) if op[0] in "iu" else (_signed(a[0], w), _signed(a[1], w))
            res = {"equal": x == y, "not_equal": x != y, "less_than": x < y, "less_than_equal": x <= y, "greater_than": x > y,
                   "greater_than_equal": x >= y}[op[2:]]
            return self.put(lane, r, int(res))
        if op.startswith("f_ord_") or op == "f_unord_not_equal":
            f = _f32 if FLOAT_TYPES[t0.name] == 32 else _f64
            x, y = f(a[0]), f(a[1])
            with np.errstate(all="ignore"):
                if op == "f_unord_not_equal":
                    res = bool(x != y)  # true when either is NaN
                else:
                    res = bool({"equal": x == y, "less_than": x < y, "less_than_equal": x <= y, "greater_than": x > y,
                                "greater_than_equal": x >= y}[op[6:]])
            return self.put(lane, r, int(res))
        if op == "logical_and":
            return self.put(lane, r, a[0] & a[1] & 1)
        if op == "logical_or":
            return self.put(lane, r, (a[0] | a[1]) & 1)
        if op == "logical_not":
            return self.put(lane, r, (a[0] ^ 1) & 1)
        if op == "select":
            return self.put(lane, r, a[1] if a[0] else a[2])
        if op in ("convert_f_to_u", "convert_f_to_s"):
            w = INT_TYPES[r.type.name]
            f = float(_f32(a[0])) if FLOAT_TYPES[t0.name] == 32 else float(_f64(a[0]))
            lo, hi = (0, (1 << w) - 1) if op.endswith("u") else (-(1 << (w - 1)), (1 << (w - 1)) - 1)
            if f != f:
                v = 0
            elif f == float("inf"):
                v = hi
            elif f == float("-inf"):
                v = lo
            else:
                v = min(max(int(f), lo), hi)  # int() truncates toward zero
            return self.put(lane, r, v & ((1 << w) - 1))
        if op in ("convert_s_to_f", "convert_u_to_f"):
            w = INT_TYPES[t0.name]
            n = _signed(a[0], w) if op == "convert_s_to_f" else a[0]
            if FLOAT_TYPES[r.type.name] == 32:
                return self.put(lane, r, _round_int_to_float(n, 24, np.float32, _bits32))
            return self.put(lane, r, _round_int_to_float(n, 53, np.float64, _bits64))
        if op in ("u_convert", "s_convert"):
            wa, wr = INT_TYPES[t0.name], INT_TYPES[r.type.name]
            n = _signed(a[0], wa) if op == "s_convert" else a[0]
            return self.put(lane, r, n & ((1 << wr) - 1))
        if op == "f_convert":
            src, dst = FLOAT_TYPES[t0.name], FLOAT_TYPES[r.type.name]
            if src == dst:
                return self.put(lane, r, a[0])
            with np.errstate(all="ignore"):
                if dst == 64:
                    v = _bits64(np.float64(_f32(a[0])))
                else:
                    v = _bits32(np.float32(_f64(a[0])))
            return self.put(lane, r, v)
        if op == "bitcast":
            return self.put(lane, r, a[0])
        raise IrError(f"instruction {op} cannot be executed")

    @staticmethod
    def float_op(op: str, tname: str, a: list) -> int:
        w = FLOAT_TYPES[tname]
        f, bits, isnan = (_f32, _bits32, _is_nan32) if w == 32 else (_f64, _bits64, _is_nan64)
        sign = 1 << (w - 1)
        if op == "f_negate":
            return a[0] ^ sign
        if op == "f_abs":
            return a[0] & (sign - 1)
        x = f(a[0])
        y = f(a[1]) if len(a) > 1 else None
        with np.errstate(all="ignore"):
            if op == "f_add":
                return bits(x + y)
            if op == "f_sub":
                return bits(x - y)
            if op == "f_mul":
                return bits(x * y)
            if op == "f_div":
                return bits(x / y)
            if op == "fma":
                z = f(a[2])
                if w == 32:
                    return _bits32(_libm.fmaf(float(x), float(y), float(z)))
                return _bits64(_libm.fma(float(x), float(y), float(z)))
            # f_min / f_max: the other operand when one is NaN; -0 < +0
            if isnan(a[0]):
                return a[1]
            if isnan(a[1]):
                return a[0]
            if x == y:  # equal values: +-0 are told apart by the sign bit
                neg0 = a[0] if a[0] & sign else a[1]
                pos0 = a[1] if a[0] & sign else a[0]
                return neg0 if op == "f_min" else pos0
            return a[0] if (x < y) == (op == "f_min") else a[1]

    # ---- subgroup operations (kExtOps: group_non_uniform_*) ----
    def subgroup(self, ins, lanes: list):
        if lanes != self.uniform_lanes:
            raise IrError(f"{ins.op} executed in non-uniform control flow (defined in subgroup-uniform control flow only)")
        op = ins.op[len("group_non_uniform_"):]
        r = ins.results[0]
        first = lanes[0]
        if op == "elect":
            for l in lanes:
                self.put(l, r, int(l == first))
            return
        vals = {l: self.val(l, ins.args[0]) for l in lanes}
        if op == "all":
            v = int(all(vals.values()))
        elif op == "any":
            v = int(any(vals.values()))
        elif op == "ballot":
            v = sum(1 << (l % SUBGROUP) for l in lanes if vals[l])
        elif op == "broadcast_first":
            v = vals[first]
        elif op == "i_add":
            v = sum(vals.values()) & 0xFFFFFFFF
        elif op in ("shuffle", "shuffle_xor"):
            base = first - first % SUBGROUP
            for l in lanes:
                sel = self.val(l, ins.args[1])
                src = (sel if op == "shuffle" else (l % SUBGROUP) ^ sel) & 31
                if base + src not in vals:
                    raise IrError(f"{ins.op}: lane {src} is not an active invocation (the result is undefined)")
                self.put(l, r, vals[base + src])
            return
        else:
            raise IrError(f"unknown subgroup operation {ins.op}")
        for l in lanes:
            self.put(l, r, v)

    # ---- structured control flow over a set of lanes ----
    def exec_block(self, blk, lanes: list) -> list:
        """Runs `blk` for `lanes` (ascending).  Returns the exits that leave a construct outside blk: [(target, lane, values)];
        lanes that `break` to blk itself have its results bound and simply continue after the block."""
        pending, active = [], list(lanes)
        for ins in blk.body:
            if not active:
                break
            self.steps += len(active)
            if self.steps > self.limit:
                raise IrError("step limit exceeded")
            op = ins.op
            if op in ("break", "continue"):
                pending += [(ins.target, l, [self.val(l, x) for x in ins.args]) for l in active]
                active = []
            elif op == "branch":
                still = []
                for l in active:
                    x = self.val(l, ins.args[0])
                    for c, target, vals in ins.cases:
                        if int(c.value) == x:
                            pending.append((target, l, [self.val(l, u) for u in vals]))
                            break
                    else:
                        still.append(l)
                active = still
            elif op == "block":
                out = self.exec_block(ins.target, active)
                gone = {l for _, l, _ in out}
                pending += out
                active = [l for l in active if l not in gone]
            elif op == "loop":
                out = self.exec_loop(ins.target, active)
                gone = {l for _, l, _ in out}
                pending += out
                active = [l for l in active if l not in gone]
            elif op.startswith("group_non_uniform_"):
                self.subgroup(ins, active)
            else:
                for l in active:
                    self.simple(l, ins)
        if active:
            raise IrError(f"block {blk.name!r} fell off its end")
        outer = []
        for target, l, vals in pending:
            if target is blk:
                for d, v in zip(blk.results or [], vals):
                    self.put(l, d, v)
            else:
                outer.append((target, l, vals))
        return outer

    def exec_loop(self, lp, lanes: list) -> list:
        vals = {l: [self.val(l, a) for a in lp.arguments] for l in lanes}
        current, outer = list(lanes), []
        while current:
            for l in current:
                for d, v in zip(lp.params, vals[l]):
                    self.put(l, d, v)
            out = self.exec_block(lp.body, current)  # lanes that break to the body block have left the loop
            again = []
            for target, l, v in out:
                if target is lp:
                    vals[l] = v
                    again.append(l)
                else:
                    outer.append((target, l, v))
            current = sorted(again)
        return outer

    # ---- entry points ----
    def _variables(self, lane_slot: int, built_ins: dict):
        """allocate the invocation's variables in its private memory; returns {id(pointer value): address}"""
        self.var_address = getattr(self, "var_address", {})
        m, f = self.m, self.m.entry_point
        base = LOCAL_BASE + lane_slot * LANE_STRIDE
        region = bytearray(LANE_STRIDE)
        self.mem.regions[base] = region
        ptrs, at = {}, 0
        for v in list(m.built_in_inputs) + list(m.built_in_outputs) + list(m.globals) + list(f.locals):
            al = max(v.align, 1)
            at = (at + al - 1) // al * al
            ptrs[id(v.pointer)] = base + at
            self.var_address[(lane_slot, id(v.pointer))] = base + at
            if v.kind in built_ins:
                data = built_ins[v.kind]
                region[at:at + len(data)] = data
            at += max(_type_bytes(v.type), 1)
        if at > LANE_STRIDE:
            raise IrError("variables of one invocation do not fit the interpreter's private memory")
        return ptrs

    def run_group(self, lanes: list, lane_env: dict) -> dict:
        """lanes: global invocation indices of ONE subgroup (ascending); lane_env[lane] = initial {id(value): bits}.
        Returns {lane: [values the entry point's body was left with]}"""
        self.env = {l: dict(lane_env[l]) for l in lanes}
        self.uniform_lanes = list(lanes)
        body = self.m.entry_point.body
        out = self.exec_block(body, lanes)
        if out:
            raise IrError("a break or continue left the entry point's body")
        return {l: [self.env[l][id(d)] for d in (body.results or [])] for l in lanes}


def output_slot_size(module) -> int:
    """bytes of one invocation's output slot: user_outputs_block, then the built-in output variables (vertex_position), each at
    a multiple of 16 bytes (kazan_b200/csrc/ptx_emit.cpp emit_ptx)"""
    size = module.user_outputs.size
    for v in module.built_in_outputs:
        size = (size + 15) // 16 * 16 + (_type_bytes(v.type) + 15) // 16 * 16
    return size


def run_dispatch(module, count: int, memory: Memory, user_inputs: int = 0, user_outputs: int = 0, first_invocation: int = 0,
                 global_size=(None, 1, 1), global_base=(0, 0, 0), push_constants: bytes = b"", descriptor_table: int = 0) -> None:
    """The dispatch of kzn_cmd_dispatch: invocation g (0 <= g < count) runs with index first_invocation + g; its interface
    blocks are base + index * block.size; the values its body is left with are stored to the members of user_outputs_block."""
    sim = Simt(module, memory)
    sx = global_size[0] if global_size[0] else max(count, 1)
    sy = global_size[1] or 1
    out_stride = output_slot_size(module)
    pc_addr = 0x6FFF_0000_0000
    memory.regions[pc_addr] = bytearray(push_constants.ljust(128, b"\0"))
    for g0 in range(0, count, SUBGROUP):
        lanes = list(range(g0, min(g0 + SUBGROUP, count)))
        env = {}
        for g in lanes:
            idx = first_invocation + g
            gx, q = idx % sx, idx // sx
            gy, gz = q % sy, q // sy
            gid = [(gx + global_base[0]) & 0xFFFFFFFF, (gy + global_base[1]) & 0xFFFFFFFF, (gz + global_base[2]) & 0xFFFFFFFF]
            built = {"global_invocation_id": struct.pack("<3I", *gid), "subgroup_local_invocation_id": struct.pack("<I", g % SUBGROUP),
                     "subgroup_size": struct.pack("<I", SUBGROUP), "push_constants": struct.pack("<Q", pc_addr),
                     "descriptor_table": struct.pack("<Q", descriptor_table)}
            e = sim._variables(g - g0, built)
            e[id(module.user_inputs.pointer)] = user_inputs + idx * module.user_inputs.size
            e[id(module.user_outputs.pointer)] = user_outputs + idx * out_stride
            env[g] = e
        res = sim.run_group(lanes, env)
        for g in lanes:
            base = user_outputs + (first_invocation + g) * out_stride
            for (off, ty), v in zip(module.user_outputs.members, res[g]):
                memory.write(base + off, iro.type_size(ty), int(v))
            at = module.user_outputs.size
            for v in module.built_in_outputs:  # the built-in output variables follow the user outputs
                at = (at + 15) // 16 * 16
                n = _type_bytes(v.type)
                src = sim.var_address[(g - g0, id(v.pointer))]
                memory.write(base + at, n, memory.read(src, n))
                at += (n + 15) // 16 * 16
    for k in [k for k in memory.regions if k >= LOCAL_BASE or k == pc_addr]:
        del memory.regions[k]


def run_fragments(module, coords, inputs, push_constants: bytes = b"", descriptor_table: int = 0, memory: Memory | None = None):
    """A fragment shader (kzn_fragment_program_create) over fragments: coords = [(x, y, z, w) floats], inputs = [(4 floats)].
    Returns [[4 output bit patterns]].  Fragments are independent; they run in groups of 32 only to share the machinery."""
    memory = memory or Memory()
    sim = Simt(module, memory)
    pc_addr, io = 0x6FFF_0000_0000, 0x6FFE_0000_0000
    memory.regions[pc_addr] = bytearray(push_constants.ljust(128, b"\0"))
    out = []
    for g0 in range(0, len(coords), SUBGROUP):
        lanes = list(range(g0, min(g0 + SUBGROUP, len(coords))))
        blocks = bytearray(64 * SUBGROUP)
        memory.regions[io] = blocks
        env = {}
        for g in lanes:
            k = g - g0
            built = {"frag_coord": np.asarray(coords[g], dtype=np.float32).tobytes(), "push_constants": struct.pack("<Q", pc_addr),
                     "descriptor_table": struct.pack("<Q", descriptor_table)}
            e = sim._variables(k, built)
            blocks[64 * k:64 * k + 16] = np.asarray(inputs[g], dtype=np.float32).tobytes()
            e[id(module.user_inputs.pointer)] = io + 64 * k
            e[id(module.user_outputs.pointer)] = io + 64 * k + 32
            env[g] = e
        res = sim.run_group(lanes, env)
        for g in lanes:
            k = g - g0
            vals = [int.from_bytes(blocks[64 * k + 32 + 4 * c:64 * k + 36 + 4 * c], "little") for c in range(4)]
            for c, v in enumerate(res[g][:4]):  # values the body returns override what it stored
                vals[c] = int(v)
            out.append(vals)
    for k in [k for k in memory.regions if k >= LOCAL_BASE or k in (pc_addr, io)]:
        del memory.regions[k]
    return out
