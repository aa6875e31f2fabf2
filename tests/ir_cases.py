"""IR programs for the compute-dispatch path (BASELINE configs[1]) with their invocation layouts and closed-form answers.

The programs (tests/ir_programs/*.kazan-ir) are written by hand from the 7 instructions shader-compiler-ir has
(instructions_impl.rs:395-433); `trivial_test` is the reference's own golden (translate/test_data/trivial_test.kazan-ir).
Each case: numpy dtypes of one user_inputs_block / user_outputs_block slot, a seeded input generator, and `expect`, the
closed form of what the program computes (size-independent check used at 2^24 invocations).
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
PROGRAMS = HERE / "ir_programs"
GOLDENS = json.loads((HERE / "golden" / "ir_goldens.json").read_text())


def program_text(name: str) -> str:
    if name == "trivial_test":
        return GOLDENS["spirv_to_ir"][0]["expected_ir"]
    return (PROGRAMS / f"{name}.kazan-ir").read_text()


def _rng(seed):
    return np.random.default_rng(seed)


def _fib64():
    a, b = 0, 1
    for _ in range(64):
        a, b = b, (a + b) & (2**64 - 1)
    return a


def _branchy(x):
    x = x.astype(np.uint32)
    y = x + np.uint32(1)
    r = np.where(x == 0, np.uint32(100), np.where(x == 1, np.uint32(200), y))
    s = np.where(x == 0, x, np.where(x == 1, np.uint32(0xFFFFFFFF), np.uint32(0)))
    t = r + s
    o0 = np.where(x == 7, x, t)
    o1 = np.where(x == 7, x, r)
    return o0, o1


IN_I32 = np.dtype([("x", "<u4")])
CASES = {
    # name: (input dtype, output dtype, make_inputs(n, seed) -> structured array, expect(inputs) -> structured array)
    "trivial_test": (np.dtype([]), np.dtype([]), None, None),
    "load_i32": (IN_I32, np.dtype([("y", "<u4")]), None, lambda i: i["x"]),
    "sum_loop": (IN_I32, np.dtype([("y", "<u4")]), None, lambda i: (i["x"].astype(np.uint64) * 256).astype(np.uint32)),
    "branchy": (IN_I32, np.dtype([("t", "<u4"), ("r", "<u4")]), "small", _branchy),
    "widths": (np.dtype([("c", "<u8")]), np.dtype([("c2", "<u8"), ("d2", "<u4"), ("b2", "<u2"), ("a3", "u1"), ("pad", "u1")]), None, None),
    "pointer_chase": (np.dtype([("p", "<u8")]), np.dtype([("w", "<u4")]), "pointers", None),
    "fibonacci": (np.dtype([("count", "u1")]), np.dtype([("f", "<u8")]), None, lambda i: np.full(len(i), _fib64(), dtype=np.uint64)),
    "bool_float": (np.dtype([("f", "<u4")]), np.dtype([("f", "<u4"), ("flag", "u1"), ("neg", "u1"), ("pad", "<u2")]), "bool_byte", None),
}


def make_inputs(name: str, n: int, seed: int = 1, table_address: int = 0, table_len: int = 0) -> np.ndarray:
    dt, _, kind, _ = CASES[name]
    r = _rng(seed)
    out = np.zeros(n, dtype=dt)
    if dt.itemsize == 0:
        return out
    if kind == "small":
        out["x"] = r.integers(0, 10, n, dtype=np.uint32)
    elif kind == "pointers":
        out["p"] = table_address + 4 * r.integers(0, table_len, n, dtype=np.uint64)
    else:
        raw = r.integers(0, 256, n * dt.itemsize, dtype=np.uint8)
        out = raw.view(dt).copy()
        if kind == "bool_byte":  # a bool in memory is the byte 0 or 1; anything else has no defined meaning
            out["f"] = (out["f"] & np.uint32(0xFFFFFF00)) | r.integers(0, 2, n, dtype=np.uint32)
    return out


def widths_expect(i):
    c = i["c"]
    a = (c & np.uint64(0xFF)).astype(np.uint8)
    b = (c & np.uint64(0xFFFF)).astype(np.uint16)
    d = (c & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    return (c + c, d + np.uint32(0xFFFFFFFF), b + np.uint16(0x7FFF), (a + a + np.uint8(0xFF)).astype(np.uint8))


CASES["widths"] = CASES["widths"][:3] + (widths_expect,)


def bool_float_expect(i):
    flag = ((i["f"] & np.uint32(0xFF)) != 0).astype(np.uint8)
    return (i["f"], flag, (1 - flag).astype(np.uint8))


CASES["bool_float"] = CASES["bool_float"][:3] + (bool_float_expect,)


def interpret(name: str, inputs: np.ndarray, extra_memory: dict | None = None) -> np.ndarray:
    """Run the oracle interpreter once per invocation; returns the outputs as a structured array."""
    from oracle import ir_oracle as iro
    m = iro.parse_module(program_text(name))
    dt_in, dt_out, _, _ = CASES[name]
    n = len(inputs)
    out = np.zeros(n, dtype=dt_out)
    raw_in = inputs.view(np.uint8).reshape(n, dt_in.itemsize) if dt_in.itemsize else None
    raw_out = out.view(np.uint8).reshape(n, dt_out.itemsize) if dt_out.itemsize else None
    members = m.user_outputs.members
    base_in, base_out = 0x10000000, 0x20000000
    for g in range(n):
        mem = dict(extra_memory or {})
        if raw_in is not None:
            mem[base_in] = raw_in[g].tobytes()
        vals = iro.Interpreter(m, mem).run_entry(base_in, base_out) or []
        for (off, ty), v in zip(members, vals):
            sz = iro.type_size(ty)
            raw_out[g, off:off + sz] = np.frombuffer(int(v).to_bytes(sz, "little"), dtype=np.uint8)
    return out
