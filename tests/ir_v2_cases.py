"""IR programs of version 2 (kazan_b200/csrc/ir.h `kExtOps`) with their invocation layouts and seeded inputs.

tests/ir_programs_v2/*.kazan-ir are written by hand; the checker is the SIMT interpreter (oracle/ir_simt.py), and where a
closed form exists (saxpy on exactly representable values, the ballot of a known pattern) it is checked as well."""
from __future__ import annotations

from pathlib import Path

import numpy as np

PROGRAMS = Path(__file__).resolve().parent / "ir_programs_v2"


def program_text(name: str) -> str:
    return (PROGRAMS / f"{name}.kazan-ir").read_text()


CASES = {
    "int_ops": (np.dtype([("a", "<u4"), ("b", "<u4"), ("c", "<u8"), ("d", "<u2"), ("e", "u1"), ("f", "u1"), ("sh", "<u4")]),
                np.dtype([("r0", "<u4"), ("r1", "<u4"), ("r2", "<u4"), ("r3", "<u4"), ("r4", "<u4"), ("r5", "<u4"), ("r6", "<u8"),
                          ("r7", "<u8"), ("r8", "<u2"), ("r9", "u1"), ("r10", "u1"), ("r11", "<u4"), ("r12", "<u4"), ("r13", "<u4"),
                          ("r14", "<u8")])),
    "float_ops": (np.dtype([("a", "<f4"), ("b", "<f4"), ("c", "<f8"), ("n", "<u4"), ("w", "<f4")]),
                  np.dtype([("r0", "<u4"), ("r1", "<u4"), ("r2", "<u4"), ("r3", "<u4"), ("r4", "<u8"), ("r5", "<u8"), ("r6", "<u4"),
                            ("r7", "<u4"), ("r8", "<u8"), ("r9", "<u4"), ("r10", "<u4"), ("r11", "<u4"), ("r12", "<u2"), ("r13", "u1"),
                            ("r14", "u1")])),
    "subgroup_ops": (np.dtype([("v", "<u4")]), np.dtype([(f"r{i}", "<u4") for i in range(9)])),
}


def make_inputs(name: str, n: int, seed: int = 1) -> np.ndarray:
    dt = CASES[name][0]
    r = np.random.default_rng(seed)
    out = np.zeros(n, dtype=dt)
    if name == "int_ops":
        raw = r.integers(0, 256, n * dt.itemsize, dtype=np.uint8)
        out = raw.view(dt).copy()
        out["sh"] = r.choice(np.array([0, 1, 5, 31, 32, 33, 63, 64, 200, 2**31], dtype=np.uint32), n)  # over-wide shifts included
        out["a"][:4] = [0, 1, 0x80000000, 0xFFFFFFFF]
        out["b"][:4] = [0xFFFFFFFF, 0x7FFFFFFF, 0x80000000, 1]
    elif name == "float_ops":
        sign = r.choice(np.array([-1.0, 1.0], dtype=np.float32), n)
        out["a"] = (sign * np.exp(r.uniform(-3, 5, n))).astype(np.float32)
        out["b"] = (r.choice(np.array([-1.0, 1.0], dtype=np.float32), n) * np.exp(r.uniform(-3, 5, n))).astype(np.float32)
        out["w"] = r.uniform(-50, 50, n).astype(np.float32)
        out["c"] = r.uniform(-4, 4, n)
        out["n"] = r.integers(0, 2**32, n, dtype=np.uint32)
        out["a"][:3] = [0.0, -0.0, 1.0]
        out["b"][:3] = [-0.0, 0.0, 3.0]
    elif name == "subgroup_ops":
        out["v"] = r.integers(0, 2**32, n, dtype=np.uint32)
        out["v"][32:64] |= 1  # one subgroup whose values are all odd
        out["v"][64:96] &= ~np.uint32(1)  # ... and one with none
    return out


def interpret(name: str, inputs: np.ndarray) -> np.ndarray:
    """The SIMT interpreter over all invocations; outputs as a structured array."""
    from oracle import ir_oracle as iro
    from oracle import ir_simt as sim
    dt_in, dt_out = CASES[name]
    m = iro.parse_module(program_text(name), 2)
    n = len(inputs)
    base_in, base_out = 0x10000000, 0x20000000
    out = bytearray(n * dt_out.itemsize)
    mem = sim.Memory({base_in: bytearray(inputs.tobytes()), base_out: out})
    sim.run_dispatch(m, n, mem, base_in, base_out)
    return np.frombuffer(bytes(out), dtype=dt_out)


def same_floats(a: np.ndarray, b: np.ndarray) -> bool:
    """bit patterns equal, any NaN equal to any NaN (NaN payloads are not part of the contract: ir.h)"""
    w = a.dtype.itemsize
    fa, fb = a.view(f"<f{w}"), b.view(f"<f{w}")
    return bool(np.all((a == b) | (np.isnan(fa) & np.isnan(fb))))
