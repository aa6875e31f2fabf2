"""Extract the reference's own known-answer vectors for the .kazan-ir text format and the SPIR-V -> IR translator.

    python tests/golden/extract_ir_goldens.py        # needs /root/reference (read-only); writes ir_goldens.json

The reference (kazan-3d/kazan @ f2820899) pins this part of the hot path with in-crate Rust unit tests
(SURVEY.md §8c): text round trips `parse(text).display() == formatted` in
    shader-compiler-ir/src/{types,consts,target_properties,block,function,instructions_impl}.rs
and two golden files
    shader-compiler-ir/test_data/test_assert_ir_matches_file.kazan-ir
    shader-compiler-translate-spirv-to-ir/test_data/trivial_test.{spv,kazan-ir}.
This script pulls the (input text, expected canonical text) pairs out of those test sources — test vectors only, no
implementation code — so that both the Python oracle (oracle/ir_oracle.py) and the C++ reader in the product can be
checked against them on machines where /root/reference does not exist.
"""
from __future__ import annotations

import json
import re
import sys
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "ir_goldens.json"


def rust_string(lit: str) -> str:
    """Evaluate a Rust "..." literal (the escapes used in the tests: \\n \\" \\\\)."""
    assert lit[0] == '"' and lit[-1] == '"'
    out, i, s = [], 0, lit[1:-1]
    while i < len(s):
        if s[i] == "\\":
            n = s[i + 1]
            out.append({"n": "\n", '"': '"', "\\": "\\", "t": "\t", "0": "\0", "r": "\r", "'": "'"}[n])
            i += 2
        else:
            out.append(s[i])
            i += 1
    return "".join(out)


STR = r'"(?:[^"\\]|\\.)*"'


def macro_calls(text: str, name: str):
    """Yield the top-level argument strings of every `name!( ... )` invocation."""
    for m in re.finditer(re.escape(name) + r"!\s*\(", text):
        i = m.end()
        depth, args, cur = 1, [], []
        while depth:
            c = text[i]
            if c == '"':
                j = re.compile(STR).match(text, i).end()
                cur.append(text[i:j])
                i = j
                continue
            if c in "([{":
                depth += 1
            elif c in ")]}":
                depth -= 1
                if depth == 0:
                    break
            if c == "," and depth == 1:
                args.append("".join(cur).strip())
                cur = []
            else:
                cur.append(c)
            i += 1
        if "".join(cur).strip():
            args.append("".join(cur).strip())
        yield args


def literal_value(arg: str):
    """String value of a `"..."` or `concat!("...", ...)` argument, else None."""
    arg = arg.strip()
    if arg.startswith("concat!"):
        return "".join(rust_string(x) for x in re.findall(STR, arg))
    if re.fullmatch(STR, arg, flags=re.S):
        return rust_string(arg)
    return None


def pairs_from(path: str, macro: str, kind: str, kind_arg: int | None = None, first_line: int = 0):
    src = (REF / path).read_text()
    src = "\n".join(src.splitlines()[first_line:])
    # drop the macro_rules! definition itself
    out = []
    for args in macro_calls(src, macro):
        if not args or "$" in args[0]:
            continue
        text = literal_value(args[1])
        if text is None:
            continue
        formatted = literal_value(args[-1]) if len(args) > 3 and literal_value(args[-1]) is not None else text
        k = kind if kind_arg is None else args[kind_arg].strip().lower()
        out.append({"kind": k, "text": text, "formatted": formatted, "source": path})
    return out


def main() -> int:
    if not REF.exists():
        print("reference tree not present; keeping the committed ir_goldens.json", file=sys.stderr)
        return 1
    g = []
    g += pairs_from("shader-compiler-ir/src/types.rs", "test_type", "type", first_line=1230)
    g += pairs_from("shader-compiler-ir/src/consts.rs", "test_const", "const", first_line=595)
    g += pairs_from("shader-compiler-ir/src/target_properties.rs", "test_target_properties", "target_properties")
    g += pairs_from("shader-compiler-ir/src/block.rs", "test_from_to_text", "", kind_arg=2, first_line=780)
    g += pairs_from("shader-compiler-ir/src/function.rs", "test_from_to_text", "", kind_arg=2, first_line=490)
    # instructions_impl.rs:611-623 — `[value definitions]\ninstruction`
    src = (REF / "shader-compiler-ir/src/instructions_impl.rs").read_text().splitlines()[605:]
    for m in re.finditer(r"test_instruction_from_to_text!\s*\{(.*?)\n    \}", "\n".join(src), flags=re.S):
        text = "".join(rust_string(x) for x in re.findall(STR, m.group(1)))
        g.append({"kind": "definitions_then_instruction", "text": text, "formatted": text,
                  "source": "shader-compiler-ir/src/instructions_impl.rs"})
    for rel in ("shader-compiler-ir/test_data/test_assert_ir_matches_file.kazan-ir",
                "shader-compiler-translate-spirv-to-ir/test_data/trivial_test.kazan-ir"):
        t = (REF / rel).read_text()
        g.append({"kind": "module", "text": t, "formatted": t, "source": rel})
    spv = (REF / "shader-compiler-translate-spirv-to-ir/test_data/trivial_test.spv").read_bytes()
    doc = {"reference": "kazan-3d/kazan @ f2820899", "text_round_trips": g,
           "spirv_to_ir": [{"source": "shader-compiler-translate-spirv-to-ir/src/lib.rs:232-258 (trivial_test)",
                            "spv_hex": spv.hex(), "entry_point": "main", "execution_model": "GLCompute",
                            "expected_ir": (REF / "shader-compiler-translate-spirv-to-ir/test_data/trivial_test.kazan-ir").read_text()}]}
    OUT.write_text(json.dumps(doc, indent=1) + "\n")
    kinds = {}
    for e in g:
        kinds[e["kind"]] = kinds.get(e["kind"], 0) + 1
    print(f"wrote {len(g)} text vectors {kinds} + 1 SPIR-V vector to {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
