"""Extract the reference's SPIR-V disassembly known-answer tests.

    python tests/golden/extract_spirv_goldens.py     # needs /root/reference (read-only); writes spirv_goldens.json

spirv-parser/src/lib.rs:46-215 parses test_inputs/test.spv, test2.spv and test3.spv and compares the dump of the header
and of every instruction with an expected string, byte for byte.  This script pulls the three (module bytes, expected
text) pairs out of the reference — test vectors only, no implementation code — so that the C++ reader
(kazan_b200/csrc/spirv_dis.cpp, through kzn_spirv_disassemble) can be checked against them where /root/reference does
not exist.
"""
from __future__ import annotations

import base64
import hashlib
import json
import re
from pathlib import Path

REF = Path("/root/reference/spirv-parser")
OUT = Path(__file__).resolve().parent / "spirv_goldens.json"


def main():
    src = (REF / "src/lib.rs").read_text()
    cases = []
    for m in re.finditer(r'fn (parse_test\d*)\(\).*?include_bytes!\("\.\./test_inputs/([^"]+)"\).*?let expected = r#"(.*?)"#;', src, re.S):
        name, file, expected = m.groups()
        data = (REF / "test_inputs" / file).read_bytes()
        line = src[:m.start()].count("\n") + 1
        cases.append({"test": name, "source": f"spirv-parser/src/lib.rs:{line}", "file": file,
                      "sha256": hashlib.sha256(data).hexdigest(), "spv_base64": base64.b64encode(data).decode(),
                      "expected": expected})
    assert len(cases) == 3, len(cases)
    OUT.write_text(json.dumps({"reference": "kazan-3d/kazan spirv-parser/src/lib.rs:46-215", "cases": cases}, indent=1) + "\n")
    print(f"wrote {OUT} ({len(cases)} cases)")


if __name__ == "__main__":
    main()
