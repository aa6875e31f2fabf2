"""CPU: the C++ readers of the library must answer damaged input with a result code, never with a crash or a hang.

tests/fuzz_worker.py mutates the reference's SPIR-V fixtures and IR texts (bit flips, truncation, insertions, token swaps)
and feeds them to kzn_spirv_disassemble, kzn_translate_spirv_to_ir, kzn_ir_parse + kzn_ir_to_text + kzn_ptx_from_ir.  It
runs in a child process so that a segmentation fault or an endless loop shows up as a failed test, not as a dead test run.
(The first run of this test found two: an unchecked id bound that sized the id table, and literal strings that were not
checked for UTF-8 as the reference's String::from_utf8 does.)
"""
import json
import subprocess
import sys
from pathlib import Path

import pytest

WORKER = Path(__file__).resolve().parent / "fuzz_worker.py"


@pytest.mark.parametrize("seed", [20261017, 7])
def test_mutated_inputs_never_crash_the_readers(seed):
    r = subprocess.run([sys.executable, str(WORKER), str(seed), "700"], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, f"worker died with {r.returncode}: {r.stderr[-2000:]}"
    outcome = json.loads(r.stdout.strip().splitlines()[-1])
    assert outcome["error"] > 500 and outcome["ok"] > 10  # damaged input is rejected; harmless mutations still pass
