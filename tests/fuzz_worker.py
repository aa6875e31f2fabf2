"""Worker of tests/test_fuzz_parsers.py: feeds mutated inputs to the C++ readers of the library (SPIR-V disassembler, SPIR-V ->
IR translator, .kazan-ir text reader, IR -> PTX emitter).  Every call must come back with a result code; a crash of the
process is the failure the parent test looks for.     python tests/fuzz_worker.py <seed> <iterations>"""
import base64
import json
import random
import struct
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from kazan_b200 import backend, host  # noqa: E402


def words_of(data: bytes):
    data = data[:len(data) - len(data) % 4]
    return list(struct.unpack(f"<{len(data) // 4}I", data))


def mutate_words(rng, words):
    w = list(words)
    for _ in range(rng.randint(1, 4)):
        kind = rng.random()
        i = rng.randrange(len(w)) if w else 0
        if kind < 0.35 and w:
            w[i] ^= 1 << rng.randrange(32)
        elif kind < 0.55 and w:
            w[i] = rng.choice([0, 1, 0xFFFFFFFF, 0x7FFFFFFF, 0x00010000, 0xFFFF0000 | (w[i] & 0xFFFF), rng.getrandbits(32)])
        elif kind < 0.7 and w:
            del w[i:i + rng.randint(1, 4)]
        elif kind < 0.85:
            w[i:i] = [rng.getrandbits(32) for _ in range(rng.randint(1, 3))]
        elif w:
            w = w[:rng.randrange(len(w) + 1)]
    return w


def mutate_text(rng, text):
    t = list(text)
    for _ in range(rng.randint(1, 4)):
        kind = rng.random()
        i = rng.randrange(len(t)) if t else 0
        if kind < 0.3 and t:
            t[i] = rng.choice("{}[]<>:;,-=@\"x0123456789 \n\\%")
        elif kind < 0.5 and t:
            del t[i:i + rng.randint(1, 12)]
        elif kind < 0.7:
            t[i:i] = list(rng.choice(["loop", "block", "i32", "<4 x f32>", "data_ptr", "->", "fn", "0x", "-1", "struct", "\"", "{", "}",
                                         "18446744073709551616", "<4294967296 x i8>", "vscale x ", "<0 x i8>"]))
        elif kind < 0.85 and t:
            j = rng.randrange(len(t))
            t[i], t[j] = t[j], t[i]
        elif t:
            t = t[:rng.randrange(len(t) + 1)]
    return "".join(t)


def main(seed, iterations):
    rng = random.Random(seed)
    spv = [words_of(base64.b64decode(c["spv_base64"])) for c in
           json.loads((ROOT / "tests/golden/spirv_goldens.json").read_text())["cases"]]
    ir = json.loads((ROOT / "tests/golden/ir_goldens.json").read_text())
    from tests import spirv_loop_shaders as loops  # hand-assembled shaders with loops: the version 2 translator's corpus
    spv2 = spv + [getattr(loops, k)() for k in ("variables", "phis", "do_while", "nested", "header_values_after_loop", "calls", "recursive")]
    texts2 = [p.read_text() for p in sorted((ROOT / "tests/ir_programs_v2").glob("*.kazan-ir"))]
    texts = [p.read_text() for p in sorted((ROOT / "tests/ir_programs").glob("*.kazan-ir"))]
    texts += [v for v in json.dumps(ir).split('"') if "fn " in v or "block" in v][:20]
    pairs = [(v["kind"], v["text"]) for v in ir["text_round_trips"]]  # (grammar production, text) of the reference's vectors
    lib, C = backend.load_library(), backend.C
    outcomes = {"ok": 0, "error": 0}

    def attempt(fn):
        try:
            fn()
            outcomes["ok"] += 1
        except host.KazanError:
            outcomes["error"] += 1

    for _ in range(iterations):
        w = mutate_words(rng, rng.choice(spv))
        if w:
            attempt(lambda: backend.disassemble_spirv(w))
            attempt(lambda: backend.translate_spirv_to_ir(w, "main", rng.choice([0, 4, 5]))[0].destroy())
        w2 = mutate_words(rng, rng.choice(spv2))
        if w2:  # version 2: loads, stores, selections, loops, the scope repair
            def translate_v2():
                m = backend.translate_spirv_to_ir(w2, "main", rng.choice([0, 4, 5]), 2)[0]
                try:
                    m.to_text()
                    m.to_ptx()
                finally:
                    m.destroy()
            attempt(translate_v2)
        t2 = mutate_text(rng, rng.choice(texts2))

        def parse_and_emit_v2():
            m = backend.IrModule.parse(t2, 2)
            try:
                m.to_text()
                m.to_ptx()
            finally:
                m.destroy()
        attempt(parse_and_emit_v2)
        t = mutate_text(rng, rng.choice(texts))

        def parse_and_emit():
            m = backend.IrModule.parse(t)
            try:
                m.to_text()
                m.to_ptx()
            finally:
                m.destroy()
        attempt(parse_and_emit)
        kind, text = rng.choice(pairs)
        damaged = mutate_text(rng, text).encode()

        def round_trip():
            out, n = C.c_void_p(), C.c_size_t()
            backend._check(lib.kzn_ir_round_trip(kind.encode(), damaged, len(damaged), C.byref(out), C.byref(n)))
            backend._take_text(out, n)
        attempt(round_trip)
    print(json.dumps(outcomes))


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]))
