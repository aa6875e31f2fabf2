import sys, json, random, base64, struct
import os; R=os.path.abspath(os.path.join(os.path.dirname(__file__),'..','..')); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import fuzz_worker as fw
seed=int(sys.argv[1]); n=int(sys.argv[2])
rng=random.Random(seed)
ROOT=fw.ROOT
spv=[fw.words_of(base64.b64decode(c["spv_base64"])) for c in json.loads((ROOT/"tests/golden/spirv_goldens.json").read_text())["cases"]]
ir=json.loads((ROOT/"tests/golden/ir_goldens.json").read_text())
spv.append(fw.words_of(bytes.fromhex(ir["spirv_to_ir"][0]["spv_hex"])))
texts=[p.read_text() for p in sorted((ROOT/"tests/ir_programs").glob("*.kazan-ir"))]
import tests.ir_random as irr
texts+=[irr.make_program(s)[0] for s in range(10)]
pairs=[(v["kind"],v["text"]) for v in ir["text_round_trips"]]
from tests import spirv_loop_shaders as loops
spv2=spv+[getattr(loops,k)() for k in ("variables","phis","do_while","nested","header_values_after_loop","calls","recursive")]
texts2=[p.read_text() for p in sorted((ROOT/"tests/ir_programs_v2").glob("*.kazan-ir"))]
out=open(sys.argv[3],'wb')
def rec(kind,payload):
    out.write(bytes([kind])+struct.pack('<I',len(payload))+payload)
for _ in range(n):
    w=fw.mutate_words(rng,rng.choice(spv))
    b=struct.pack(f'<{len(w)}I',*w)
    rec(0,b); rec(1,b)
    rec(2,fw.mutate_text(rng,rng.choice(texts)).encode())
    w2=fw.mutate_words(rng,rng.choice(spv2))
    rec(4,struct.pack(f'<{len(w2)}I',*w2))  # version 2 translator (loads, stores, selections, loops, scope repair) + printer + PTX
    rec(5,fw.mutate_text(rng,rng.choice(texts2)).encode())
    k,t=rng.choice(pairs)
    rec(3,k.encode()+b'\0'+fw.mutate_text(rng,t).encode())
