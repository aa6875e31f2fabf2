#!/bin/bash
# Short GPU visit: parity tests + the default bench line.  gpurun --timeout 900 -- 'bash tools/gpu_visit.sh <tag>'
TAG=${1:-v}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 700 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -6 $OUT/pytest_gpu.log
timeout 500 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -3 $OUT/bench.err
python - <<P
import json
d=json.load(open("$OUT/bench.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["roofline"]["frac"], d["e2e"]["value"])
for k,v in d["extra"]["transfer"].items(): print(k, {a:b for a,b in v.items() if "submit" in a})
print(d["extra"]["mesh_1M_1080p"]["ms_per_frame"], d["extra"]["split_frame_8k_4Mtri"]["ms_per_frame"])
P
