#!/bin/bash
# Transfer iteration visit: transfer parity tests + the transfer sweep of bench.py.  gpurun --timeout 900 -- 'bash tools/gpu_transfer_iter.sh <tag>'
TAG=${1:-t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_transfer.py tests/test_gpu_command_buffer.py -m gpu -x -q > $OUT/pytest_transfer.log 2>&1; tail -4 $OUT/pytest_transfer.log
timeout 600 python - > $OUT/transfer.json 2> $OUT/err.txt <<P
import sys, json; sys.path.insert(0,".")
import torch, bench
from kazan_b200 import abi, host
dev=host.Device(0); s=torch.cuda.Stream(); q=dev.wrap_stream(s.cuda_stream)
peak,_=bench.measured_peaks()
with torch.cuda.stream(s):
    print(json.dumps(bench.transfer_extra(torch, abi, host, dev, q, s, peak)))
P
tail -3 $OUT/err.txt
python - <<P
import json
d=json.load(open("$OUT/transfer.json"))
for k,v in d.items():
    print(k, {a:b for a,b in v.items() if "frac" in a})
    for kk in ("copy_b2i_variants_submit","blit_and_swizzle"):
        if kk in v: print("   ", {a:b["frac_hbm"] for a,b in v[kk].items()})
P
