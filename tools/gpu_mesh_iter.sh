#!/bin/bash
# Raster iteration visit: raster parity tests, the 1M-triangle and 8K frames with full-size parity, launch lists.
# gpurun --timeout 900 -- 'bash tools/gpu_mesh_iter.sh <tag>'
TAG=${1:-m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_raster.py tests/test_gpu_split_frame.py tests/test_gpu_command_buffer.py -m gpu -x -q > $OUT/pytest_raster.log 2>&1; tail -4 $OUT/pytest_raster.log
timeout 300 python - > $OUT/mesh.log 2>&1 <<P
import sys; sys.path.insert(0,".")
import torch, bench, json
from kazan_b200 import abi, host
dev=host.Device(0); s=torch.cuda.Stream(); q=dev.wrap_stream(s.cuda_stream)
with torch.cuda.stream(s):
    print(json.dumps(bench.mesh_extra(torch, abi, host, q, s, check_oracle=True)))
    print(json.dumps(bench.mesh_extra(torch, abi, host, q, s, quads=(2000, 1000), size=(7680, 4320), seed=5, check_oracle=("$2" != "fast"))))
P
cut -c1-900 $OUT/mesh.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_mesh_raw.csv python tools/prof_mesh.py > $OUT/ncu_mesh.log 2>&1
python tools/ncu_summary.py list $OUT/launches_mesh_raw.csv $OUT/launches_mesh.csv; cat $OUT/launches_mesh.csv
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_8k_raw.csv python tools/prof_8k.py > $OUT/ncu_8k.log 2>&1
python tools/ncu_summary.py list $OUT/launches_8k_raw.csv $OUT/launches_8k.csv; cat $OUT/launches_8k.csv
