#!/bin/bash
# Build-and-bench sweep of the direct kernel's tunables on the GPU box.  gpurun --timeout 1200 -- 'bash tools/gpu_sweep_direct.sh <tag>'
TAG=${1:-sweep}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
[ $# -eq 0 ] && set -- "4 16" "4 20" "4 24" "2 20" "8 20" "4 12"
for cfg in "$@"; do
  rows=${cfg% *}; warps=${cfg#* }
  KZN_NVCC_EXTRA="-DKZN_DIRECT_ROWS=$rows -DKZN_DIRECT_WARPS=$warps" python -m kazan_b200.build --force > $OUT/build_${rows}_${warps}.log 2>&1 || { echo "build failed $cfg"; continue; }
  timeout 300 python bench.py --profile --steps 400 --warmup 10 > $OUT/bench_${rows}_${warps}.json 2>>$OUT/bench.err
  python - $OUT/bench_${rows}_${warps}.json "$cfg" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[2], 'value', round(d['value'],1), 'kernel_ms', round(d['roofline']['kernel_ms']*1e3,2), 'frac', round(d['roofline']['frac'],3))
PY
done
python -m kazan_b200.build --force > /dev/null 2>&1
