#!/bin/bash
# Build-and-bench sweep of -D variants on the 4K fill.  gpurun --timeout 1200 -- 'bash tools/gpu_sweep_fill.sh <tag> "<flags 1>" "<flags 2>" ...'
TAG=${1:-sweep}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
i=0
for cfg in "$@"; do
  i=$((i+1))
  KZN_NVCC_EXTRA="$cfg" python -m kazan_b200.build --force > $OUT/build_$i.log 2>&1 || { echo "build failed: $cfg"; tail -5 $OUT/build_$i.log; continue; }
  timeout 300 python bench.py --profile --steps 400 --warmup 10 > $OUT/bench_$i.json 2>>$OUT/bench.err
  python - $OUT/bench_$i.json "$cfg" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[2] or '(default)', 'value', round(d['value'],1), 'kernel_us', round(d['roofline']['kernel_ms']*1e3,2), 'frac', round(d['roofline']['frac'],3))
PY
done
python -m kazan_b200.build --force > /dev/null 2>&1
