#!/bin/bash
# Quick GPU visit: parity tests, fill bench (profile mode), microbenchmarks.  gpurun --timeout 900 -- 'bash tools/gpu_quick.sh <tag>'
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
timeout 300 python bench.py --profile --steps 500 --warmup 10 > $OUT/bench_profile.json 2>$OUT/bench.err; cat $OUT/bench_profile.json | cut -c1-1500
timeout 300 python tools/prof_mesh.py > $OUT/mesh.log 2>&1; tail -2 $OUT/mesh.log | cut -c1-800
for m in tools/micro/*.cu; do b=${m%.cu}; if [ -x $b ]; then echo "== $b"; timeout 120 $b 2>&1 | tee $OUT/$(basename $b).log; fi; done
