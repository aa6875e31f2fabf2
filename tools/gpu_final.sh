#!/bin/bash
# Short closing visit: parity tests, smoke, the bench line and the reference arm (no ncu).  gpurun --timeout 900 -- 'bash tools/gpu_final.sh <tag>'
TAG=${1:-final}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 400 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cut -c1-400 $OUT/bench.json
timeout 200 python bench.py --impl reference --steps 5 --warmup 3 > $OUT/bench_reference.json 2>> $OUT/bench.err; cut -c1-200 $OUT/bench_reference.json
