#!/bin/bash
# compute-sanitizer over a subset of the GPU parity tests (memcheck, then racecheck on the raster tests)
OUT=gpurun_out/${1:-san}
mkdir -p $OUT
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_raster.py tests/test_gpu_transfer.py tests/test_gpu_dispatch.py tests/test_gpu_command_buffer.py tests/test_gpu_ir_v2.py tests/test_gpu_present.py -m gpu -x -q -k "not 2_pow_24 and not 4k and not full_size" > $OUT/memcheck.log 2>&1; echo "memcheck exit $?" | tee -a $OUT/memcheck.log; tail -5 $OUT/memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_raster.py -m gpu -x -q > $OUT/racecheck.log 2>&1; echo "racecheck exit $?" | tee -a $OUT/racecheck.log; tail -5 $OUT/racecheck.log
