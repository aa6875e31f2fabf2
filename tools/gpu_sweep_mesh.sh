#!/bin/bash
# Build-and-time sweep of a raster_kernel tunable on the mesh frame.  gpurun --timeout 1200 -- 'bash tools/gpu_sweep_mesh.sh <tag> "<-D flags 1>" "<-D flags 2>" ...'
TAG=${1:-sweep}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
i=0
for cfg in "$@"; do
  i=$((i+1))
  KZN_NVCC_EXTRA="$cfg" python -m kazan_b200.build --force > $OUT/build_$i.log 2>&1 || { echo "build failed: $cfg"; tail -5 $OUT/build_$i.log; continue; }
  echo "== $cfg"
  timeout 300 python tools/prof_mesh.py 2>>$OUT/err.log | cut -c1-120
done
python -m kazan_b200.build --force > /dev/null 2>&1
