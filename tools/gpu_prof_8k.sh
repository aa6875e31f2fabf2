#!/bin/bash
# ncu full capture of the 8K / 4M-triangle frame's raster kernel.  gpurun --timeout 900 -- 'bash tools/gpu_prof_8k.sh <tag>'
TAG=${1:-p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 800 ncu --set full --clock-control none --import-source on -k regex:'raster_kernel' -s 2 -c 1 -f -o $OUT/prof_mesh python tools/prof_8k.py > $OUT/ncu_8k_full.log 2>&1
tail -3 $OUT/ncu_8k_full.log
