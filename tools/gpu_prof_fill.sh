#!/bin/bash
# ncu full capture of the fill kernel.  gpurun --timeout 900 -- 'bash tools/gpu_prof_fill.sh <tag>'
TAG=${1:-p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:raster_ -s 3 -c 1 -f -o $OUT/prof_fill python bench.py --steps 5 --warmup 3 --profile > $OUT/ncu_fill_full.log 2>&1
tail -3 $OUT/ncu_fill_full.log
