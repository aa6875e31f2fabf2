#!/bin/bash
# ncu full capture of the mesh draw's kernels.  gpurun --timeout 900 -- 'bash tools/gpu_prof_mesh.sh <tag>'
TAG=${1:-p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 800 ncu --set full --clock-control none --import-source on -k regex:'raster_|setup_kernel|bin_' -s 4 -c 4 -f -o $OUT/prof_mesh python tools/prof_mesh.py > $OUT/ncu_mesh_full.log 2>&1
tail -3 $OUT/ncu_mesh_full.log
