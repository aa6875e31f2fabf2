#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list, ncu full capture of the dominant kernels.
# Usage (from the repo root): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cat $OUT/bench.json | cut -c1-3000
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $OUT/bench_reference.json 2>> $OUT/bench.err; cat $OUT/bench_reference.json | cut -c1-600
# launch list of the same command (short run)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_fill.csv python bench.py --steps 20 --warmup 3 --profile > $OUT/ncu_fill.log 2>&1
# full capture of the dominant kernel of the fill bench
timeout 600 ncu --set full --clock-control none --import-source on -k regex:raster_ -s 3 -c 2 -f -o $OUT/prof_fill python bench.py --steps 5 --warmup 3 --profile > $OUT/ncu_fill_full.log 2>&1
# mesh: launch list + full capture of its kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_mesh.csv python tools/prof_mesh.py > $OUT/ncu_mesh.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'raster_kernel|setup_kernel|bin_' -s 4 -c 4 -f -o $OUT/prof_mesh python tools/prof_mesh.py > $OUT/ncu_mesh_full.log 2>&1
# transfer kernels: full capture
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fill_|copy_bulk|copy_realign|box_kernel|blit_kernel|merged_transfer' -c 28 -f -o $OUT/prof_transfer python tools/prof_transfer.py > $OUT/ncu_transfer_full.log 2>&1
# the 8K / 4M-triangle frame: launch list
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_8k.csv python tools/prof_8k.py > $OUT/ncu_8k.log 2>&1
# one band (1/8 of the rows) of the same frame: what each rank of an 8-GPU split frame runs
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_8k_band.csv python tools/prof_8k_band.py > $OUT/ncu_8k_band.log 2>&1
# DRAM traffic of every kernel of the two mesh frames (roofline.traffic of extra.mesh_1M_1080p / split_frame_8k)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s 40 -c 14 --csv --log-file $OUT/traffic_mesh.csv python tools/prof_mesh.py > $OUT/ncu_traffic_mesh.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s 40 -c 14 --csv --log-file $OUT/traffic_8k.csv python tools/prof_8k.py > $OUT/ncu_traffic_8k.log 2>&1
# the same with the caches as the stream leaves them (what roofline.traffic quotes): frames are 5 kernels now (the clears happen inside raster_kernel)
timeout 900 ncu --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s 40 -c 15 --csv --log-file $OUT/traffic_mesh_in_stream.csv python tools/prof_mesh.py > $OUT/ncu_traffic_mesh2.log 2>&1
timeout 900 ncu --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s 40 -c 15 --csv --log-file $OUT/traffic_8k_in_stream.csv python tools/prof_8k.py > $OUT/ncu_traffic_8k2.log 2>&1
ls -la $OUT
